# ncu --set full of site_fast_kernel at the initial point of C3 (deferred Adam active), source-correlated page exported on the box
mkdir -p gpurun_out
CMD="python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:site_fast_kernel -s 4 -c 1 \
    -o gpurun_out/sites_c3_r2 $CMD > gpurun_out/ncu_sites_c3_r2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_sites_c3_r2.log
ncu -i gpurun_out/sites_c3_r2.ncu-rep --page raw --csv > gpurun_out/sites_c3_r2_raw.csv 2>/dev/null
ncu -i gpurun_out/sites_c3_r2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/sites_c3_r2_src.csv 2>/dev/null
rm -f gpurun_out/sites_c3_r2.ncu-rep
ls -la gpurun_out/sites_c3_r2*

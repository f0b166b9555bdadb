# ncu --set full of the single-bin likelihood kernel at C2 with the source-correlated page exported on the box
mkdir -p gpurun_out
CMD="python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:ksmogn_stream_kernel -s 4 -c 1 \
    -o gpurun_out/ksmogn_src_r2 $CMD > gpurun_out/ncu_ksmogn_src_r2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_ksmogn_src_r2.log
ncu -i gpurun_out/ksmogn_src_r2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/ksmogn_src_r2.csv 2>/dev/null
rm -f gpurun_out/ksmogn_src_r2.ncu-rep
ls -la gpurun_out/ksmogn_src_r2*

# ncu capture recipe of round 2 (B200_PROFILING.md): launch list + one full capture of the dominant kernels.
# Run on the GPU box:  bash profiles/capture_r2.sh   (outputs land in gpurun_out/)
mkdir -p gpurun_out
KREG='regex:ksmogn|site_|local_post|globals_|adam_kernel|step_advance|subsample|cosmos_fused'
for W in c3 c2; do
  CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0"
  # (1) every launch of this library with its device time (cold-cache, serialised: compare SHARES)
  $CMD > gpurun_out/plain_r2_$W.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 33 -c 33 --csv \
      --log-file gpurun_out/launches_r2_$W.csv $CMD > gpurun_out/ncu_launch_r2_$W.log 2>&1
  echo "$W launch list rc=$?"
  # (2) full capture of the likelihood, site and post kernels (one launch each, after warm-up)
  ncu --set full --clock-control none --import-source on -k regex:'ksmogn_stream_kernel|site_fast_kernel|local_post_kernel' -s 12 -c 3 \
      -o gpurun_out/step_r2_$W $CMD > gpurun_out/ncu_full_r2_$W.log 2>&1
  echo "$W full rc=$?"
  ncu -i gpurun_out/step_r2_$W.ncu-rep --page raw --csv > gpurun_out/step_r2_${W}_raw.csv 2>/dev/null
  [ "$W" = c2 ] && rm -f gpurun_out/step_r2_$W.ncu-rep   # gpurun_out/ travels back only below 64 MiB: keep the C3 report, CSVs of the rest
done
# (3) the many-bins form of the likelihood kernel: C2 with a 64-bin offset histogram
CMD3="python bench.py --workload c2 --offset-hist 64 --steps 3 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0"
$CMD3 > gpurun_out/plain_r2_o64.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ksmogn_stream_kernel' -s 4 -c 1 \
    -o gpurun_out/ksmogn_o64_r2 $CMD3 > gpurun_out/ncu_o64_r2.log 2>&1
echo "o64 rc=$?"
ncu -i gpurun_out/ksmogn_o64_r2.ncu-rep --page raw --csv > gpurun_out/ksmogn_o64_r2_raw.csv 2>/dev/null; rm -f gpurun_out/ksmogn_o64_r2.ncu-rep
# (4) three distinct bins (register-cached form)
CMD4="python bench.py --workload c2 --keep-offset-bins --steps 3 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0"
ncu --set full --clock-control none --import-source on -k regex:'ksmogn_stream_kernel' -s 4 -c 1 \
    -o gpurun_out/ksmogn_o3_r2 $CMD4 > gpurun_out/ncu_o3_r2.log 2>&1
echo "o3 rc=$?"
ncu -i gpurun_out/ksmogn_o3_r2.ncu-rep --page raw --csv > gpurun_out/ksmogn_o3_r2_raw.csv 2>/dev/null; rm -f gpurun_out/ksmogn_o3_r2.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | head -30

# one full ncu capture each of the likelihood kernel and the site kernel (warm-up launches skipped)
mkdir -p gpurun_out
CMD="python profiles/kernel_times.py c2 2"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ksmogn_fast_kernel|site_kernel|local_post_kernel' -s 9 -c 3 -o gpurun_out/step_r1b $CMD > gpurun_out/ncu_full2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full2.log

# ncu capture recipe (B200_PROFILING.md): launch list + one full capture of the dominant kernels.
# Run on the GPU box:  bash profiles/capture.sh <tag>   (outputs land in gpurun_out/)
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
# (1) every launch of this library with its device time (cold-cache, serialised: compare SHARES)
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ksmogn|site_|local_post|reduce_|globals_|adam_kernel|finalize_loss|step_advance|subsample' -s 40 -c 40 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
# (2) full capture of the likelihood, site and post kernels (one launch each, after warm-up)
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ksmogn_stream_kernel|site_fast_kernel|local_post_kernel' -s 12 -c 3 \
    -o gpurun_out/step_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_full_${TAG}.log
# (3) the same likelihood kernel with the simulator's three offset bins kept distinct (O = 3 forms)
CMD3="$CMD --keep-offset-bins"
$CMD3 > gpurun_out/plain3_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ksmogn_stream_kernel' -s 4 -c 1 \
    -o gpurun_out/ksmogn_o3_${TAG} $CMD3 > gpurun_out/ncu_o3_${TAG}.log 2>&1
echo "o3 rc=$?"
# (4) the guide-site kernels of a TRAINED model (2000 SVI iterations before the capture)
CMD4="$CMD --train-iters 2000"
ncu --set full --clock-control none --import-source on -k regex:'site_fast_kernel|site_worklist_kernel' -s 4010 -c 2 \
    -o gpurun_out/sites_trained_${TAG} $CMD4 > gpurun_out/ncu_sites_trained_${TAG}.log 2>&1
echo "trained rc=$?"

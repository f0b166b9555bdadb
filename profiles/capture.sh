# ncu capture recipe (B200_PROFILING.md): launch list + one full capture of the dominant kernel.
# Run on the GPU box:  bash profiles/capture.sh   (outputs land in gpurun_out/)
mkdir -p gpurun_out
set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tq|ksmogn|adam|local|globals|reduce|step_advance|subsample' -s 60 -c 40 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ksmogn_kernel -s 4 -c 2 -o gpurun_out/ksmogn_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log

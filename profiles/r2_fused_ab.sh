# fused step kernel vs the per-stage kernels: tests, then A/B timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q > gpurun_out/r2d_fused_tests.log 2>&1; echo "fused tests rc=$?"; tail -15 gpurun_out/r2d_fused_tests.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "all tests rc=$?"; tail -5 gpurun_out/r2d_tests.log
for w in c2 c3; do
  for f in 1 0; do
    TQ_FUSED=$f timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 1000 > gpurun_out/r2d_bench_${w}_fused$f.json 2> gpurun_out/r2d_bench_${w}_fused$f.err; echo "$w fused=$f rc=$?"
  done
done

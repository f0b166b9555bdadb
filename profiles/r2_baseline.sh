# round-2 starting point: never-measured configurations of round 1 (O=64, c2mb, c3) + per-kernel times
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"
for w in c2 c2mb c3; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --trained-iters 0 > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err; echo "$w rc=$?"
done
python bench.py --offset-hist 64 --steps 10 --warmup 3 --no-cpu-baseline --trained-iters 0 > gpurun_out/r2a_bench_o64.json 2> gpurun_out/r2a_bench_o64.err; echo "o64 rc=$?"
python bench.py --keep-offset-bins --steps 10 --warmup 3 --no-cpu-baseline --trained-iters 0 > gpurun_out/r2a_bench_o3.json 2> gpurun_out/r2a_bench_o3.err; echo "o3 rc=$?"
python profiles/kernel_times.py c2 10 > gpurun_out/r2a_ktimes_c2.txt 2>&1
python profiles/kernel_times.py c2mb 10 > gpurun_out/r2a_ktimes_c2mb.txt 2>&1
python profiles/kernel_times.py c3 5 > gpurun_out/r2a_ktimes_c3.txt 2>&1
tail -3 gpurun_out/r2a_tests.log

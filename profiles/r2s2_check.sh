# session-2 check of round 2: GPU suite, per-kernel times of one rank's C3 shard at 8 GPUs (c3s8, on one GPU), its step time
mkdir -p gpurun_out
TAG=${1:-s2}
(timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log)
python profiles/kernel_times.py c3s8 10 > gpurun_out/${TAG}_kt_c3s8.txt 2>&1; tail -20 gpurun_out/${TAG}_kt_c3s8.txt
timeout 300 python bench.py --workload c3s8 --steps 10 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0 > gpurun_out/${TAG}_bench_c3s8.json 2> gpurun_out/${TAG}_bench_c3s8.err
tail -c 1500 gpurun_out/${TAG}_bench_c3s8.json

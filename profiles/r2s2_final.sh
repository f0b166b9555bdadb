# what the driver runs at round end, on one GPU: smoke, GPU suite, the default bench line and the reference arm (wall times printed)
mkdir -p gpurun_out
TAG=${1:-final}
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/${TAG}_smoke.log
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/${TAG}_tests.log
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "reference arm rc=$? $(( $(date +%s) - t0 )) s"
tail -c 600 gpurun_out/${TAG}_bench_ref.json

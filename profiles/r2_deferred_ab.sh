# deferred AOI-local Adam (engine.deferred_adam, tq_cosmos_sites_adam) against the separate dense Adam launch:
# tests, then step time on C3 / C2 / C4 / one 8-GPU shard with TQ_DEFERRED_ADAM=1 / 0
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deferred_adam_gpu.py -x -q > gpurun_out/deferred_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/deferred_tests.log
for w in ${WORKLOADS:-c3 c2 c4 c3s8}; do
  for d in 1 0; do
    TQ_DEFERRED_ADAM=$d timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-subs --trained-iters 0 \
        > gpurun_out/dab_${w}_$d.json 2> gpurun_out/dab_${w}_$d.err
    python - <<P
import json
try:
    r = json.loads(open("gpurun_out/dab_${w}_$d.json").read().strip().splitlines()[-1])
    print("$w deferred=$d ms_per_step", r["ms_per_step"], "value", r["value"], "launches", r["gpu_launches"], "kernel_ms", r["roofline"].get("kernel_ms"))
except Exception as e:
    print("$w deferred=$d failed", e)
P
  done
done

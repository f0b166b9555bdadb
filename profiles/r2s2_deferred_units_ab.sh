# threshold of the deferred AOI-local Adam (TQ_DEFERRED_ADAM_MIN_UNITS): one 8-GPU / 4-GPU / 2-GPU rank's C3 shard and C2 on one GPU,
# deferred for every launch size (1) against the default threshold (2^20 units)
mkdir -p gpurun_out
for w in ${WORKLOADS:-c3s8 c3s4 c2}; do
  for m in 1 1048576; do
    TQ_DEFERRED_ADAM_MIN_UNITS=$m timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-subs --trained-iters 0 \
        > gpurun_out/dmu_${w}_$m.json 2> gpurun_out/dmu_${w}_$m.err
    python - <<P
import json
try:
    r = json.loads(open("gpurun_out/dmu_${w}_$m.json").read().strip().splitlines()[-1])
    print("$w min_units=$m ms_per_step", round(r["ms_per_step"], 5), "value", round(r["value"] / 1e6, 1), "M launches", r["gpu_launches"], "kernel_ms", r["roofline"].get("kernel_ms"))
except Exception as e:
    print("$w min_units=$m failed", e)
P
  done
done

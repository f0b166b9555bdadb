#!/bin/bash
# A/B of the likelihood kernel's scheduling forms (TQ_KSMOGN_MODE: 0 one block per 16 patches, 1 persistent warps
# with static striding + cp.async prefetch, 2 the same with a dynamic work counter).  Run on the GPU box.
mkdir -p gpurun_out
MODES="${MODES:-0 1 2}"
for m in $MODES; do
  [ "$m" = 0 ] && continue
  TQ_KSMOGN_MODE=$m timeout 600 python -m pytest tests -m gpu -x -q -k "ksmogn or step or api" > gpurun_out/ab_test_m$m.log 2>&1
  tail -3 gpurun_out/ab_test_m$m.log
done
for m in $MODES; do
  for w in c2 c3; do
    TQ_KSMOGN_MODE=$m timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --trained-iters 0 > gpurun_out/ab_${w}_m$m.json 2> gpurun_out/ab_${w}_m$m.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${w}_m$m.json").read().strip().splitlines()[-1])
    print("mode $m $w", d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e:
    print("mode $m $w failed", e)
PY
  done
done

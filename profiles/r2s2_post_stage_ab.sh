# staged (cp.async prefetch of the next unit's inputs) against unstaged local_post: GPU suite on the default build, then
# per-kernel times (torch profiler) and step times of C3 and of one 8-GPU rank's shard with both builds
mkdir -p gpurun_out
TAG=${1:-stage}
(timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log)
for v in default nostage; do
  LIBV=""; [ "$v" != "default" ] && LIBV="$PWD/tapqir_b200/lib/libtapqir_b200.$v.so"
  for w in c3 c3s8 c5; do
    echo "== $v $w"; TQ_LIB=$LIBV timeout 600 python profiles/kernel_times.py $w 10 2>&1 | grep -E "local_post|sum of kernels"
  done
  for w in c3 c3s8; do
    TQ_LIB=$LIBV timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-subs --trained-iters 0 > gpurun_out/${TAG}_${v}_$w.json 2> gpurun_out/${TAG}_${v}_$w.err
    python - gpurun_out/${TAG}_${v}_$w.json $v $w <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], sys.argv[3], '| step ms', round(d['ms_per_step'],4), '| value', round(d['value']/1e6,1), 'M | loss', d['final_loss'])
except Exception as e: print(sys.argv[2], sys.argv[3], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
P
  done
done

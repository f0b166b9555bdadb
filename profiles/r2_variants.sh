# A/B of likelihood-kernel build variants (csrc/build.py --variant=...): C2 / C3 step + kernel time
mkdir -p gpurun_out
for v in "$@"; do
  LIBV=""; [ "$v" != "default" ] && LIBV="$PWD/tapqir_b200/lib/libtapqir_b200.$v.so"
  for w in c3 c2; do
    TQ_LIB=$LIBV timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0 > gpurun_out/var_${v}_$w.json 2> gpurun_out/var_${v}_$w.err
    python - gpurun_out/var_${v}_$w.json $v $w <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[2], sys.argv[3], '| step ms', round(d['ms_per_step'],4), '| kernel ms', round(r['kernel_ms'],4), '| loss', d['final_loss'])
except Exception as e: print(sys.argv[2], sys.argv[3], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
P
  done
done

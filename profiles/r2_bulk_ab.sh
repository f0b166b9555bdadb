# A/B on the box: (1) bulk-copy (TMA) staging of the pixels vs seven cp.async per lane (TQ_NO_BULK=1);
# (2) site kernel with 4 units per thread (default build) vs 1 (variant build upt1), initial point and trained state
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_ksmogn_gpu.py tests/test_step_gpu.py tests/test_baseline_sizes_gpu.py tests/test_vs_reference_code_gpu.py -x -q > gpurun_out/bulk_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/bulk_tests.log
run() {  # name, workload, extra env
  timeout -s KILL 400 env $3 python bench.py --workload $2 --steps 10 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 1000 > gpurun_out/ab_$1_$2.json 2> gpurun_out/ab_$1_$2.err
  python - gpurun_out/ab_$1_$2.json $1 $2 <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(f"{sys.argv[2]:10s} {sys.argv[3]} | step ms {d['ms_per_step']:.4f} | kernel ms {r['kernel_ms']:.4f} | trained step ms {d['trained_state']['ms_per_step']:.4f} | loss {d['final_loss']}")
except Exception as e: print(sys.argv[2], sys.argv[3], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
P
}
for w in c3 c2; do
  run default $w "TQ_X=0"
  run nobulk $w "TQ_NO_BULK=1"
  run upt1 $w "TQ_LIB=$PWD/tapqir_b200/lib/libtapqir_b200.upt1.so"
done

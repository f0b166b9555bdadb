# quick check after a kernel change: GPU suite, then C2 / C3 step + likelihood-kernel time
mkdir -p gpurun_out
TAG=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
for w in c2 c3; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "$w rc=$?"
  python - gpurun_out/${TAG}_bench_$w.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], '| ms', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1),'M | kern ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'pipe', round(r['frac_pipe_slots'],3), 'e2e', round(d['e2e']['value']/1e6,1))
except Exception as e: print('ERR', e)
P
done
timeout 600 python bench.py --workload c2 --offset-hist 64 --steps 10 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0 > gpurun_out/${TAG}_bench_o64.json 2> gpurun_out/${TAG}_bench_o64.err; echo "o64 rc=$?"
python - gpurun_out/${TAG}_bench_o64.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print('O=64 c2 | ms', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1),'M | kern ms', round(r['kernel_ms'],4), r['bound'], 'frac', round(r['frac'],3))
except Exception as e: print('ERR', e)
P

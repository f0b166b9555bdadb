# strong-scaling bench at N GPUs only (no fit / collective checks): N = first argument
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err; echo "bench rc=$?"
python - gpurun_out/mg${N}_bench.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print('N', d['n_gpus'], d['scaling'], '| ms', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1),'M | kern ms', round(r['kernel_ms'],4), 'e2e', round(d['e2e']['value']/1e6,1), 'trained', d['trained_state'])
except Exception as e: print('ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-1500:])
P

# N-GPU validation (N = first argument): multi-rank fit path, then the strong-scaling bench at N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 profiles/r2_multigpu_fit.py > gpurun_out/mg${N}_fit.log 2>&1; echo "fit rc=$?"; grep -E "multi-rank fit OK|Error|error|assert" gpurun_out/mg${N}_fit.log | head -5
timeout 300 $TR --master-port 29512 profiles/multigpu_check.py cosmos > gpurun_out/mg${N}_check.log 2>&1; grep "allreduce=" gpurun_out/mg${N}_check.log
timeout 300 $TR --master-port 29513 profiles/multigpu_check.py cosmos+hmm >> gpurun_out/mg${N}_check.log 2>&1; grep "hmm allreduce=" gpurun_out/mg${N}_check.log
timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err; echo "bench rc=$?"
python - gpurun_out/mg${N}_bench.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print('N', d['n_gpus'], d['scaling'], '| ms', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1),'M | kern ms', round(r['kernel_ms'],4), 'e2e', round(d['e2e']['value']/1e6,1), 'trained', d['trained_state'])
except Exception as e: print('ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-1500:])
P

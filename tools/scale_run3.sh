# usage (8-GPU box): bash tools/scale_run3.sh <tag>  -- bench.py at N = 8 only (launched as the driver does)
TAG=$1
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/scale_${TAG}_n8.json 2> gpurun_out/scale_${TAG}_n8.err
tail -1 gpurun_out/scale_${TAG}_n8.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N=8 value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['e2e']['roofline'].get('frac'), 'failed', d['failed_tiles'], d['status_gather'][:40], d['accuracy'])" || tail -5 gpurun_out/scale_${TAG}_n8.err

# usage (multi-GPU box): bash tools/scale_run4.sh <tag> <N>  -- bench.py at N GPUs (launched as the driver does)
TAG=$1; N=$2
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
tail -1 gpurun_out/scale_${TAG}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N=$N value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['e2e']['roofline'].get('frac'), 'failed', d['failed_tiles'])" || tail -5 gpurun_out/scale_${TAG}_n$N.err

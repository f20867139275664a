# usage (multi-GPU box): bash tools/scale_run.sh <tag> <N...>   -- bench.py at each N the way the driver launches it
TAG=$1; shift
for N in "$@"; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  fi
  tail -1 gpurun_out/scale_${TAG}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N=$N value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'failed', d['failed_tiles'], 'clocks', d['clocks']['sm_mhz'])" || tail -3 gpurun_out/scale_${TAG}_n$N.err
done

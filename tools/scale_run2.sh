# usage (8-GPU box): bash tools/scale_run2.sh <tag>  -- raw H2D/D2H ceiling at 1/2/4/8 ranks, then bench.py at N = 8, 4, 2
TAG=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
numactl -H > gpurun_out/numa_$TAG.txt 2>&1 || lscpu | grep -i numa > gpurun_out/numa_$TAG.txt
free -g | head -2 >> gpurun_out/numa_$TAG.txt; nproc >> gpurun_out/numa_$TAG.txt
timeout 200 python tools/pcie_ceiling.py gpurun_out/pcie_ceiling_n1_$TAG.json | tail -1
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2941$N tools/pcie_ceiling.py gpurun_out/pcie_ceiling_n${N}_$TAG.json 2> gpurun_out/pcie_n$N.err | tail -1
done
for N in 8 4 2; do
  cp gpurun_out/pcie_ceiling_n${N}_$TAG.json profiles/pcie_ceiling_n$N.json 2>/dev/null
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  tail -1 gpurun_out/scale_${TAG}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N=$N value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['e2e']['roofline'].get('frac'), 'failed', d['failed_tiles'], d['status_gather'][:40], d['accuracy'])" || tail -5 gpurun_out/scale_${TAG}_n$N.err
done

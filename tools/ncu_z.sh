# GPU box: full-set capture of the complex pipeline kernels with per-line samples.  usage: bash tools/ncu_z.sh <tag>
TAG=$1; O=gpurun_out; T=/tmp/ncuz_$TAG; mkdir -p $O $T
ncu --set full --clock-control none --import-source on -k regex:ztlr_ -s 4 -c 4 -f -o $T/z python bench.py --config zgemm --steps 1 --warmup 1 > $T/l.log 2>&1
python tools/ncu_summary.py $T/z.ncu-rep > $O/ncu_z_${TAG}_summary.txt 2>&1
for k in stack qr_blocked svd apply; do python tools/ncu_lines.py $T/z.ncu-rep ztlr_$k 14 > $O/ncu_z_${TAG}_lines_$k.txt 2>&1; done
tail -2 $T/l.log

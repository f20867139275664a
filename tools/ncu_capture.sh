# GPU box: ncu evidence of the round, summarised ON the box (the reports themselves are too large to bring back).
# usage: bash tools/ncu_capture.sh <tag>
TAG=$1; O=gpurun_out; mkdir -p $O; T=/tmp/ncu_$TAG; mkdir -p $T
# (1) which counter lights up on a pure DMMA loop / a pure DFMA loop (tools/fp64_peak.cu)
ncu --set full --clock-control none -k regex:"dmma884|dfma_kernel" -c 2 -f -o $T/peak ./tools/fp64_peak > /dev/null 2>&1
ncu -i $T/peak.ncu-rep --page raw --csv > $O/ncu_peak_raw_$TAG.csv 2>/dev/null
# (2) launch list of one chunk of the batched update (shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tlr_|fill_" -c 60 --csv --log-file $O/ncu_launches_$TAG.csv python tools/profile_run.py 2368 1024 64 2 > $T/l.log 2>&1
# (3) the four pipeline kernels, full set
ncu --set full --clock-control none --import-source on -k regex:tlr_ -s 4 -c 4 -f -o $T/pipe python tools/profile_run.py 296 1024 64 2 > $T/p.log 2>&1
python tools/ncu_summary.py $T/pipe.ncu-rep > $O/ncu_${TAG}_summary.txt 2>&1
ncu -i $T/pipe.ncu-rep --page raw --csv > $O/ncu_${TAG}_raw.csv 2>/dev/null
python tools/ncu_lines.py $T/pipe.ncu-rep tlr_tsqr 30 > $O/ncu_${TAG}_lines_tsqr.txt 2>&1
python tools/ncu_lines.py $T/pipe.ncu-rep tlr_svd 30 > $O/ncu_${TAG}_lines_svd.txt 2>&1
# (4) the secondary kernels
ncu --set full --clock-control none -k regex:"dense_|trsm_|potrf_|zdense" -s 20 -c 40 -f -o $T/ops python tools/profile_ops.py 296 > $T/o.log 2>&1
python tools/ncu_summary.py $T/ops.ncu-rep > $O/ncu_ops_${TAG}_summary.txt 2>&1
tail -2 $T/o.log; ls -la $T; du -sh $O

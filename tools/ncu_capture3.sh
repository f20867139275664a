# GPU box: launch list + full-set capture of the four pipeline kernels, summarised ON the box.  usage: bash tools/ncu_capture3.sh <tag>
TAG=$1; O=gpurun_out; mkdir -p $O; T=/tmp/ncu3_$TAG; mkdir -p $T
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tlr_|fill_" -c 60 --csv --log-file $O/ncu_launches_$TAG.csv python tools/profile_run.py 2368 1024 64 2 > $T/l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tlr_ -s 4 -c 4 -f -o $T/pipe python tools/profile_run.py 296 1024 64 2 > $T/p.log 2>&1
python tools/ncu_summary.py $T/pipe.ncu-rep > $O/ncu_${TAG}_summary.txt 2>&1
ncu -i $T/pipe.ncu-rep --page raw --csv > $O/ncu_${TAG}_raw.csv 2>/dev/null
python tools/ncu_lines.py $T/pipe.ncu-rep tlr_tsqr 30 > $O/ncu_${TAG}_lines_tsqr.txt 2>&1
python tools/ncu_lines.py $T/pipe.ncu-rep tlr_svd 40 > $O/ncu_${TAG}_lines_svd.txt 2>&1
python tools/ncu_traffic.py $O/ncu_${TAG}_raw.csv 296 1024 64 $O/roofline_traffic_$TAG.json "ncu --set full (tools/ncu_capture3.sh $TAG, tools/profile_run.py 296 1024 64): dram__bytes_read.sum + dram__bytes_write.sum per launch / 296 tiles"
tail -2 $T/p.log; du -sh $O

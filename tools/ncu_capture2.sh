# GPU box: the two captures ncu_capture.sh got wrong (kernel filters), summarised on the box.  usage: bash tools/ncu_capture2.sh <tag>
TAG=$1; O=gpurun_out; mkdir -p $O; T=/tmp/ncu2_$TAG; mkdir -p $T
ncu --set full --clock-control none -k regex:"dmma884_kernel" -c 1 -f -o $T/peak ./tools/fp64_peak > /dev/null 2>&1
ncu -i $T/peak.ncu-rep --page raw --csv > $O/ncu_peak_dmma_raw_$TAG.csv 2>/dev/null
ncu --set full --clock-control none -k regex:"dense_update_kernel|dense_prep_kernel|dense_bdtav_kernel|trsm_inv_kernel|trsm_solve_kernel|potrf_diag_kernel|potrf_panel_kernel|potrf_trail_kernel|zdense_update_kernel" -s 20 -c 24 -f -o $T/ops python tools/profile_ops.py 296 > $T/o.log 2>&1
python tools/ncu_summary.py $T/ops.ncu-rep > $O/ncu_ops_${TAG}_summary.txt 2>&1
tail -2 $T/o.log; du -sh $O

# usage (GPU box): bash tools/gpu_round_check.sh <tag> [stagecheck] [tests] [bench] [ncu]
set -x
TAG=${1:-x}; shift
for what in "$@"; do
case $what in
 stagecheck) timeout 600 python tools/gpu_stage_check.py 256 8 2 500 25 2 1024 64 2 > gpurun_out/stage_$TAG.log 2>&1; tail -40 gpurun_out/stage_$TAG.log ;;
 tests) timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ;;
 bench) timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 2500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err ;;
 quick) timeout 600 python bench.py --tiles 4736 --no-e2e --no-cpu-baseline --steps 2 > gpurun_out/quick_$TAG.json 2> gpurun_out/quick_$TAG.err; python -c "
import json;d=json.load(open('gpurun_out/quick_$TAG.json'));print(d['value'],{k:round(v['ms'],2) for k,v in d['stages'].items()},d['failed_tiles'],d['mean_new_rank'],d['sample_abs_err_fro'])"; tail -5 gpurun_out/quick_$TAG.err ;;
 ncu) ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tlr_|dense_|fill_' -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_run.py 2368 1024 64 2 > gpurun_out/prof_launch.log 2>&1
      ncu --set full --clock-control none --import-source on -k regex:tlr_ -s 4 -c 4 -f -o gpurun_out/prof_$TAG python tools/profile_run.py 296 1024 64 2 > gpurun_out/prof_full.log 2>&1; tail -3 gpurun_out/prof_full.log ;;
 quick4) HCORE_B200_TSQR_WARPS=4 timeout 600 python bench.py --tiles 4736 --no-e2e --no-cpu-baseline --steps 2 > gpurun_out/quick4_$TAG.json 2> gpurun_out/quick4_$TAG.err; python -c "
import json;d=json.load(open('gpurun_out/quick4_$TAG.json'));print(d['value'],{k:round(v['ms'],2) for k,v in d['stages'].items()},d['failed_tiles'],d['mean_new_rank'],d['sample_abs_err_fro'])"; tail -5 gpurun_out/quick4_$TAG.err ;;
 quickw) for W in 8 4 2; do HCORE_B200_TSQR_WARPS=$W timeout 600 python bench.py --tiles 4736 --no-e2e --no-cpu-baseline --steps 2 > gpurun_out/quickw${W}_$TAG.json 2> gpurun_out/quickw${W}_$TAG.err; python -c "
import json;d=json.load(open('gpurun_out/quickw${W}_$TAG.json'));print('W=$W',d['value'],{k:round(v['ms'],2) for k,v in d['stages'].items()},d['failed_tiles'],d['mean_new_rank'],d['sample_abs_err_fro'])"; tail -3 gpurun_out/quickw${W}_$TAG.err; done ;;
 ops) timeout 1200 python tools/bench_ops.py gpurun_out/ops_$TAG.json 2>&1 | tail -20 ;;
 dense) timeout 600 python -m pytest tests -m gpu -x -q -k "dgemmcd or dsyrk" 2>&1 | tail -8; OPS_ONLY=dense timeout 600 python tools/bench_ops.py gpurun_out/dense_$TAG.json 2>&1 | tail -4
        ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dense_' -c 12 --csv --log-file gpurun_out/dense_launches_$TAG.csv env OPS_ONLY=dense python tools/bench_ops.py > /dev/null 2>&1; grep -v "^==" gpurun_out/dense_launches_$TAG.csv | awk -F'","' '{print $5, $(NF)}' | tail -6 ;;
 sanitize) for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_run.py 2>&1 | tail -4; done ;;
esac
done

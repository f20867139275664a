set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -c 3000 gpurun_out/bench_r1a.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tlr_|dense_|fill_' -c 60 --csv --log-file gpurun_out/launches_r1a.csv python tools/profile_run.py 2368 1024 64 2 > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tlr_ -s 4 -c 4 -f -o gpurun_out/prof_r1a python tools/profile_run.py 296 1024 64 2 > gpurun_out/prof_full.log 2>&1
tail -3 gpurun_out/prof_full.log

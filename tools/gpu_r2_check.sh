# Round-2 GPU check (under gpurun): bash tools/gpu_r2_check.sh <tag> [tests] [probes] [pcie] [quick] [bench] [ops]
TAG=${1:-x}; shift
mkdir -p gpurun_out
for what in "$@"; do
case $what in
 tests)  timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -40 | tee gpurun_out/tests_$TAG.log ;;
 testsall) timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -80 | tee gpurun_out/tests_$TAG.log ;;
 newtests) timeout 1800 python -m pytest tests/test_gpu_cholstep.py tests/test_gpu_stages.py -m gpu -q --timeout 600 2>&1 | tail -80 | tee gpurun_out/newtests_$TAG.log ;;
 probes) timeout 300 ./tools/fp64_peak 2>&1 | grep -v "1684\|16816" | tee gpurun_out/fp64_mix_$TAG.txt ;;
 pcie)   timeout 300 python tools/pcie_ceiling.py gpurun_out/pcie_ceiling_n1_$TAG.json 2>&1 | tail -3 ;;
 quick)  timeout 600 python bench.py --tiles 4736 --no-e2e --no-cpu-baseline --steps 2 > gpurun_out/quick_$TAG.json 2> gpurun_out/quick_$TAG.err; python -c "
import json;d=json.load(open('gpurun_out/quick_$TAG.json'));print(d['value'],{k:round(v['ms'],2) for k,v in d['stages'].items()},d['failed_tiles'],d['mean_new_rank'],d['accuracy'])"; tail -5 gpurun_out/quick_$TAG.err ;;
 bench)  timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err ;;
 ops)    timeout 1500 python tools/bench_ops.py gpurun_out/ops_$TAG.json 2>&1 | tail -30 ;;
 chol)   OPS_ONLY=chol timeout 900 python tools/bench_ops.py gpurun_out/chol_$TAG.json 2>&1 | tail -20 ;;
 choldemo) timeout 900 python tools/chol_step_demo.py 16 512 256 1e-8 gpurun_out/chol_demo_$TAG.json 2>&1 | tail -6 ;;
 legacy) timeout 600 python tools/legacy_latency.py 2>&1 | tail -12 | tee gpurun_out/legacy_$TAG.txt ;;
 ctests) timeout 600 python -m pytest tests/test_c_boundary.py -m gpu -q 2>&1 | tail -15 ;;
 sanitize) for tool in memcheck racecheck; do timeout 1200 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_run.py 2>&1 | tail -6; done | tee gpurun_out/sanitize_$TAG.txt ;;
esac
done

#!/bin/bash
# Evidence run on one GPU box (via gpurun): GPU tests, bench (+ its traced twin), ncu captures.
# usage: bash tools/run_gpu_checks.sh <tag>      outputs under gpurun_out/; copy what matters to profiles/
mkdir -p gpurun_out
TAG=${1:-rXX}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
export J2P_EXPECT_GPU=1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 2700 gpurun_out/bench_${TAG}.json
tail -3 gpurun_out/bench_${TAG}.err
J2P_TRACE=1 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err
grep "j2p trace: \(create\|upload\|queue\|device\|download\|destroy\)" gpurun_out/bench_trace.err | tail -12
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 32 --csv --log-file gpurun_out/launches_${TAG}.csv python tools/prof_driver.py > gpurun_out/ncu_list_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gradient -s 4 -c 1 -o gpurun_out/prof_gradient_${TAG} -f python tools/prof_driver.py > gpurun_out/ncu_grad_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_project -s 4 -c 1 -o gpurun_out/prof_project_${TAG} -f python tools/prof_driver.py > gpurun_out/ncu_proj_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_proj_${TAG}.log

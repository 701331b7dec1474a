# usage: bash tools/run_bench_profile.sh <round-tag>   (on the GPU box, via gpurun)
TAG=${1:-r01}
mkdir -p gpurun_out
export J2P_EXPECT_GPU=1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 2500 gpurun_out/bench_${TAG}.json
tail -5 gpurun_out/bench_${TAG}.err
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 32 --csv --log-file gpurun_out/launches_${TAG}.csv python tools/prof_driver.py > gpurun_out/ncu_list_${TAG}.log 2>&1
# full capture of one launch of each kernel
ncu --set full --clock-control none --import-source on -k regex:k_gradient -s 4 -c 1 -o gpurun_out/prof_gradient_${TAG} -f python tools/prof_driver.py > gpurun_out/ncu_grad_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_project -s 12 -c 1 -o gpurun_out/prof_project_${TAG} -f python tools/prof_driver.py > gpurun_out/ncu_proj_${TAG}.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_ref_${TAG}.json

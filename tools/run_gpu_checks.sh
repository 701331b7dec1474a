export J2P_EXPECT_GPU=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for v in 0 2; do J2P_PROJ_VARIANT=$v python tools/quick_time.py 2>&1 | tail -1; done
bash tools/run_profile_only.sh ${1:-x} > /dev/null 2>&1

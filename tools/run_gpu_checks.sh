set -x
export J2P_EXPECT_GPU=1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in 0 1 2; do J2P_PROJ_VARIANT=$v python tools/quick_time.py 2>&1 | tail -1; done
bash tools/run_profile_only.sh ${1:-x} > /dev/null 2>&1

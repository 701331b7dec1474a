set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
export J2P_EXPECT_GPU=1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
./tools/microbench 2>&1 | tee gpurun_out/microbench_r01.txt

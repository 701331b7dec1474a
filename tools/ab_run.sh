mkdir -p gpurun_out
export J2P_EXPECT_GPU=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tma or callbacks or reset" > gpurun_out/pytest_tma.log 2>&1
tail -8 gpurun_out/pytest_tma.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2

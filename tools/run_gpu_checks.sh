export J2P_EXPECT_GPU=1
mkdir -p gpurun_out
./tools/rootcheck | tee gpurun_out/rootcheck_r01.txt
python tools/quick_time.py build_ab/lib_roots0.so build_ab/lib_roots1.so 2>&1 | grep lib_
python tools/e2e_trace.py 2>&1 | grep -E "trace|call"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3

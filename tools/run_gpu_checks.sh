export J2P_EXPECT_GPU=1
mkdir -p gpurun_out
python tools/quick_time.py build_ab/lib_base.so build_ab/lib_cur.so build_ab/lib_g4.so build_ab/lib_t5.so build_ab/lib_t6.so 2>&1 | grep lib_
J2P_TRACE=1 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err
grep -E "trace" gpurun_out/bench_trace.err | tail -14
python -m pytest tests -m gpu -x -q 2>&1 | tail -3

#!/bin/bash
# A/B of k_gradient builds (per-kernel device time), then the GPU test suite on the default build
mkdir -p gpurun_out
timeout 600 python tools/quick_time.py build_ab/lib_v6.so build_ab/lib_new_g3.so build_ab/lib_new_g4.so > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log

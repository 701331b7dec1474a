#!/bin/bash
# Round-2 GPU call: parity suite with the cp.async projection (validates the packed gradient kernel
# on its own), again with the TMA-fed projection, A/B timings, ncu captures.  One GPU.
mkdir -p gpurun_out
rm -f gpurun_out/grad_ab.log
export J2P_EXPECT_GPU=1
J2P_PROJ_TMA=0 timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_notma.log 2>&1
tail -4 gpurun_out/pytest_gpu_notma.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
for f in "3840 2160 50 4:4:4" "1920 1080 10 4:2:0" "7680 4320 10 4:2:0"; do
  echo "== frame $f: default (packed gradient, TMA projection) | J2P_PROJ_TMA=0 | J2P_GRAD_SCALAR=1 J2P_PROJ_TMA=0" >> gpurun_out/grad_ab.log
  timeout 300 python tools/quick_time.py --frame $f >> gpurun_out/grad_ab.log 2>&1
  J2P_PROJ_TMA=0 timeout 300 python tools/quick_time.py --frame $f >> gpurun_out/grad_ab.log 2>&1
  J2P_GRAD_SCALAR=1 J2P_PROJ_TMA=0 timeout 300 python tools/quick_time.py --frame $f >> gpurun_out/grad_ab.log 2>&1
done
cat gpurun_out/grad_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gradient_packed -s 4 -c 1 -o gpurun_out/prof_gradient_r02c -f python tools/prof_driver.py > gpurun_out/ncu_grad_r02c.log 2>&1
tail -2 gpurun_out/ncu_grad_r02c.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_project_tma -s 4 -c 1 -o gpurun_out/prof_project_r02c -f python tools/prof_driver.py > gpurun_out/ncu_proj_r02c.log 2>&1
tail -2 gpurun_out/ncu_proj_r02c.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
tail -c 4000 gpurun_out/bench_r02c.json; tail -3 gpurun_out/bench_r02c.err

mkdir -p gpurun_out
rm -f gpurun_out/ab7.log
export J2P_EXPECT_GPU=1
for f in "3840 2160 50 4:4:4" "1920 1080 10 4:2:0" "7680 4320 10 4:2:0"; do
  echo "== $f  (w4c2 = previous projection kernels; default = packed branch-free stepper/clamp in the tile kernel)" >> gpurun_out/ab7.log
  timeout 600 python tools/quick_time.py --frame $f build_ab/w4c2.so jpeg2png_b200/csrc/libjpeg2png_b200.so >> gpurun_out/ab7.log 2>&1
done
cat gpurun_out/ab7.log
timeout 300 ncu --section LaunchStats --section Occupancy --section SpeedOfLight --clock-control none -k regex:k_gradient_packed -s 4 -c 1 python tools/prof_driver.py --lib build_ab/w5r200.so > gpurun_out/ncu_w5r200.log 2>&1
grep -E "Registers|Block Size|Grid Size|Waves|Occupancy|Active Warps|Duration|Shared Memory|Block Limit" gpurun_out/ncu_w5r200.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_project_tile -s 4 -c 1 -o gpurun_out/prof_project_tile_packed -f python tools/prof_driver.py > gpurun_out/ncu_proj_tile.log 2>&1
tail -2 gpurun_out/ncu_proj_tile.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_ab7.log 2>&1
tail -5 gpurun_out/pytest_gpu_ab7.log

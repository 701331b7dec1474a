mkdir -p gpurun_out
rm -f gpurun_out/ab8.log
export J2P_EXPECT_GPU=1
timeout 300 tools/divcheck > gpurun_out/divcheck_r02.txt 2>&1; cat gpurun_out/divcheck_r02.txt
for f in "3840 2160 50 4:4:4" "1920 1080 10 4:2:0" "7680 4320 10 4:2:0"; do
  echo "== $f  (w4c2 = five-operation quotients; q4 = four-operation quotients on a two-term reciprocal)" >> gpurun_out/ab8.log
  timeout 600 python tools/quick_time.py --frame $f build_ab/w4c2.so build_ab/q4.so >> gpurun_out/ab8.log 2>&1
done
cat gpurun_out/ab8.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_ab8.log 2>&1
tail -5 gpurun_out/pytest_gpu_ab8.log

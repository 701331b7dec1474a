mkdir -p gpurun_out
rm -f gpurun_out/ab9.log
export J2P_EXPECT_GPU=1
for f in "3840 2160 50 4:4:4" "1920 1080 10 4:2:0" "7680 544 10 4:2:0"; do
  echo "== $f  (q4 = previous commit; default = release-ordered ticket; q4c3 = 3 CTAs/SM)" >> gpurun_out/ab9.log
  timeout 600 python tools/quick_time.py --frame $f build_ab/q4.so jpeg2png_b200/csrc/libjpeg2png_b200.so build_ab/q4c3.so >> gpurun_out/ab9.log 2>&1
done
cat gpurun_out/ab9.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_ab9.log 2>&1
tail -3 gpurun_out/pytest_gpu_ab9.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err
tail -c 2500 gpurun_out/bench_r02d.json; tail -3 gpurun_out/bench_r02d.err

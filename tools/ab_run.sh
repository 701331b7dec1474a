mkdir -p gpurun_out
rm -f gpurun_out/ab15.log
for f in "3840 2160 50 4:4:4" "1920 1080 10 4:2:0" "7680 544 10 4:2:0"; do
  echo "== $f  (tree: 32-block CTA tiles, 4 x 8 warps per SM; tile16: 16-block tiles, 8 x 4 warps)" >> gpurun_out/ab15.log
  timeout 600 python tools/quick_time.py --frame $f jpeg2png_b200/csrc/libjpeg2png_b200.so build_ab/tile16.so >> gpurun_out/ab15.log 2>&1
done
cat gpurun_out/ab15.log

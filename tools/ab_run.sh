mkdir -p gpurun_out
rm -f gpurun_out/ab12.log
export J2P_EXPECT_GPU=1
timeout 300 python tools/cli_batch.py 64 > gpurun_out/cli_batch_r02.txt 2>&1; head -3 gpurun_out/cli_batch_r02.txt
for f in "3840 2160 50 4:4:4" "1920 1080 10 4:2:0"; do
  echo "== $f: joint (tree, depth2 = row ring 2 deep), then -s mode" >> gpurun_out/ab12.log
  timeout 600 python tools/quick_time.py --frame $f jpeg2png_b200/csrc/libjpeg2png_b200.so build_ab/depth2.so >> gpurun_out/ab12.log 2>&1
  timeout 600 python tools/quick_time.py --separate --frame $f >> gpurun_out/ab12.log 2>&1
done
cat gpurun_out/ab12.log

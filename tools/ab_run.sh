mkdir -p gpurun_out
rm -f gpurun_out/ab11.log
export J2P_EXPECT_GPU=1
for rep in 1 2; do
for f in "3840 2160 50 4:4:4" "7680 4320 10 4:2:0" "1920 1080 10 4:2:0"; do
  echo "== $f  (nopdl.so = previous commit; then the tree with J2P_PDL=0; then the tree, default)" >> gpurun_out/ab11.log
  timeout 600 python tools/quick_time.py --frame $f build_ab/nopdl.so >> gpurun_out/ab11.log 2>&1
  J2P_PDL=0 timeout 600 python tools/quick_time.py --frame $f >> gpurun_out/ab11.log 2>&1
  timeout 600 python tools/quick_time.py --frame $f >> gpurun_out/ab11.log 2>&1
done
done
cat gpurun_out/ab11.log

mkdir -p gpurun_out
rm -f gpurun_out/ab3.log
for f in "3840 2160 50 4:4:4"; do
  echo "== $f  J2P_PROJ_TMA=0" >> gpurun_out/ab3.log
  J2P_PROJ_TMA=0 timeout 600 python tools/quick_time.py --frame $f build_ab/g_w4c2.so build_ab/g_w5c1.so >> gpurun_out/ab3.log 2>&1
done
cat gpurun_out/ab3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_project_tma -s 4 -c 1 -o gpurun_out/prof_project_tma3 -f python tools/prof_driver.py > gpurun_out/ncu_proj_tma3.log 2>&1
tail -2 gpurun_out/ncu_proj_tma3.log

mkdir -p gpurun_out
rm -f gpurun_out/final_configs.log
for f in "1920 1080 10 4:2:0" "3840 2160 50 4:4:4" "7680 4320 10 4:2:0"; do
  echo "== $f" >> gpurun_out/final_configs.log
  timeout 300 python tools/quick_time.py --frame $f >> gpurun_out/final_configs.log 2>&1
done
timeout 200 python tools/quick_time.py --separate --frame 1920 1080 10 4:2:0 >> gpurun_out/final_configs.log 2>&1
cat gpurun_out/final_configs.log
timeout 200 python tools/cli_batch.py 64 > gpurun_out/cli_batch_final2.txt 2>&1; head -2 gpurun_out/cli_batch_final2.txt

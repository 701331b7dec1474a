mkdir -p gpurun_out
timeout 300 python tools/cli_batch.py 32 > gpurun_out/cli_batch_trace.txt 2>&1; head -2 gpurun_out/cli_batch_trace.txt
grep "session create" gpurun_out/cli_batch_trace.txt | head -20
grep "session create" gpurun_out/cli_batch_trace.txt | tail -6

mkdir -p gpurun_out
export J2P_EXPECT_GPU=1
timeout 200 python -m pytest tests/test_gpu_cli.py -m gpu -q -x > gpurun_out/pytest_cli.log 2>&1; tail -2 gpurun_out/pytest_cli.log
timeout 100 python tools/cli_batch.py 64 > gpurun_out/cli_batch_final3.txt 2>&1; head -2 gpurun_out/cli_batch_final3.txt; grep "read+parse" gpurun_out/cli_batch_final3.txt | tail -3 | cut -c1-200

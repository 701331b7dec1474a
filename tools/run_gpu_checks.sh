#!/bin/bash
# after the clean-up: full GPU suite; CLI batch (config 5 style) at two host thread counts
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
make -C jpeg2png_b200/cli jpeg2png > /dev/null 2>&1
timeout 300 python tools/cli_batch.py 16 > gpurun_out/cli_batch.log 2>&1
timeout 300 python tools/cli_batch.py 16 16 >> gpurun_out/cli_batch.log 2>&1
cat gpurun_out/cli_batch.log

#!/bin/bash
mkdir -p gpurun_out
make -C jpeg2png_b200/cli jpeg2png > /dev/null 2>&1
timeout 300 python tools/cli_batch.py 16 > gpurun_out/cli_batch.log 2>&1
cat gpurun_out/cli_batch.log

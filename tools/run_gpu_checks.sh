#!/bin/bash
# 2 GPUs: the peer-memory strip protocol (J2P_STRIP_P2P=1) against the single-GPU result, then its timing
mkdir -p gpurun_out
export J2P_STRIP_P2P=1
timeout 150 python -m pytest tests/test_gpu_strips.py -m gpu -q -x -k "2 and native" > gpurun_out/strips_p2p.log 2>&1
echo "pytest exit $?" >> gpurun_out/strips_p2p.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/strip_bench.py >> gpurun_out/strips_p2p.log 2> gpurun_out/strips_p2p_err.log
grep -v "^$" gpurun_out/strips_p2p.log | tail -14; tail -3 gpurun_out/strips_p2p_err.log

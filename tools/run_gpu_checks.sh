#!/bin/bash
# multi-GPU strip checks: parity tests, then strong scaling of one 8K frame (N from $1...)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG" > gpurun_out/strips.log
timeout 600 python -m pytest tests/test_gpu_strips.py -m gpu -q -x ${PYTEST_K:+-k "$PYTEST_K"} >> gpurun_out/strips.log 2>&1
echo "pytest exit $?" >> gpurun_out/strips.log
for n in "$@"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      tools/strip_bench.py >> gpurun_out/strips.log 2> gpurun_out/strips_err_$n.log
  echo "strip_bench N=$n exit $?" >> gpurun_out/strips.log
done
grep -v "^$" gpurun_out/strips.log | tail -40

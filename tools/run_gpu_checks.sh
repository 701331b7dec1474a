#!/bin/bash
# 2 GPUs: strip parity (both drivers) with the final kernels, strong scaling N=2, bench N=2 under torchrun
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strips.py -m gpu -q -x > gpurun_out/strips.log 2>&1
echo "pytest exit $?" >> gpurun_out/strips.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/strip_bench.py >> gpurun_out/strips.log 2> gpurun_out/strips_err_2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep -v "^$" gpurun_out/strips.log | tail -12
python -c "import json;d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]);print('N=2 value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"

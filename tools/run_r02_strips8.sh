#!/bin/bash
# Round-2 multi-GPU call, trimmed for an 8-GPU box (charged 8x): the in-kernel peer-memory protocol on
# 8 strips against the reference (small frame and the 8K frame of config 4), then strong-scaling timings.
# usage: gpurun --gpus 8 -- bash tools/run_r02_strips8.sh [bench]     (bench: also bench.py under torchrun)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export J2P_EXPECT_GPU=1
timeout 500 python -m pytest tests/test_gpu_strips.py -m gpu -q -x -k "(test_strips_match_reference and native and 8) or (test_8k and 8)" \
    > gpurun_out/pytest_strips_n$NG.log 2>&1
tail -4 gpurun_out/pytest_strips_n$NG.log
rm -f gpurun_out/strips_time_n$NG.log
for n in 1 2 4 8; do
  [ "$n" -le "$NG" ] || continue
  echo "== N=$n default (peer memory, exchanges inside the kernels)" >> gpurun_out/strips_time_n$NG.log
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      tools/strip_bench.py 2>> gpurun_out/strips_err_n$NG.log | grep -E "native" >> gpurun_out/strips_time_n$NG.log
done
cat gpurun_out/strips_time_n$NG.log
if [ "$1" = "bench" ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $NG --steps 3 --warmup 3 \
      > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
  tail -c 1500 gpurun_out/bench_n$NG.json; tail -3 gpurun_out/bench_n$NG.err
fi

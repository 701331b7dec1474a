#!/bin/bash
# Round-2 multi-GPU call: strip parity against the reference, strong-scaling timings of one 8K frame
# for the three protocols, and bench.py under torchrun (strong block).  usage: gpurun --gpus N -- bash tools/run_r02_strips.sh
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export J2P_EXPECT_GPU=1
timeout 1500 python -m pytest tests/test_gpu_strips.py -m gpu -q -x > gpurun_out/pytest_strips_n$NG.log 2>&1
tail -6 gpurun_out/pytest_strips_n$NG.log
rm -f gpurun_out/strips_time_n$NG.log
for n in 1 2 4 8; do
  [ "$n" -le "$NG" ] || continue
  for mode in "" "J2P_STRIP_FUSED_HALO=0" "J2P_STRIP_P2P=0"; do
    echo "== N=$n ${mode:-default (peer memory, exchanges inside the kernels)}" >> gpurun_out/strips_time_n$NG.log
    env $mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
        tools/strip_bench.py 2>> gpurun_out/strips_err_n$NG.log | grep -E "native|torchdist" >> gpurun_out/strips_time_n$NG.log
  done
done
cat gpurun_out/strips_time_n$NG.log
if [ "$NG" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $NG --steps 3 --warmup 3 \
      > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
  tail -c 3000 gpurun_out/bench_n$NG.json; tail -5 gpurun_out/bench_n$NG.err
fi

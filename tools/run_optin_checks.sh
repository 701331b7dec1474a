#!/bin/bash
# Round-2 starter (one GPU box, via gpurun): validate and time the opt-in paths written at the end
# of round 1.  usage: bash tools/run_optin_checks.sh            (add --gpus N to gpurun for the strip part)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
# 1. the coalesced-staging projection for 2x2 planes: every GPU parity test with it switched on
J2P_PROJ_TILE22=1 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/optin_tile22_tests.log 2>&1
tail -3 gpurun_out/optin_tile22_tests.log
for f in "1920 1080 10 4:2:0" "7680 4320 10 4:2:0"; do
  echo "== frame $f: default, then J2P_PROJ_TILE22=1" >> gpurun_out/optin_tile22_time.log
  timeout 300 python tools/quick_time.py --frame $f >> gpurun_out/optin_tile22_time.log 2>&1
  J2P_PROJ_TILE22=1 timeout 300 python tools/quick_time.py --frame $f >> gpurun_out/optin_tile22_time.log 2>&1
done
cat gpurun_out/optin_tile22_time.log
# 2. the peer-memory strip protocol on every GPU count the box offers
if [ "$NG" -ge 2 ]; then
  J2P_STRIP_P2P=1 timeout 600 python -m pytest tests/test_gpu_strips.py -m gpu -q -x -k native > gpurun_out/optin_p2p_tests.log 2>&1
  tail -3 gpurun_out/optin_p2p_tests.log
  for n in 2 4 8; do
    [ "$n" -le "$NG" ] || continue
    for mode in 0 1; do
      J2P_STRIP_P2P=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
          tools/strip_bench.py 2> gpurun_out/optin_p2p_err.log | grep "native" >> gpurun_out/optin_p2p_time.log
    done
  done
  cat gpurun_out/optin_p2p_time.log
fi

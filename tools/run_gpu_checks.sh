export J2P_EXPECT_GPU=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n2.json'))
    print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
except Exception as e:
    print('bench n2 failed', e); print(open('gpurun_out/bench_n2.err').read()[-2000:])
PY

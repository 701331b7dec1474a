export J2P_EXPECT_GPU=1
mkdir -p gpurun_out
python tools/e2e_trace.py 2>&1 | grep -E "trace|call" | tail -8
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k in d['roofline']['kernels']: print(k['name'], k['ms'], k['frac'], k['traffic'])
print('iter', d['roofline']['iteration'], 'launches', d['gpu_launches'])
PY
tail -3 gpurun_out/bench_quick.err

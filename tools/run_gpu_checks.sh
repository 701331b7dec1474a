set -x
export J2P_EXPECT_GPU=1
mkdir -p gpurun_out
./tools/divcheck 2>&1 | tee gpurun_out/divcheck.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 1500 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err

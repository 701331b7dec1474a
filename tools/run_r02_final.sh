#!/bin/bash
# Round-2 closing GPU call (one GPU): the whole GPU suite, bench.py, the ncu launch list of the bench
# command, and one `ncu --set full` capture of each solver kernel.  (The command-line batch and the
# reference arm were recorded by the previous closing call: profiles/r02_cli_batch.txt, r02_bench_final_reference_arm.json.)
mkdir -p gpurun_out
export J2P_EXPECT_GPU=1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_final.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 1500 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
J2P_BENCH_STRONG=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches_final.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_project_tile -s 4 -c 1 -o gpurun_out/prof_project_final -f python tools/prof_driver.py > gpurun_out/ncu_proj_final.log 2>&1
tail -1 gpurun_out/ncu_proj_final.log

#!/bin/bash
# Round-2 closing GPU call (one GPU): the whole GPU suite, the command-line batch, bench.py (both arms),
# the ncu launch list of the bench command, and one `ncu --set full` capture of each solver kernel.
mkdir -p gpurun_out
export J2P_EXPECT_GPU=1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_final.log
timeout 300 python tools/cli_batch.py 64 > gpurun_out/cli_batch_final.txt 2>&1; head -2 gpurun_out/cli_batch_final.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 1800 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
tail -c 600 gpurun_out/bench_final_reference.json
J2P_BENCH_STRONG=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches_final.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gradient_packed -s 4 -c 1 -o gpurun_out/prof_gradient_final -f python tools/prof_driver.py > gpurun_out/ncu_grad_final.log 2>&1
tail -1 gpurun_out/ncu_grad_final.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_project_tile -s 4 -c 1 -o gpurun_out/prof_project_final -f python tools/prof_driver.py > gpurun_out/ncu_proj_final.log 2>&1
tail -1 gpurun_out/ncu_proj_final.log

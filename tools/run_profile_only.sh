TAG=${1:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_gradient -s 4 -c 1 -o gpurun_out/prof_gradient_${TAG} -f python tools/prof_driver.py > gpurun_out/ncu_grad_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_project -s 12 -c 1 -o gpurun_out/prof_project_${TAG} -f python tools/prof_driver.py > gpurun_out/ncu_proj_${TAG}.log 2>&1
ls -la gpurun_out | tail -5

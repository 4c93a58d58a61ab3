mkdir -p gpurun_out
./profiles/ubench_fly > gpurun_out/ubench_fly.txt 2>&1
cat gpurun_out/ubench_fly.txt

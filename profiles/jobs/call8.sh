mkdir -p gpurun_out
./profiles/ubench_fly32 > gpurun_out/ubench_fly32.txt 2>&1
cat gpurun_out/ubench_fly32.txt

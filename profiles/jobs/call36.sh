mkdir -p gpurun_out


python profiles/quick_time.py modes > gpurun_out/c36_quick.txt 2>&1
cat gpurun_out/c36_quick.txt

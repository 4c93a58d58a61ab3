mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'fast|strided|n13' -c 12 -f -o gpurun_out/prof_c3c4u12 python profiles/prof_plan.py c3 c4 u12 > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/c3_ncu.log
ls -la gpurun_out

mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'fast|strided|n13' -s 1 -c 1 -f -o gpurun_out/prof_c5 python profiles/prof_plan.py c5 > gpurun_out/c4_ncu.log 2>&1
tail -3 gpurun_out/c4_ncu.log

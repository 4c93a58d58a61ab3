mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'n13' -s 1 -c 1 -f -o gpurun_out/prof_c5b python profiles/prof_plan.py c5 > gpurun_out/c32_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fast16' -s 1 -c 1 -f -o gpurun_out/prof_c2b python profiles/prof_plan.py c2 >> gpurun_out/c32_ncu.log 2>&1
tail -2 gpurun_out/c32_ncu.log; ls -la gpurun_out

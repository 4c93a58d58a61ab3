# round 2, job H: source-level ncu of the c5 kernel after the pre-shift change; c2 kernel for the SASS evidence
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:'n13' -s 1 -c 1 -f -o gpurun_out/r02h_c5 python profiles/prof_plan.py c5 > gpurun_out/r02h_c5.log 2>&1; tail -2 gpurun_out/r02h_c5.log
ls -la gpurun_out/r02h_c5.ncu-rep

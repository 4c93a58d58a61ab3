# round 2, job B: source-level ncu captures (one launch each) of the c5 kernel and the two c3 kernels
mkdir -p gpurun_out
export INTFFT_GROUP_MB=0
ncu --set full --import-source on --clock-control none -k regex:'n13' -s 1 -c 1 -f -o gpurun_out/r02b_c5 python profiles/prof_plan.py c5 > gpurun_out/r02b_c5.log 2>&1; tail -2 gpurun_out/r02b_c5.log
ncu --set full --import-source on --clock-control none -k regex:'fast' -s 2 -c 2 -f -o gpurun_out/r02b_c3 python profiles/prof_plan.py c3 > gpurun_out/r02b_c3.log 2>&1; tail -2 gpurun_out/r02b_c3.log
ls -la gpurun_out/*.ncu-rep

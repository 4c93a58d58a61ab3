# round 2, job O: 32-bit-lane butterfly forms in a register-only loop (is a three-multiply complex product worth it?)
mkdir -p gpurun_out
./profiles/bin/ubench_fly32 | tee gpurun_out/r02o_ubench_fly32.txt

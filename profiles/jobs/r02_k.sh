# round 2, job K: FP64-pipe issue rates beside IMAD.WIDE (is DFMA usable as an exact wide-integer multiplier?)
mkdir -p gpurun_out
./profiles/bin/ubench_fp64 | tee gpurun_out/r02k_ubench_fp64.txt

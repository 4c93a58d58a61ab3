// 32-bit-lane kernels, contiguous pass, DIT instantiations (see intfft_fast32.cuh)
#include "intfft_fast32.cuh"
namespace intfft {
int f32_launch_contig_dit(const f32::Fast32Params &p, int bits, int mode, int kind, int grid, void *stream)
{
    return (int)f32::launch_contig_n<true>(p, bits, mode, kind, grid, reinterpret_cast<cudaStream_t>(stream));
}
}  // namespace intfft

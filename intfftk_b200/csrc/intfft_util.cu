// Data-movement and test-support kernels: USE_FLY = 0 bypass, bit-reversal reorder (f1),
// on-device stimulus generator and checksum.
//
// Reference behaviour reproduced:
//   int_fftNk.vhd:260-277, 178-182   USE_FLY = '0': samples pass every stage untouched; in UNSCALED
//                                     mode they sit zero-extended in the wider data bus
//   buffers/int_bitrev_order.vhd:82-104  write linearly, read with the address bits reversed
#include <cuda_runtime.h>

#include <atomic>

#include "intfft_internal.h"
#include "intfft_taylor.cuh"

namespace intfft {

namespace {

std::atomic<long long> g_launches{0};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ long long ld_scalar(const void *p, long long i, int sb)
{
    if (sb == 2) return reinterpret_cast<const short *>(p)[i];
    if (sb == 4) return reinterpret_cast<const int *>(p)[i];
    return reinterpret_cast<const long long *>(p)[i];
}
__device__ __forceinline__ void st_scalar(void *p, long long i, int sb, long long v)
{
    if (sb == 2) reinterpret_cast<short *>(p)[i] = (short)v;
    else if (sb == 4) reinterpret_cast<int *>(p)[i] = (int)v;
    else reinterpret_cast<long long *>(p)[i] = v;
}

__global__ void bypass_kernel(const void *in, void *out, long long n, int in_sb, int out_sb, int dw, int zext)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        long long v = ld_scalar(in, i, in_sb);
        if (zext) v = dw >= 64 ? v : (long long)((unsigned long long)v & ((1ull << dw) - 1ull));
        else v = (long long)((unsigned long long)v << (64 - dw)) >> (64 - dw);
        st_scalar(out, i, out_sb, v);
    }
}

__global__ void fill_kernel(void *buf, long long n, int sb, int width, unsigned long long seed)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long h = mix64(seed + (unsigned long long)i * 0x9E3779B97F4A7C15ull);
        const long long v = (long long)(h << (64 - width)) >> (64 - width);
        st_scalar(buf, i, sb, v);
    }
}

__global__ void checksum_kernel(const void *buf, long long n, int sb, unsigned long long *sum)
{
    unsigned long long acc = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc += (unsigned long long)ld_scalar(buf, i, sb) * (mix64((unsigned long long)i) | 1ull);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, acc);
}

__device__ __forceinline__ unsigned rev_bits(unsigned v, int bits) { return bits ? (__brev(v) >> (32 - bits)) : 0u; }

// One CTA moves a 2^h x 2^h tile: rows = top h index bits, columns = low h index bits, for a fixed
// middle field; after reversal rows and columns swap roles, so both the read and the write touch
// runs of 2^h consecutive samples.  E = complex element (short2 / int2 / longlong2).
template <typename E>
__global__ void bitrev_kernel(const E *in, E *out, int n, int h, long long n_tiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *tile = reinterpret_cast<E *>(smem_raw);
    const int side = 1 << h, pitch = side + 1;
    const int mid_bits = n - 2 * h;
    const unsigned tx = threadIdx.x & (side - 1), ty = threadIdx.x >> h;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned mid = (unsigned)(t & ((1ll << mid_bits) - 1));
        const long long frame = t >> mid_bits;
        const long long src = (frame << n) + ((long long)ty << (n - h)) + ((long long)mid << h) + tx;
        tile[ty * pitch + tx] = in[src];
        __syncthreads();
        // destination row = reversed source column, destination column = reversed source row
        const unsigned s_col = rev_bits(ty, h), s_row = rev_bits(tx, h);
        const long long dst = (frame << n) + ((long long)ty << (n - h)) +
                              ((long long)rev_bits(mid, mid_bits) << h) + tx;
        out[dst] = tile[s_row * pitch + s_col];
        __syncthreads();
    }
}

// Same permutation with 16-byte global accesses: a 2^h x 2^h tile (h = 6 for NFFT >= 12) is read as
// 16-byte vectors along its rows (runs of 2^h consecutive samples), parked element-wise in a padded
// shared tile, and written as 16-byte vectors along the rows of the transposed, index-reversed tile.
// 256 threads move 4096 samples per trip, so every thread keeps 64..256 bytes in flight.
template <typename E>
__global__ void __launch_bounds__(256) bitrev_vec_kernel(const E *in, E *out, int n, int h, long long n_tiles)
{
    constexpr int V = 16 / (int)sizeof(E);          // elements per 16-byte vector: 4 / 2 / 1
    union Vec { uint4 q; E e[V]; };
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *tile = reinterpret_cast<E *>(smem_raw);
    const int side = 1 << h, pitch = side + 1;
    const int vpr_log2 = h - (V == 4 ? 2 : (V == 2 ? 1 : 0));     // log2 vectors per row
    const int n_vec = side << vpr_log2;
    const int mid_bits = n - 2 * h;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned mid = (unsigned)(t & ((1ll << mid_bits) - 1));
        const long long frame = t >> mid_bits;
        const E *src = in + (frame << n) + ((long long)mid << h);
        for (int v = threadIdx.x; v < n_vec; v += blockDim.x) {
            const int row = v >> vpr_log2, c0 = (v & ((1 << vpr_log2) - 1)) * V;
            Vec x;
            x.q = __ldg(reinterpret_cast<const uint4 *>(src + ((long long)row << (n - h)) + c0));
#pragma unroll
            for (int e = 0; e < V; ++e) tile[row * pitch + c0 + e] = x.e[e];
        }
        __syncthreads();
        E *dst = out + (frame << n) + ((long long)rev_bits(mid, mid_bits) << h);
        for (int v = threadIdx.x; v < n_vec; v += blockDim.x) {
            const int row = v >> vpr_log2, c0 = (v & ((1 << vpr_log2) - 1)) * V;
            const unsigned s_col = rev_bits((unsigned)row, h);
            Vec x;
#pragma unroll
            for (int e = 0; e < V; ++e) x.e[e] = tile[rev_bits((unsigned)(c0 + e), h) * pitch + s_col];
            *reinterpret_cast<uint4 *>(dst + ((long long)row << (n - h)) + c0) = x.q;
        }
        __syncthreads();
    }
}

template <typename E>
cudaError_t launch_bitrev_vec(const void *in, void *out, int n, int h, long long n_tiles, cudaStream_t st)
{
    const int side = 1 << h;
    const size_t smem = (size_t)side * (side + 1) * sizeof(E);
    cudaError_t e = cudaFuncSetAttribute(bitrev_vec_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = n_tiles < 148 * 8 ? n_tiles : 148 * 8;
    bitrev_vec_kernel<E><<<(int)grid, 256, smem, st>>>((const E *)in, (E *)out, n, h, n_tiles);
    return cudaGetLastError();
}

}  // namespace

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

static int grid_for(long long n, int block)
{
    long long g = (n + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;
    return g < 1 ? 1 : (int)g;
}

int launch_bypass(const void *in, void *out, long long n_scalars, int in_sb, int out_sb, int dw,
                  int zero_extend, void *stream)
{
    bypass_kernel<<<grid_for(n_scalars, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n_scalars, in_sb, out_sb, dw, zero_extend);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_fill_random(void *buf, long long n_scalars, int sb, int width, uint64_t seed, void *stream)
{
    fill_kernel<<<grid_for(n_scalars, 256), 256, 0, (cudaStream_t)stream>>>(buf, n_scalars, sb, width, seed);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_checksum(const void *buf, long long n_scalars, int sb, uint64_t *d_sum, void *stream)
{
    checksum_kernel<<<grid_for(n_scalars, 256), 256, 0, (cudaStream_t)stream>>>(buf, n_scalars, sb, (unsigned long long *)d_sum);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_bitrev(int n, int sb, long long batch, const void *in, void *out, void *stream)
{
    // 16-byte path: rows of 2^h samples must hold whole, aligned vectors
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    if (aligned && n >= 8) {
        const int hv = n / 2 < 6 ? n / 2 : 6;
        const long long tiles = batch << (n - 2 * hv);
        cudaStream_t sv = (cudaStream_t)stream;
        cudaError_t e = sb == 2 ? launch_bitrev_vec<short2>(in, out, n, hv, tiles, sv)
                      : sb == 4 ? launch_bitrev_vec<int2>(in, out, n, hv, tiles, sv)
                                : launch_bitrev_vec<longlong2>(in, out, n, hv, tiles, sv);
        count_launch();
        return (int)e;
    }
    const int h = n / 2 < 5 ? n / 2 : 5;
    const int side = 1 << h;
    const long long n_tiles = batch << (n - 2 * h);
    long long grid = n_tiles < 148 * 8 ? n_tiles : 148 * 8;
    if (grid < 1) grid = 1;
    const size_t smem = (size_t)side * (side + 1) * 2 * sb;
    cudaStream_t st = (cudaStream_t)stream;
    if (sb == 2) bitrev_kernel<short2><<<(int)grid, side * side, smem, st>>>((const short2 *)in, (short2 *)out, n, h, n_tiles);
    else if (sb == 4) bitrev_kernel<int2><<<(int)grid, side * side, smem, st>>>((const int2 *)in, (int2 *)out, n, h, n_tiles);
    else bitrev_kernel<longlong2><<<(int)grid, side * side, smem, st>>>((const longlong2 *)in, (longlong2 *)out, n, h, n_tiles);
    count_launch();
    return (int)cudaGetLastError();
}

// test hook: the device Taylor function over a whole stage (intfft_twiddles_device)
namespace {
__global__ void taylor_table_kernel(const __grid_constant__ TaylorDev t, int stage, int2 *out)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < (1u << stage)) out[k] = taylor_twiddle(t, stage, k);
}
}  // namespace

int launch_taylor_table(const TaylorDev &tay, int stage, int2 *d_out, void *stream)
{
    const unsigned n = 1u << stage;
    taylor_table_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tay, stage, d_out);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace intfft

// 2-D TMA helpers shared by the strided kernels: the column block of a frame (2^G rows of 2^(12-G) samples at a
// pitch of 2^(NFFT-G) samples) is one box of a tensor map over the whole batch, moved by the copy engine in both
// directions (cp.async.bulk.tensor.2d: UTMALDG / UTMASTG in SASS) instead of per-thread cp.async / STG.
// This is the "stride-2^s shuffle" of the reference's delay lines (delay/int_delay_line.vhd:52-104) done in hardware.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace intfft {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void store_2d(const CUtensorMap *map, int c0, int c1, const void *src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// at most N of the most recent tensor stores may still be READING their shared-memory source
template <int N> __device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the TMA engine
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}
// A batch of frames as a 2-D tensor of samples: `row_samples` samples per row, `rows` rows, box = box_samples x box_rows.
// A sample is one 32-bit word (packed-16 containers) or one 64-bit word (32-bit containers); coordinates are in samples.
// No swizzle: the landing tile is dense.
inline bool make_map(CUtensorMap *map, const void *base, int sample_words, uint64_t row_samples, uint64_t rows,
                     uint32_t box_samples, uint32_t box_rows)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || box_samples > 256 || box_rows > 256) return false;
    const cuuint64_t dims[2] = {row_samples, rows};
    const cuuint64_t strides[1] = {row_samples * 4 * (uint64_t)sample_words};
    const cuuint32_t box[2] = {box_samples, box_rows}, estr[2] = {1, 1};
    return fn(map, sample_words == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2,
              const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tma
}  // namespace intfft

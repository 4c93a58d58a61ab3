/*
 * intfft.h — C-ABI of libintfft_b200: a B200-native (sm_100a) integer FFT/IFFT engine that
 * reproduces, bit for bit, the arithmetic of the intfftk FPGA cores `int_fftNk` / `int_ifftNk`.
 *
 * Drop-in boundary.  The reference has no software API; its boundary for this path is the VHDL
 * entity interface (generics + stream ports):
 *     int_fftNk   src/vhdl/fft/int_fftNk.vhd:72-103
 *     int_ifftNk  src/vhdl/fft/int_ifftNk.vhd:71-102
 * `intfft_generics` mirrors the entity generics 1:1 (same names, same meaning); a "plan" is an
 * elaborated entity; `intfft_exec*` is "clock N/2 beats per frame through DI_* and collect DO_*"
 * for a whole batch of frames.  Illegal generic combinations fail at plan creation with
 * INTFFT_EINVAL, mirroring "does not elaborate" in the reference.
 *
 * Stream contract / flat layout:
 *   A frame is N = 2^NFFT complex samples stored interleaved {re, im} (re at the lower address,
 *   matching the `im & re` packing of int_fftNk.vhd:285), transforms contiguous in the batch.
 *   FFT  (direction 0, int_fftNk,  DIF): in[i]  = x[i] natural order (lane 0 = DI_*0 = first half,
 *        lane 1 = DI_*1 = second half, int_fftNk.vhd:15-17); out[q] = X[bitrev_NFFT(q)]
 *        (lane 0 = even q, lane 1 = odd q, int_fftNk.vhd:19-21).
 *   IFFT (direction 1, int_ifftNk, DIT): in[q] in that same bit-reversed order (lane 0 = even q,
 *        lane 1 = odd q, int_ifftNk.vhd:15-17); out[i] natural order (lanes = halves, :19-21).
 *   Scalar container: int16 when the width is <= 16 bits, int32 when <= 32, else int64;
 *   sign-extended two's complement.  Input width = DATA_WIDTH, output width =
 *   DATA_WIDTH + FORMAT*NFFT (int_fftNk.vhd:97-100).  Input values outside DATA_WIDTH bits are
 *   wrapped (low DATA_WIDTH bits kept), like conv_std_logic_vector in tb/fft_signle_test.vhd:163.
 *
 * No torch types, no C++ types, no exceptions cross this boundary.  All functions return
 * INTFFT_OK (0) or a negative status.  The library has NO CPU fallback: every exec entry point
 * requires a CUDA device and fails with INTFFT_ECUDA when none is usable.
 */
#ifndef INTFFT_H_
#define INTFFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INTFFT_OK            0
#define INTFFT_EINVAL       (-1)  /* generics do not elaborate in the reference / bad argument      */
#define INTFFT_ECUDA        (-2)  /* CUDA runtime error (no device, launch failure, ...)            */
#define INTFFT_ENOMEM       (-3)  /* host or device allocation failed                               */
#define INTFFT_EUNSUPPORTED (-4)  /* legal in the reference but outside this build's 64-bit lanes   */

/* Entity generics of int_fftNk / int_ifftNk (int_fftNk.vhd:73-84, int_ifftNk.vhd:72-83). */
typedef struct intfft_generics {
    int32_t nfft_log2;   /* NFFT: number of stages, N = 2^NFFT. Reference: 3..19; 20 = documented
                            extension of row_twiddle_tay.                         */
    int32_t data_width;  /* DATA_WIDTH: input sample width in bits (per re / im).                  */
    int32_t twdl_width;  /* TWDL_WIDTH: twiddle width, 8..27 (XSER NEW) / 8..25 (XSER OLD).        */
    int32_t format;      /* FORMAT: 1 = UNSCALED (1 bit growth per stage), 0 = SCALED.             */
    int32_t rndmode;     /* RNDMODE: 0 = TRUNCATE, 1 = ROUNDING (scaled mode only).                */
    int32_t xser;        /* XSER: 0 = "OLD" (DSP48E1), 1 = "NEW" (DSP48E2). Changes the numbers.   */
    int32_t use_fly;     /* USE_FLY port held constant: 1 = butterflies on, 0 = bypass.            */
    int32_t direction;   /* 0 = int_fftNk (DIF, natural in, bit-reversed out),
                            1 = int_ifftNk (DIT, bit-reversed in, natural out).                    */
} intfft_generics;

/* What a plan reads and writes (per scalar = one of re / im). */
typedef struct intfft_layout {
    int64_t n;               /* points per transform, 2^NFFT                                       */
    int64_t batch;           /* transforms per exec                                                */
    int32_t in_width;        /* DATA_WIDTH                                                         */
    int32_t out_width;       /* DATA_WIDTH + FORMAT*NFFT                                           */
    int32_t in_scalar_bytes; /* 2, 4 or 8                                                          */
    int32_t out_scalar_bytes;/* 2, 4 or 8                                                          */
    int64_t in_bytes;        /* batch * n * 2 * in_scalar_bytes                                    */
    int64_t out_bytes;       /* batch * n * 2 * out_scalar_bytes                                   */
    int32_t n_passes;        /* kernel launches per exec (1 = whole transform in one tile)         */
    int32_t lane_bits;       /* arithmetic lane used on the device: 32 or 64                       */
} intfft_layout;

typedef struct intfft_plan intfft_plan;

/* Elaboration check only; no device needed. Mirrors which generic combinations the reference can
 * elaborate (int_cmult_dsp48.vhd:182-434, int_dif2_fly.vhd:86-116, row_twiddle_tay.vhd:156). */
int intfft_validate(const intfft_generics *g);

/* Stands in for elaborating int_fftNk / int_ifftNk with these generics on `device`;
 * builds the fixed-point twiddle ROM/Taylor tables (rom_twiddle_int.vhd:135-159,
 * row_twiddle_tay.vhd:123-268) and uploads them. */
int intfft_plan_create(intfft_plan **out, const intfft_generics *g, int64_t batch, int device);
int intfft_plan_destroy(intfft_plan *p);
int intfft_query(const intfft_plan *p, intfft_layout *l);

/* Run `batch` frames. d_in / d_out are DEVICE pointers in the flat layout above; asynchronous on
 * `cuda_stream` (a cudaStream_t, NULL = default stream).  (Environment knob INTFFT_GROUP_MB=<n>: run a two-pass plan
 * in groups of frames whose intermediate fits n MB of L2; off by default — as separate launches it measured slower.) d_in == d_out is allowed when the input
 * and output containers have the same size. Both pointers must be 16-byte aligned (the kernels move
 * frames with 16-byte vector and TMA bulk accesses; cudaMalloc memory and any whole-frame offset into
 * it qualify), otherwise INTFFT_EINVAL. Replaces driving DI_RE0/IM0/RE1/IM1 + DI_ENA and
 * sampling DO_* + DO_VAL (int_fftNk.vhd:86-102). */
int intfft_exec(intfft_plan *p, const void *d_in, void *d_out, void *cuda_stream);

/* Same through HOST buffers: H2D copy, exec, D2H copy, synchronised on return.  This is the call a
 * testbench-style host makes (tb/fft_signle_test.vhd:154-358 replays a file through the core).
 * The batch streams through a ring of three ~32 MiB device staging buffers on three streams (copy in /
 * kernels / copy out), so the device footprint does not depend on the batch; pinned host buffers
 * (intfft_host_alloc) are needed for the copies to overlap.  On any error every internal stream is
 * synchronised before returning: no copy touches h_in / h_out afterwards. */
int intfft_exec_host(intfft_plan *p, const void *h_in, void *h_out);

/* Page-locked host memory, usable from every device of the process (cudaHostAllocPortable). */
int intfft_host_alloc(void **h_ptr, size_t bytes);
int intfft_host_free(void *h_ptr);

/* Threading.  A plan is immutable after creation as far as the DEVICE path is concerned: intfft_exec and
 * intfft_exec_natural copy the per-call state (pointers, frame counts, the natural-order flag) and may be
 * called concurrently from several host threads on different streams (the caller orders accesses to the
 * buffers, and to the plan's own intermediate for multi-pass plans and intfft_exec_natural, by using one
 * stream per plan or events).  The HOST path (intfft_exec_host, intfft_pair_exec_host) owns staging buffers
 * and streams that are created on first use; those calls are serialised per plan by an internal lock.
 * intfft_exec / intfft_exec_natural / intfft_pair_exec only enqueue kernels on the given stream (no allocation, no
 * synchronisation, tensor maps travel by value in the kernel arguments), so they may be captured into a CUDA graph and
 * replayed with the same buffers. */

/* Twiddle read-back: the W(k), k = 0 .. 2^stage - 1, that rom_twiddle_int(STAGE = stage) streams
 * (rom_twiddle_int.vhd:98-248); stage >= 2.  Host-only, needs no device. */
int intfft_twiddles(const intfft_generics *g, int stage, int32_t *h_re, int32_t *h_im);
/* Test hook for the on-device Taylor path (strided kernels of NFFT >= 17 plans recompute their STAGE >= 11
 * twiddles from the 512-entry coarse ROM instead of reading 2^NFFT-entry tables): the same device function,
 * evaluated for every k of `stage` (11..19) on `device`.  Must equal intfft_twiddles bit for bit.
 * Replaces: rom_twiddle_int.vhd:215-246 + row_twiddle_tay.vhd:123-382 as they run in hardware (per sample). */
int intfft_twiddles_device(const intfft_generics *g, int stage, int32_t *h_re, int32_t *h_im, int device);

/* f2: the FFT -> IFFT loop-back of int_fft_ifft_pair (main/int_fft_ifft_pair.vhd:209-283):
 * int_fftNk(generics) feeding int_ifftNk with DATA_WIDTH + FORMAT*NFFT input bits (:261), natural order in,
 * natural order out, output width DATA_WIDTH + 2*FORMAT*NFFT (:98-101).  g->use_fly is FLY_FWD, fly_inv is
 * FLY_INV (:91-92); g->direction is ignored.  Specified on the core lanes (flat in-place order), not on the
 * wrapper's interleave-2 I/O buffers.  The spectrum stays on the device between the two cores. */
typedef struct intfft_pair intfft_pair;
int intfft_pair_create(intfft_pair **out, const intfft_generics *g, int fly_inv, int64_t batch, int device);
int intfft_pair_destroy(intfft_pair *p);
int intfft_pair_query(const intfft_pair *p, intfft_layout *l);
int intfft_pair_exec(intfft_pair *p, const void *d_in, void *d_out, void *cuda_stream);
/* The same through host buffers (what tb/fft_double_test.vhd:154-217 does with di_double.dat -> dout_pair.dat):
 * the spectrum never leaves the device. */
int intfft_pair_exec_host(intfft_pair *p, const void *h_in, void *h_out);

/* Multi-GPU: ONE process driving several devices.  Frames are independent (the reference
 * streams frame after frame through one core, int_fftNk.vhd:15-21), so the batch is cut into contiguous
 * shards — shard i = frames [batch*i/n, batch*(i+1)/n) on devices[i] — and every device runs the same plan
 * on its shard; there is no exchange step and no collective.  intfft_multi_exec_host takes ONE host batch
 * (pinned: intfft_host_alloc) and drives one copy/compute pipeline per device from its own host thread.
 * intfft_multi_exec runs device-resident shards: d_in[i] / d_out[i] / cuda_streams[i] live on devices[i]
 * (cuda_streams may be NULL = default streams); asynchronous like intfft_exec. */
typedef struct intfft_multi intfft_multi;
int intfft_multi_create(intfft_multi **out, const intfft_generics *g, int64_t batch, const int *devices, int n_devices);
int intfft_multi_destroy(intfft_multi *m);
int intfft_multi_devices(const intfft_multi *m);
int intfft_multi_query(const intfft_multi *m, intfft_layout *l);   /* layout of the WHOLE batch */
int intfft_multi_shard(const intfft_multi *m, int i, int *device, int64_t *first_frame, int64_t *frames);
int intfft_multi_exec_host(intfft_multi *m, const void *h_in, void *h_out);
int intfft_multi_exec(intfft_multi *m, const void *const *d_in, void *const *d_out, void *const *cuda_streams);

/* f1: int_fft_single_path semantics — natural order in AND out
 * (main/int_fft_single_path.vhd:157-268: inbuf_half_path -> int_fftNk -> outbuf_half_path -> int_bitrev_order).
 * FFT plans: exec, then the bit-reversal reorder; IFFT plans: reorder, then exec.  d_in != d_out.
 * The plan owns the intermediate buffer (allocated on first use). */
int intfft_exec_natural(intfft_plan *p, const void *d_in, void *d_out, void *cuda_stream);

/* f1: bit-reversal reorder of a batch of frames, the job int_bitrev_order does in
 * int_fft_single_path (buffers/int_bitrev_order.vhd:82-104): out[bitrev(q)] = in[q].
 * scalar_bytes in {2,4,8}; d_in != d_out. */
int intfft_bitrev(int nfft_log2, int scalar_bytes, int64_t batch,
                  const void *d_in, void *d_out, int device, void *cuda_stream);

/* Synthetic stimulus, generated on the device: re/im i.i.d. uniform over the full `width`-bit
 * two's-complement range from a counter-based hash of (seed, scalar index).  Stand-in for the
 * stimulus files of math/fft_single.m:94-98.  scalar_bytes in {2,4,8}. */
int intfft_fill_random(void *d_buf, int64_t n_scalars, int scalar_bytes, int width,
                       uint64_t seed, int device, void *cuda_stream);

/* 64-bit order-sensitive checksum of a device buffer of scalars (sum of value * odd hash(index)). */
int intfft_checksum(const void *d_buf, int64_t n_scalars, int scalar_bytes,
                    uint64_t *h_sum, int device, void *cuda_stream);

/* Which kernels a plan for these generics would run, as text ("fast32_strided[bits 8..15] -> fast64[bits 0..7,
 * instance 1]"), written to buf (at most len bytes, NUL-terminated).  Host-only, needs no device: the stand-in for
 * reading the elaboration log of the reference (which multiplier / delay-line variants were generated), and
 * what the CPU test suite uses to pin the kernel selection of every BASELINE configuration. */
int intfft_describe(const intfft_generics *g, int64_t batch, char *buf, size_t len);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t intfft_launch_count(void);

const char *intfft_strerror(int status);
int intfft_version(void);

#ifdef __cplusplus
}
#endif
#endif /* INTFFT_H_ */

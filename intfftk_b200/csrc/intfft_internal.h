// Internal declarations shared by the host-side plan code and the sm_100a kernels.
// Nothing here crosses the C-ABI (include/intfft.h).
#pragma once
#include <cstdint>
#include <mutex>
#include <vector>

#include <vector_types.h>

#include "../../include/intfft.h"

namespace intfft {

// ---- complex-multiplier arrangement (int_cmult_dsp48.vhd:182-434), plan-wide constants --------
// variant chosen per stage by the multiplier data width dtwc:
//   dtwc < lim_single           -> one DSP48 pair:   (P2 +- P1) >> sh_single
//   lim_single <= dtwc < lim_dbl-> double:           wrap48((P2>>k_pre) +- (P1>>k_pre)) >> sh_post
//   lim_dbl <= dtwc             -> triple:           wrap((P2>>sh_single)) +- wrap((P1>>sh_single))
struct CmultConsts {
    int lim_single, lim_dbl, lim_none;
    int sh_single, k_pre, sh_post;
    int trpl_awd;     // triple arrangement: the wide multiplier's data port is this many bits (61 / 59 for trpl18:
                      // dspA_M1 <= SXT(M1_AA, AWD), int_cmult_trpl18_dsp48.vhd:161-162); wider data is cut there
    int trpl_pwd;     // and its product has this many bits (79 / 77): DTW + TWD - 2 must stay below it (:151)
};
CmultConsts cmult_consts(int twdl_width, int xser);

// ---- twiddle generator: rom_twiddle_int.vhd + row_twiddle_tay.vhd, host side ------------------
// fills re/im[0 .. 2^stage) with the stream rom_twiddle_int(STAGE = stage) produces.
void twiddle_stage_table(int stage, int twdl_width, int xser, int32_t *re, int32_t *im);

// ---- on-device Taylor twiddles (intfft_taylor.cuh): what a kernel needs to recompute W_s[k], s >= 11 ----
struct TaylorDev {
    const int2 *rom9;      // rom_twiddle_int's coarse ROM of depth 9 (512 entries, device memory)
    int on;                // 0: every twiddle comes from the tables
    int tw, xs;            // TWDL_WIDTH, XSHIFT (21 NEW / 23 OLD)
    int e;                 // left shift applied to the result (pre-shifted-table kernels), else 0
    int mathpi[9];         // MATHPI of STAGE 11 .. 19
};
// fills the 512-entry coarse ROM (as (c_i, s_i) pairs) and the nine MATHPI constants
void taylor_consts(int twdl_width, int xser, int32_t *rom_c, int32_t *rom_s, int *mathpi, int *xshift);

// ---- kernel-facing description of one pass over the data --------------------------------------
enum LaneKind { LANE_I32_P64 = 0, LANE_I64_P64 = 1, LANE_I64_P128 = 2 };
enum ModeKind { MODE_TRUNC = 0, MODE_ROUND = 1, MODE_UNSCALED = 2 };

struct PassParams {
    const void *in;
    void *out;
    const int2 *tw;        // twiddles, entry (1 << s) + k for stage s >= 2
    long long total;       // batch * N complex samples
    long long n_tiles;
    int n;                 // NFFT
    int L;                 // log2 samples per tile (threads = 2^(L-4))
    int c;                 // log2 contiguous run ("column") length inside a tile
    int pb;                // lowest global bit handled by this pass
    int g;                 // number of stage bits handled by this pass (local bits [c, c+g))
    int dw;                // DATA_WIDTH
    int format;            // FORMAT
    int in_sb, out_sb;     // scalar bytes of the containers this pass reads / writes
    int in_wrap;           // wrap loaded scalars to dw bits (first pass only)
    CmultConsts cm;
    int nrounds;
    signed char r_lo[8];   // local bit where round r starts
    signed char r_n[8];    // stage bits in round r (1..4)
};

struct PassDesc {
    PassParams kp;         // in/out/tw/n_tiles filled at exec time
    int lane;              // LaneKind
    int threads;
    size_t smem_bytes;
    int scratch_in, scratch_out;  // -1 = user buffer, else index of plan scratch buffer
    int path;              // 0 generic tile kernel, 1 packed-16 kernels, 2 32-bit-lane kernels, 3 64-bit-lane low-8 kernel,
                           // 4 64-bit-lane strided kernel
    int natural = 0;       // exec time (set on a per-call COPY of the descriptor, never on the plan's): this single
                           // 4096-point packed-16 DIF pass also applies int_bitrev_order
};

// Host-buffer pipeline state (intfft_exec_host and friends): a ring of chunk-sized device staging buffers and
// three streams (H2D copies / kernels / D2H copies).  Created on first use, guarded by `mu`: host-path calls on
// one plan are serialised, device-path calls (intfft_exec) never touch it.
struct HostPipe {
    static constexpr int kSlots = 3;
    std::mutex mu;
    void *d_in[kSlots] = {nullptr, nullptr, nullptr}, *d_out[kSlots] = {nullptr, nullptr, nullptr};
    long long chunk_frames = 0;
    void *s_in = nullptr, *s_k = nullptr, *s_out = nullptr;   // cudaStream_t
    void *ev_in[kSlots] = {nullptr, nullptr, nullptr};        // cudaEvent_t: chunk landed in its slot
    void *ev_k[kSlots] = {nullptr, nullptr, nullptr};         //              kernels of the slot's chunk done
    void *ev_out[kSlots] = {nullptr, nullptr, nullptr};       //              slot's result copied back (slot free)
};

struct Plan {
    intfft_generics g;
    int64_t batch;
    int device;
    int mode;              // ModeKind
    int in_width, out_width, in_sb, out_sb;
    std::vector<PassDesc> passes;
    int2 *d_tw = nullptr;        // device twiddle table (int32 pairs)
    int2 *d_twp = nullptr;       // twiddles pre-shifted for the 32-bit-product kernel (fast16 plans only)
    int lw_r[16] = {0}, lw_i[16] = {0};  // its lowest-round twiddles (stages 2, 3)
    int lw32_r[16] = {0}, lw32_i[16] = {0};  // same, not pre-shifted (32-bit-lane kernels)
    int2 *d_twp32 = nullptr;     // twiddles << (31 - sh_single) for the 32-bit-lane TRUNCATE kernels (KIND_SINGLE_PRE)
    int lwp32_r[16] = {0}, lwp32_i[16] = {0};
    int2 *d_rom9 = nullptr;      // coarse ROM of the on-device Taylor path
    TaylorDev tay{};             // .on = 1: STAGE >= 12 tables are not even built (NFFT >= 17 plans on the strided kernels)
    TaylorDev tay16{};           // the same with the packed-16 kernels' pre-shift
    unsigned *d_tw12p = nullptr; // STAGE-12 twiddles packed {re:16 | im:16} (one-pass 8192-point kernel, TWDL_WIDTH <= 16)
    void *scratch[2] = {nullptr, nullptr};
    size_t scratch_bytes[2] = {0, 0};
    // Two-pass plans run group by group: `group_frames` frames go through BOTH passes before the next group
    // starts, so what the first pass wrote is still in the 126 MB L2 when the second pass reads it (and, for
    // in-place intermediates, is overwritten there before it is ever written back): HBM sees the algorithmic
    // bytes once.  0 = whole batch at once.
    long long group_frames = 0;
    void *nat = nullptr;                  // bit-reversed intermediate of intfft_exec_natural (allocated at creation
    size_t nat_bytes = 0;                 // time when the plan needs one, so exec never mutates the plan)
    HostPipe pipe;
    int num_sms = 0;
};

// kernels (intfft_tile.cu / intfft_fast16.cu / intfft_util.cu); all return cudaError_t as int
int launch_tile_pass(const PassDesc &pd, int mode, bool dit, int num_sms, void *stream);
int launch_fast16(const PassDesc &pd, int mode, bool dit, const int2 *twp, const int *lw_r, const int *lw_i,
                  int num_sms, void *stream);
bool fast16_supported(const intfft_generics &g);
bool fast16_pair_supported(const intfft_generics &g);
int launch_fast16_n13(const PassDesc &pd, int mode, bool dit, const int2 *twp, const int *lw_r, const int *lw_i,
                      int num_sms, void *stream);
int launch_fast16_n14(const PassDesc &pd, int mode, bool dit, const int2 *twp, const int *lw_r, const int *lw_i,
                      int num_sms, void *stream);
int launch_fast16_pair(const PassDesc &pd, int mode, const int2 *twp, const int *lw_r, const int *lw_i, int num_sms,
                       void *stream);
int launch_fast16_strided(const PassDesc &pd, int mode, bool dit, const int2 *twp, int num_sms, void *stream,
                          const TaylorDev *tay = nullptr);
bool fast32_supported(const intfft_generics &g);
int launch_fast32(const PassDesc &pd, int mode, bool dit, const int2 *tw, const int *lw_r, const int *lw_i,
                  int num_sms, void *stream, const int2 *twp = nullptr, const int *lwp_r = nullptr,
                  const int *lwp_i = nullptr, const unsigned *tw16 = nullptr, const TaylorDev *tay = nullptr);
int launch_taylor_table(const TaylorDev &tay, int stage, int2 *d_out, void *stream);     // test hook (intfft_util.cu)
// 64-bit-lane kernel for the lowest eight stage bits (intfft_fast64.cu)
int fast64_uniform_kind(const PassParams &kp, bool dit);
int launch_fast64_strided(const PassDesc &pd, int mode, bool dit, const int2 *tw, int num_sms, void *stream);
int launch_fast64(const PassDesc &pd, int mode, bool dit, const int2 *tw, const int *lw_r, const int *lw_i,
                  int num_sms, void *stream);
int launch_bypass(const void *in, void *out, long long n_scalars, int in_sb, int out_sb, int dw,
                  int zero_extend, void *stream);
int launch_bitrev(int nfft_log2, int scalar_bytes, long long batch, const void *in, void *out, void *stream);
int launch_fill_random(void *buf, long long n_scalars, int sb, int width, uint64_t seed, void *stream);
int launch_checksum(const void *buf, long long n_scalars, int sb, uint64_t *d_sum, void *stream);
void count_launch(int n = 1);
long long launches();

}  // namespace intfft

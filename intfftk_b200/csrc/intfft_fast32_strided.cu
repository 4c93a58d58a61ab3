// 32-bit-lane kernels: strided-pass instantiations and the host-side dispatcher (see intfft_fast32.cuh)
#include <cstdlib>

#include "intfft_fast32.cuh"

namespace intfft {

int f32_launch_strided(const f32::Fast32Params &p, int g, bool dit, int mode, int kind, int grid, void *stream)
{
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const char *tma_env = std::getenv("INTFFT_STRIDED_TMA");              // =0: the cp.async / STG form
    const bool use_tma = !(tma_env && tma_env[0] == '0');
    // cudaErrorNotSupported from the TMA launchers = no tensor map could be encoded: the cp.async / STG kernels do the same work
    if (g == 4) {
        cudaError_t e = cudaErrorNotSupported;
        if (use_tma) e = dit ? f32::launch_strided_tma<4, true>(p, mode, kind, grid, st) : f32::launch_strided_tma<4, false>(p, mode, kind, grid, st);
        if (e != cudaErrorNotSupported) return (int)e;
        return (int)(dit ? f32::launch_strided<4, true>(p, mode, kind, grid, st) : f32::launch_strided<4, false>(p, mode, kind, grid, st));
    }
    if (g == 8) {
        cudaError_t e = cudaErrorNotSupported;
        if (use_tma) e = dit ? f32::launch_strided_tma<8, true>(p, mode, kind, grid, st) : f32::launch_strided_tma<8, false>(p, mode, kind, grid, st);
        if (e != cudaErrorNotSupported) return (int)e;
        return (int)(dit ? f32::launch_strided<8, true>(p, mode, kind, grid, st) : f32::launch_strided<8, false>(p, mode, kind, grid, st));
    }
    return (int)cudaErrorInvalidValue;
}

// plans whose every intermediate fits 32-bit lanes (ROUNDING needs one carry bit)
bool fast32_supported(const intfft_generics &g)
{
    if (!g.use_fly || g.nfft_log2 < 3) return false;
    const int worst = g.data_width + g.format * g.nfft_log2 + ((!g.format && g.rndmode) ? 1 : 0);
    return worst <= 32;
}

// twp / lwp_r / lwp_i: the same twiddles pre-shifted by 31 - sh_single (nullptr when the plan has none)
int launch_fast32(const PassDesc &pd, int mode, bool dit, const int2 *tw, const int *lw_r, const int *lw_i,
                  int num_sms, void *stream, const int2 *twp, const int *lwp_r, const int *lwp_i, const unsigned *tw16,
                  const TaylorDev *tay)
{
    f32::Fast32Params p{};
    p.in = pd.kp.in;
    p.out = pd.kp.out;
    p.tw = tw;
    p.total = pd.kp.total;
    p.n = pd.kp.n;
    p.batch = pd.kp.total >> pd.kp.n;
    p.n_tiles = (pd.kp.total + 4095) >> 12;
    p.dw = pd.kp.dw;
    p.format = pd.kp.format;
    p.in_sb = pd.kp.in_sb;
    p.out_sb = pd.kp.out_sb;
    p.in_wrap = pd.kp.in_wrap;
    p.cm = pd.kp.cm;
    for (int i = 0; i < 16; ++i) { p.lw_r[i] = lw_r[i]; p.lw_i[i] = lw_i[i]; }
    // multiplier arrangement policy: all stages of this pass single-DSP, or decided per stage
    int kind = f32::KIND_SINGLE, kind_lo = f32::KIND_SINGLE;     // kind_lo: the same question for STAGE < 8 only
    {
        const int n = pd.kp.n, fmt = pd.kp.format;
        for (int b = pd.kp.pb; b < pd.kp.pb + pd.kp.g; ++b) {
            const int ii = dit ? b : n - 1 - b;
            const int dtw = pd.kp.dw + ii * fmt;
            const int dtwc = dit ? dtw : dtw + fmt;
            if (b >= 2 && dtwc >= pd.kp.cm.lim_single) {
                kind = f32::KIND_MIXED;
                if (b < 8) kind_lo = f32::KIND_MIXED;
            }
        }
    }
    long long grid = 2ll * num_sms;
    int e = (int)cudaErrorInvalidValue;
    // DIT, TRUNCATE, every stage of the pass single-DSP, TWDL_WIDTH < 19: the pre-shifted-twiddle instances
    const bool pre = mode == MODE_TRUNC && kind == f32::KIND_SINGLE && twp && !std::getenv("INTFFT_NO_PRESHIFT");
    if (pre && !(pd.kp.c == 0 && pd.kp.g == 13)) {
        p.tw = twp;
        for (int i = 0; i < 16; ++i) { p.lw_r[i] = lwp_r[i]; p.lw_i[i] = lwp_i[i]; }
        kind = f32::KIND_SINGLE_PRE;
    }
    if (pd.kp.c == 0 && pd.kp.g == 13) {               // one-pass 8192-point kernel
        p.n_tiles = pd.kp.total >> 13;
        grid = 3ll * num_sms;                          // 80 registers, 70 KB of shared memory: three CTAs per SM
        if (grid > p.n_tiles) grid = p.n_tiles;
        if (grid < 1) grid = 1;
        if (mode == MODE_TRUNC && kind == f32::KIND_SINGLE && twp && tw16 && !std::getenv("INTFFT_NO_PRESHIFT")) {
            p.tw = twp;
            p.tw16 = tw16;
            for (int i = 0; i < 16; ++i) { p.lw_r[i] = lwp_r[i]; p.lw_i[i] = lwp_i[i]; }
            kind = kind_lo = f32::KIND_SINGLE_PRE;
        }
        e = f32_launch_n13(p, dit, mode, kind, kind_lo, (int)grid, stream);
    } else if (pd.kp.c == 0) {
        if (grid > p.n_tiles) grid = p.n_tiles;
        if (grid < 1) grid = 1;
        e = dit ? f32_launch_contig_dit(p, pd.kp.g, mode, kind, (int)grid, stream)
                : f32_launch_contig_dif(p, pd.kp.g, mode, kind, (int)grid, stream);
    } else {
        const int G = pd.kp.g, C = 12 - G, mid_bits = p.n - G - C;
        const long long mids = 1ll << mid_bits;
        p.n_units = mids * p.batch;
        if (grid > p.n_units) grid = p.n_units;
        if (tay) {
            p.tay = *tay;
            p.tay.e = kind == f32::KIND_SINGLE_PRE ? 31 - pd.kp.cm.sh_single : 0;    // what the table in p.tw carries
        }
        e = f32_launch_strided(p, G, dit, mode, kind, (int)grid, stream);
    }
    count_launch();
    return (int)e;
}

}  // namespace intfft

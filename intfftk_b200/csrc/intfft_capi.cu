// C-ABI of libintfft_b200 (include/intfft.h): plan = elaborated int_fftNk / int_ifftNk entity,
// exec = a batch of frames clocked through it.  Host logic only; kernels live in intfft_tile.cu,
// intfft_fast16.cu and intfft_util.cu.  There is no CPU compute path in this file.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "intfft_internal.h"

using namespace intfft;

struct intfft_plan : public Plan {};

namespace {

int scalar_bytes(int width) { return width <= 16 ? 2 : (width <= 32 ? 4 : 8); }

// "does it elaborate" — see the citations on intfft_validate in include/intfft.h
int validate(const intfft_generics *g)
{
    if (!g) return INTFFT_EINVAL;
    if (g->nfft_log2 < 3 || g->nfft_log2 > 20) return INTFFT_EINVAL;
    if ((g->format | 1) != 1 || (g->rndmode | 1) != 1 || (g->xser | 1) != 1 || (g->use_fly | 1) != 1 ||
        (g->direction | 1) != 1)
        return INTFFT_EINVAL;
    if (g->twdl_width < 8 || g->twdl_width > (g->xser ? 27 : 25)) return INTFFT_EINVAL;
    if (g->data_width < 8) return INTFFT_EINVAL;
    // int_dif2_fly.vhd:331-338: RNDMODE = 1 with SCALE = 0 generates two drivers for wz_re / wz_im
    if (g->direction == 0 && g->format == 1 && g->rndmode == 1) return INTFFT_EINVAL;
    const CmultConsts cm = cmult_consts(g->twdl_width, g->xser);
    const int n = g->nfft_log2;
    for (int ii = 0; ii < n; ++ii) {
        const int s = g->direction ? ii : n - 1 - ii;
        const int dtw = g->data_width + ii * g->format;
        const int dtwc = g->direction ? dtw : dtw + g->format;
        if (s > 1 && dtwc >= cm.lim_none) return INTFFT_EINVAL;   // no multiplier is generated
        // trpl18: the product slice dspP_M1(MAW+MBW-2 downto MBW-1) must lie inside the 79 / 77-bit product
        // (int_cmult_trpl18_dsp48.vhd:151-152); beyond that the entity does not elaborate
        if (s > 1 && dtwc >= cm.lim_dbl && dtwc + g->twdl_width - 2 > cm.trpl_pwd - 1) return INTFFT_EINVAL;
        if (dtw + 1 > 96) return INTFFT_EINVAL;                  // int_addsub_dsp48.vhd:16-22
    }
    const int worst = g->data_width + g->format * n + ((!g->format && g->rndmode) ? 1 : 0);
    if (worst > 64) return INTFFT_EUNSUPPORTED;
    return INTFFT_OK;
}

// split `g` stage bits starting at local bit `c` into rounds of <= 4, short round at the bottom;
// DIF walks the bits downwards, DIT upwards
void make_rounds(PassParams &kp, bool dit)
{
    const int full = kp.g / 4, rem = kp.g % 4;
    int lo[8], cnt[8], nr = 0;
    int bit = kp.c;
    if (rem) { lo[nr] = bit; cnt[nr] = rem; ++nr; bit += rem; }
    for (int i = 0; i < full; ++i) { lo[nr] = bit; cnt[nr] = 4; ++nr; bit += 4; }
    kp.nrounds = nr;
    for (int i = 0; i < nr; ++i) {
        const int src = dit ? i : nr - 1 - i;
        kp.r_lo[i] = (signed char)lo[src];
        kp.r_n[i] = (signed char)cnt[src];
    }
}

int pick_lane(int max_width, int max_dtwc, int tw)
{
    if (max_width <= 32) return LANE_I32_P64;
    return (max_dtwc + tw <= 64) ? LANE_I64_P64 : LANE_I64_P128;
}

void build_passes(Plan &pl)
{
    const intfft_generics &g = pl.g;
    const int n = g.nfft_log2;
    const bool dit = g.direction != 0;
    const int rnd_extra = (pl.mode == MODE_ROUND) ? 1 : 0;
    const CmultConsts cm = cmult_consts(g.twdl_width, g.xser);

    struct Span { int lo_bit, bits; bool strided; };
    std::vector<Span> spans;
    const bool no_fast = std::getenv("INTFFT_DISABLE_FAST16") != nullptr;   // tests: force the generic kernel
    const bool f16 = !no_fast && g.use_fly && fast16_supported(g);
    const bool f32 = !no_fast && !f16 && fast32_supported(g);
    // wide plans (some stage beyond 32 bits): STAGE 7..0 on the 64-bit-lane warp-centric kernel, the
    // n - 8 stage bits above them as one strided pass (c3: 8 strided bits on 32-bit lanes + 8 here)
    bool wide8 = false;
    if (!no_fast && g.use_fly && !f16 && !f32 && (n == 8 || (n >= 12 && n <= 16))) {
        // the plan as a whole needs 64-bit lanes with 64-bit products; each of its two passes then takes the
        // narrowest kernel family its own widths allow (32-bit lanes where they still fit)
        const int w_final = g.data_width + n * g.format;
        const int max_dtwc = dit ? (w_final - g.format) : w_final;
        wide8 = pick_lane(w_final + rnd_extra, max_dtwc, g.twdl_width) == LANE_I64_P64;
    }
    if (wide8) {
        if (n == 8) spans.push_back({0, 8, false});
        else if (!dit) { spans.push_back({8, n - 8, true}); spans.push_back({0, 8, false}); }
        else           { spans.push_back({0, 8, false}); spans.push_back({8, n - 8, true}); }
    } else if ((f32 || f16) && n == 13 && !std::getenv("INTFFT_N13_TWO_PASS")) {
        spans.push_back({0, 13, false});              // one-pass 8192-point kernels (intfft_fast32_n13.cu, fast16_n13_kernel)
    } else if (f16 && n == 14 && !std::getenv("INTFFT_N14_TWO_PASS")) {
        spans.push_back({0, 14, false});              // one-pass 16384-point packed-16 kernel (fast16_n14_kernel)
    } else if ((f16 || f32) && n >= 13) {
        // packed-16 kernels: top 4 or 8 bits as a strided pass, the rest (9..12 bits) contiguous
        const int g_hi = n <= 16 ? 4 : 8, g_lo = n - g_hi;
        if (!dit) { spans.push_back({g_lo, g_hi, true}); spans.push_back({0, g_lo, false}); }
        else      { spans.push_back({0, g_lo, false}); spans.push_back({g_lo, g_hi, true}); }
    } else if (n <= 13) {
        spans.push_back({0, n, false});
    } else {
        const int g_hi = (n - 12) < 4 ? 4 : (n - 12);
        const int g_lo = n - g_hi;
        if (!dit) { spans.push_back({g_lo, g_hi, true}); spans.push_back({0, g_lo, false}); }
        else      { spans.push_back({0, g_lo, false}); spans.push_back({g_lo, g_hi, true}); }
    }

    int stages_done = 0;
    for (size_t i = 0; i < spans.size(); ++i) {
        const Span &sp = spans[i];
        PassDesc pd{};
        PassParams &kp = pd.kp;
        kp.n = n;
        kp.L = (!wide8 && !sp.strided && (n == 13 || n == 14) && sp.bits == n) ? n : 12;
        kp.g = sp.bits;
        kp.pb = sp.lo_bit;
        kp.c = sp.strided ? kp.L - sp.bits : 0;
        kp.dw = g.data_width;
        kp.format = g.format;
        kp.cm = cm;
        kp.total = pl.batch << n;
        kp.n_tiles = sp.strided ? (pl.batch << (n - kp.L)) : ((kp.total + (1ll << kp.L) - 1) >> kp.L);
        make_rounds(kp, dit);
        const int w_in = g.data_width + stages_done * g.format;
        const int w_out = g.data_width + (stages_done + sp.bits) * g.format;
        kp.in_sb = scalar_bytes(w_in);
        kp.out_sb = scalar_bytes(w_out);
        kp.in_wrap = (i == 0) ? 1 : 0;
        // widest multiplier operand inside this pass: DIF multiplies the stage output, DIT its input
        const int max_dtwc = dit ? (w_out - g.format) : w_out;
        pd.lane = pick_lane(w_out + rnd_extra, max_dtwc, g.twdl_width);
        pd.threads = 1 << (kp.L - 4);
        pd.smem_bytes = ((size_t)1 << kp.L) * (pd.lane == LANE_I32_P64 ? 8 : 16);
        pd.scratch_in = pd.scratch_out = -1;
        // a pass of a wide plan can still run on the 32-bit-lane kernels when its own widths fit them
        // (c3: the strided top pass works on 24..28 bits, only the contiguous pass needs 64-bit lanes)
        const bool geom32 = sp.strided ? (sp.bits == 4 || sp.bits == 8) : (sp.bits >= 8 && sp.bits <= 12);
        const bool span32 = !no_fast && g.use_fly && geom32 && kp.L == 12 && (w_out + rnd_extra) <= 32 &&
                            kp.in_sb <= 4 && kp.out_sb <= 4;
        pd.path = f16 ? 1 : ((f32 || span32) ? 2 : 0);
        if (wide8 && !sp.strided && pd.path == 0) pd.path = 3;
        // the strided pass of a wide plan whose widths no longer fit 32-bit lanes: 64-bit-lane strided kernel
        if (wide8 && sp.strided && pd.path == 0 && geom32 && kp.L == 12 && pd.lane == LANE_I64_P64) pd.path = 4;
        stages_done += sp.bits;
        pl.passes.push_back(pd);
    }
    // wire intermediates: a pass writes straight into the user's output buffer when its container
    // already has the final size (later passes then run in place), otherwise into plan scratch
    for (size_t i = 0; i + 1 < pl.passes.size(); ++i) {
        PassDesc &a = pl.passes[i], &b = pl.passes[i + 1];
        if (a.kp.out_sb != pl.out_sb) {
            a.scratch_out = 0;
            b.scratch_in = 0;
            const size_t need = (size_t)a.kp.total * 2 * a.kp.out_sb;
            if (need > pl.scratch_bytes[0]) pl.scratch_bytes[0] = need;
        }
    }
}

// N > 64K on the strided kernels: STAGE >= 11 twiddles of the strided pass are recomputed on the device (coarse ROM +
// Taylor MACs, intfft_taylor.cuh) where the kernel hoists them, so the tables stop at STAGE 11 (2^12 entries instead
// of 2^NFFT: 32 KB instead of 8 MB per table at NFFT = 20) and plan creation skips the host-side Taylor sweep.
// INTFFT_TAYLOR_MIN_NFFT moves the threshold (13 = every two-pass plan on those kernels, 99 = tables only).
static bool wants_device_taylor(const Plan &pl)
{
    int min_nfft = 17;
    if (const char *e = std::getenv("INTFFT_TAYLOR_MIN_NFFT")) min_nfft = std::atoi(e);
    if (pl.g.nfft_log2 < min_nfft || pl.g.nfft_log2 < 13 || pl.passes.size() != 2) return false;
    for (const PassDesc &pd : pl.passes) {
        if (pd.path != 1 && pd.path != 2) return false;                   // packed-16 / 32-bit-lane kernels only
        if (pd.kp.c == 0 && pd.kp.g > 12) return false;                   // contiguous pass reads STAGE < 12 tables
    }
    return true;
}

int upload_twiddles(Plan &pl)
{
    const int n = pl.g.nfft_log2;
    const bool tay = wants_device_taylor(pl);
    const size_t cnt = (size_t)1 << (tay ? 12 : n);
    if (tay) {
        std::vector<int32_t> rc(512), rs(512);
        taylor_consts(pl.g.twdl_width, pl.g.xser, rc.data(), rs.data(), pl.tay.mathpi, &pl.tay.xs);
        std::vector<int2> rom(512);
        for (int i = 0; i < 512; ++i) rom[i] = make_int2(rc[i], rs[i]);
        if (cudaMalloc(&pl.d_rom9, 512 * sizeof(int2)) != cudaSuccess) return INTFFT_ENOMEM;
        if (cudaMemcpy(pl.d_rom9, rom.data(), 512 * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess) return INTFFT_ECUDA;
        pl.tay.rom9 = pl.d_rom9;
        pl.tay.tw = pl.g.twdl_width;
        pl.tay.e = 0;
        pl.tay.on = 1;
    }
    std::vector<int2> tab(cnt, make_int2(0, 0));
    std::vector<int32_t> re(cnt / 2), im(cnt / 2);
    for (int s = 2; s < n && ((size_t)2 << s) <= cnt; ++s) {
        twiddle_stage_table(s, pl.g.twdl_width, pl.g.xser, re.data(), im.data());
        for (size_t k = 0; k < ((size_t)1 << s); ++k) tab[((size_t)1 << s) + k] = make_int2(re[k], im[k]);
    }
    if (cudaMalloc(&pl.d_tw, cnt * sizeof(int2)) != cudaSuccess) return INTFFT_ENOMEM;
    if (cudaMemcpy(pl.d_tw, tab.data(), cnt * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess) return INTFFT_ECUDA;
    for (int s = 2; s <= 3 && s < n; ++s)
        for (int k = 0; k < (1 << s); ++k) {
            pl.lw32_r[(1 << s) - 1 + k] = tab[((size_t)1 << s) + k].x;
            pl.lw32_i[(1 << s) - 1 + k] = tab[((size_t)1 << s) + k].y;
        }
    if (!pl.passes.empty() && pl.passes[0].path == 2 && pl.mode == MODE_TRUNC && pl.g.twdl_width < 19) {
        // 32-bit-lane TRUNCATE kernels (KIND_SINGLE_PRE): W << e with e = 31 - sh_single = 32 - TWDL_WIDTH puts the
        // multiplier's output slice at bit 32 of the 64-bit sum of products; a twiddle is a TWDL_WIDTH-bit two's
        // complement number, so W << e always fits an int32.  (Not for TWDL_WIDTH >= 19: there the slice starts one bit
        // lower relative to the container, sh_single = TWDL_WIDTH - 2, and the Taylor refinement can push |W| to
        // 2^(TWDL_WIDTH-2) and beyond, which W << (33 - TWDL_WIDTH) would no longer hold.)
        const CmultConsts cm = cmult_consts(pl.g.twdl_width, pl.g.xser);
        const int e = 31 - cm.sh_single;
        std::vector<int2> tp(cnt);
        for (size_t i = 0; i < cnt; ++i)
            tp[i] = make_int2((int)((unsigned)tab[i].x << e), (int)((unsigned)tab[i].y << e));
        for (int s = 2; s <= 3 && s < n; ++s)
            for (int k = 0; k < (1 << s); ++k) {
                pl.lwp32_r[(1 << s) - 1 + k] = tp[(1 << s) + k].x;
                pl.lwp32_i[(1 << s) - 1 + k] = tp[(1 << s) + k].y;
            }
        if (cudaMalloc(&pl.d_twp32, cnt * sizeof(int2)) != cudaSuccess) return INTFFT_ENOMEM;
        if (cudaMemcpy(pl.d_twp32, tp.data(), cnt * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess) return INTFFT_ECUDA;
        // STAGE 12 of the one-pass 8192-point kernel, 4 bytes per twiddle.  TWDL_WIDTH = 16 only: the kernel rebuilds the
        // pre-shifted pair as (x & 0xffff0000, x << 16), which IS W << (32 - TWDL_WIDTH) only for 16-bit twiddles (narrower
        // ones fall back to the raw-twiddle instance; found by the generics fuzz test)
        if (n == 13 && !tay && pl.g.twdl_width == 16) {
            std::vector<unsigned> pk(4096);
            for (size_t k = 0; k < 4096; ++k)
                pk[k] = ((unsigned)tab[4096 + k].x << 16) | ((unsigned)tab[4096 + k].y & 0xffffu);
            if (cudaMalloc(&pl.d_tw12p, pk.size() * sizeof(unsigned)) != cudaSuccess) return INTFFT_ENOMEM;
            if (cudaMemcpy(pl.d_tw12p, pk.data(), pk.size() * sizeof(unsigned), cudaMemcpyHostToDevice) != cudaSuccess) return INTFFT_ECUDA;
        }
    }
    if (!pl.passes.empty() && pl.passes[0].path == 1) {
        // 32-bit-product kernel: W << e with e = 33 - TWDL_WIDTH - DATA_WIDTH puts the multiplier's
        // output slice P(DTW+TWD-2 downto TWD-1) (int_cmult_dsp48.vhd:189-190) at bits 31 .. 32-DTW
        const int e = 33 - pl.g.twdl_width - pl.g.data_width;
        pl.tay16 = pl.tay;
        pl.tay16.e = e;
        std::vector<int2> tp(cnt);
        for (size_t i = 0; i < cnt; ++i) tp[i] = make_int2(tab[i].x * (1 << e), tab[i].y * (1 << e));
        for (int s = 2; s <= 3 && s < n; ++s)
            for (int k = 0; k < (1 << s); ++k) {
                pl.lw_r[(1 << s) - 1 + k] = tp[(1 << s) + k].x;
                pl.lw_i[(1 << s) - 1 + k] = tp[(1 << s) + k].y;
            }
        if (cudaMalloc(&pl.d_twp, cnt * sizeof(int2)) != cudaSuccess) return INTFFT_ENOMEM;
        if (cudaMemcpy(pl.d_twp, tp.data(), cnt * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess) return INTFFT_ECUDA;
    }
    return INTFFT_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (ok && prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int intfft_validate(const intfft_generics *g) { return validate(g); }

// frames * 2^n complex samples must stay below 2^40 (checked without shifting the caller's value)
static bool batch_ok(int64_t batch, int nfft_log2) { return batch >= 1 && batch <= ((1ll << 40) >> nfft_log2); }

// Group size of a two-pass plan (see Plan::group_frames).  INTFFT_GROUP_MB overrides the budget (0 = no groups).
static long long pick_group_frames(const Plan &pl)
{
    if (pl.passes.size() != 2) return 0;
    // Measured (profiles/r02/group_sweep_multilaunch.jsonl): as separate launches per group this LOSES on every
    // two-pass plan — c4 1.28 -> 1.77 ms at 64 MB groups, 3.7 ms at 8 MB — because a launch boundary costs 10-15 us of
    // drain / ramp / per-unit twiddle loads against 20-80 us of work per group.  Off unless asked for.
    long long mb = 0;
    if (const char *e = std::getenv("INTFFT_GROUP_MB")) mb = std::atoll(e);
    if (mb <= 0) return 0;
    const PassParams &a = pl.passes[0].kp, &b = pl.passes[1].kp;
    // bytes per frame that should survive in L2 between the passes: what the first pass writes, plus the
    // streaming traffic that competes with it (its own input)
    const long long n = 1ll << pl.g.nfft_log2;
    const long long per_frame = n * 2 * (a.out_sb + a.in_sb);
    (void)b;
    long long gf = (mb << 20) / per_frame;
    return gf < 1 ? 1 : gf;
}

int intfft_plan_create(intfft_plan **out, const intfft_generics *g, int64_t batch, int device)
{
    if (!out) return INTFFT_EINVAL;
    *out = nullptr;
    int st = validate(g);
    if (st) return st;
    if (!batch_ok(batch, g->nfft_log2)) return INTFFT_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return INTFFT_ECUDA;
    DeviceGuard guard(device);
    if (!guard.ok) return INTFFT_ECUDA;

    intfft_plan *pl = new (std::nothrow) intfft_plan();
    if (!pl) return INTFFT_ENOMEM;
    pl->g = *g;
    pl->batch = batch;
    pl->device = device;
    pl->mode = g->format ? MODE_UNSCALED : (g->rndmode ? MODE_ROUND : MODE_TRUNC);
    pl->in_width = g->data_width;
    pl->out_width = g->data_width + g->format * g->nfft_log2;
    pl->in_sb = scalar_bytes(pl->in_width);
    pl->out_sb = scalar_bytes(pl->out_width);
    cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, device);
    build_passes(*pl);
    pl->group_frames = pick_group_frames(*pl);
    if (pl->group_frames && pl->scratch_bytes[0]) {        // the intermediate only ever holds one group
        const long long gf = pl->group_frames < batch ? pl->group_frames : batch;
        pl->scratch_bytes[0] = (size_t)(gf << g->nfft_log2) * 2 * pl->passes[0].kp.out_sb;
    }
    st = upload_twiddles(*pl);
    if (st == INTFFT_OK && pl->scratch_bytes[0]) {
        if (cudaMalloc(&pl->scratch[0], pl->scratch_bytes[0]) != cudaSuccess) st = INTFFT_ENOMEM;
    }
    if (st != INTFFT_OK) { intfft_plan_destroy(pl); return st; }
    *out = pl;
    return INTFFT_OK;
}

static void pipe_destroy(HostPipe &hp)
{
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        cudaFree(hp.d_in[i]);
        cudaFree(hp.d_out[i]);
        if (hp.ev_in[i]) cudaEventDestroy((cudaEvent_t)hp.ev_in[i]);
        if (hp.ev_k[i]) cudaEventDestroy((cudaEvent_t)hp.ev_k[i]);
        if (hp.ev_out[i]) cudaEventDestroy((cudaEvent_t)hp.ev_out[i]);
    }
    if (hp.s_in) cudaStreamDestroy((cudaStream_t)hp.s_in);
    if (hp.s_k) cudaStreamDestroy((cudaStream_t)hp.s_k);
    if (hp.s_out) cudaStreamDestroy((cudaStream_t)hp.s_out);
}

int intfft_plan_destroy(intfft_plan *p)
{
    if (!p) return INTFFT_EINVAL;
    DeviceGuard guard(p->device);
    cudaFree(p->d_tw);
    cudaFree(p->d_twp);
    cudaFree(p->d_twp32);
    cudaFree(p->d_tw12p);
    cudaFree(p->d_rom9);
    cudaFree(p->scratch[0]);
    cudaFree(p->scratch[1]);
    cudaFree(p->nat);
    pipe_destroy(p->pipe);
    delete p;
    return INTFFT_OK;
}

int intfft_query(const intfft_plan *p, intfft_layout *l)
{
    if (!p || !l) return INTFFT_EINVAL;
    l->n = 1ll << p->g.nfft_log2;
    l->batch = p->batch;
    l->in_width = p->in_width;
    l->out_width = p->out_width;
    l->in_scalar_bytes = p->in_sb;
    l->out_scalar_bytes = p->out_sb;
    l->in_bytes = p->batch * l->n * 2 * p->in_sb;
    l->out_bytes = p->batch * l->n * 2 * p->out_sb;
    l->n_passes = p->g.use_fly ? (int32_t)p->passes.size() : 1;
    int lane = 32;
    for (const PassDesc &pd : p->passes) if (pd.lane != LANE_I32_P64) lane = 64;
    l->lane_bits = lane;
    return INTFFT_OK;
}

// One pass over `frames` frames.  The plan's descriptor is COPIED and the per-call fields are set on the copy:
// nothing in the plan changes at exec time, so one plan can be driven from several host threads / streams.
static int run_pass(const intfft_plan *p, size_t i, const void *in, void *out, long long frames, int natural,
                    void *cuda_stream)
{
    const int n = p->g.nfft_log2;
    const bool dit = p->g.direction != 0;
    PassDesc pd = p->passes[i];
    pd.kp.in = in;
    pd.kp.out = out;
    pd.kp.tw = p->d_tw;
    pd.kp.total = frames << n;
    pd.kp.n_tiles = pd.kp.c > 0 ? (frames << (n - pd.kp.L)) : ((pd.kp.total + (1ll << pd.kp.L) - 1) >> pd.kp.L);
    pd.natural = natural;
    int e;
    if (pd.path == 1)
        e = pd.kp.c > 0 ? launch_fast16_strided(pd, p->mode, dit, p->d_twp, p->num_sms, cuda_stream, &p->tay16)
            : (pd.kp.g == 14 ? launch_fast16_n14(pd, p->mode, dit, p->d_twp, p->lw_r, p->lw_i, p->num_sms, cuda_stream)
               : pd.kp.g == 13 ? launch_fast16_n13(pd, p->mode, dit, p->d_twp, p->lw_r, p->lw_i, p->num_sms, cuda_stream)
                             : launch_fast16(pd, p->mode, dit, p->d_twp, p->lw_r, p->lw_i, p->num_sms, cuda_stream));
    else if (pd.path == 2)
        e = launch_fast32(pd, p->mode, dit, p->d_tw, p->lw32_r, p->lw32_i, p->num_sms, cuda_stream, p->d_twp32,
                          p->lwp32_r, p->lwp32_i, p->d_tw12p, &p->tay);
    else if (pd.path == 3)
        e = launch_fast64(pd, p->mode, dit, p->d_tw, p->lw32_r, p->lw32_i, p->num_sms, cuda_stream);
    else if (pd.path == 4)
        e = launch_fast64_strided(pd, p->mode, dit, p->d_tw, p->num_sms, cuda_stream);
    else
        e = launch_tile_pass(pd, p->mode, dit, p->num_sms, cuda_stream);
    return e ? INTFFT_ECUDA : INTFFT_OK;
}

// run `frames` frames (<= plan batch) starting at d_in / d_out
static int exec_frames(const intfft_plan *p, const void *d_in, void *d_out, long long frames, void *cuda_stream,
                       int natural = 0)
{
    const int n = p->g.nfft_log2;
    if (!p->g.use_fly) {
        const int e = launch_bypass(d_in, d_out, (frames << n) * 2, p->in_sb, p->out_sb, p->g.data_width,
                                    p->g.format, cuda_stream);
        return e ? INTFFT_ECUDA : INTFFT_OK;
    }
    const size_t np = p->passes.size();
    if (np == 1) return run_pass(p, 0, d_in, d_out, frames, natural, cuda_stream);
    const long long in_frame = (1ll << n) * 2 * p->in_sb, out_frame = (1ll << n) * 2 * p->out_sb;
    const long long gf = (np == 2 && p->group_frames > 0) ? p->group_frames : frames;
    for (long long f0 = 0; f0 < frames; f0 += gf) {
        const long long nf = f0 + gf <= frames ? gf : frames - f0;
        const char *gin = (const char *)d_in + f0 * in_frame;
        char *gout = (char *)d_out + f0 * out_frame;
        for (size_t i = 0; i < np; ++i) {
            const PassDesc &pd = p->passes[i];
            // intermediates: the plan's scratch (whole batch without groups, else one group, reused) or in place
            // in the caller's output buffer
            const long long so = (np == 2 && p->group_frames > 0) ? 0 : f0;
            const void *in = (i == 0) ? (const void *)gin
                                      : (pd.scratch_in >= 0 ? (const void *)((char *)p->scratch[pd.scratch_in] + so * (1ll << n) * 2 * pd.kp.in_sb)
                                                            : (const void *)gout);
            void *out = (i + 1 == np) ? (void *)gout
                                      : (pd.scratch_out >= 0 ? (void *)((char *)p->scratch[pd.scratch_out] + so * (1ll << n) * 2 * pd.kp.out_sb)
                                                             : (void *)gout);
            const int st = run_pass(p, i, in, out, nf, 0, cuda_stream);
            if (st) return st;
        }
    }
    return INTFFT_OK;
}

int intfft_exec(intfft_plan *p, const void *d_in, void *d_out, void *cuda_stream)
{
    if (!p || !d_in || !d_out) return INTFFT_EINVAL;
    if (d_in == d_out && p->in_sb != p->out_sb) return INTFFT_EINVAL;
    if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15u) return INTFFT_EINVAL;
    DeviceGuard guard(p->device);
    if (!guard.ok) return INTFFT_ECUDA;
    return exec_frames(p, d_in, d_out, p->batch, cuda_stream);
}

int intfft_exec_natural(intfft_plan *p, const void *d_in, void *d_out, void *cuda_stream)
{
    if (!p || !d_in || !d_out || d_in == d_out) return INTFFT_EINVAL;
    if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15u) return INTFFT_EINVAL;
    DeviceGuard guard(p->device);
    if (!guard.ok) return INTFFT_ECUDA;
    intfft_layout l;
    intfft_query(p, &l);
    const bool dit = p->g.direction != 0;
    // the reorder acts on the bit-reversed side: the FFT's output / the IFFT's input
    const int n = p->g.nfft_log2;
    // a 4096-point packed-16 FFT reorders inside its own kernel (natural-order stores): no second pass
    if (!dit && p->g.use_fly && p->passes.size() == 1 && p->passes[0].path == 1 && p->passes[0].kp.g == 12)
        return exec_frames(p, d_in, d_out, p->batch, cuda_stream, /*natural=*/1);
    {   // the intermediate is allocated once, under the plan's lock; afterwards it is only read here
        std::lock_guard<std::mutex> lk(p->pipe.mu);
        const size_t need = (size_t)(dit ? l.in_bytes : l.out_bytes);
        if (!p->nat) {
            if (cudaMalloc(&p->nat, need) != cudaSuccess) return INTFFT_ENOMEM;
            p->nat_bytes = need;
        }
    }
    if (!dit) {
        const int st = exec_frames(p, d_in, p->nat, p->batch, cuda_stream);
        if (st) return st;
        return launch_bitrev(n, p->out_sb, p->batch, p->nat, d_out, cuda_stream) ? INTFFT_ECUDA : INTFFT_OK;
    }
    if (launch_bitrev(n, p->in_sb, p->batch, d_in, p->nat, cuda_stream)) return INTFFT_ECUDA;
    return exec_frames(p, p->nat, d_out, p->batch, cuda_stream);
}

// ---- host-buffer pipeline -------------------------------------------------------------------------
// The batch is cut into chunks of ~32 MiB that flow through a ring of kSlots device staging buffers and three
// streams (H2D copy, kernels, D2H copy): the two PCIe directions and the SMs work at the same time, and the
// device footprint is kSlots chunks, not the batch (so a batch larger than HBM streams through as well).
}  // extern "C" (templates cannot have C linkage)
template <typename Exec>
static int host_pipeline(HostPipe &hp, long long frames, long long in_per_frame, long long out_per_frame,
                         const void *h_in, void *h_out, Exec &&exec_chunk)
{
    std::lock_guard<std::mutex> lk(hp.mu);
    if (!hp.s_in) {
        long long chunk = (32ll << 20) / in_per_frame;
        if (const char *e = std::getenv("INTFFT_CHUNK_MB")) chunk = ((long long)std::atoll(e) << 20) / in_per_frame;
        if (chunk < 1) chunk = 1;
        if (chunk > frames) chunk = frames;
        hp.chunk_frames = chunk;
        cudaStream_t a = nullptr, b = nullptr, c = nullptr;
        if (cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking) != cudaSuccess)
            return INTFFT_ECUDA;
        hp.s_in = a; hp.s_k = b; hp.s_out = c;
        for (int i = 0; i < HostPipe::kSlots; ++i) {
            cudaEvent_t e1, e2, e3;
            if (cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&e3, cudaEventDisableTiming) != cudaSuccess)
                return INTFFT_ECUDA;
            hp.ev_in[i] = e1; hp.ev_k[i] = e2; hp.ev_out[i] = e3;
            if (cudaMalloc(&hp.d_in[i], (size_t)(chunk * in_per_frame)) != cudaSuccess ||
                cudaMalloc(&hp.d_out[i], (size_t)(chunk * out_per_frame)) != cudaSuccess)
                return INTFFT_ENOMEM;
        }
    }
    cudaStream_t s_in = (cudaStream_t)hp.s_in, s_k = (cudaStream_t)hp.s_k, s_out = (cudaStream_t)hp.s_out;
    const long long chunk = hp.chunk_frames;
    int rc = INTFFT_OK;
    long long c = 0;
    for (long long f0 = 0; f0 < frames && rc == INTFFT_OK; f0 += chunk, ++c) {
        const long long nf = f0 + chunk <= frames ? chunk : frames - f0;
        const int slot = (int)(c % HostPipe::kSlots);
        // the slot is free once the result of the chunk that used it last has left for the host
        if (c >= HostPipe::kSlots && cudaStreamWaitEvent(s_in, (cudaEvent_t)hp.ev_out[slot], 0) != cudaSuccess) { rc = INTFFT_ECUDA; break; }
        if (cudaMemcpyAsync(hp.d_in[slot], (const char *)h_in + f0 * in_per_frame, (size_t)(nf * in_per_frame),
                            cudaMemcpyHostToDevice, s_in) != cudaSuccess ||
            cudaEventRecord((cudaEvent_t)hp.ev_in[slot], s_in) != cudaSuccess ||
            cudaStreamWaitEvent(s_k, (cudaEvent_t)hp.ev_in[slot], 0) != cudaSuccess) { rc = INTFFT_ECUDA; break; }
        rc = exec_chunk(hp.d_in[slot], hp.d_out[slot], nf, (void *)s_k);
        if (rc) break;
        if (cudaEventRecord((cudaEvent_t)hp.ev_k[slot], s_k) != cudaSuccess ||
            cudaStreamWaitEvent(s_out, (cudaEvent_t)hp.ev_k[slot], 0) != cudaSuccess ||
            cudaMemcpyAsync((char *)h_out + f0 * out_per_frame, hp.d_out[slot], (size_t)(nf * out_per_frame),
                            cudaMemcpyDeviceToHost, s_out) != cudaSuccess ||
            cudaEventRecord((cudaEvent_t)hp.ev_out[slot], s_out) != cudaSuccess) { rc = INTFFT_ECUDA; break; }
    }
    // success or not, no copy may still be touching the caller's buffers when this returns
    const cudaError_t e1 = cudaStreamSynchronize(s_in), e2 = cudaStreamSynchronize(s_k), e3 = cudaStreamSynchronize(s_out);
    if (rc == INTFFT_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) rc = INTFFT_ECUDA;
    if (rc == INTFFT_OK && cudaGetLastError() != cudaSuccess) rc = INTFFT_ECUDA;
    return rc;
}
extern "C" {

int intfft_exec_host(intfft_plan *p, const void *h_in, void *h_out)
{
    if (!p || !h_in || !h_out) return INTFFT_EINVAL;
    DeviceGuard guard(p->device);
    if (!guard.ok) return INTFFT_ECUDA;
    const long long n = 1ll << p->g.nfft_log2;
    return host_pipeline(p->pipe, p->batch, n * 2 * p->in_sb, n * 2 * p->out_sb, h_in, h_out,
                         [p](const void *di, void *dout, long long nf, void *st) { return exec_frames(p, di, dout, nf, st); });
}

int intfft_host_alloc(void **h_ptr, size_t bytes)
{
    if (!h_ptr || bytes == 0) return INTFFT_EINVAL;
    *h_ptr = nullptr;
    // portable: page-locked for every device of the process (intfft_multi_* copies from several devices at once)
    const cudaError_t e = cudaHostAlloc(h_ptr, bytes, cudaHostAllocPortable);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return INTFFT_ENOMEM; }
    return e == cudaSuccess ? INTFFT_OK : INTFFT_ECUDA;
}

int intfft_host_free(void *h_ptr)
{
    if (!h_ptr) return INTFFT_EINVAL;
    return cudaFreeHost(h_ptr) == cudaSuccess ? INTFFT_OK : INTFFT_ECUDA;
}

// ---- f2: int_fft_ifft_pair ----------------------------------------------------------------------
struct intfft_pair {
    intfft_plan *fwd = nullptr, *inv = nullptr;
    void *mid = nullptr;           // spectrum between the two cores (bit-reversed order), device
    long long mid_frames = 0;      // frames `mid` holds: the pair runs group by group so the spectrum stays in L2
    bool fused = false;            // both cores in one kernel (packed-16 plans of 2^8 .. 2^12 points): no `mid` at all
    HostPipe pipe;
};

int intfft_pair_create(intfft_pair **out, const intfft_generics *g, int fly_inv, int64_t batch, int device)
{
    if (!out || !g || (fly_inv | 1) != 1) return INTFFT_EINVAL;
    *out = nullptr;
    intfft_generics gf = *g, gi = *g;
    gf.direction = 0;
    gi.direction = 1;
    gi.use_fly = fly_inv;
    gi.data_width = g->data_width + g->format * g->nfft_log2;      // int_fft_ifft_pair.vhd:261
    int st = validate(&gf);
    if (!st) st = validate(&gi);
    if (st) return st;
    intfft_pair *p = new (std::nothrow) intfft_pair();
    if (!p) return INTFFT_ENOMEM;
    st = intfft_plan_create(&p->fwd, &gf, batch, device);
    if (!st) st = intfft_plan_create(&p->inv, &gi, batch, device);
    if (!st && gf.use_fly && gi.use_fly && fast16_pair_supported(gf) && fast16_supported(gi) && gi.data_width == gf.data_width &&
        p->fwd->passes.size() == 1 && p->fwd->passes[0].path == 1 && !std::getenv("INTFFT_PAIR_UNFUSED")) {
        p->fused = true;           // the spectrum stays in registers between the two cores (intfft_fast16.cu, PAIR)
        p->mid_frames = batch;
    } else if (!st) {
        DeviceGuard guard(device);
        // the spectrum of one group of frames: small enough to be read back from L2 by the inverse core
        const long long frame_bytes = (1ll << g->nfft_log2) * 2 * p->fwd->out_sb;
        long long mb = 0;          // whole batch: per-group launches measured slower (see pick_group_frames)
        if (const char *e = std::getenv("INTFFT_PAIR_GROUP_MB")) mb = std::atoll(e);
        long long gfm = mb > 0 ? (mb << 20) / frame_bytes : batch;
        if (gfm < 1) gfm = 1;
        if (gfm > batch) gfm = batch;
        p->mid_frames = gfm;
        if (cudaMalloc(&p->mid, (size_t)(gfm * frame_bytes)) != cudaSuccess) st = INTFFT_ENOMEM;
    }
    if (st) { intfft_pair_destroy(p); return st; }
    *out = p;
    return INTFFT_OK;
}

int intfft_pair_destroy(intfft_pair *p)
{
    if (!p) return INTFFT_EINVAL;
    if (p->fwd) {
        DeviceGuard guard(p->fwd->device);
        cudaFree(p->mid);
        pipe_destroy(p->pipe);
    }
    if (p->fwd) intfft_plan_destroy(p->fwd);
    if (p->inv) intfft_plan_destroy(p->inv);
    delete p;
    return INTFFT_OK;
}

int intfft_pair_query(const intfft_pair *p, intfft_layout *l)
{
    if (!p || !l) return INTFFT_EINVAL;
    intfft_layout a, b;
    intfft_query(p->fwd, &a);
    intfft_query(p->inv, &b);
    *l = a;
    l->out_width = b.out_width;
    l->out_scalar_bytes = b.out_scalar_bytes;
    l->out_bytes = b.out_bytes;
    l->n_passes = p->fused ? 1 : a.n_passes + b.n_passes;
    l->lane_bits = a.lane_bits > b.lane_bits ? a.lane_bits : b.lane_bits;
    return INTFFT_OK;
}

static int pair_frames(const intfft_pair *p, const void *d_in, void *d_out, long long frames, void *cuda_stream)
{
    if (p->fused) {
        const intfft_plan *f = p->fwd;
        PassDesc pd = f->passes[0];
        pd.kp.in = d_in;
        pd.kp.out = d_out;
        pd.kp.total = frames << f->g.nfft_log2;
        return launch_fast16_pair(pd, f->mode, f->d_twp, f->lw_r, f->lw_i, f->num_sms, cuda_stream) ? INTFFT_ECUDA : INTFFT_OK;
    }
    const long long n = 1ll << p->fwd->g.nfft_log2;
    const long long in_frame = n * 2 * p->fwd->in_sb, out_frame = n * 2 * p->inv->out_sb;
    for (long long f0 = 0; f0 < frames; f0 += p->mid_frames) {
        const long long nf = f0 + p->mid_frames <= frames ? p->mid_frames : frames - f0;
        int st = exec_frames(p->fwd, (const char *)d_in + f0 * in_frame, p->mid, nf, cuda_stream);
        if (!st) st = exec_frames(p->inv, p->mid, (char *)d_out + f0 * out_frame, nf, cuda_stream);
        if (st) return st;
    }
    return INTFFT_OK;
}

int intfft_pair_exec(intfft_pair *p, const void *d_in, void *d_out, void *cuda_stream)
{
    if (!p || !d_in || !d_out) return INTFFT_EINVAL;
    if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15u) return INTFFT_EINVAL;
    DeviceGuard guard(p->fwd->device);
    if (!guard.ok) return INTFFT_ECUDA;
    return pair_frames(p, d_in, d_out, p->fwd->batch, cuda_stream);
}

int intfft_pair_exec_host(intfft_pair *p, const void *h_in, void *h_out)
{
    if (!p || !h_in || !h_out) return INTFFT_EINVAL;
    DeviceGuard guard(p->fwd->device);
    if (!guard.ok) return INTFFT_ECUDA;
    const long long n = 1ll << p->fwd->g.nfft_log2;
    return host_pipeline(p->pipe, p->fwd->batch, n * 2 * p->fwd->in_sb, n * 2 * p->inv->out_sb, h_in, h_out,
                         [p](const void *di, void *dout, long long nf, void *st) { return pair_frames(p, di, dout, nf, st); });
}

// ---- multi-GPU: one process, several devices ---------------------------------------
// Frames are independent, so the batch is cut into contiguous shards, one per device, and every device runs
// the same plan on its shard: no exchange step, no collective.  The host path drives one pipeline per device
// from its own host thread.
struct intfft_multi {
    std::vector<intfft_plan *> plans;
    std::vector<long long> first;      // first frame of each shard
    long long batch = 0;
    intfft_generics g{};
};

int intfft_multi_create(intfft_multi **out, const intfft_generics *g, int64_t batch, const int *devices, int n_devices)
{
    if (!out) return INTFFT_EINVAL;
    *out = nullptr;
    int st = validate(g);
    if (st) return st;
    if (!devices || n_devices < 1 || n_devices > 64 || !batch_ok(batch, g->nfft_log2) || batch < n_devices) return INTFFT_EINVAL;
    intfft_multi *m = new (std::nothrow) intfft_multi();
    if (!m) return INTFFT_ENOMEM;
    m->batch = batch;
    m->g = *g;
    for (int i = 0; i < n_devices && st == INTFFT_OK; ++i) {
        // same split as intfftk_b200/sharding.py: shards differ by at most one frame and tile the batch
        const long long lo = batch * i / n_devices, hi = batch * (i + 1) / n_devices;
        intfft_plan *pl = nullptr;
        st = intfft_plan_create(&pl, g, hi - lo, devices[i]);
        if (st == INTFFT_OK) { m->plans.push_back(pl); m->first.push_back(lo); }
    }
    if (st) { intfft_multi_destroy(m); return st; }
    *out = m;
    return INTFFT_OK;
}

int intfft_multi_destroy(intfft_multi *m)
{
    if (!m) return INTFFT_EINVAL;
    for (intfft_plan *p : m->plans) intfft_plan_destroy(p);
    delete m;
    return INTFFT_OK;
}

int intfft_multi_query(const intfft_multi *m, intfft_layout *l)
{
    if (!m || !l || m->plans.empty()) return INTFFT_EINVAL;
    intfft_query(m->plans[0], l);
    l->batch = m->batch;
    l->in_bytes = m->batch * l->n * 2 * l->in_scalar_bytes;
    l->out_bytes = m->batch * l->n * 2 * l->out_scalar_bytes;
    return INTFFT_OK;
}

int intfft_multi_shard(const intfft_multi *m, int i, int *device, int64_t *first_frame, int64_t *frames)
{
    if (!m || i < 0 || i >= (int)m->plans.size()) return INTFFT_EINVAL;
    if (device) *device = m->plans[i]->device;
    if (first_frame) *first_frame = m->first[i];
    if (frames) *frames = m->plans[i]->batch;
    return INTFFT_OK;
}

int intfft_multi_devices(const intfft_multi *m) { return m ? (int)m->plans.size() : INTFFT_EINVAL; }

int intfft_multi_exec_host(intfft_multi *m, const void *h_in, void *h_out)
{
    if (!m || !h_in || !h_out) return INTFFT_EINVAL;
    const long long n = 1ll << m->g.nfft_log2;
    const size_t nd = m->plans.size();
    std::vector<int> rc(nd, INTFFT_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < nd; ++i) {
        intfft_plan *p = m->plans[i];
        const char *hi = (const char *)h_in + m->first[i] * n * 2 * p->in_sb;
        char *ho = (char *)h_out + m->first[i] * n * 2 * p->out_sb;
        th.emplace_back([p, hi, ho, &rc, i] { rc[i] = intfft_exec_host(p, hi, ho); });
    }
    for (std::thread &t : th) t.join();
    for (int r : rc) if (r) return r;
    return INTFFT_OK;
}

int intfft_multi_exec(intfft_multi *m, const void *const *d_in, void *const *d_out, void *const *cuda_streams)
{
    if (!m || !d_in || !d_out) return INTFFT_EINVAL;
    for (size_t i = 0; i < m->plans.size(); ++i) {
        const int st = intfft_exec(m->plans[i], d_in[i], d_out[i], cuda_streams ? cuda_streams[i] : nullptr);
        if (st) return st;
    }
    return INTFFT_OK;
}

int intfft_twiddles(const intfft_generics *g, int stage, int32_t *h_re, int32_t *h_im)
{
    if (!g || !h_re || !h_im || stage < 2 || stage > 19) return INTFFT_EINVAL;
    if (g->twdl_width < 8 || g->twdl_width > ((g->xser & 1) ? 27 : 25)) return INTFFT_EINVAL;
    twiddle_stage_table(stage, g->twdl_width, g->xser & 1, h_re, h_im);
    return INTFFT_OK;
}

// Test hook for the on-device Taylor path: the kernels' own device function evaluated for every k of the stage
int intfft_twiddles_device(const intfft_generics *g, int stage, int32_t *h_re, int32_t *h_im, int device)
{
    if (!g || !h_re || !h_im || stage < 11 || stage > 19) return INTFFT_EINVAL;
    if (g->twdl_width < 8 || g->twdl_width > ((g->xser & 1) ? 27 : 25)) return INTFFT_EINVAL;
    DeviceGuard guard(device);
    if (!guard.ok) return INTFFT_ECUDA;
    std::vector<int32_t> rc(512), rs(512);
    TaylorDev t{};
    taylor_consts(g->twdl_width, g->xser & 1, rc.data(), rs.data(), t.mathpi, &t.xs);
    std::vector<int2> rom(512);
    for (int i = 0; i < 512; ++i) rom[i] = make_int2(rc[i], rs[i]);
    const size_t cnt = (size_t)1 << stage;
    int2 *d_rom = nullptr, *d_out = nullptr;
    int st = INTFFT_OK;
    if (cudaMalloc(&d_rom, 512 * sizeof(int2)) != cudaSuccess || cudaMalloc(&d_out, cnt * sizeof(int2)) != cudaSuccess) st = INTFFT_ENOMEM;
    std::vector<int2> out(cnt);
    if (st == INTFFT_OK) {
        t.rom9 = d_rom; t.tw = g->twdl_width; t.e = 0; t.on = 1;
        if (cudaMemcpy(d_rom, rom.data(), 512 * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess ||
            launch_taylor_table(t, stage, d_out, nullptr) != 0 ||
            cudaMemcpy(out.data(), d_out, cnt * sizeof(int2), cudaMemcpyDeviceToHost) != cudaSuccess)
            st = INTFFT_ECUDA;
    }
    cudaFree(d_rom);
    cudaFree(d_out);
    if (st != INTFFT_OK) return st;
    for (size_t k = 0; k < cnt; ++k) { h_re[k] = out[k].x; h_im[k] = out[k].y; }
    return INTFFT_OK;
}

int intfft_bitrev(int nfft_log2, int scalar_bytes_, int64_t batch, const void *d_in, void *d_out,
                  int device, void *cuda_stream)
{
    if (nfft_log2 < 1 || nfft_log2 > 30 || batch < 1 || !d_in || !d_out || d_in == d_out) return INTFFT_EINVAL;
    if (scalar_bytes_ != 2 && scalar_bytes_ != 4 && scalar_bytes_ != 8) return INTFFT_EINVAL;
    DeviceGuard guard(device);
    if (!guard.ok) return INTFFT_ECUDA;
    return launch_bitrev(nfft_log2, scalar_bytes_, batch, d_in, d_out, cuda_stream) ? INTFFT_ECUDA : INTFFT_OK;
}

int intfft_fill_random(void *d_buf, int64_t n_scalars, int sb, int width, uint64_t seed, int device, void *cuda_stream)
{
    if (!d_buf || n_scalars < 0 || (sb != 2 && sb != 4 && sb != 8) || width < 1 || width > 8 * sb) return INTFFT_EINVAL;
    DeviceGuard guard(device);
    if (!guard.ok) return INTFFT_ECUDA;
    return launch_fill_random(d_buf, n_scalars, sb, width, seed, cuda_stream) ? INTFFT_ECUDA : INTFFT_OK;
}

int intfft_checksum(const void *d_buf, int64_t n_scalars, int sb, uint64_t *h_sum, int device, void *cuda_stream)
{
    if (!d_buf || !h_sum || n_scalars < 0 || (sb != 2 && sb != 4 && sb != 8)) return INTFFT_EINVAL;
    DeviceGuard guard(device);
    if (!guard.ok) return INTFFT_ECUDA;
    uint64_t *d_sum = nullptr;
    if (cudaMalloc(&d_sum, sizeof(uint64_t)) != cudaSuccess) return INTFFT_ENOMEM;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = INTFFT_OK;
    if (cudaMemsetAsync(d_sum, 0, sizeof(uint64_t), st) != cudaSuccess) rc = INTFFT_ECUDA;
    if (!rc && launch_checksum(d_buf, n_scalars, sb, d_sum, cuda_stream)) rc = INTFFT_ECUDA;
    if (!rc && cudaMemcpyAsync(h_sum, d_sum, sizeof(uint64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = INTFFT_ECUDA;
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = INTFFT_ECUDA;
    cudaFree(d_sum);
    return rc;
}

int intfft_describe(const intfft_generics *g, int64_t batch, char *buf, size_t len)
{
    if (!g || !buf || len == 0) return INTFFT_EINVAL;
    const int st = validate(g);
    if (st) return st;
    if (!batch_ok(batch, g->nfft_log2)) return INTFFT_EINVAL;
    Plan pl;                                   // host-side description only: nothing is allocated on a device
    pl.g = *g;
    pl.batch = batch;
    pl.mode = g->format ? MODE_UNSCALED : (g->rndmode ? MODE_ROUND : MODE_TRUNC);
    pl.in_width = g->data_width;
    pl.out_width = g->data_width + g->format * g->nfft_log2;
    pl.in_sb = scalar_bytes(pl.in_width);
    pl.out_sb = scalar_bytes(pl.out_width);
    std::string out;
    if (!g->use_fly) {
        out = "bypass";
    } else {
        build_passes(pl);
        const bool dit = g->direction != 0;
        for (size_t i = 0; i < pl.passes.size(); ++i) {
            const PassDesc &pd = pl.passes[i];
            const PassParams &kp = pd.kp;
            const bool strided = kp.c > 0;
            const char *fam = "tile";
            char extra[48] = "";
            if (pd.path == 1) fam = strided ? "fast16_strided" : (kp.g == 14 ? "fast16_n14" : kp.g == 13 ? "fast16_n13" : "fast16");
            else if (pd.path == 2) fam = strided ? "fast32_strided" : (kp.g == 13 ? "fast32_n13" : "fast32");
            else if (pd.path == 3) { fam = "fast64"; std::snprintf(extra, sizeof extra, ", instance %d", fast64_uniform_kind(kp, dit)); }
            else if (pd.path == 4) fam = "fast64_strided";
            else std::snprintf(extra, sizeof extra, ", lane %d", pd.lane == LANE_I32_P64 ? 32 : (pd.lane == LANE_I64_P64 ? 64 : 128));
            char item[128];
            std::snprintf(item, sizeof item, "%s%s[bits %d..%d, %d->%d B%s]", i ? " -> " : "", fam, kp.pb, kp.pb + kp.g - 1,
                          2 * kp.in_sb, 2 * kp.out_sb, extra);
            out += item;
        }
    }
    std::snprintf(buf, len, "%s", out.c_str());
    return INTFFT_OK;
}

int64_t intfft_launch_count(void) { return (int64_t)launches(); }

const char *intfft_strerror(int status)
{
    switch (status) {
    case INTFFT_OK: return "ok";
    case INTFFT_EINVAL: return "invalid argument, or generics that do not elaborate in the reference";
    case INTFFT_ECUDA: return "CUDA error (no usable device, or a launch / copy failed)";
    case INTFFT_ENOMEM: return "out of memory";
    case INTFFT_EUNSUPPORTED: return "legal in the reference but wider than this build's 64-bit lanes";
    default: return "unknown status";
    }
}

int intfft_version(void) { return 100; }

}  // extern "C"

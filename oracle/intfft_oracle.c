/*
 * intfft_oracle.c — CPU restatement of the intfftk integer FFT/IFFT arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (intfftk_b200/, include/) may call,
 * link or load this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU arm.
 *
 * PARITY UNPINNED by the reference itself: hukenovs/intfftk ships no golden vectors, no asserting
 * testbench and no runnable software model of the integer arithmetic (the VHDL needs a simulator
 * plus Xilinx unisim; math/fn_radix2.m is a floating-point structural model and needs Octave —
 * neither tool exists in the build image).  This file is therefore a restatement written from the
 * VHDL text, pinned by (a) the derived known-answer vectors in tests/golden/kat_survey.json
 * (produced by an independent throw-away model), (b) an independent second
 * restatement in oracle/intfft_oracle.py that follows the lane/delay-line formulation instead of
 * the in-place one used here, and (c) property tests (tests/test_oracle_*.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef __int128 i128;

typedef struct orc_generics {
    int32_t nfft_log2, data_width, twdl_width, format, rndmode, xser, use_fly, direction;
} orc_generics;

#define ORC_OK 0
#define ORC_EINVAL (-1)
#define ORC_EUNSUPPORTED (-4)

/* keep the low w bits of v as a two's-complement number — what a VHDL slice (w-1 downto 0) does */
static inline int64_t wrap_w(i128 v, int w)
{
    if (w >= 64) return (int64_t)v;
    uint64_t u = (uint64_t)v << (64 - w);
    return (int64_t)u >> (64 - w);
}
static inline i128 wrap48(i128 v)
{
    uint64_t u = (uint64_t)v << 16;
    return (i128)((int64_t)u >> 16);
}

/* ------------------------------------------------------------------------------------------ */
/* Complex multiplier variant selection: src/vhdl/math/cmult/int_cmult_dsp48.vhd:115-156,182-434 */
enum { CM_NONE = 0, CM_SINGLE, CM_DBL18, CM_TRPL18, CM_SINGLE25, CM_DBL35, CM_TRPL52 };

static int cmult_variant(int dtw, int twd, int xser_new)
{
    const int sngl = xser_new ? 28 : 26, dbl = xser_new ? 45 : 43, trpl = xser_new ? 79 : 77;
    const int twd_dsp = xser_new ? 28 : 26;
    if (twd < 19) {                                   /* xGEN_TWD18, :182 */
        if (dtw < sngl) return CM_SINGLE;             /* :184 */
        if (dtw < dbl) return CM_DBL18;               /* :226 */
        /* :266; and its product slice dspP_M1(MAW+MBW-2 downto MBW-1) must exist in the 79 / 77-bit product
         * (int_cmult_trpl18_dsp48.vhd:151-152, PWD :131-143): beyond that the entity does not elaborate */
        if (dtw < trpl) return (dtw + twd - 2 <= trpl - 1) ? CM_TRPL18 : CM_NONE;
        return CM_NONE;
    }
    if (twd < twd_dsp) {                              /* xGEN_TWD25, :307 */
        if (dtw < 19) return CM_SINGLE25;             /* :309 */
        if (dtw < 36) return CM_DBL35;                /* :354 */
        if (dtw < 53) return CM_TRPL52;               /* :395 */
        return CM_NONE;
    }
    return CM_NONE;
}

/*
 * DO = DI * WW, scaled and truncated the way each DSP48 arrangement does it.
 *   real part: P2 = DI_RE*WW_RE, P1 = DI_IM*WW_IM, P2 - P1   (int_cmult_dsp48.vhd:192-207, XALU "SUB")
 *   imag part: P2 = DI_RE*WW_IM, P1 = DI_IM*WW_RE, P2 + P1   (:209-224, XALU "ADD")
 * ALUMODE "0011" = Z - (X+Y) = PCIN - M1 (int_cmult18x25_dsp48.vhd:111-116).
 */
static int64_t cmult_half(int variant, i128 p2, i128 p1, int sub, int dtw, int twd, int xser_new)
{
    i128 r;
    switch (variant) {
    case CM_SINGLE:   /* P(DTW+TWD-2 downto TWD-1), int_cmult_dsp48.vhd:189-190 */
        r = (sub ? p2 - p1 : p2 + p1) >> (twd - 1);
        break;
    case CM_SINGLE25: /* P(DTW+TWD-3 downto TWD-2), int_cmult_dsp48.vhd:316-317 */
        r = (sub ? p2 - p1 : p2 + p1) >> (twd - 2);
        break;
    case CM_DBL18: {  /* int_cmult_dbl18_dsp48.vhd:129 (AWD), :174-175 (pre-shift), :163 (post) */
        const int awd = xser_new ? 44 : 42, pwd = xser_new ? 62 : 60;
        const int k = pwd - 48 - (18 - twd);
        i128 a = wrap48(p2 >> k), b = wrap48(p1 >> k);
        r = wrap48(sub ? a - b : a + b) >> (47 - awd);
        break;
    }
    case CM_DBL35: {  /* int_cmult_dbl35_dsp48.vhd:155-156 (pre-shift), :160 (post) */
        const int pwd = xser_new ? 62 : 60, bwd = xser_new ? 27 : 25;
        const int k = pwd - 48 - (bwd - twd) - 1;
        i128 a = wrap48(p2 >> k), b = wrap48(p1 >> k);
        r = wrap48(sub ? a - b : a + b) >> (47 - 35);
        break;
    }
    case CM_TRPL18: { /* dspP(MAW+MBW-2 downto MBW-1), int_cmult_trpl18_dsp48.vhd:151-152 */
        int64_t a = wrap_w(p2 >> (twd - 1), dtw), b = wrap_w(p1 >> (twd - 1), dtw);
        r = sub ? (i128)a - b : (i128)a + b;
        break;
    }
    case CM_TRPL52: { /* dspP(MAW+MBW-3 downto MBW-2), int_cmult_trpl52_dsp48.vhd:167-168 */
        int64_t a = wrap_w(p2 >> (twd - 2), dtw), b = wrap_w(p1 >> (twd - 2), dtw);
        r = sub ? (i128)a - b : (i128)a + b;
        break;
    }
    default:
        r = 0;
    }
    return wrap_w(r, dtw);
}

static void cmult(int64_t d_re, int64_t d_im, int64_t w_re, int64_t w_im,
                  int dtw, int twd, int xser_new, int64_t *o_re, int64_t *o_im)
{
    const int v = cmult_variant(dtw, twd, xser_new);
    if (v == CM_TRPL18) {
        /* dspA_M1 <= SXT(M1_AA, AWD) with AWD = 61 / 59 (int_cmult_trpl18_dsp48.vhd:129, :161-162): the 61x18 / 59x18
         * multiplier only ever sees the low AWD bits of the data operand, so a DTW above AWD is cut there.
         * (Found by the primitive-level netlist in oracle/rtl/, tests/test_oracle_rtl.py.) */
        const int awd = xser_new ? 61 : 59;
        d_re = wrap_w(d_re, awd);
        d_im = wrap_w(d_im, awd);
    }
    *o_re = cmult_half(v, (i128)d_re * w_re, (i128)d_im * w_im, 1, dtw, twd, xser_new);
    *o_im = cmult_half(v, (i128)d_re * w_im, (i128)d_im * w_re, 0, dtw, twd, xser_new);
}

/* ------------------------------------------------------------------------------------------ */
/* Twiddle ROM: src/vhdl/twiddle/rom_twiddle_int.vhd:135-159 (function rom_twiddle).
 * VHDL INTEGER(real) rounds to nearest -> llround. */
static void rom_entry(int depth, int64_t i, int awd, int64_t *c, int64_t *s)
{
    const double mg = (awd < 18) ? ldexp(1.0, awd - 1) - 1.0 : ldexp(1.0, awd - 2) - 1.0; /* :143-147 */
    const double ang = ((double)i * M_PI) / ldexp(1.0, depth + 1);                        /* :149 */
    *c = llround(mg * cos(ang));                                                          /* :151 */
    *s = llround(mg * sin(-ang));                                                         /* :152 */
}

/* W(k) streamed by rom_twiddle_int(STAGE = s), k = value of the STAGE-bit counter (:187-202). */
static void twiddle(int s, int64_t k, int awd, int xser_new, int64_t *w_re, int64_t *w_im)
{
    const int64_t q = k >> (s - 1);                       /* div = cnt(STAGE-1), :189 */
    const int64_t a = k & (((int64_t)1 << (s - 1)) - 1);  /* addr = cnt(STAGE-2 downto 0), :188 */
    int64_t c, sn, L, H;
    int64_t cnt = 0;
    if (s <= 10) {
        rom_entry(s - 1, a, awd, &c, &sn);                /* depth = STAGE-1, :122-123; :211 */
    } else {
        rom_entry(9, a >> (s - 10), awd, &c, &sn);        /* addrx = addr(STAGE-2 downto STAGE-10), :221 */
        cnt = a & (((int64_t)1 << (s - 10)) - 1);         /* count = addr(STAGE-11 downto 0), :225 */
    }
    if (q == 0) { L = c; H = sn; }                        /* ww_rom <= ram, :177-178 */
    else { L = sn; H = wrap_w(-(i128)c, awd); }           /* im <= not(re)+1, re <= im, :179-182 */
    if (s <= 10) { *w_re = L; *w_im = H; return; }        /* :207-208 */

    /* Taylor refinement: src/vhdl/twiddle/row_twiddle_tay.vhd */
    const int ii = s - 11;                                /* rom_twiddle_int.vhd:234 */
    const int xs = xser_new ? 21 : 23;                    /* find_widthA, :123-132 */
    const int del = xser_new ? 2 : 0;                     /* const_pi, :134-148 */
    const int64_t mathpi = llround(M_PI * ldexp(1.0, 13 - ii - del)); /* :146 */
    const int64_t mpi = (mathpi * cnt) & 0xFFFF;          /* conv_std_logic_vector(MATHPI*jj,16), :213 */
    const int64_t mpx = mpi >> 1;                         /* mpx <= '0' & mpi(17 downto 1), :247 */
    /* MULT_ADD: A = sin_aa (= low half = re), C = cos_cc (= high half = im << XSHIFT),
     * ALUMODE "0011" -> P = C - A*B (:304-312, :454-462); P -> cos_prod -> rom_im (:174). */
    const i128 p_im = wrap48(((i128)H << xs) - (i128)L * mpx);
    /* MULT_SUB: A = cos_aa (= im), C = sin_cc (= re << XSHIFT), ALUMODE "0000" -> P = C + A*B
     * (:374-382, :524-530); P -> sin_prod -> rom_re (:175). */
    const i128 p_re = wrap48(((i128)L << xs) + (i128)H * mpx);
    /* pr_rnd: pdt = prod(47 downto XSHIFT-1); rnd = pdt(.. downto 1) + pdt(0)  (:178-199) */
    const i128 t_im = p_im >> (xs - 1), t_re = p_re >> (xs - 1);
    *w_im = wrap_w((t_im >> 1) + (t_im & 1), awd);
    *w_re = wrap_w((t_re >> 1) + (t_re & 1), awd);
}

/* ------------------------------------------------------------------------------------------ */
/* multiply by -j / +j without a multiplier: "for positive values use not(X)+1, for negative
 * values use not(X)" — int_dif2_fly.vhd:281-304, int_dit2_fly.vhd:252-281 */
static inline int64_t negq(int64_t v, int w)
{
    return wrap_w(v >= 0 ? -(i128)v : ~(i128)v, w);
}

/* (v >> 1) + v(0): pr_rnd in int_dif2_fly.vhd:190-217 / int_dit2_fly.vhd:190-215 */
static inline int64_t rnd_half(i128 v, int w)
{
    return wrap_w((v >> 1) + (v & 1), w);
}

/* DIF butterfly int_dif2_fly(STAGE = s, DTW = dtw): src/vhdl/fft/int_dif2_fly.vhd:142-373 */
static void fly_dif(const orc_generics *g, int s, int dtw, int64_t k,
                    int64_t *a_re, int64_t *a_im, int64_t *b_re, int64_t *b_im)
{
    const int scale = g->format ? 0 : 1;
    const int ow = dtw + 1 - scale;
    int64_t ad_re, ad_im, su_re, su_im;
    if (scale && g->rndmode == 0) {        /* xTRUNC :144-164: inputs sliced (DTW-1 downto 1) */
        ad_re = (*a_re >> 1) + (*b_re >> 1); ad_im = (*a_im >> 1) + (*b_im >> 1);
        su_re = (*a_re >> 1) - (*b_re >> 1); su_im = (*a_im >> 1) - (*b_im >> 1);
    } else if (scale) {                    /* xROUND :167-219 */
        ad_re = rnd_half((i128)*a_re + *b_re, ow); ad_im = rnd_half((i128)*a_im + *b_im, ow);
        su_re = rnd_half((i128)*a_re - *b_re, ow); su_im = rnd_half((i128)*a_im - *b_im, ow);
    } else {                               /* xUNSCALED :221-241 */
        ad_re = wrap_w((i128)*a_re + *b_re, ow); ad_im = wrap_w((i128)*a_im + *b_im, ow);
        su_re = wrap_w((i128)*a_re - *b_re, ow); su_im = wrap_w((i128)*a_im - *b_im, ow);
    }
    *a_re = ad_re; *a_im = ad_im;
    if (s == 0) {                          /* xST0 :245-255 */
        *b_re = su_re; *b_im = su_im;
    } else if (s == 1) {                   /* xST1 :259-318, dt_sw toggles per valid beat */
        if ((k & 1) == 0) { *b_re = su_re; *b_im = su_im; }
        else { *b_re = su_im; *b_im = negq(su_re, ow); }
    } else {                               /* xSTn :322-373, cmult generic DTW+1-SCALE */
        int64_t w_re, w_im;
        twiddle(s, k, g->twdl_width, g->xser, &w_re, &w_im);
        cmult(su_re, su_im, w_re, w_im, ow, g->twdl_width, g->xser, b_re, b_im);
    }
}

/* DIT butterfly int_dit2_fly(STAGE = s, DTW = dtw): src/vhdl/fft/int_dit2_fly.vhd:140-325 */
static void fly_dit(const orc_generics *g, int s, int dtw, int64_t k,
                    int64_t *a_re, int64_t *a_im, int64_t *b_re, int64_t *b_im)
{
    const int scale = g->format ? 0 : 1;
    const int ow = dtw + 1 - scale;
    int64_t bw_re, bw_im;
    if (s == 0) {                          /* xST0 :221-230 */
        bw_re = *b_re; bw_im = *b_im;
    } else if (s == 1) {                   /* xST1 :234-286 */
        if ((k & 1) == 0) { bw_re = *b_re; bw_im = *b_im; }
        else { bw_im = *b_re; bw_re = negq(*b_im, dtw); }
    } else {                               /* xSTn :289-325: DI_RE<=IB_IM, DI_IM<=IB_RE, DO_RE=>bw_im, DO_IM=>bw_re */
        int64_t w_re, w_im, o_re, o_im;
        twiddle(s, k, g->twdl_width, g->xser, &w_re, &w_im);
        cmult(*b_im, *b_re, w_re, w_im, dtw, g->twdl_width, g->xser, &o_re, &o_im);
        bw_im = o_re; bw_re = o_im;
    }
    const int64_t ar = *a_re, ai = *a_im;
    if (scale && g->rndmode == 1) {        /* xROUND :164-217 */
        *a_re = rnd_half((i128)ar + bw_re, ow); *a_im = rnd_half((i128)ai + bw_im, ow);
        *b_re = rnd_half((i128)ar - bw_re, ow); *b_im = rnd_half((i128)ai - bw_im, ow);
    } else if (scale) {                    /* xUNSCALED with SCALE=1 :142-162: slices (DTW-1 downto 1) */
        *a_re = (ar >> 1) + (bw_re >> 1); *a_im = (ai >> 1) + (bw_im >> 1);
        *b_re = (ar >> 1) - (bw_re >> 1); *b_im = (ai >> 1) - (bw_im >> 1);
    } else {
        *a_re = wrap_w((i128)ar + bw_re, ow); *a_im = wrap_w((i128)ai + bw_im, ow);
        *b_re = wrap_w((i128)ar - bw_re, ow); *b_im = wrap_w((i128)ai - bw_im, ow);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* "does it elaborate": int_cmult_dsp48.vhd:182-434 (a multiplier variant must exist for every
 * stage with STAGE > 1), int_dif2_fly.vhd:331-338 (RNDMODE=1 with SCALE=0 double-drives wz_re),
 * row_twiddle_tay.vhd:156 (cnt_exp is 8 bits: NFFT <= 19; 20 is this project's extension). */
int orc_validate(const orc_generics *g)
{
    if (!g) return ORC_EINVAL;
    if (g->nfft_log2 < 3 || g->nfft_log2 > 20) return ORC_EINVAL;
    if (g->format != 0 && g->format != 1) return ORC_EINVAL;
    if (g->rndmode != 0 && g->rndmode != 1) return ORC_EINVAL;
    if (g->xser != 0 && g->xser != 1) return ORC_EINVAL;
    if (g->use_fly != 0 && g->use_fly != 1) return ORC_EINVAL;
    if (g->direction != 0 && g->direction != 1) return ORC_EINVAL;
    if (g->twdl_width < 8 || g->twdl_width > (g->xser ? 27 : 25)) return ORC_EINVAL;
    if (g->data_width < 8) return ORC_EINVAL;
    if (g->direction == 0 && g->format == 1 && g->rndmode == 1) return ORC_EINVAL;
    const int n = g->nfft_log2, scale = g->format ? 0 : 1;
    for (int ii = 0; ii < n; ii++) {
        const int s = g->direction ? ii : n - 1 - ii;
        const int dtw = g->data_width + ii * g->format;
        const int dtwc = g->direction ? dtw : dtw + 1 - scale;
        if (s > 1 && cmult_variant(dtwc, g->twdl_width, g->xser) == CM_NONE) return ORC_EINVAL;
        if (dtw + 1 > 96) return ORC_EINVAL;   /* int_addsub_dsp48.vhd:16-22: DSPW up to 96 */
    }
    const int worst = g->data_width + g->format * n + ((scale && g->rndmode) ? 1 : 0);
    if (worst > 64) return ORC_EUNSUPPORTED;
    return ORC_OK;
}

int orc_twiddle(const orc_generics *g, int stage, int64_t k, int64_t *w_re, int64_t *w_im)
{
    if (!g || stage < 2 || stage > 19 || k < 0 || k >= ((int64_t)1 << stage)) return ORC_EINVAL;
    twiddle(stage, k, g->twdl_width, g->xser, w_re, w_im);
    return ORC_OK;
}

/* Test hooks for tests/test_oracle_rtl.py: the multiplier and the butterflies on their own, so that they can
 * be compared, operand by operand, with the DSP48-primitive-level transcription in oracle/rtl/. */
int orc_cmult(int dtw, int twd, int xser_new, int64_t d_re, int64_t d_im, int64_t w_re, int64_t w_im,
              int64_t *o_re, int64_t *o_im)
{
    if (cmult_variant(dtw, twd, xser_new) == CM_NONE || dtw > 64) return ORC_EINVAL;
    cmult(d_re, d_im, w_re, w_im, dtw, twd, xser_new, o_re, o_im);
    return ORC_OK;
}

/* one butterfly int_dif2_fly / int_dit2_fly (direction 0 / 1) of STAGE s at input width dtw, beat k; ab = {a_re, a_im,
 * b_re, b_im} in and out; the twiddle is generated inside (rom_twiddle_int) */
int orc_fly(const orc_generics *g, int s, int dtw, int64_t k, int64_t *ab)
{
    if (g->direction == 0) fly_dif(g, s, dtw, k, &ab[0], &ab[1], &ab[2], &ab[3]);
    else fly_dit(g, s, dtw, k, &ab[0], &ab[1], &ab[2], &ab[3]);
    return ORC_OK;
}

int orc_twiddle_table(const orc_generics *g, int stage, int32_t *re, int32_t *im)
{
    if (!g || stage < 2 || stage > 19) return ORC_EINVAL;
    const int64_t cnt = (int64_t)1 << stage;
    for (int64_t k = 0; k < cnt; k++) {
        int64_t r, i;
        twiddle(stage, k, g->twdl_width, g->xser, &r, &i);
        re[k] = (int32_t)r; im[k] = (int32_t)i;
    }
    return ORC_OK;
}

/* per-plan twiddle cache so that batch runs do not call cos/sin per butterfly */
typedef struct { int64_t *re[20], *im[20]; } tw_cache;

static int tw_cache_build(const orc_generics *g, tw_cache *c)
{
    memset(c, 0, sizeof(*c));
    for (int s = 2; s < g->nfft_log2; s++) {
        const int64_t cnt = (int64_t)1 << s;
        c->re[s] = (int64_t *)malloc(sizeof(int64_t) * cnt);
        c->im[s] = (int64_t *)malloc(sizeof(int64_t) * cnt);
        if (!c->re[s] || !c->im[s]) return -1;
        for (int64_t k = 0; k < cnt; k++)
            twiddle(s, k, g->twdl_width, g->xser, &c->re[s][k], &c->im[s][k]);
    }
    return 0;
}
static void tw_cache_free(tw_cache *c)
{
    for (int s = 0; s < 20; s++) { free(c->re[s]); free(c->im[s]); }
}

/* Butterfly with cached twiddles (same arithmetic as fly_dif / fly_dit; the uncached versions
 * above are kept as the readable statement and used by orc_transform_uncached). */
static void fly_cached(const orc_generics *g, const tw_cache *c, int s, int dtw, int64_t k,
                       int64_t *a_re, int64_t *a_im, int64_t *b_re, int64_t *b_im)
{
    if (s < 2) {
        if (g->direction) fly_dit(g, s, dtw, k, a_re, a_im, b_re, b_im);
        else fly_dif(g, s, dtw, k, a_re, a_im, b_re, b_im);
        return;
    }
    const int scale = g->format ? 0 : 1;
    const int ow = dtw + 1 - scale;
    const int64_t w_re = c->re[s][k], w_im = c->im[s][k];
    if (!g->direction) {
        int64_t ad_re, ad_im, su_re, su_im;
        if (scale && g->rndmode == 0) {
            ad_re = (*a_re >> 1) + (*b_re >> 1); ad_im = (*a_im >> 1) + (*b_im >> 1);
            su_re = (*a_re >> 1) - (*b_re >> 1); su_im = (*a_im >> 1) - (*b_im >> 1);
        } else if (scale) {
            ad_re = rnd_half((i128)*a_re + *b_re, ow); ad_im = rnd_half((i128)*a_im + *b_im, ow);
            su_re = rnd_half((i128)*a_re - *b_re, ow); su_im = rnd_half((i128)*a_im - *b_im, ow);
        } else {
            ad_re = wrap_w((i128)*a_re + *b_re, ow); ad_im = wrap_w((i128)*a_im + *b_im, ow);
            su_re = wrap_w((i128)*a_re - *b_re, ow); su_im = wrap_w((i128)*a_im - *b_im, ow);
        }
        *a_re = ad_re; *a_im = ad_im;
        cmult(su_re, su_im, w_re, w_im, ow, g->twdl_width, g->xser, b_re, b_im);
    } else {
        int64_t o_re, o_im;
        cmult(*b_im, *b_re, w_re, w_im, dtw, g->twdl_width, g->xser, &o_re, &o_im);
        const int64_t bw_im = o_re, bw_re = o_im, ar = *a_re, ai = *a_im;
        if (scale && g->rndmode == 1) {
            *a_re = rnd_half((i128)ar + bw_re, ow); *a_im = rnd_half((i128)ai + bw_im, ow);
            *b_re = rnd_half((i128)ar - bw_re, ow); *b_im = rnd_half((i128)ai - bw_im, ow);
        } else if (scale) {
            *a_re = (ar >> 1) + (bw_re >> 1); *a_im = (ai >> 1) + (bw_im >> 1);
            *b_re = (ar >> 1) - (bw_re >> 1); *b_im = (ai >> 1) - (bw_im >> 1);
        } else {
            *a_re = wrap_w((i128)ar + bw_re, ow); *a_im = wrap_w((i128)ai + bw_im, ow);
            *b_re = wrap_w((i128)ar - bw_re, ow); *b_im = wrap_w((i128)ai - bw_im, ow);
        }
    }
}

/*
 * One frame, in place in (re[], im[]) of length N.  Stage wiring:
 *   int_fftNk.vhd:184-215  stage ii -> int_dif2_fly(STAGE = NFFT-1-ii, DTW = DATA_WIDTH+ii*FORMAT)
 *   int_ifftNk.vhd:183-214 stage ii -> int_dit2_fly(STAGE = ii,        DTW = DATA_WIDTH+ii*FORMAT)
 * Pairing: the cross-commutation of int_delay_line(STAGE=ii) swaps blocks of 2^(NFFT-ii-2)
 * between the lanes (int_delay_line.vhd:52-104,201; math/fn_radix2.m:51-69), which makes the
 * chain the in-place radix-2 DIF with `half` = N >> (ii+1); for the IFFT the delay lines are
 * instantiated with STAGE = NFFT-ii-2 (int_ifftNk.vhd:289-312) giving `half` = 1 << ii.
 * Twiddle index = beat number mod 2^STAGE (rom_twiddle_int.vhd:187-202), which in the in-place
 * picture is ia mod half.  USE_FLY = 0 bypasses every butterfly (int_fftNk.vhd:260-277); in
 * UNSCALED mode the bypassed data is zero-extended into the wider bus (ia_re(0) drives only
 * DATA_WIDTH bits of a zero-initialised signal, int_fftNk.vhd:178-182).
 */
static void transform(const orc_generics *g, const tw_cache *c, int64_t *re, int64_t *im)
{
    const int n = g->nfft_log2;
    const int64_t N = (int64_t)1 << n;
    for (int64_t i = 0; i < N; i++) {
        re[i] = wrap_w(re[i], g->data_width);
        im[i] = wrap_w(im[i], g->data_width);
    }
    if (!g->use_fly) {
        if (g->format) {
            const uint64_t m = g->data_width >= 64 ? ~0ull : ((1ull << g->data_width) - 1);
            for (int64_t i = 0; i < N; i++) { re[i] = (int64_t)((uint64_t)re[i] & m); im[i] = (int64_t)((uint64_t)im[i] & m); }
        }
        return;
    }
    for (int ii = 0; ii < n; ii++) {
        const int s = g->direction ? ii : n - 1 - ii;
        const int64_t half = (int64_t)1 << s;
        const int dtw = g->data_width + ii * g->format;
        for (int64_t p = 0; p < N / 2; p++) {
            const int64_t j = p & (half - 1);
            const int64_t ia = ((p >> s) << (s + 1)) + j, ib = ia + half;
            if (c) fly_cached(g, c, s, dtw, j, &re[ia], &im[ia], &re[ib], &im[ib]);
            else if (g->direction) fly_dit(g, s, dtw, j, &re[ia], &im[ia], &re[ib], &im[ib]);
            else fly_dif(g, s, dtw, j, &re[ia], &im[ia], &re[ib], &im[ib]);
        }
    }
}

/* One frame on int64 scalars; twiddles recomputed per butterfly (slow, most literal path). */
int orc_transform(const orc_generics *g, const int64_t *in_re, const int64_t *in_im,
                  int64_t *out_re, int64_t *out_im)
{
    int st = orc_validate(g);
    if (st) return st;
    const int64_t N = (int64_t)1 << g->nfft_log2;
    memmove(out_re, in_re, sizeof(int64_t) * N);
    memmove(out_im, in_im, sizeof(int64_t) * N);
    transform(g, NULL, out_re, out_im);
    return ORC_OK;
}

static int scalar_bytes(int width) { return width <= 16 ? 2 : (width <= 32 ? 4 : 8); }

static inline int64_t load_scalar(const void *p, int64_t idx, int sb)
{
    switch (sb) {
    case 2: return ((const int16_t *)p)[idx];
    case 4: return ((const int32_t *)p)[idx];
    default: return ((const int64_t *)p)[idx];
    }
}
static inline void store_scalar(void *p, int64_t idx, int sb, int64_t v)
{
    switch (sb) {
    case 2: ((int16_t *)p)[idx] = (int16_t)v; break;
    case 4: ((int32_t *)p)[idx] = (int32_t)v; break;
    default: ((int64_t *)p)[idx] = v; break;
    }
}

/* worker: frames [b0, b1) of a batch */
typedef struct {
    const orc_generics *g; const tw_cache *c; const void *in; void *out;
    int64_t b0, b1; int isb, osb; int failed;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const int64_t N = (int64_t)1 << j->g->nfft_log2;
    int64_t *re = (int64_t *)malloc(sizeof(int64_t) * N);
    int64_t *im = (int64_t *)malloc(sizeof(int64_t) * N);
    if (!re || !im) { j->failed = 1; free(re); free(im); return NULL; }
    for (int64_t b = j->b0; b < j->b1; b++) {
        const int64_t base = b * N * 2;
        for (int64_t i = 0; i < N; i++) {
            re[i] = load_scalar(j->in, base + 2 * i, j->isb);
            im[i] = load_scalar(j->in, base + 2 * i + 1, j->isb);
        }
        transform(j->g, j->c, re, im);
        for (int64_t i = 0; i < N; i++) {
            store_scalar(j->out, base + 2 * i, j->osb, re[i]);
            store_scalar(j->out, base + 2 * i + 1, j->osb, im[i]);
        }
    }
    free(re); free(im);
    return NULL;
}

/* `batch` frames in the flat layout of include/intfft.h (interleaved {re,im}, containers chosen
 * by width).  threads <= 0 -> one per online core.  Returns the number of threads used (>0) or a
 * negative status. */
int orc_batch(const orc_generics *g, int64_t batch, const void *in, void *out, int threads)
{
    int st = orc_validate(g);
    if (st) return st;
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((int64_t)threads > batch) threads = batch > 0 ? (int)batch : 1;
    tw_cache c;
    if (tw_cache_build(g, &c)) { tw_cache_free(&c); return -3; }
    batch_job jobs[256];
    pthread_t tid[256];
    for (int t = 0; t < threads; t++) {
        jobs[t] = (batch_job){ g, &c, in, out, batch * t / threads, batch * (t + 1) / threads,
                               scalar_bytes(g->data_width),
                               scalar_bytes(g->data_width + g->format * g->nfft_log2), 0 };
    }
    int started = 0, failed = 0;
    for (int t = 1; t < threads; t++) {
        if (pthread_create(&tid[t], NULL, batch_worker, &jobs[t])) { batch_worker(&jobs[t]); tid[t] = 0; }
        else started++;
    }
    (void)started;
    batch_worker(&jobs[0]);
    for (int t = 1; t < threads; t++) if (tid[t]) pthread_join(tid[t], NULL);
    for (int t = 0; t < threads; t++) failed |= jobs[t].failed;
    tw_cache_free(&c);
    return failed ? -3 : threads;
}

/* ------------------------------------------------------------------------------------------ */
/* Synthetic stimulus and checksum — the same counter-based functions the device uses
 * (intfftk_b200/csrc/intfft_util.cu), restated here so tests can regenerate any input on the CPU.
 * splitmix64 finaliser of (seed + index * golden gamma). */
static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void orc_fill_random(void *buf, int64_t n_scalars, int sb, int width, uint64_t seed)
{
    for (int64_t i = 0; i < n_scalars; i++) {
        const uint64_t h = mix64(seed + (uint64_t)i * 0x9E3779B97F4A7C15ull);
        store_scalar(buf, i, sb, wrap_w((i128)(int64_t)h, width));
    }
}

uint64_t orc_checksum(const void *buf, int64_t n_scalars, int sb)
{
    uint64_t sum = 0;
    for (int64_t i = 0; i < n_scalars; i++) {
        const uint64_t w = mix64((uint64_t)i) | 1ull;
        sum += (uint64_t)load_scalar(buf, i, sb) * w;
    }
    return sum;
}

/* bit-reversal reorder: buffers/int_bitrev_order.vhd:82-104 (out[bitrev(q)] = in[q]) */
void orc_bitrev(int nfft_log2, int sb, int64_t batch, const void *in, void *out)
{
    const int64_t N = (int64_t)1 << nfft_log2;
    for (int64_t b = 0; b < batch; b++)
        for (int64_t q = 0; q < N; q++) {
            int64_t r = 0;
            for (int t = 0; t < nfft_log2; t++) r |= ((q >> t) & 1) << (nfft_log2 - 1 - t);
            for (int c = 0; c < 2; c++)
                store_scalar(out, (b * N + r) * 2 + c, sb, load_scalar(in, (b * N + q) * 2 + c, sb));
        }
}

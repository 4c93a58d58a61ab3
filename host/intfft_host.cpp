// intfft_host — small C++ host that stands in for the reference's VHDL testbenches
// (src/vhdl/tb/fft_signle_test.vhd:154-358, fft_double_test.vhd:154-217): it replays a stimulus file
// through an elaborated core and dumps the output stream.
//
// Stimulus format = math/di_single.dat as written by math/fft_single.m:94-98 and read by
// tb/fft_signle_test.vhd:158-165: one "re im" pair of decimal integers per line, frames back to back.
// Output: same format, flat stream order (FFT: bit-reversed, IFFT: natural), one frame after another.
//
// --lanes switches both files to the two-lane testbench format of tb/fft_double_test.vhd:154-161,207-214
// (math/di_double.dat / dout_pair.dat): one beat per line, "lane0_re lane1_re lane0_im lane1_im"; the
// lanes are the core's own (FFT: halves in, even/odd out; IFFT: even/odd in, halves out; pair: halves
// both ways).  --top17 dumps only the 17 most significant bits of every output, as that testbench does.
// --pair runs int_fft_ifft_pair (FFT then IFFT on DATA_WIDTH + FORMAT*NFFT bits) instead of one core.
//
// --describe prints what the generics elaborate to (output width, kernel chain) and exits; it needs neither files
// nor a GPU — the counterpart of reading the synthesis log for the multiplier / buffer variants generated.
//
// usage: intfft_host [--ifft | --pair] [--nfft N] [--dw W] [--tw W] [--mode UNSCALED|ROUNDING|TRUNCATE]
//                    [--xser OLD|NEW] [--no-fly | --no-fly-fwd] [--no-fly-inv] [--lanes] [--top17] <in.dat> <out.dat>
//        intfft_host --describe [--ifft] [--nfft N] [--dw W] [--tw W] [--mode ...] [--xser ...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "intfft.h"

static int fail(const char *what, int st)
{
    std::fprintf(stderr, "intfft_host: %s: %s\n", what, intfft_strerror(st));
    return 1;
}

int main(int argc, char **argv)
{
    intfft_generics g{7, 16, 16, 1, 0, 1, 1, 0};   // the testbench defaults: NFFT=7, 16/16, XSERIES="NEW"
    std::string in_path, out_path;
    bool lanes = false, top17 = false, pair = false, describe = false;
    int fly_inv = 1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); std::exit(2); }
            return argv[++i];
        };
        if (a == "--ifft") g.direction = 1;
        else if (a == "--pair") pair = true;
        else if (a == "--describe") describe = true;
        else if (a == "--lanes") lanes = true;
        else if (a == "--top17") top17 = true;
        else if (a == "--nfft") g.nfft_log2 = std::atoi(next("--nfft"));
        else if (a == "--dw") g.data_width = std::atoi(next("--dw"));
        else if (a == "--tw") g.twdl_width = std::atoi(next("--tw"));
        else if (a == "--xser") g.xser = std::strcmp(next("--xser"), "OLD") ? 1 : 0;
        else if (a == "--no-fly" || a == "--no-fly-fwd") g.use_fly = 0;      // USE_FLY / FLY_FWD = '0'
        else if (a == "--no-fly-inv") fly_inv = 0;                            // FLY_INV = '0' (--pair only)
        else if (a == "--mode") {
            std::string m = next("--mode");                 // set_mode, tb/fft_signle_test.vhd:81-88
            if (m == "UNSCALED") { g.format = 1; g.rndmode = 0; }
            else if (m == "ROUNDING") { g.format = 0; g.rndmode = 1; }
            else if (m == "TRUNCATE") { g.format = 0; g.rndmode = 0; }
            else { std::fprintf(stderr, "bad --mode\n"); return 2; }
        } else if (in_path.empty()) in_path = a;
        else out_path = a;
    }
    if (describe) {
        char chain[512];
        const int sd = intfft_describe(&g, 1, chain, sizeof chain);
        if (sd) return fail("generics", sd);
        std::printf("%s NFFT=%d DATA_WIDTH=%d TWDL_WIDTH=%d FORMAT=%d RNDMODE=%d XSER=%s USE_FLY=%d: %d-bit out, %s\n",
                    g.direction ? "int_ifftNk" : "int_fftNk", g.nfft_log2, g.data_width, g.twdl_width, g.format, g.rndmode,
                    g.xser ? "NEW" : "OLD", g.use_fly, g.data_width + g.format * g.nfft_log2, chain);
        return 0;
    }
    if (in_path.empty() || out_path.empty()) {
        std::fprintf(stderr, "usage: intfft_host [options] <in.dat> <out.dat>   |   intfft_host --describe [options]\n");
        return 2;
    }
    int st = intfft_validate(&g);
    if (st) return fail("generics", st);

    const long long n = 1ll << g.nfft_log2;
    // lane <-> flat index maps of the cores (int_fftNk.vhd:15-21, int_ifftNk.vhd:15-21)
    auto flat_of = [&](bool halves, int lane, long long beat) { return halves ? lane * (n / 2) + beat : 2 * beat + lane; };
    const bool in_halves = pair || g.direction == 0, out_halves = pair || g.direction == 1;

    std::vector<long long> vals;
    if (FILE *f = std::fopen(in_path.c_str(), "r")) {
        long long a, b, c, d;
        if (!lanes) {
            while (std::fscanf(f, "%lld %lld", &a, &b) == 2) { vals.push_back(a); vals.push_back(b); }
        } else {
            std::vector<long long> beats;
            while (std::fscanf(f, "%lld %lld %lld %lld", &a, &b, &c, &d) == 4) { beats.insert(beats.end(), {a, b, c, d}); }
            const long long nb = (long long)beats.size() / 4, fr = nb / (n / 2);
            vals.assign((size_t)fr * n * 2, 0);
            for (long long fi = 0; fi < fr; ++fi)
                for (long long p = 0; p < n / 2; ++p) {
                    const long long *q = &beats[(size_t)(fi * (n / 2) + p) * 4];
                    const long long i0 = fi * n + flat_of(in_halves, 0, p), i1 = fi * n + flat_of(in_halves, 1, p);
                    vals[2 * i0] = q[0]; vals[2 * i1] = q[1]; vals[2 * i0 + 1] = q[2]; vals[2 * i1 + 1] = q[3];
                }
        }
        std::fclose(f);
    } else { std::perror(in_path.c_str()); return 1; }
    const long long frames = (long long)vals.size() / (2 * n);
    if (frames < 1) { std::fprintf(stderr, "need at least one frame of %lld samples\n", n); return 1; }

    intfft_plan *plan = nullptr;
    intfft_pair *pr = nullptr;
    intfft_layout lay;
    if (pair) {
        st = intfft_pair_create(&pr, &g, fly_inv, frames, 0);
        if (st) return fail("pair_create", st);
        intfft_pair_query(pr, &lay);
    } else {
        st = intfft_plan_create(&plan, &g, frames, 0);
        if (st) return fail("plan_create", st);
        intfft_query(plan, &lay);
    }
    std::vector<unsigned char> hin((size_t)lay.in_bytes), hout((size_t)lay.out_bytes);
    for (long long i = 0; i < frames * n * 2; ++i) {
        if (lay.in_scalar_bytes == 2) reinterpret_cast<int16_t *>(hin.data())[i] = (int16_t)vals[i];
        else if (lay.in_scalar_bytes == 4) reinterpret_cast<int32_t *>(hin.data())[i] = (int32_t)vals[i];
        else reinterpret_cast<int64_t *>(hin.data())[i] = vals[i];
    }
    if (pair) st = intfft_pair_exec_host(pr, hin.data(), hout.data());     // the spectrum stays on the device
    else st = intfft_exec_host(plan, hin.data(), hout.data());
    if (st) return fail("exec", st);
    auto out_scalar = [&](long long i) -> long long {
        long long v;
        if (lay.out_scalar_bytes == 2) v = reinterpret_cast<int16_t *>(hout.data())[i];
        else if (lay.out_scalar_bytes == 4) v = reinterpret_cast<int32_t *>(hout.data())[i];
        else v = reinterpret_cast<int64_t *>(hout.data())[i];
        return top17 && lay.out_width > 17 ? v >> (lay.out_width - 17) : v;   // slice (W-1 downto W-17)
    };
    FILE *o = std::fopen(out_path.c_str(), "w");
    if (!o) { std::perror(out_path.c_str()); return 1; }
    if (!lanes) {
        for (long long i = 0; i < frames * n; ++i) std::fprintf(o, "%lld %lld\n", out_scalar(2 * i), out_scalar(2 * i + 1));
    } else {
        for (long long fi = 0; fi < frames; ++fi)
            for (long long p = 0; p < n / 2; ++p) {
                const long long i0 = fi * n + flat_of(out_halves, 0, p), i1 = fi * n + flat_of(out_halves, 1, p);
                std::fprintf(o, "%lld    %lld    %lld    %lld\n", out_scalar(2 * i0), out_scalar(2 * i1),
                             out_scalar(2 * i0 + 1), out_scalar(2 * i1 + 1));
            }
    }
    std::fclose(o);
    if (plan) intfft_plan_destroy(plan);
    if (pr) intfft_pair_destroy(pr);
    std::fprintf(stderr, "intfft_host: %lld frame(s) of %lld points, %d-bit in, %d-bit out\n", frames, n,
                 lay.in_width, lay.out_width);
    return 0;
}

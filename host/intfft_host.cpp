// intfft_host — small C++ host that stands in for the reference's VHDL testbenches
// (src/vhdl/tb/fft_signle_test.vhd:154-358, fft_double_test.vhd:154-217): it replays a stimulus file
// through an elaborated core and dumps the output stream.
//
// Stimulus format = math/di_single.dat as written by math/fft_single.m:94-98 and read by
// tb/fft_signle_test.vhd:158-165: one "re im" pair of decimal integers per line, frames back to back.
// Output: same format, flat stream order (FFT: bit-reversed, IFFT: natural), one frame after another.
//
// usage: intfft_host [--ifft] [--nfft N] [--dw W] [--tw W] [--mode UNSCALED|ROUNDING|TRUNCATE]
//                    [--xser OLD|NEW] [--no-fly] <in.dat> <out.dat>
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
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); std::exit(2); }
            return argv[++i];
        };
        if (a == "--ifft") g.direction = 1;
        else if (a == "--nfft") g.nfft_log2 = std::atoi(next("--nfft"));
        else if (a == "--dw") g.data_width = std::atoi(next("--dw"));
        else if (a == "--tw") g.twdl_width = std::atoi(next("--tw"));
        else if (a == "--xser") g.xser = std::strcmp(next("--xser"), "OLD") ? 1 : 0;
        else if (a == "--no-fly") g.use_fly = 0;
        else if (a == "--mode") {
            std::string m = next("--mode");                 // set_mode, tb/fft_signle_test.vhd:81-88
            if (m == "UNSCALED") { g.format = 1; g.rndmode = 0; }
            else if (m == "ROUNDING") { g.format = 0; g.rndmode = 1; }
            else if (m == "TRUNCATE") { g.format = 0; g.rndmode = 0; }
            else { std::fprintf(stderr, "bad --mode\n"); return 2; }
        } else if (in_path.empty()) in_path = a;
        else out_path = a;
    }
    if (in_path.empty() || out_path.empty()) {
        std::fprintf(stderr, "usage: intfft_host [options] <in.dat> <out.dat>\n");
        return 2;
    }
    int st = intfft_validate(&g);
    if (st) return fail("generics", st);

    std::vector<long long> vals;
    if (FILE *f = std::fopen(in_path.c_str(), "r")) {
        long long re, im;
        while (std::fscanf(f, "%lld %lld", &re, &im) == 2) { vals.push_back(re); vals.push_back(im); }
        std::fclose(f);
    } else { std::perror(in_path.c_str()); return 1; }
    const long long n = 1ll << g.nfft_log2;
    const long long frames = (long long)vals.size() / (2 * n);
    if (frames < 1) { std::fprintf(stderr, "need at least one frame of %lld samples\n", n); return 1; }

    intfft_plan *plan = nullptr;
    st = intfft_plan_create(&plan, &g, frames, 0);
    if (st) return fail("plan_create", st);
    intfft_layout lay;
    intfft_query(plan, &lay);
    std::vector<unsigned char> hin((size_t)lay.in_bytes), hout((size_t)lay.out_bytes);
    for (long long i = 0; i < frames * n * 2; ++i) {
        if (lay.in_scalar_bytes == 2) reinterpret_cast<int16_t *>(hin.data())[i] = (int16_t)vals[i];
        else if (lay.in_scalar_bytes == 4) reinterpret_cast<int32_t *>(hin.data())[i] = (int32_t)vals[i];
        else reinterpret_cast<int64_t *>(hin.data())[i] = vals[i];
    }
    st = intfft_exec_host(plan, hin.data(), hout.data());
    if (st) return fail("exec_host", st);
    FILE *o = std::fopen(out_path.c_str(), "w");
    if (!o) { std::perror(out_path.c_str()); return 1; }
    for (long long i = 0; i < frames * n; ++i) {
        long long re, im;
        if (lay.out_scalar_bytes == 2) { re = reinterpret_cast<int16_t *>(hout.data())[2 * i]; im = reinterpret_cast<int16_t *>(hout.data())[2 * i + 1]; }
        else if (lay.out_scalar_bytes == 4) { re = reinterpret_cast<int32_t *>(hout.data())[2 * i]; im = reinterpret_cast<int32_t *>(hout.data())[2 * i + 1]; }
        else { re = reinterpret_cast<int64_t *>(hout.data())[2 * i]; im = reinterpret_cast<int64_t *>(hout.data())[2 * i + 1]; }
        std::fprintf(o, "%lld %lld\n", re, im);
    }
    std::fclose(o);
    intfft_plan_destroy(plan);
    std::fprintf(stderr, "intfft_host: %lld frame(s) of %lld points, %d-bit in, %d-bit out\n", frames, n,
                 lay.in_width, lay.out_width);
    return 0;
}

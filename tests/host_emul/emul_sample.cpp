// CPU emulation of the device sample conversion (brutefir_b200/csrc/bf_sample.cuh) against the oracle
// restatement (oracle/bf_oracle.c): raw->real for every format, and the quantiser / packer / overflow
// accounting on adversarial values, must agree bit for bit.  Built and run by tests/test_host_emulation.py.
#define BF_HOST_EMULATION 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../../brutefir_b200/csrc/bf_sample.cuh"
extern "C" {
#include "../../oracle/bf_oracle.h"
}

using namespace bf;

struct Fmt { const char *name; int isfloat, swap, bytes, sbytes; };
static const Fmt FORMATS[] = {
    {"S8", 0, 1, 1, 1}, {"S16_LE", 0, 0, 2, 2}, {"S16_BE", 0, 1, 2, 2}, {"S24_LE", 0, 0, 3, 3}, {"S24_BE", 0, 1, 3, 3},
    {"S24_4LE", 0, 0, 4, 3}, {"S24_4BE", 0, 1, 4, 3}, {"S32_LE", 0, 0, 4, 4}, {"S32_BE", 0, 1, 4, 4},
    {"FLOAT_LE", 1, 0, 4, 4}, {"FLOAT_BE", 1, 1, 4, 4}, {"FLOAT64_LE", 1, 0, 8, 8}, {"FLOAT64_BE", 1, 1, 8, 8},
};

template <typename T>
static int run(int L)
{
    int bad = 0;
    const int rs = (int)sizeof(T);
    orc_init(NULL, L, rs);
    for (const Fmt &f : FORMATS) {
        const int spacing = 3, offset = f.bytes;   // channel 1 of 3, interleaved
        std::vector<uint8_t> raw((size_t)L * spacing * f.bytes + 64);
        for (auto &b : raw) b = (uint8_t)(rand() & 0xff);
        if (f.isfloat) {    // keep float inputs finite
            for (int n = 0; n < L; n++) {
                double v = ((double)rand() / RAND_MAX - 0.5) * 4.0;
                uint8_t t[8];
                if (f.bytes == 4) { float x = (float)v; memcpy(t, &x, 4); } else { memcpy(t, &v, 8); }
                for (int i = 0; i < f.bytes; i++) raw[offset + (size_t)n * spacing * f.bytes + i] = f.swap ? t[f.bytes - 1 - i] : t[i];
            }
        }
        orc_buffer_format bf;
        memset(&bf, 0, sizeof(bf));
        bf.sf.isfloat = f.isfloat; bf.sf.swap = f.swap; bf.sf.bytes = f.bytes; bf.sf.sbytes = f.sbytes;
        bf.sample_spacing = spacing; bf.byte_offset = offset;
        std::vector<T> cbuf(2 * L), next(2 * L);
        orc_raw2cbuf(raw.data(), cbuf.data(), next.data(), &bf, NULL, NULL);
        for (int n = 0; n < L; n++) {
            T v = raw_to_real<T>(raw.data() + offset + (size_t)n * spacing * f.bytes, f.bytes, f.isfloat, f.swap);
            if (memcmp(&v, &next[n], sizeof(T)) != 0) { bad++; break; }
        }
        // output side
        const double of_max = f.isfloat ? 1.0 : (double)(((uint64_t)1 << ((f.sbytes << 3) - 1)) - 1);
        std::vector<T> y(2 * L, (T)0);
        for (int n = 0; n < L; n++) {
            double u = (double)rand() / RAND_MAX - 0.5;
            switch (n % 8) {
            case 0:                                                           // around and beyond full scale
                y[n] = (T)(u * 2.2 * of_max);
                if (f.sbytes == 4 && (double)y[n] < -of_max * 0.999 && (double)y[n] > -of_max - 2.0) y[n] = (T)0;
                break;
            case 1: y[n] = (T)(floor(u * 100.0) + 0.5); break;               // exact halves
            case 2: y[n] = (T)floor(u * 1000.0); break;                      // exact integers (negative quirk)
            case 3: y[n] = (T)(of_max + (u > 0 ? 0.4 : 0.6)); break;         // clip edge, positive
            case 4:                                                           // clip edge, negative
                // 32-bit: a value quantising to exactly INT32_MIN makes the reference negate INT32_MIN
                // (dither_funs.h:93-95, undefined behaviour), so stay one LSB inside there
                y[n] = f.sbytes == 4 ? (T)(-of_max * 0.999) : (T)(-of_max - 1.0 - (u > 0 ? 0.4 : 0.6));
                break;
            case 5: y[n] = (T)(u * 1e-3); break;
            case 6: y[n] = (T)(u * of_max); break;
            default: y[n] = (T)(-0.5 - 1e-7 * n); break;
            }
        }
        orc_overflow of_o = {2, 5, 0.25, of_max};
        std::vector<uint8_t> out_o((size_t)L * spacing * f.bytes + 64, 0x55), out_e(out_o);
        orc_cbuf2raw(y.data(), out_o.data(), &bf, 0, NULL, &of_o);
        QuantStats st;
        quant_stats_init(st);
        for (int n = 0; n < L; n++) {
            real_to_raw<T>(y[n], out_e.data() + offset + (size_t)n * spacing * f.bytes, f.bytes, f.sbytes, f.isfloat,
                           f.swap, 0.0, of_max, st);
        }
        // merge like reduce_stats does
        unsigned n_over = 2 + st.n_overflows;
        int32_t intl = st.intlargest > 5 ? st.intlargest : 5;
        double larg = st.largest > 0.25 ? st.largest : 0.25;
        const bool same = out_o == out_e && n_over == of_o.n_overflows && intl == of_o.intlargest && larg == of_o.largest;
        if (!same) {
            printf("MISMATCH %s rs=%d: overflows %u/%u intlargest %d/%d largest %.17g/%.17g bytes %s\n", f.name, rs, n_over,
                   of_o.n_overflows, intl, of_o.intlargest, larg, of_o.largest, out_o == out_e ? "same" : "DIFFER");
            bad++;
        }
    }
    return bad;
}

// the single-precision form of the quantiser (k_pack's 4-byte integer path) against the full double chain:
// every float in a dense sweep around the integers and halves, random values over the whole fast range and beyond it
static int fast_path()
{
    int bad = 0;
    long n_fast = 0;
    for (int sbytes = 3; sbytes <= 4; sbytes++) {
        const int bits_n = sbytes << 3;
        const int32_t imin = (int32_t)(-((uint64_t)1 << (bits_n - 1)));
        const int32_t imax = (int32_t)(((uint64_t)1 << (bits_n - 1)) - 1);
        const double rmin = (double)(float)imin, rmax = (double)(float)imax;
        const float thr = fminf(4194303.0f, (float)(imax - 1));
        auto check = [&](float x) {
            int32_t q, cand;
            if (!real_to_int_fast(x, thr, q, cand)) {
                return;
            }
            n_fast++;
            QuantStats st;
            quant_stats_init(st);
            const int32_t want = real_to_int<float>(x, rmin, rmax, imin, imax, st);
            if (q != want || cand != st.intlargest || st.n_overflows != 0 || st.largest != 0.0) {
                if (bad < 10) printf("fast path: x = %.9g -> %d (cand %d), full chain %d (intlargest %d)\n", x, q, cand, want, st.intlargest);
                bad++;
            }
        };
        for (int k = -5000; k <= 5000; k++) {           // integers, halves, quarters and their float neighbours
            for (int f4 = 0; f4 < 4; f4++) {
                const float c = (float)k + 0.25f * f4;
                check(c);
                check(nextafterf(c, 1e30f));
                check(nextafterf(c, -1e30f));
                check(-c);
            }
        }
        const float edges[] = { 4194303.0f, 4194302.75f, 4194302.5f, 4194303.25f, 4194304.0f, 8388606.0f, 8388607.0f, 0.0f, -0.0f,
                                -0.5f, 0.5f, -1.0f, -1.5f, 1e-30f, -1e-30f, 2097151.75f, -2097152.0f, INFINITY, -INFINITY, NAN };
        for (float e : edges) {
            check(e);
            check(-e);
        }
        for (int i = 0; i < 4000000; i++) {
            const double u = (double)rand() / RAND_MAX - 0.5;
            const int sh = rand() % 24;
            check((float)(u * 2.2 * (double)(1 << sh)));
            check((float)(floor(u * (double)(1 << sh)) + ((rand() & 1) ? 0.5 : 0.0)));
        }
    }
    printf("fast path: %ld values checked, %d mismatches\n", n_fast, bad);
    return bad;
}

// the packed S24_LE tile path of k_unpack / k_pack (bf_fft2_kernels.cu), one warp emulated lane by lane: 3 nc / 4 lanes
// hold the words of a tile row, shuffles + a funnel shift pick (or place) every channel's three bytes -- against the
// generic per-sample routines, for every tile width nc = 4 .. 32
static uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) { return sh == 0 ? lo : (uint32_t)((((uint64_t)hi << 32) | lo) >> sh); }
static int packed_tiles()
{
    int bad = 0;
    for (int nc = 4; nc <= 32; nc += 4) {
        for (int trial = 0; trial < 200; trial++) {
            uint8_t rowb[96 + 8] = {0}, outb[96 + 8];
            for (int i = 0; i < 3 * nc; i++) rowb[i] = (uint8_t)(rand() & 0xff);
            uint32_t wv[32];
            for (int lane = 0; lane < 32; lane++) {
                wv[lane] = 0;
                if (4 * lane < 3 * nc) memcpy(&wv[lane], rowb + 4 * lane, 4);
            }
            int32_t q[32];
            for (int lane = 0; lane < 32; lane++) {             // unpack
                const int w0 = (3 * lane) >> 2, sh = ((3 * lane) & 3) * 8;
                const uint32_t v3 = funnel_r(wv[w0], wv[(w0 + 1) & 31], sh);
                const float got = (float)((int32_t)(v3 << 8) >> 8);
                q[lane] = (int32_t)got;
                if (lane < nc) {
                    const float want = raw_to_real<float>(rowb + 3 * lane, 3, 0, 0);
                    if (got != want) bad++;
                } else {
                    q[lane] = 0;
                }
            }
            memset(outb, 0xaa, sizeof(outb));
            for (int lane = 0; lane < 32; lane++) {             // pack
                const int ca = (4 * lane) / 3, sh = ((4 * lane) % 3) * 8;
                const uint32_t qa = (uint32_t)q[ca & 31] & 0xffffffu, qb = (uint32_t)q[(ca + 1) & 31] & 0xffffffu;
                if (4 * lane < 3 * nc) {
                    const uint32_t word = (qa >> sh) | (qb << (24 - sh));
                    memcpy(outb + 4 * lane, &word, 4);
                }
            }
            if (memcmp(outb, rowb, 3 * nc) != 0 || outb[3 * nc] != 0xaa) bad++;
        }
    }
    printf("packed S24_LE tiles: %d mismatches\n", bad);
    return bad;
}

int main()
{
    srand(12345);
    int bad = run<float>(256) + run<double>(256) + fast_path() + packed_tiles();
    printf("%s\n", bad ? "FAIL" : "emul_sample ok");
    return bad;
}

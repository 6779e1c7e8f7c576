// bf_convolver.cu -- the reference's per-call interface (include/bfcuda_convolver.h == convolver.h) on
// the GPU: host pointers in, one or a few kernel launches, host pointers out.  See the header for why
// this exists next to the block-level engine.  One process-wide context, like the reference's file-scope
// statics (fftw_convolver.c:36-49); not re-entrant, like the reference.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/bfcuda_convolver.h"
#include "bf_kernels.h"

using namespace bf;

struct _td_conv_t_ {            // fftw_convolver.c:682-687, device edition
    FftPlan plan;
    bool has_plan;              // transforms of 2 and 4 points go through launch_td_small
    char *d_coeffs, *d_a, *d_b;
    int blocklen, rs;
};

namespace {

struct CvContext {
    bool ready;
    int L, N, rs;
    FftPlan plan;
    cudaStream_t stream;
    char *d_a, *d_b, *d_c;      // three cbuf-sized (x2) work buffers
    void **d_ptrs;              // mixnscale input pointer table
    double *d_scales;
    char *d_mix;                // mixnscale input staging, grown on demand
    size_t mix_cap;
    uint8_t *d_raw;             // raw sample staging, grown on demand
    size_t raw_cap;
    Overflow *d_of;
    unsigned int *d_status;
    // dither (convolver_cbuf2raw with apply_dither)
    int8_t *d_tab;              // this block's slice of the host's dither table, [previous byte | L bytes]
    size_t tab_cap;
    void *d_map;                // 512 reals: dither value per table difference (dither.c:113-131)
    DitherChan *d_chan;
    SampleFormat *d_fmt;
};

CvContext g_cv;
void (*g_exit_hook)(int) = nullptr;
int g_quiet = 0;
double g_safety_limit = 0.0;
int8_t *g_dither_tab = nullptr;     // the host's dither_randtab / dither_randtab_size
int g_dither_tab_size = 0;
char g_cv_err[512] = "";

void cv_exit(int status)
{
    if (g_exit_hook != nullptr) {
        g_exit_hook(status);
    } else {
        exit(status);       // stands in for bf_exit(), bfrun.c:2790-2820
    }
}

bool cv_fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_cv_err, sizeof(g_cv_err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "%s\n", g_cv_err);
    return false;
}

// CUDA failures have no counterpart in the reference: report and leave through the exit hook
bool cu_ok(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) {
        return true;
    }
    cv_fail("bfcuda convolver: %s failed: %s", what, cudaGetErrorString(e));
    cv_exit(1 /* BF_EXIT_OTHER */);
    return false;
}
#define CVCU(call)                    \
    do {                              \
        if (!cu_ok((call), #call)) {  \
            return;                   \
        }                             \
    } while (0)

size_t cbytes() { return (size_t)g_cv.N * g_cv.rs; }

bool need_ready()
{
    if (!g_cv.ready) {
        cv_fail("bfcuda convolver: convolver_init() has not been called");
        cv_exit(1);
        return false;
    }
    return true;
}

bool grow(char **p, size_t *cap, size_t want)
{
    if (*cap >= want) {
        return true;
    }
    if (*p != nullptr) {
        cudaFree(*p);
        *p = nullptr;
    }
    if (!cu_ok(cudaMalloc((void **)p, want), "cudaMalloc")) {
        *cap = 0;
        return false;
    }
    *cap = want;
    return true;
}

SampleFormat dev_format(const bfcuda_buffer_format *bf, int byte_offset)
{
    SampleFormat f;
    f.isfloat = bf->sf.isfloat;
    f.swap = bf->sf.swap;
    f.bytes = bf->sf.bytes;
    f.sbytes = bf->sf.sbytes;
    f.sample_spacing = bf->sample_spacing;
    f.byte_offset = byte_offset;
    return f;
}

// bytes from the first to the last byte this channel touches inside a raw block
size_t raw_span(const bfcuda_buffer_format *bf)
{
    return ((size_t)(g_cv.L - 1) * bf->sample_spacing + 1) * bf->sf.bytes;
}

void h2d(void *dst, const void *src, size_t n) { cu_ok(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, g_cv.stream), "H2D"); }
void d2h(void *dst, const void *src, size_t n) { cu_ok(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, g_cv.stream), "D2H"); }
void sync() { cu_ok(cudaStreamSynchronize(g_cv.stream), "cudaStreamSynchronize"); }

// convolver_cbuf2raw with apply_dither (fftw_convolver.c:489-499): the preloop on the host, exactly as
// dither_preloop_real2int_hp_tpdf (dither.h:28-38) -- it moves the state's table pointer and, on a wrap, copies the
// last used byte into slot 0 of the HOST's table --, then this block's L + 1 table bytes, the error-feedback state and
// the samples go to the device, one lane runs the reference's sequential loop (k_dither, dither_funs.h:7-68), and the
// raw samples, the overflow record and the state come back.
void cv_cbuf2raw_dither(void *cbuf, void *outbuf, struct bfcuda_buffer_format *bf, struct bfcuda_dither_state *state,
                        struct bfcuda_overflow *overflow)
{
    const int L = g_cv.L, rs = g_cv.rs;
    if (state == nullptr || g_dither_tab == nullptr) {
        cv_fail("bfcuda convolver: dither needs the host's table: call bfcuda_convolver_set_dither_table() after "
                "dither_init(), and pass the channel's struct dither_state");
        cv_exit(1);
        return;
    }
    if (bf->sf.bytes < 1 || bf->sf.bytes > 4) {
        fprintf(stderr, "Sample byte size %d is not supported.\n", bf->sf.bytes);   // real2raw.h:245-249
        cv_exit(1);
        return;
    }
    if (L + 1 >= g_dither_tab_size) {
        cv_fail("bfcuda convolver: dither table of %d entries is shorter than a block", g_dither_tab_size);
        cv_exit(1);
        return;
    }
    // dither.h:28-38
    if (state->randtab_ptr + L >= g_dither_tab_size) {
        g_dither_tab[0] = g_dither_tab[state->randtab_ptr - 1];
        state->randtab_ptr = 1;
    }
    state->randtab = &g_dither_tab[state->randtab_ptr];
    state->randtab_ptr += L;

    if (g_cv.d_map == nullptr) {
        // dither.c:113-131: table difference -> dither in (-1, +1) plus the +0.5 of the mid-tread requantiser
        std::vector<double> mapd(512);
        std::vector<float> mapf(512);
        mapf[0] = -0.5f;
        mapd[0] = -0.5;
        for (int k = -255; k < 254; k++) {
            mapf[k + 256] = (float)(0.5 + 1.0 / 255.0 + 1.0 / 255.0 * (float)k);
            mapd[k + 256] = 0.5 + 1.0 / 255.0 + 1.0 / 255.0 * (double)k;
        }
        mapf[510] = 1.5f;
        mapd[510] = 1.5;
        mapf[511] = (float)(1.5 + 1.0 / 255.0);
        mapd[511] = 1.5 + 1.0 / 255.0;
        if (!cu_ok(cudaMalloc(&g_cv.d_map, 512 * (size_t)rs), "cudaMalloc") ||
            !cu_ok(cudaMalloc((void **)&g_cv.d_chan, sizeof(DitherChan)), "cudaMalloc") ||
            !cu_ok(cudaMalloc((void **)&g_cv.d_fmt, sizeof(SampleFormat)), "cudaMalloc") ||
            !cu_ok(cudaMemcpy(g_cv.d_map, rs == 4 ? (const void *)mapf.data() : (const void *)mapd.data(),
                              512 * (size_t)rs, cudaMemcpyHostToDevice), "cudaMemcpy")) {
            return;
        }
    }
    if (!grow((char **)&g_cv.d_tab, &g_cv.tab_cap, (size_t)L + 1)) return;
    const size_t span = raw_span(bf);
    uint8_t *host_raw = (uint8_t *)outbuf + bf->byte_offset;
    if (!grow((char **)&g_cv.d_raw, &g_cv.raw_cap, span)) return;
    Overflow of;
    of.n_overflows = overflow->n_overflows;
    of.intlargest = overflow->intlargest;
    of.largest = overflow->largest;
    of.max = overflow->max;
    unsigned int status = 0;
    DitherChan ch;
    ch.out = 0;
    ch.randtab_ptr = 1;
    ch.e0 = rs == 4 ? (double)state->sf[0] : state->sd[0];
    ch.e1 = rs == 4 ? (double)state->sf[1] : state->sd[1];
    const SampleFormat fmt = dev_format(bf, 0);
    h2d(g_cv.d_tab, state->randtab - 1, (size_t)L + 1);
    h2d(g_cv.d_raw, host_raw, span);    // neighbouring channels' bytes inside the span must survive
    h2d(g_cv.d_a, cbuf, (size_t)L * rs);
    h2d(g_cv.d_of, &of, sizeof(of));
    h2d(g_cv.d_status, &status, sizeof(status));
    h2d(g_cv.d_chan, &ch, sizeof(ch));
    h2d(g_cv.d_fmt, &fmt, sizeof(fmt));
    InverseArgs ia;
    memset(&ia, 0, sizeof(ia));
    ia.out_time = g_cv.d_a;
    ia.raw_out = g_cv.d_raw;
    ia.fmt = g_cv.d_fmt;
    ia.overflow = g_cv.d_of;
    ia.status = g_cv.d_status;
    ia.n_out = 1;
    ia.batch = 1;
    ia.safety_limit = g_safety_limit;
    DitherArgs da;
    da.chans = g_cv.d_chan;
    da.randtab = g_cv.d_tab;
    da.randmap = g_cv.d_map;
    da.randtab_size = L + 2;        // never wraps on the device: the host did the preloop
    da.n_dither = 1;
    CVCU(launch_dither(g_cv.plan, ia, da, g_cv.stream));
    d2h(host_raw, g_cv.d_raw, span);
    d2h(&of, g_cv.d_of, sizeof(of));
    d2h(&status, g_cv.d_status, sizeof(status));
    d2h(&ch, g_cv.d_chan, sizeof(ch));
    sync();
    overflow->n_overflows = of.n_overflows;
    overflow->intlargest = of.intlargest;
    overflow->largest = of.largest;
    if (rs == 4) {
        state->sf[0] = (float)ch.e0;
        state->sf[1] = (float)ch.e1;
    } else {
        state->sd[0] = ch.e0;
        state->sd[1] = ch.e1;
    }
    if (status & 1u) {
        fprintf(stderr, "NaN or Inf values in the output! Bad output. Aborting.\n");    // real2raw.h:27-31
        cv_exit(-5);
    } else if (status & 2u) {
        fprintf(stderr, "Safety limit exceeded on output. Aborting.\n");                // real2raw.h:32-41
        cv_exit(1);
    }
}

// device-side mixnscale over device-resident inputs
bool dev_mixnscale(char *const in[], const double scales[], int n, char *out, int mode)
{
    std::vector<void *> ptrs(in, in + n);
    if (!cu_ok(cudaMemcpyAsync(g_cv.d_ptrs, ptrs.data(), sizeof(void *) * n, cudaMemcpyHostToDevice, g_cv.stream), "H2D")) return false;
    if (!cu_ok(cudaMemcpyAsync(g_cv.d_scales, scales, sizeof(double) * n, cudaMemcpyHostToDevice, g_cv.stream), "H2D")) return false;
    return cu_ok(launch_cv_mixnscale(g_cv.plan, (const void *const *)g_cv.d_ptrs, g_cv.d_scales, n, out, mode, g_cv.stream), "mixnscale");
}

}  // namespace

extern "C" {

void bfcuda_convolver_set_host(void (*bf_exit_hook)(int), int quiet, double safety_limit)
{
    g_exit_hook = bf_exit_hook;
    g_quiet = quiet;
    g_safety_limit = safety_limit;
}

const char *bfcuda_convolver_last_error(void)
{
    return g_cv_err;
}

bool_t convolver_init(const char config_filename[], int length, int realsize)
{
    (void)config_filename;  // FFTW wisdom: "Some convolvers may ignore 'config_filename'" (convolver.h:147)
    if (realsize != 4 && realsize != 8) {
        fprintf(stderr, "Invalid real size %d.\n", realsize);       // fftw_convolver.c:796-799
        return 0;
    }
    if (length < 1 || (length & (length - 1)) != 0) {
        fprintf(stderr, "Invalid length %d.\n", length);           // fftw_convolver.c:800-803
        return 0;
    }
    if (!fft_single_block_supported(2 * length, realsize)) {    // the per-call surface runs one-block transforms only
        fprintf(stderr, "Invalid length %d (single-block FFT supports 4..%d at realsize %d).\n", length,
                realsize == 4 ? 16384 : 8192, realsize);
        return 0;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fprintf(stderr, "No CUDA device available (there is no CPU fallback).\n");
        return 0;
    }
    if (g_cv.ready) {
        cudaStreamSynchronize(g_cv.stream);
        fft_plan_destroy(&g_cv.plan);
        cudaFree(g_cv.d_a); cudaFree(g_cv.d_b); cudaFree(g_cv.d_c); cudaFree(g_cv.d_ptrs); cudaFree(g_cv.d_scales);
        cudaFree(g_cv.d_of); cudaFree(g_cv.d_status);
        if (g_cv.d_mix) cudaFree(g_cv.d_mix);
        if (g_cv.d_raw) cudaFree(g_cv.d_raw);
        if (g_cv.d_tab) cudaFree(g_cv.d_tab);
        if (g_cv.d_map) cudaFree(g_cv.d_map);
        if (g_cv.d_chan) cudaFree(g_cv.d_chan);
        if (g_cv.d_fmt) cudaFree(g_cv.d_fmt);
        cudaStreamDestroy(g_cv.stream);
        memset(&g_cv, 0, sizeof(g_cv));
    }
    g_cv.L = length;
    g_cv.N = 2 * length;
    g_cv.rs = realsize;
    const size_t cb = 2 * cbytes();
    if (cudaStreamCreateWithFlags(&g_cv.stream, cudaStreamNonBlocking) != cudaSuccess ||
        fft_plan_create(&g_cv.plan, g_cv.N, realsize) != cudaSuccess ||
        cudaMalloc((void **)&g_cv.d_a, cb) != cudaSuccess || cudaMalloc((void **)&g_cv.d_b, cb) != cudaSuccess ||
        cudaMalloc((void **)&g_cv.d_c, cb) != cudaSuccess ||
        cudaMalloc((void **)&g_cv.d_ptrs, sizeof(void *) * (BFCUDA_MAXCHANNELS + BFCUDA_MAXFILTERS)) != cudaSuccess ||
        cudaMalloc((void **)&g_cv.d_scales, sizeof(double) * (BFCUDA_MAXCHANNELS + BFCUDA_MAXFILTERS)) != cudaSuccess ||
        cudaMalloc((void **)&g_cv.d_of, sizeof(Overflow)) != cudaSuccess ||
        cudaMalloc((void **)&g_cv.d_status, sizeof(unsigned int)) != cudaSuccess) {
        fprintf(stderr, "CUDA initialisation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    g_cv.ready = true;
    return 1;
}

int convolver_cbufsize(void)
{
    return g_cv.N * g_cv.rs;    // fftw_convolver.c:520-524
}

void convolver_raw2cbuf(void *rawbuf, void *cbuf, void *next_cbuf, struct bfcuda_buffer_format *bf,
                        void (*postprocess)(void *, int, void *), void *pp_arg)
{
    if (!need_ready()) return;
    const bool okf = bf->sf.isfloat ? (bf->sf.bytes == 4 || bf->sf.bytes == 8) : (bf->sf.bytes >= 1 && bf->sf.bytes <= 4);
    if (!okf) {
        fprintf(stderr, "Sample byte size %d is not supported.\n", bf->sf.bytes);   // raw2real.h:154-158
        cv_exit(1);
        return;
    }
    const size_t span = raw_span(bf), half = (size_t)g_cv.L * g_cv.rs;
    if (!grow((char **)&g_cv.d_raw, &g_cv.raw_cap, span)) return;
    h2d(g_cv.d_raw, (const uint8_t *)rawbuf + bf->byte_offset, span);
    CVCU(launch_cv_raw2real(g_cv.plan, g_cv.d_raw, dev_format(bf, 0), g_cv.d_a, g_cv.stream));
    d2h(next_cbuf, g_cv.d_a, half);
    sync();
    if (postprocess != nullptr) {
        postprocess(next_cbuf, g_cv.L, pp_arg);
    }
    memcpy((uint8_t *)cbuf + half, next_cbuf, half);    // fftw_convolver.c:193
}

void convolver_time2freq(void *input_cbuf, void *output_cbuf)
{
    if (!need_ready()) return;
    h2d(g_cv.d_a, input_cbuf, cbytes());
    CVCU(launch_r2hc(g_cv.plan, g_cv.d_a, g_cv.d_b, 1, g_cv.stream));
    d2h(output_cbuf, g_cv.d_b, cbytes());
    sync();
}

void convolver_freq2time(void *input_cbuf, void *output_cbuf)
{
    if (!need_ready()) return;
    h2d(g_cv.d_a, input_cbuf, cbytes());
    CVCU(launch_hc2r(g_cv.plan, g_cv.d_a, g_cv.d_b, 1, g_cv.stream));
    d2h(output_cbuf, g_cv.d_b, cbytes());
    sync();
}

void convolver_mixnscale(void *input_cbufs[], void *output_cbuf, double scales[], int n_bufs, int mixmode)
{
    if (!need_ready()) return;
    if (mixmode != CONVOLVER_MIXMODE_INPUT && mixmode != CONVOLVER_MIXMODE_OUTPUT) {
        fprintf(stderr, "Invalid mixmode: %d.\n", mixmode);     // fftw_convfuns.h:496-499
        cv_exit(1);
        return;
    }
    if (n_bufs < 1 || n_bufs > BFCUDA_MAXCHANNELS + BFCUDA_MAXFILTERS) {
        cv_fail("bfcuda convolver: mixnscale with %d buffers", n_bufs);
        cv_exit(1);
        return;
    }
    if (!grow(&g_cv.d_mix, &g_cv.mix_cap, (size_t)n_bufs * cbytes())) return;
    std::vector<char *> in(n_bufs);
    for (int i = 0; i < n_bufs; i++) {
        in[i] = g_cv.d_mix + (size_t)i * cbytes();
        h2d(in[i], input_cbufs[i], cbytes());
    }
    if (!dev_mixnscale(in.data(), scales, n_bufs, g_cv.d_a, mixmode)) return;
    d2h(output_cbuf, g_cv.d_a, cbytes());
    sync();
}

static void cv_convolve(void *input_cbuf, void *coeffs, void *output_cbuf, int op)
{
    if (!need_ready()) return;
    h2d(g_cv.d_a, input_cbuf, cbytes());
    h2d(g_cv.d_b, coeffs, cbytes());
    if (op) {
        h2d(g_cv.d_c, output_cbuf, cbytes());
    }
    CVCU(launch_cv_convolve(g_cv.plan, g_cv.d_a, g_cv.d_b, g_cv.d_c, op, g_cv.stream));
    d2h(output_cbuf, g_cv.d_c, cbytes());
    sync();
}

void convolver_convolve(void *input_cbuf, void *coeffs, void *output_cbuf)
{
    cv_convolve(input_cbuf, coeffs, output_cbuf, 0);
}

void convolver_convolve_inplace(void *cbuf, void *coeffs)
{
    cv_convolve(cbuf, coeffs, cbuf, 0);     // fftw_convfuns.h:503-532: same arithmetic, result over the input
}

void convolver_convolve_add(void *input_cbuf, void *coeffs, void *output_cbuf)
{
    cv_convolve(input_cbuf, coeffs, output_cbuf, 1);
}

void convolver_dirac_convolve(void *input_cbuf, void *output_cbuf)
{
    if (!need_ready()) return;
    h2d(g_cv.d_a, input_cbuf, cbytes());
    CVCU(launch_cv_dirac(g_cv.plan, g_cv.d_a, g_cv.d_b, g_cv.stream));
    d2h(output_cbuf, g_cv.d_b, cbytes());
    sync();
}

void convolver_dirac_convolve_inplace(void *cbuf)
{
    convolver_dirac_convolve(cbuf, cbuf);
}

// fftw_convolver.c:330-368, float branch semantics for both precisions (see DESIGN.md on the reference's
// broken double branch).  All five transforms stay on the device.
void convolver_crossfade_inplace(void *input_cbuf, void *crossfade_cbuf, void *buffer_cbuf)
{
    if (!need_ready()) return;
    const size_t cb = cbytes();
    const double one = 1.0, inv_n = 1.0 / (double)g_cv.N;
    char *d_new = g_cv.d_a, *d_old = g_cv.d_b, *d_tmp = g_cv.d_c;
    char *d_old_t = g_cv.d_b + cb, *d_new_t = g_cv.d_a + cb;
    h2d(d_new, input_cbuf, cb);
    h2d(d_old, crossfade_cbuf, cb);
    char *in[1];
    in[0] = d_old;
    if (!dev_mixnscale(in, &one, 1, d_tmp, CONVOLVER_MIXMODE_OUTPUT)) return;
    CVCU(launch_hc2r(g_cv.plan, d_tmp, d_old_t, 1, g_cv.stream));
    in[0] = d_new;
    if (!dev_mixnscale(in, &one, 1, d_tmp, CONVOLVER_MIXMODE_OUTPUT)) return;
    CVCU(launch_hc2r(g_cv.plan, d_tmp, d_new_t, 1, g_cv.stream));
    CVCU(launch_cv_xfade_blend(g_cv.plan, d_old_t, d_new_t, g_cv.stream));
    CVCU(launch_r2hc(g_cv.plan, d_new_t, d_tmp, 1, g_cv.stream));
    in[0] = d_tmp;
    if (!dev_mixnscale(in, &inv_n, 1, d_new, CONVOLVER_MIXMODE_INPUT)) return;
    d2h(input_cbuf, d_new, cb);
    d2h(crossfade_cbuf, d_old_t, cb);   // the reference leaves the old signal's time domain here (l.345)
    d2h(buffer_cbuf, d_tmp, cb);        // and the blended spectrum, half-complex, here (l.364)
    sync();
}

// fftw_convolver.c:411-433
void convolver_convolve_eval(void *input_cbuf, void *buffer_cbuf, void *output_cbuf)
{
    if (!need_ready()) return;
    const size_t cb = cbytes(), half = cb / 2;
    // d_b holds the 1.5 x cbuf state
    h2d(g_cv.d_a, input_cbuf, cb);
    h2d(g_cv.d_b, buffer_cbuf, half);
    CVCU(launch_hc2r(g_cv.plan, g_cv.d_a, g_cv.d_b + half, 1, g_cv.stream));
    CVCU(launch_r2hc(g_cv.plan, g_cv.d_b, g_cv.d_c, 1, g_cv.stream));
    d2h(output_cbuf, g_cv.d_c, cb);
    d2h((char *)buffer_cbuf + half, g_cv.d_b + half, cb);
    sync();
    memcpy(buffer_cbuf, (char *)buffer_cbuf + half, half);
}

void convolver_cbuf2raw(void *cbuf, void *outbuf, struct bfcuda_buffer_format *bf, bool_t apply_dither,
                        void *dither_state, struct bfcuda_overflow *overflow)
{
    if (!need_ready()) return;
    if (apply_dither && !bf->sf.isfloat) {
        cv_cbuf2raw_dither(cbuf, outbuf, bf, (struct bfcuda_dither_state *)dither_state, overflow);
        return;
    }
    const bool okf = bf->sf.isfloat ? (bf->sf.bytes == 4 || bf->sf.bytes == 8) : (bf->sf.bytes >= 1 && bf->sf.bytes <= 4);
    if (!okf) {
        fprintf(stderr, "Sample byte size %d is not supported.\n", bf->sf.bytes);   // real2raw.h:245-249
        cv_exit(1);
        return;
    }
    const size_t span = raw_span(bf);
    uint8_t *host_raw = (uint8_t *)outbuf + bf->byte_offset;
    if (!grow((char **)&g_cv.d_raw, &g_cv.raw_cap, span)) return;
    Overflow of;
    of.n_overflows = overflow->n_overflows;
    of.intlargest = overflow->intlargest;
    of.largest = overflow->largest;
    of.max = overflow->max;
    unsigned int status = 0;
    h2d(g_cv.d_raw, host_raw, span);    // neighbouring channels' bytes inside the span must survive
    h2d(g_cv.d_a, cbuf, (size_t)g_cv.L * g_cv.rs);
    h2d(g_cv.d_of, &of, sizeof(of));
    h2d(g_cv.d_status, &status, sizeof(status));
    CVCU(launch_cv_real2raw(g_cv.plan, g_cv.d_a, g_cv.d_raw, dev_format(bf, 0), g_cv.d_of, g_cv.d_status,
                            g_safety_limit, g_cv.stream));
    d2h(host_raw, g_cv.d_raw, span);
    d2h(&of, g_cv.d_of, sizeof(of));
    d2h(&status, g_cv.d_status, sizeof(status));
    sync();
    overflow->n_overflows = of.n_overflows;
    overflow->intlargest = of.intlargest;
    overflow->largest = of.largest;
    if (status & 1u) {
        fprintf(stderr, "NaN or Inf values in the output! Bad output. Aborting.\n");    // real2raw.h:27-31
        cv_exit(-5);
    } else if (status & 2u) {
        fprintf(stderr, "Safety limit exceeded on output. Aborting.\n");                // real2raw.h:32-41
        cv_exit(1);
    }
}

void *convolver_coeffs2cbuf(void *coeffs, int n_coeffs, double scale, void *optional_dest)
{
    if (!need_ready()) return nullptr;
    const int len = n_coeffs > g_cv.L ? g_cv.L : n_coeffs;
    std::vector<unsigned char> padded((size_t)g_cv.L * g_cv.rs, 0);
    // fftw_convolver.c:539-557: reject NaN/Inf among the scaled taps
    for (int n = 0; n < len; n++) {
        const double v = g_cv.rs == 4 ? (double)(((float *)coeffs)[n] * (float)scale) : ((double *)coeffs)[n] * scale;
        if (!std::isfinite(v)) {
            fprintf(stderr, "NaN or Inf value among coefficients.\n");
            return nullptr;
        }
    }
    memcpy(padded.data(), coeffs, (size_t)len * g_cv.rs);
    h2d(g_cv.d_a, padded.data(), padded.size());
    if (!cu_ok(launch_coeff_fft(g_cv.plan, g_cv.d_a, 1, scale, g_cv.d_b, 0, g_cv.stream), "coeff_fft")) return nullptr;
    if (!cu_ok(launch_permute(g_cv.plan, g_cv.d_b, g_cv.d_c, 1, PLANAR_TO_BLOCKED, g_cv.stream), "permute")) return nullptr;
    void *dest = optional_dest;
    if (dest == nullptr && posix_memalign(&dest, 32, cbytes()) != 0) {     // emallocaligned, sysarch.h:10
        return nullptr;
    }
    d2h(dest, g_cv.d_c, cbytes());
    sync();
    return dest;
}

void convolver_runtime_coeffs2cbuf(void *src, void *dest)
{
    convolver_coeffs2cbuf(src, g_cv.L, 1.0, dest);  // fftw_convolver.c:575-596: same transform, scale 1, no check
}

bool_t convolver_verify_cbuf(void *cbufs[], int n_cbufs)
{
    // fftw_convolver.c:598-622; a scan of host memory, nothing to accelerate
    for (int n = 0; n < n_cbufs; n++) {
        for (int i = 0; i < g_cv.N; i++) {
            const double v = g_cv.rs == 4 ? (double)((float *)cbufs[n])[i] : ((double *)cbufs[n])[i];
            if (!std::isfinite(v)) {
                fprintf(stderr, "NaN or Inf value among coefficients.\n");
                return 0;
            }
        }
    }
    return 1;
}

void convolver_debug_dump_cbuf(const char filename[], void *cbufs[], int n_cbufs)
{
    // fftw_convolver.c:624-660: back to taps (second half of the inverse transform), one per line
    if (!need_ready()) return;
    FILE *stream = fopen(filename, "wt");
    if (stream == nullptr) {
        fprintf(stderr, "Could not open \"%s\" for writing.", filename);
        return;
    }
    std::vector<unsigned char> host(cbytes());
    const double one = 1.0;
    for (int n = 0; n < n_cbufs; n++) {
        char *in[1] = { g_cv.d_a };
        h2d(g_cv.d_a, cbufs[n], cbytes());
        if (!dev_mixnscale(in, &one, 1, g_cv.d_b, CONVOLVER_MIXMODE_OUTPUT)) break;
        if (!cu_ok(launch_hc2r(g_cv.plan, g_cv.d_b, g_cv.d_c, 1, g_cv.stream), "hc2r")) break;
        d2h(host.data(), g_cv.d_c, cbytes());
        sync();
        for (int i = 0; i < g_cv.L; i++) {
            const double v = g_cv.rs == 4 ? (double)((float *)host.data())[g_cv.L + i] : ((double *)host.data())[g_cv.L + i];
            fprintf(stream, "%.16e\n", v);
        }
    }
    fclose(stream);
}

void bfcuda_convolver_set_dither_table(int8_t *dither_randtab, int dither_randtab_size)
{
    g_dither_tab = dither_randtab;
    g_dither_tab_size = dither_randtab_size;
}

void *convolver_fftplan(int order, int invert, int inplace)
{
    (void)order; (void)invert; (void)inplace;
    cv_fail("bfcuda convolver: convolver_fftplan() hands out FFTW plans (fftw_convolver.c:662-680); there is no "
            "FFTW behind this convolver");
    return nullptr;
}

int convolver_td_block_length(int n_coeffs)
{
    // fftw_convolver.c:689-696 / log2.h:28-43: next power of two (n_coeffs = 1 is undefined there: 1 << -1; here 1)
    if (n_coeffs < 1) {
        return -1;
    }
    int len = 1;
    while (len < n_coeffs) {
        len <<= 1;
    }
    return len;
}

// The sub-sample-delay convolver (fftw_convolver.c:682-782; caller: delay.c:415-506, 199-tap sinc filters, block
// length 256).  An object owns its transform plan, its scaled coefficient spectrum and two work buffers on the device.
td_conv_t *convolver_td_new(void *coeffs, int n_coeffs)
{
    if (!need_ready()) return nullptr;
    const int blocklen = convolver_td_block_length(n_coeffs);
    if (blocklen == -1 || coeffs == nullptr) {
        return nullptr;                                                     // fftw_convolver.c:707-709
    }
    const int n = 2 * blocklen, rs = g_cv.rs;
    if (n >= 8 && !fft_single_block_supported(n, rs)) {
        cv_fail("bfcuda convolver: convolver_td_new: %d coefficients need a %d-point transform, above the single-block "
                "limit", n_coeffs, n);
        return nullptr;
    }
    td_conv_t *tdc = (td_conv_t *)calloc(1, sizeof(td_conv_t));
    if (tdc == nullptr) {
        cv_fail("bfcuda convolver: out of memory");
        return nullptr;
    }
    tdc->blocklen = blocklen;
    tdc->rs = rs;
    const size_t bytes = (size_t)n * rs;
    bool ok = cu_ok(cudaMalloc((void **)&tdc->d_coeffs, bytes), "cudaMalloc") &&
              cu_ok(cudaMalloc((void **)&tdc->d_a, bytes), "cudaMalloc") &&
              cu_ok(cudaMalloc((void **)&tdc->d_b, bytes), "cudaMalloc");
    if (ok && n >= 8) {
        ok = cu_ok(fft_plan_create(&tdc->plan, n, rs), "fft_plan_create");
        tdc->has_plan = ok;
    }
    if (ok) {
        // [0_blocklen | taps | 0] -> R2HC -> * 1/(2 blocklen)                     fftw_convolver.c:716-732
        std::vector<char> frame(bytes, 0);
        memcpy(frame.data() + (size_t)blocklen * rs, coeffs, (size_t)n_coeffs * rs);
        ok = cu_ok(cudaMemcpyAsync(tdc->d_a, frame.data(), bytes, cudaMemcpyHostToDevice, g_cv.stream), "H2D") &&
             cu_ok(tdc->has_plan ? launch_r2hc(tdc->plan, tdc->d_a, tdc->d_coeffs, 1, g_cv.stream)
                                 : launch_td_small(rs, tdc->d_a, tdc->d_coeffs, n, 0, g_cv.stream), "td forward") &&
             cu_ok(launch_td_scale(rs, tdc->d_coeffs, n, rs == 4 ? (double)(1.0f / (float)n) : 1.0 / (double)n,
                                   g_cv.stream), "td scale") &&
             cu_ok(cudaStreamSynchronize(g_cv.stream), "cudaStreamSynchronize");
    }
    if (!ok) {
        bfcuda_convolver_td_delete(tdc);
        return nullptr;
    }
    return tdc;
}

void convolver_td_convolve(td_conv_t *tdc, void *overlap_block)
{
    if (!need_ready()) return;
    if (tdc == nullptr || overlap_block == nullptr) {
        cv_fail("bfcuda convolver: convolver_td_convolve: null argument");
        return;
    }
    const int n = 2 * tdc->blocklen, rs = tdc->rs;
    const size_t bytes = (size_t)n * rs;
    // R2HC, ordered product, HC2R, all in place on the caller's block              fftw_convolver.c:765-781
    h2d(tdc->d_a, overlap_block, bytes);
    if (tdc->has_plan) {
        CVCU(launch_r2hc(tdc->plan, tdc->d_a, tdc->d_b, 1, g_cv.stream));
    } else {
        CVCU(launch_td_small(rs, tdc->d_a, tdc->d_b, n, 0, g_cv.stream));
    }
    CVCU(launch_td_mul(rs, tdc->d_b, tdc->d_coeffs, n, g_cv.stream));
    if (tdc->has_plan) {
        CVCU(launch_hc2r(tdc->plan, tdc->d_b, tdc->d_a, 1, g_cv.stream));
    } else {
        CVCU(launch_td_small(rs, tdc->d_b, tdc->d_a, n, 1, g_cv.stream));
    }
    d2h(overlap_block, tdc->d_a, bytes);
    sync();
}

// Not on the reference surface (it never frees these objects): releases the device memory of a td convolver.
void bfcuda_convolver_td_delete(td_conv_t *tdc)
{
    if (tdc == nullptr) return;
    if (tdc->has_plan) fft_plan_destroy(&tdc->plan);
    if (tdc->d_coeffs) cudaFree(tdc->d_coeffs);
    if (tdc->d_a) cudaFree(tdc->d_a);
    if (tdc->d_b) cudaFree(tdc->d_b);
    free(tdc);
}

}  // extern "C"

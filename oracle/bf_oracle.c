/*
 * oracle/bf_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.  See bf_oracle.h.
 *
 * Non-generic part of the restatement: global sizes (fftw_convolver.c:36-49, 784-851), the two FFT
 * wrappers (196-214, 391-409), raw2cbuf / cbuf2raw (170-194, 482-518), convolve_eval (411-433),
 * runtime_coeffs2cbuf (575-596), verify_cbuf (598-622) and the realsize dispatch.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "bf_oracle.h"
#include "shim/fftw3.h"

static int g_realsize, g_n_fft, g_n_fft2;
static double g_safety_limit;
static void *g_plan_fwd, *g_plan_inv;
static void (*g_fail_handler)(int);

static void
orc_fail(int code)
{
    if (g_fail_handler != NULL) {
        g_fail_handler(code);
        return;
    }
    abort();
}

void
orc_set_fail_handler(void (*handler)(int code))
{
    g_fail_handler = handler;
}

static void *
orc_alloc(size_t size)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, size < 64 ? 64 : size) != 0) {
        fprintf(stderr, "oracle: out of memory\n");
        abort();
    }
    return p;
}

/* ---- dither.c / dither.h ------------------------------------------------------------------------------------ */
static int8_t *g_dither_randtab;
static int g_dither_randtab_size;
static void *g_dither_randmap;

/* dither.c:37-73: "maximally equidistributed combined Tausworthe generator" (GSL's taus), default seed */
static uint32_t
orc_tausrand(uint32_t state[3])
{
#define ORC_TAUSWORTHE(s, a, b, c, d) ((s & c) << d) ^ (((s << a) ^ s) >> b)
    state[0] = ORC_TAUSWORTHE(state[0], 13, 19, (uint32_t)4294967294U, 12);
    state[1] = ORC_TAUSWORTHE(state[1], 2, 25, (uint32_t)4294967288U, 4);
    state[2] = ORC_TAUSWORTHE(state[2], 3, 11, (uint32_t)4294967280U, 17);
    return state[0] ^ state[1] ^ state[2];
}

static void
orc_tausinit(uint32_t state[3], uint32_t seed)
{
    int i;
    if (seed == 0) {
        seed = 1;
    }
#define ORC_LCG(n) ((69069 * n) & 0xFFFFFFFFU)
    state[0] = ORC_LCG(seed);
    state[1] = ORC_LCG(state[0]);
    state[2] = ORC_LCG(state[1]);
    for (i = 0; i < 6; i++) {
        orc_tausrand(state);
    }
}

/* dither.c:75-139.  RANDTAB_SPACING 10 s, MIN_RANDTAB_SPACING 1 s (dither.c:19-21).  Returns 0 on failure.
 * NB the reference's map has 511 entries for differences -256..254, but two int8 values can differ by +255
 * (-128 followed by 127): the reference then reads one element past its allocation.  Here the map has that
 * 512th entry and it holds the linear continuation 1.5 + 1/255; tests keep away from such pairs. */
int
orc_dither_init(int n_channels, int sample_rate, int realsize, int max_size, int max_samples_per_loop,
                struct orc_dither_state *states)
{
    int n, spacing = 10 * sample_rate, minspacing;
    uint32_t st[3];

    minspacing = (1 * sample_rate > max_samples_per_loop) ? 1 * sample_rate : max_samples_per_loop;
    if (spacing < minspacing) {
        spacing = minspacing;
    }
    if (max_size > 0 && n_channels * spacing > max_size) {
        spacing = max_size / n_channels;
    }
    if (spacing < minspacing) {
        fprintf(stderr, "Maximum dither table size %d bytes is too small, must at least be %d bytes.\n", max_size,
                n_channels * sample_rate * minspacing);
        return 0;
    }
    g_dither_randtab_size = n_channels * spacing + 1;
    orc_tausinit(st, 0);
    free(g_dither_randtab);
    g_dither_randtab = orc_alloc((size_t)g_dither_randtab_size);
    for (n = 0; n < g_dither_randtab_size; n++) {
        g_dither_randtab[n] = (int8_t)(orc_tausrand(st) & 0x000000FF);
    }
    {
        static void *map_base;
        free(map_base);
        map_base = orc_alloc((size_t)realsize * 512);
        g_dither_randmap = (uint8_t *)map_base + 256 * realsize;
    }
    if (realsize == 4) {
        ((float *)g_dither_randmap)[-256] = -0.5;
        for (n = -255; n < 254; n++) {
            ((float *)g_dither_randmap)[n] = 0.5 + 1.0 / 255.0 + 1.0 / 255.0 * (float)n;
        }
        ((float *)g_dither_randmap)[254] = 1.5;
        ((float *)g_dither_randmap)[255] = 1.5 + 1.0 / 255.0;
    } else {
        ((double *)g_dither_randmap)[-256] = -0.5;
        for (n = -255; n < 254; n++) {
            ((double *)g_dither_randmap)[n] = 0.5 + 1.0 / 255.0 + 1.0 / 255.0 * (double)n;
        }
        ((double *)g_dither_randmap)[254] = 1.5;
        ((double *)g_dither_randmap)[255] = 1.5 + 1.0 / 255.0;
    }
    for (n = 0; n < n_channels; n++) {
        memset(&states[n], 0, sizeof(states[n]));
        states[n].randtab_ptr = n * spacing + 1;
    }
    return 1;
}

/* dither.h:28-38 */
static void
orc_dither_preloop(struct orc_dither_state *state, int samples_per_loop)
{
    if (state->randtab_ptr + samples_per_loop >= g_dither_randtab_size) {
        g_dither_randtab[0] = g_dither_randtab[state->randtab_ptr - 1];
        state->randtab_ptr = 1;
    }
    state->randtab = &g_dither_randtab[state->randtab_ptr];
    state->randtab_ptr += samples_per_loop;
}

const int8_t *
orc_dither_table(int *size)
{
    *size = g_dither_randtab_size;
    return g_dither_randtab;
}

#define REAL float
#define SFX f
#define ORC_REAL_IS_FLOAT 1
#include "bf_oracle_funs.inc"
#undef REAL
#undef SFX
#undef ORC_REAL_IS_FLOAT

#define REAL double
#define SFX d
#define ORC_REAL_IS_FLOAT 0
#include "bf_oracle_funs.inc"
#undef REAL
#undef SFX
#undef ORC_REAL_IS_FLOAT

/* fftw_convolver.c:784-851: L must be a power of two, realsize 4 or 8; n_fft = 2 L. */
int
orc_init(const char *unused_config, int length, int realsize)
{
    (void)unused_config;
    if (realsize != 4 && realsize != 8) {
        fprintf(stderr, "Invalid real size %d.\n", realsize);
        return 0;
    }
    if (length < 1 || (length & (length - 1)) != 0) {
        fprintf(stderr, "Invalid length %d.\n", length);
        return 0;
    }
    g_realsize = realsize;
    g_n_fft2 = length;
    g_n_fft = 2 * length;
    if (realsize == 4) {
        g_plan_fwd = fftwf_plan_r2r_1d(g_n_fft, NULL, NULL, FFTW_R2HC, FFTW_MEASURE);
        g_plan_inv = fftwf_plan_r2r_1d(g_n_fft, NULL, NULL, FFTW_HC2R, FFTW_MEASURE);
    } else {
        g_plan_fwd = fftw_plan_r2r_1d(g_n_fft, NULL, NULL, FFTW_R2HC, FFTW_MEASURE);
        g_plan_inv = fftw_plan_r2r_1d(g_n_fft, NULL, NULL, FFTW_HC2R, FFTW_MEASURE);
    }
    return 1;
}

int
orc_cbufsize(void)
{
    return g_n_fft * g_realsize;      /* fftw_convolver.c:520-524 */
}

void
orc_set_safety_limit(double limit)
{
    g_safety_limit = limit;           /* bfconf->safety_limit, read at real2raw.h:32 */
}

void
orc_time2freq(void *in, void *out)
{
    if (g_realsize == 4) {
        fftwf_execute_r2r(g_plan_fwd, in, out);
    } else {
        fftw_execute_r2r(g_plan_fwd, in, out);
    }
}

void
orc_freq2time(void *in, void *out)
{
    if (g_realsize == 4) {
        fftwf_execute_r2r(g_plan_inv, in, out);
    } else {
        fftw_execute_r2r(g_plan_inv, in, out);
    }
}

/* fftw_convolver.c:170-194: convert L new samples into next_cbuf[0..L), copy to cbuf[L..2L). */
void
orc_raw2cbuf(void *rawbuf, void *cbuf, void *next_cbuf, struct orc_buffer_format *bf,
             void (*postprocess)(void *, int, void *), void *pp_arg)
{
    const uint8_t *raw = (const uint8_t *)rawbuf + bf->byte_offset;
    if (g_realsize == 4) {
        orc_raw2realf(next_cbuf, raw, bf->sf.bytes, bf->sf.isfloat, bf->sample_spacing, bf->sf.swap,
                      g_n_fft2);
    } else {
        orc_raw2reald(next_cbuf, raw, bf->sf.bytes, bf->sf.isfloat, bf->sample_spacing, bf->sf.swap,
                      g_n_fft2);
    }
    if (postprocess != NULL) {
        postprocess(next_cbuf, g_n_fft2, pp_arg);
    }
    memcpy((uint8_t *)cbuf + (size_t)g_n_fft2 * g_realsize, next_cbuf, (size_t)g_n_fft2 * g_realsize);
}

/* fftw_convolver.c:482-518 */
void
orc_cbuf2raw(void *cbuf, void *outbuf, struct orc_buffer_format *bf, int apply_dither,
             void *dither_state, struct orc_overflow *overflow)
{
    uint8_t *raw = (uint8_t *)outbuf + bf->byte_offset;
    struct orc_dither_state *ds = NULL;
    if (apply_dither && !bf->sf.isfloat) {
        ds = dither_state;
        orc_dither_preloop(ds, g_n_fft2);
    }
    if (g_realsize == 4) {
        orc_real2rawf(raw, cbuf, bf->sf.sbytes << 3, bf->sf.bytes, bf->sf.isfloat,
                      bf->sample_spacing, bf->sf.swap, g_n_fft2, overflow, ds);
    } else {
        orc_real2rawd(raw, cbuf, bf->sf.sbytes << 3, bf->sf.bytes, bf->sf.isfloat,
                      bf->sample_spacing, bf->sf.swap, g_n_fft2, overflow, ds);
    }
}

void
orc_mixnscale(void *inputs[], void *out, double scales[], int n_bufs, int mixmode)
{
    if (g_realsize == 4) {
        orc_mixnscalef(inputs, out, scales, n_bufs, mixmode);
    } else {
        orc_mixnscaled(inputs, out, scales, n_bufs, mixmode);
    }
}

void
orc_convolve(void *in, void *coeffs, void *out)
{
    if (g_realsize == 4) {
        orc_convolvef(in, coeffs, out);
    } else {
        orc_convolved(in, coeffs, out);
    }
}

void
orc_convolve_inplace(void *cbuf, void *coeffs)
{
    /* fftw_convfuns.h:503-532: same arithmetic as convolve with out == in (each chunk reads its
       operands before writing them) */
    orc_convolve(cbuf, coeffs, cbuf);
}

void
orc_convolve_add(void *in, void *coeffs, void *out)
{
    if (g_realsize == 4) {
        orc_convolve_addf(in, coeffs, out);
    } else {
        orc_convolve_addd(in, coeffs, out);
    }
}

void
orc_dirac_convolve(void *in, void *out)
{
    if (g_realsize == 4) {
        orc_dirac_convolvef(in, out);
    } else {
        orc_dirac_convolved(in, out);
    }
}

void
orc_dirac_convolve_inplace(void *cbuf)
{
    orc_dirac_convolve(cbuf, cbuf);
}

void
orc_crossfade_inplace(void *input_cbuf, void *crossfade_cbuf, void *buffer_cbuf)
{
    if (g_realsize == 4) {
        orc_crossfade_inplacef(input_cbuf, crossfade_cbuf, buffer_cbuf);
    } else {
        orc_crossfade_inplaced(input_cbuf, crossfade_cbuf, buffer_cbuf);
    }
}

/* fftw_convolver.c:411-433: IFFT into the upper N of a 1.5 N buffer whose lower L holds the
   previous block's second half, FFT the lower N, then slide. */
void
orc_convolve_eval(void *input_cbuf, void *buffer_cbuf, void *output_cbuf)
{
    uint8_t *buf = buffer_cbuf;
    const size_t half = (size_t)g_n_fft2 * g_realsize;
    orc_freq2time(input_cbuf, buf + half);
    orc_time2freq(buf, output_cbuf);
    memcpy(buf, buf + half, half);
}

void *
orc_coeffs2cbuf(void *coeffs, int n_coeffs, double scale, void *optional_dest)
{
    if (g_realsize == 4) {
        return orc_coeffs2cbuff(coeffs, n_coeffs, scale, optional_dest);
    }
    return orc_coeffs2cbufd(coeffs, n_coeffs, scale, optional_dest);
}

/* fftw_convolver.c:575-596 */
void
orc_runtime_coeffs2cbuf(void *src, void *dest)
{
    const size_t half = (size_t)g_n_fft2 * g_realsize;
    void *tmp = orc_alloc(2 * half), *in[1];
    double scale = 1.0 / (double)g_n_fft;

    memset(dest, 0, half);
    memcpy((uint8_t *)dest + half, src, half);
    orc_time2freq(dest, tmp);
    in[0] = tmp;
    orc_mixnscale(in, dest, &scale, 1, ORC_MIXMODE_INPUT);
    free(tmp);
}

/* fftw_convolver.c:598-622 */
int
orc_verify_cbuf(void *cbufs[], int n_cbufs)
{
    int n, i;
    for (n = 0; n < n_cbufs; n++) {
        for (i = 0; i < g_n_fft; i++) {
            const double v = g_realsize == 4 ? (double)((float *)cbufs[n])[i] : ((double *)cbufs[n])[i];
            if (!isfinite(v)) {
                fprintf(stderr, "NaN or Inf value among coefficients.\n");
                return 0;
            }
        }
    }
    return 1;
}

/* ---- the small ordered-layout convolver of the sub-sample delay (fftw_convolver.c:682-782) --------------------- */

struct orc_td {
    void *plan_fwd, *plan_inv;      /* r2r plans of 2 * blocklen points */
    void *coeffs;                   /* half-complex spectrum of [0 | taps | 0], scaled by 1 / (2 blocklen) */
    int blocklen;
};

/* fftw_convolver.c:689-696 with log2_roof (log2.h:28-43): the next power of two.  (One coefficient is undefined
 * behaviour there -- log2_roof(1) returns -1 and the result is 1 << -1; here it is a block of one sample.) */
int
orc_td_block_length(int n_coeffs)
{
    int len = 1;
    if (n_coeffs < 1) {
        return -1;
    }
    while (len < n_coeffs) {
        len <<= 1;
    }
    return len;
}

/* fftw_convolver.c:698-734 */
struct orc_td *
orc_td_new(void *coeffs, int n_coeffs)
{
    const int blocklen = orc_td_block_length(n_coeffs);
    struct orc_td *td;
    int n, size;

    if (blocklen == -1) {
        return NULL;
    }
    size = blocklen << 1;
    td = orc_alloc(sizeof(*td));
    td->blocklen = blocklen;
    td->coeffs = orc_alloc((size_t)size * g_realsize);
    memset(td->coeffs, 0, (size_t)size * g_realsize);
    memcpy((uint8_t *)td->coeffs + (size_t)blocklen * g_realsize, coeffs, (size_t)n_coeffs * g_realsize);
    if (g_realsize == 4) {
        const float scale = 1.0 / (float)size;
        td->plan_fwd = fftwf_plan_r2r_1d(size, NULL, NULL, FFTW_R2HC, FFTW_MEASURE);
        td->plan_inv = fftwf_plan_r2r_1d(size, NULL, NULL, FFTW_HC2R, FFTW_MEASURE);
        fftwf_execute_r2r(td->plan_fwd, td->coeffs, td->coeffs);
        for (n = 0; n < size; n++) {
            ((float *)td->coeffs)[n] *= scale;
        }
    } else {
        const double scale = 1.0 / (double)size;
        td->plan_fwd = fftw_plan_r2r_1d(size, NULL, NULL, FFTW_R2HC, FFTW_MEASURE);
        td->plan_inv = fftw_plan_r2r_1d(size, NULL, NULL, FFTW_HC2R, FFTW_MEASURE);
        fftw_execute_r2r(td->plan_fwd, td->coeffs, td->coeffs);
        for (n = 0; n < size; n++) {
            ((double *)td->coeffs)[n] *= scale;
        }
    }
    return td;
}

/* fftw_convolver.c:736-781: forward transform, product on the half-complex order (bin k real part at k, imaginary
 * part at size - k; DC and Nyquist real), inverse transform, all in place. */
void
orc_td_convolve(struct orc_td *td, void *overlap_block)
{
    const int size = td->blocklen << 1, half = td->blocklen;
    int n;

    if (g_realsize == 4) {
        float *b = overlap_block, *c = td->coeffs;
        fftwf_execute_r2r(td->plan_fwd, b, b);
        b[0] *= c[0];
        for (n = 1; n < half; n++) {
            const float re = b[n], im = b[size - n];
            b[n] = re * c[n] - im * c[size - n];
            b[size - n] = re * c[size - n] + im * c[n];
        }
        b[half] *= c[half];
        fftwf_execute_r2r(td->plan_inv, b, b);
    } else {
        double *b = overlap_block, *c = td->coeffs;
        fftw_execute_r2r(td->plan_fwd, b, b);
        b[0] *= c[0];
        for (n = 1; n < half; n++) {
            const double re = b[n], im = b[size - n];
            b[n] = re * c[n] - im * c[size - n];
            b[size - n] = re * c[size - n] + im * c[n];
        }
        b[half] *= c[half];
        fftw_execute_r2r(td->plan_inv, b, b);
    }
}

void
orc_td_free(struct orc_td *td)
{
    if (td == NULL) {
        return;
    }
    if (g_realsize == 4) {
        fftwf_destroy_plan(td->plan_fwd);
        fftwf_destroy_plan(td->plan_inv);
    } else {
        fftw_destroy_plan(td->plan_fwd);
        fftw_destroy_plan(td->plan_inv);
    }
    free(td->coeffs);
    free(td);
}

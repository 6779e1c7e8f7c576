/*
 * oracle/bf_blockdriver.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Replays the per-block sequence of the reference's filter_process() (/root/reference/bfrun.c:1420-2083)
 * through the convolver.h API, for a filter graph given as a struct bfcuda_config (include/bfcuda.h).
 * It is compiled twice from this one source:
 *
 *   -DDRV_REFERENCE : against the reference's own objects (fftw_convolver.c, convolver_xmm.c, ...
 *                     compiled where they lie under /root/reference) -> oracle/_ref/libbfref.so,
 *                     entry points bfref_*.  This is "the reference itself run here".
 *   (default)       : against the oracle restatement (bf_oracle.c) -> oracle/libbforacle.so,
 *                     entry points bfo_*.
 *
 * The host orchestration of bfrun.c (fork, pipes, SysV shm, dai) is replaced by plain threads with
 * the same work split: filters are grouped the way load_balance_filters() groups them
 * (bfconf.c:2227-2318), input/output channels are dealt out in contiguous chunks for the forward and
 * inverse transforms (bfrun.c:2316-2328) and the two synch_filter_processes() barriers
 * (bfrun.c:1563, 1873) become pthread barriers.  powersave, sub-sample delay, mute, virtual->physical
 * channel mixing and dither are off (SURVEY.md 8(d) benchmark settings).
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <pthread.h>

#include "../include/bfcuda.h"

#ifdef DRV_REFERENCE
#include "convolver.h"      /* the reference's header, found via -I/root/reference */
#include "bfconf.h"
#include "dither.h"
#define CV(name) convolver_##name
#define DRV(name) bfref_##name
#define CV_MIXMODE_INPUT CONVOLVER_MIXMODE_INPUT
#define CV_MIXMODE_OUTPUT CONVOLVER_MIXMODE_OUTPUT
typedef struct buffer_format cv_buffer_format;
typedef struct bfoverflow cv_overflow;
/* the two host globals the reference's convolver reads */
static struct bfconf bfconf_storage;
struct bfconf *bfconf = &bfconf_storage;
static volatile int g_failed;
void
bf_exit(int status)
{
    fprintf(stderr, "bfref: bf_exit(%d)\n", status);
    g_failed = 1;
}
#else
#include "bf_oracle.h"
#define CV(name) orc_##name
#define DRV(name) bfo_##name
#define CV_MIXMODE_INPUT ORC_MIXMODE_INPUT
#define CV_MIXMODE_OUTPUT ORC_MIXMODE_OUTPUT
typedef struct orc_buffer_format cv_buffer_format;
typedef struct orc_overflow cv_overflow;
static volatile int g_failed;
static void
drv_fail_handler(int code)
{
    (void)code;
    g_failed = 1;
}
#endif

#define IN BFCUDA_IN
#define OUT BFCUDA_OUT

struct drv_filter {
    int crossfade;
    int n_ch[2];
    int *ch[2];
    double *scale[2];
    int n_fin;
    int *fin;
    double *fscale;
    int coeff, delayblocks;         /* current control (icomm->fctrl) */
    int prevcoeff, procblocks;
    int thread;
    void **cbuf;                    /* [P] delay line, or [1] aliasing ocbuf when P == 1 */
    void *ocbuf;
    void *evalbuf;
};

struct drv {
    int L, P, N, rs, cbufsize;
    int n_ch[2], n_bytes[2];
    cv_buffer_format *bf[2];
    int n_filters, n_coeffs;
    struct drv_filter *filters;
    int *coeff_n_blocks;
    void ***coeffs;                 /* [coeff][block] in the convolver's processed layout */
    void **input_freqcbuf, **output_freqcbuf;
    void *(*input_timecbuf)[2];
    cv_overflow *overflow;
    void **dither_state;            /* per output: struct dither_state * (bfconf->dither_state), NULL = no dither */
    unsigned int blockcounter;
    int curbuf;
    int powersave;                  /* bfconf->powersave */
    double analog_powersave;        /* bfconf->analog_powersave, linear */
    double *in_scale;               /* sf.scale of every input */
    /* virtual -> physical outputs, mute, sub-sample delay (bfrun.c:1503-1526, 1918-2002) */
    int *out_phys;                  /* bfconf->virt2phys[OUT], NULL = 1:1 */
    unsigned char *muted[2];        /* icomm->ismuted */
    void **sd_filter[2];            /* per channel: td convolver of its sub-sample delay (subdelay_filter[subdelay]) */
    void **sd_rest[2];              /* per channel: input_sd_rest / output_sd_rest */
    int *sd_blocksize[2];
    void *mixbuf;
    int n_threads;
    /* per-thread scratch */
    void **crossfadebuf0, **crossfadebuf1, **timebuf;
    /* outputs: which thread mixes it (the thread owning its filters), which does the inverse FFT */
    int *out_mix_thread;
    int *in_fft_thread, *out_fft_thread;
    void **debug_time;              /* copy of the first L time-domain reals per output */
    /* run state */
    pthread_barrier_t barrier;
    const uint8_t *run_in;
    uint8_t *run_out;
    size_t in_stride, out_stride;
    int run_blocks;
};

static void *
drv_alloc(size_t size)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, size < 64 ? 64 : size) != 0) {
        fprintf(stderr, "blockdriver: out of memory\n");
        abort();
    }
    memset(p, 0, size);
    return p;
}

static void
balance(struct drv *d)
{
    /* bfconf.c:2227-2318: connected filters and filters sharing an output form a group; groups are
       dealt round-robin over the workers */
    int *group = malloc(sizeof(int) * d->n_filters);
    int n, i, j, k, changed, n_groups = 0;

    for (n = 0; n < d->n_filters; n++) {
        group[n] = -1;
    }
    for (n = 0; n < d->n_filters; n++) {
        if (group[n] != -1) {
            continue;
        }
        group[n] = n_groups;
        do {
            changed = 0;
            for (i = 0; i < d->n_filters; i++) {
                for (j = 0; j < d->n_filters; j++) {
                    int linked = 0;
                    if ((group[i] == n_groups) == (group[j] == n_groups)) {
                        continue;
                    }
                    for (k = 0; k < d->filters[i].n_fin; k++) {
                        linked |= d->filters[i].fin[k] == j;
                    }
                    for (k = 0; k < d->filters[j].n_fin; k++) {
                        linked |= d->filters[j].fin[k] == i;
                    }
                    for (k = 0; k < d->filters[i].n_ch[OUT] && !linked; k++) {
                        int m;
                        for (m = 0; m < d->filters[j].n_ch[OUT]; m++) {
                            linked |= d->filters[i].ch[OUT][k] == d->filters[j].ch[OUT][m];
                        }
                    }
                    if (linked) {
                        group[i] = group[j] = n_groups;
                        changed = 1;
                    }
                }
            }
        } while (changed);
        n_groups++;
    }
    for (n = 0; n < d->n_filters; n++) {
        d->filters[n].thread = group[n] % d->n_threads;
    }
    free(group);
    for (n = 0; n < d->n_ch[OUT]; n++) {
        d->out_mix_thread[n] = 0;
        for (i = 0; i < d->n_filters; i++) {
            for (j = 0; j < d->filters[i].n_ch[OUT]; j++) {
                if (d->filters[i].ch[OUT][j] == n) {
                    d->out_mix_thread[n] = d->filters[i].thread;
                }
            }
        }
    }
    /* bfrun.c:2316-2328: contiguous chunks, the last worker takes the remainder */
    for (j = 0; j < 2; j++) {
        int per = d->n_ch[j] / d->n_threads;
        for (n = 0; n < d->n_ch[j]; n++) {
            int t = per > 0 ? n / per : d->n_threads - 1;
            if (t > d->n_threads - 1) {
                t = d->n_threads - 1;
            }
            (j == IN ? d->in_fft_thread : d->out_fft_thread)[n] = t;
        }
    }
}

int
DRV(create)(const struct bfcuda_config *c, int n_threads, struct drv **out)
{
    struct drv *d;
    int n, i, io;

#ifdef DRV_REFERENCE
    bfconf->quiet = 1;
    bfconf->safety_limit = c->safety_limit;
    if (!CV(init)("/dev/null", c->filter_length, c->realsize)) {
        return -1;
    }
#else
    orc_set_fail_handler(drv_fail_handler);
    orc_set_safety_limit(c->safety_limit);
    if (!CV(init)(NULL, c->filter_length, c->realsize)) {
        return -1;
    }
#endif
    g_failed = 0;
    d = calloc(1, sizeof(*d));
    d->L = c->filter_length;
    d->P = c->n_blocks;
    d->N = 2 * d->L;
    d->rs = c->realsize;
    d->cbufsize = CV(cbufsize)();
    d->n_threads = n_threads < 1 ? 1 : n_threads;
    d->n_filters = c->n_filters;
    d->n_coeffs = c->n_coeffs;
    for (io = 0; io < 2; io++) {
        d->n_ch[io] = c->n_channels[io];
        d->n_bytes[io] = c->n_bytes[io];
        d->bf[io] = calloc(d->n_ch[io], sizeof(cv_buffer_format));
        for (n = 0; n < d->n_ch[io]; n++) {
            const struct bfcuda_buffer_format *s = &c->formats[io][n];
            d->bf[io][n].sf.isfloat = s->sf.isfloat;
            d->bf[io][n].sf.swap = s->sf.swap;
            d->bf[io][n].sf.bytes = s->sf.bytes;
            d->bf[io][n].sf.sbytes = s->sf.sbytes;
            d->bf[io][n].sf.scale = s->sf.scale;
            d->bf[io][n].sf.format = s->sf.format;
            d->bf[io][n].sample_spacing = s->sample_spacing;
            d->bf[io][n].byte_offset = s->byte_offset;
        }
    }
    d->filters = calloc(d->n_filters, sizeof(struct drv_filter));
    for (n = 0; n < d->n_filters; n++) {
        const struct bfcuda_filter *s = &c->filters[n];
        struct drv_filter *f = &d->filters[n];
        f->crossfade = s->crossfade;
        for (io = 0; io < 2; io++) {
            f->n_ch[io] = s->n_channels[io];
            f->ch[io] = malloc(sizeof(int) * (f->n_ch[io] + 1));
            f->scale[io] = malloc(sizeof(double) * (f->n_ch[io] + 1));
            for (i = 0; i < f->n_ch[io]; i++) {
                f->ch[io][i] = s->channels[io][i];
                f->scale[io][i] = s->scale[io][i];
            }
        }
        f->n_fin = s->n_filters_in;
        f->fin = malloc(sizeof(int) * (f->n_fin + 1));
        f->fscale = malloc(sizeof(double) * (f->n_fin + 1));
        for (i = 0; i < f->n_fin; i++) {
            f->fin[i] = s->filters_in[i];
            f->fscale[i] = s->fscale[i];
        }
        f->coeff = f->prevcoeff = s->coeff;     /* bfrun.c:1326 */
        f->delayblocks = s->delayblocks;
        f->ocbuf = drv_alloc(d->cbufsize);
        f->cbuf = calloc(d->P, sizeof(void *));
        if (d->P > 1) {
            for (i = 0; i < d->P; i++) {
                f->cbuf[i] = drv_alloc(d->cbufsize);
            }
        } else {
            f->cbuf[0] = f->ocbuf;              /* bfrun.c:1289-1291 */
        }
        f->evalbuf = f->n_fin > 0 ? drv_alloc(d->cbufsize + d->cbufsize / 2) : NULL;
    }
    d->coeff_n_blocks = malloc(sizeof(int) * (d->n_coeffs + 1));
    d->coeffs = calloc(d->n_coeffs + 1, sizeof(void **));
    for (n = 0; n < d->n_coeffs; n++) {
        d->coeff_n_blocks[n] = c->coeff_n_blocks[n];
        d->coeffs[n] = calloc(d->coeff_n_blocks[n], sizeof(void *));
        for (i = 0; i < d->coeff_n_blocks[n]; i++) {
            d->coeffs[n][i] = drv_alloc(d->cbufsize);
        }
    }
    for (io = 0; io < 2; io++) {
        d->muted[io] = calloc(d->n_ch[io] + 1, 1);
        d->sd_filter[io] = calloc(d->n_ch[io] + 1, sizeof(void *));
        d->sd_rest[io] = calloc(d->n_ch[io] + 1, sizeof(void *));
        d->sd_blocksize[io] = calloc(d->n_ch[io] + 1, sizeof(int));
    }
    d->out_phys = NULL;
    if (c->out_physical != NULL) {
        d->out_phys = malloc(sizeof(int) * (d->n_ch[OUT] + 1));
        memcpy(d->out_phys, c->out_physical, sizeof(int) * d->n_ch[OUT]);
    }
    d->mixbuf = drv_alloc((size_t)d->L * d->rs);
    d->powersave = c->powersave;
    d->analog_powersave = (c->analog_powersave <= 0.0) ? 1.0 : c->analog_powersave;
    d->in_scale = calloc(d->n_ch[IN] + 1, sizeof(double));
    for (n = 0; n < d->n_ch[IN]; n++) {
        d->in_scale[n] = c->formats[IN][n].sf.scale;
    }
    d->input_freqcbuf = calloc(d->n_ch[IN], sizeof(void *));
    d->input_timecbuf = calloc(d->n_ch[IN], sizeof(void *[2]));
    for (n = 0; n < d->n_ch[IN]; n++) {
        d->input_freqcbuf[n] = drv_alloc(d->cbufsize);
        d->input_timecbuf[n][0] = drv_alloc(d->cbufsize);
        d->input_timecbuf[n][1] = drv_alloc(d->cbufsize);
    }
    d->output_freqcbuf = calloc(d->n_ch[OUT], sizeof(void *));
    d->debug_time = calloc(d->n_ch[OUT], sizeof(void *));
    d->overflow = calloc(d->n_ch[OUT], sizeof(cv_overflow));
    d->dither_state = calloc(d->n_ch[OUT] + 1, sizeof(void *));
    if (c->apply_dither != NULL) {
        /* bfconf.c:3173-3238: which outputs are dithered, then one table for all of them */
        int j = 0, rate = c->sampling_rate > 0 ? c->sampling_rate : 44100;
        int *which = calloc(d->n_ch[OUT] + 1, sizeof(int));
        for (n = 0; n < d->n_ch[OUT]; n++) {
            const cv_buffer_format *b = &d->bf[OUT][n];
            if (!c->apply_dither[n] || b->sf.isfloat || (d->rs == 4 && b->sf.sbytes > 2) || b->sf.sbytes >= 4) {
                continue;
            }
            which[j++] = n;
        }
        if (j > 0) {
#ifdef DRV_REFERENCE
            struct dither_state **st = calloc(j, sizeof(*st));
            if (!dither_init(j, rate, d->rs, c->max_dither_table_size, d->L, st)) {
                return -1;
            }
            for (n = 0; n < j; n++) {
                d->dither_state[which[n]] = st[n];
            }
            free(st);
#else
            struct orc_dither_state *st = calloc(j, sizeof(*st));
            if (!orc_dither_init(j, rate, d->rs, c->max_dither_table_size, d->L, st)) {
                return -1;
            }
            for (n = 0; n < j; n++) {
                d->dither_state[which[n]] = &st[n];
            }
#endif
        }
        free(which);
    }
    for (n = 0; n < d->n_ch[OUT]; n++) {
        d->output_freqcbuf[n] = drv_alloc(d->cbufsize);
        d->debug_time[n] = drv_alloc(d->cbufsize);
        /* bfrun.c:2264-2279 */
        if (d->bf[OUT][n].sf.isfloat) {
            d->overflow[n].max = 1.0;
        } else {
            d->overflow[n].max = (double)((uint64_t)1 << ((d->bf[OUT][n].sf.sbytes << 3) - 1)) - 1;
        }
    }
    d->crossfadebuf0 = calloc(d->n_threads, sizeof(void *));
    d->crossfadebuf1 = calloc(d->n_threads, sizeof(void *));
    d->timebuf = calloc(d->n_threads, sizeof(void *));
    for (n = 0; n < d->n_threads; n++) {
        d->crossfadebuf0[n] = drv_alloc(d->cbufsize);
        d->crossfadebuf1[n] = drv_alloc(d->cbufsize);
        d->timebuf[n] = drv_alloc(d->cbufsize);
    }
    d->out_mix_thread = calloc(d->n_ch[OUT] + 1, sizeof(int));
    d->in_fft_thread = calloc(d->n_ch[IN] + 1, sizeof(int));
    d->out_fft_thread = calloc(d->n_ch[OUT] + 1, sizeof(int));
    balance(d);
    *out = d;
    return 0;
}

void
DRV(destroy)(struct drv *d)
{
    int n, i;
    if (d == NULL) {
        return;
    }
    for (n = 0; n < d->n_filters; n++) {
        struct drv_filter *f = &d->filters[n];
        if (d->P > 1) {
            for (i = 0; i < d->P; i++) {
                free(f->cbuf[i]);
            }
        }
        free(f->cbuf);
        free(f->ocbuf);
        free(f->evalbuf);
        free(f->ch[0]); free(f->ch[1]); free(f->scale[0]); free(f->scale[1]);
        free(f->fin); free(f->fscale);
    }
    for (n = 0; n < d->n_coeffs; n++) {
        for (i = 0; i < d->coeff_n_blocks[n]; i++) {
            free(d->coeffs[n][i]);
        }
        free(d->coeffs[n]);
    }
    for (n = 0; n < d->n_ch[IN]; n++) {
        free(d->input_freqcbuf[n]);
        free(d->input_timecbuf[n][0]);
        free(d->input_timecbuf[n][1]);
    }
    for (n = 0; n < d->n_ch[OUT]; n++) {
        free(d->output_freqcbuf[n]);
        free(d->debug_time[n]);
    }
    for (n = 0; n < d->n_threads; n++) {
        free(d->crossfadebuf0[n]); free(d->crossfadebuf1[n]); free(d->timebuf[n]);
    }
    free(d->filters); free(d->coeffs); free(d->coeff_n_blocks);
    free(d->input_freqcbuf); free(d->input_timecbuf); free(d->output_freqcbuf);
    free(d->debug_time); free(d->overflow); free(d->bf[0]); free(d->bf[1]);
    free(d->crossfadebuf0); free(d->crossfadebuf1); free(d->timebuf);
    free(d->out_mix_thread); free(d->in_fft_thread); free(d->out_fft_thread);
    free(d);
}

/* bfconf.c:1992-2019: block n of the taps (zero-extended) through convolver_coeffs2cbuf */
int
DRV(coeff_from_taps)(struct drv *d, int coeff, const void *taps, int n_taps, double scale)
{
    void *zbuf = drv_alloc((size_t)d->L * d->rs);
    int n, rc = 0;

    if (coeff < 0 || coeff >= d->n_coeffs) {
        free(zbuf);
        return -1;
    }
    for (n = 0; n < d->coeff_n_blocks[coeff]; n++) {
        const uint8_t *src = (const uint8_t *)taps + (size_t)n * d->L * d->rs;
        void *r;
        if ((long)n * d->L > n_taps) {
            r = CV(coeffs2cbuf)(zbuf, d->L, scale, d->coeffs[coeff][n]);
        } else if ((long)(n + 1) * d->L > n_taps) {
            r = CV(coeffs2cbuf)((void *)src, n_taps - n * d->L, scale, d->coeffs[coeff][n]);
        } else {
            r = CV(coeffs2cbuf)((void *)src, d->L, scale, d->coeffs[coeff][n]);
        }
        if (r == NULL) {
            rc = -5;
            break;
        }
    }
    free(zbuf);
    return rc;
}

int
DRV(coeff_set_block)(struct drv *d, int coeff, int block, const void *cbuf)
{
    if (coeff < 0 || coeff >= d->n_coeffs || block < 0 || block >= d->coeff_n_blocks[coeff]) {
        return -1;
    }
    memcpy(d->coeffs[coeff][block], cbuf, d->cbufsize);
    return 0;
}

int
DRV(coeff_get_block)(struct drv *d, int coeff, int block, void *cbuf)
{
    if (coeff < 0 || coeff >= d->n_coeffs || block < 0 || block >= d->coeff_n_blocks[coeff]) {
        return -1;
    }
    memcpy(cbuf, d->coeffs[coeff][block], d->cbufsize);
    return 0;
}

int
DRV(coeff_runtime_block)(struct drv *d, int coeff, int block, const void *taps_L)
{
    if (coeff < 0 || coeff >= d->n_coeffs || block < 0 || block >= d->coeff_n_blocks[coeff]) {
        return -1;
    }
    CV(runtime_coeffs2cbuf)((void *)taps_L, d->coeffs[coeff][block]);
    return 0;
}

int
DRV(set_control)(struct drv *d, int filter, const struct bfcuda_filter_control *c)
{
    struct drv_filter *f;
    int i, io;
    if (filter < 0 || filter >= d->n_filters || c->coeff >= d->n_coeffs) {
        return -1;
    }
    f = &d->filters[filter];
    f->coeff = c->coeff;
    f->delayblocks = c->delayblocks;
    for (io = 0; io < 2; io++) {
        if (c->scale[io] != NULL) {
            for (i = 0; i < f->n_ch[io]; i++) {
                f->scale[io][i] = c->scale[io][i];
            }
        }
    }
    if (c->fscale != NULL) {
        for (i = 0; i < f->n_fin; i++) {
            f->fscale[i] = c->fscale[i];
        }
    }
    return 0;
}

int
DRV(set_mute)(struct drv *d, int io, int channel, int muted)
{
    if (io < 0 || io > 1 || channel < 0 || channel >= d->n_ch[io]) {
        return -1;
    }
    d->muted[io][channel] = muted ? 1 : 0;
    return 0;
}

/* taps: the windowed sinc of the channel's delay step (what delay_subsample_init hands to convolver_td_new,
   delay.c:486-499); NULL switches the channel's sub-sample delay off */
int
DRV(set_subdelay)(struct drv *d, int io, int channel, void *taps, int n_taps)
{
    if (io < 0 || io > 1 || channel < 0 || channel >= d->n_ch[io]) {
        return -1;
    }
    if (taps == NULL) {
        d->sd_filter[io][channel] = NULL;
        return 0;
    }
    d->sd_blocksize[io][channel] = CV(td_block_length)(n_taps);
    if (d->L % d->sd_blocksize[io][channel] != 0) {
        return -1;      /* "Incompatible fragment/filter sizes", delay.c:465-469 */
    }
    d->sd_filter[io][channel] = CV(td_new)(taps, n_taps);
    if (d->sd_rest[io][channel] == NULL) {
        d->sd_rest[io][channel] = drv_alloc((size_t)d->sd_blocksize[io][channel] * d->rs);
    }
    return 0;
}

int
DRV(get_overflow)(struct drv *d, int out_channel, struct bfcuda_overflow *of)
{
    if (out_channel < 0 || out_channel >= d->n_ch[OUT]) {
        return -1;
    }
    of->n_overflows = d->overflow[out_channel].n_overflows;
    of->intlargest = d->overflow[out_channel].intlargest;
    of->largest = d->overflow[out_channel].largest;
    of->max = d->overflow[out_channel].max;
    return 0;
}

int
DRV(debug_read)(struct drv *d, int what, int index, int slot, void *dst)
{
    switch (what) {
    case BFCUDA_DBG_INPUT_SPECTRUM:
        memcpy(dst, d->input_freqcbuf[index], d->cbufsize);
        return 0;
    case BFCUDA_DBG_DELAYLINE:
        memcpy(dst, d->filters[index].cbuf[slot], d->cbufsize);
        return 0;
    case BFCUDA_DBG_FILTER_OUTPUT:
        memcpy(dst, d->filters[index].ocbuf, d->cbufsize);
        return 0;
    case BFCUDA_DBG_OUTPUT_TIME:
        memcpy(dst, d->debug_time[index], (size_t)d->L * d->rs);
        return 0;
    default:
        return -1;
    }
}

/* ---- one block, the part of worker `t` ------------------------------------------------------- */

/* test_silent(), bfrun.c:722-772: with analog_powersave >= 1.0 only a frame of exact zero BYTES is silent
 * (memiszero, bfrun.c:700-720); below that, a frame whose scaled peak is under the level is silent and is "made truly
 * zero".  The frame is the whole cbuf: previous block and this block. */
static int
drv_test_silent(void *buf, int size, int realsize, double analog_powersave, double scale)
{
    int n, count;
    double dmax = 0;
    if (analog_powersave >= 1.0) {
        const unsigned char *b = buf;
        for (n = 0; n < size; n++) {
            if (b[n] != 0) {
                return 0;
            }
        }
        return 1;
    }
    if (realsize == 4) {
        float fmax = 0;
        count = size >> 2;
        for (n = 0; n < count; n++) {
            const float v = ((float *)buf)[n];
            if (v < 0) {
                if (-v > fmax) fmax = -v;
            } else if (v > fmax) {
                fmax = v;
            }
        }
        dmax = fmax;
    } else {
        count = size >> 3;
        for (n = 0; n < count; n++) {
            const double v = ((double *)buf)[n];
            if (v < 0) {
                if (-v > dmax) dmax = -v;
            } else if (v > dmax) {
                dmax = v;
            }
        }
    }
    if (scale * dmax >= analog_powersave) {
        return 0;
    }
    memset(buf, 0, size);
    return 1;
}

/* delay_subsample_update(), delay.c:415-442, on the td convolver of the channel's current delay step */
struct drv_sd_params {
    struct drv *d;
    int io, ch;
};

static void
drv_subsample_update(struct drv *d, int io, int ch, void *buf)
{
    const int bs = d->sd_blocksize[io][ch];
    const size_t blocksize = (size_t)bs * d->rs;
    unsigned char *cbuffer, *rest = d->sd_rest[io][ch];
    size_t i;
    if (d->sd_filter[io][ch] == NULL) {
        return;
    }
    cbuffer = malloc(blocksize << 1);
    for (i = 0; i < (size_t)d->L * d->rs; i += blocksize) {
        memcpy(cbuffer, rest, blocksize);
        memcpy(cbuffer + blocksize, (unsigned char *)buf + i, blocksize);
        memcpy(rest, cbuffer + blocksize, blocksize);
        CV(td_convolve)(d->sd_filter[io][ch], cbuffer);
        memcpy((unsigned char *)buf + i, cbuffer, blocksize);
    }
    free(cbuffer);
}

static void
drv_apply_subdelay(void *realbuf, int n_samples, void *arg)     /* bfrun.c:971-984 */
{
    struct drv_sd_params *p = arg;
    (void)n_samples;
    drv_subsample_update(p->d, p->io, p->ch, realbuf);
}

static void
forward_part(struct drv *d, int t, const uint8_t *inbuf)
{
    int n;
    /* bfrun.c:1494-1560 */
    for (n = 0; n < d->n_ch[IN]; n++) {
        if (d->in_fft_thread[n] != t) {
            continue;
        }
        {
            struct drv_sd_params sp;
            sp.d = d;
            sp.io = IN;
            sp.ch = n;
            if (d->muted[IN][n]) {
                /* bfrun.c:1523-1525: a muted virtual input is converted from a block of zeros */
                const cv_buffer_format *bf = &d->bf[IN][n];
                const size_t span = (size_t)bf->byte_offset + ((size_t)(d->L - 1) * bf->sample_spacing + 1) * bf->sf.bytes;
                void *zeros = calloc(1, span);
                CV(raw2cbuf)(zeros, d->input_timecbuf[n][d->curbuf], d->input_timecbuf[n][!d->curbuf],
                             &d->bf[IN][n], drv_apply_subdelay, &sp);
                free(zeros);
            } else {
                CV(raw2cbuf)((void *)inbuf, d->input_timecbuf[n][d->curbuf], d->input_timecbuf[n][!d->curbuf],
                             &d->bf[IN][n], drv_apply_subdelay, &sp);
            }
        }
        /* bfrun.c:1541-1552.  What the reference then skips downstream (mixing and multiplying zero blocks,
           bfrun.c:1613-1700, 1737-1754) leaves every result as it is -- zeros times coefficients add nothing -- so
           the replay keeps computing them. */
        if (d->powersave && drv_test_silent(d->input_timecbuf[n][d->curbuf], d->cbufsize, d->rs, d->analog_powersave,
                                            d->in_scale[n])) {
            memset(d->input_freqcbuf[n], 0, d->cbufsize);
        } else {
            CV(time2freq)(d->input_timecbuf[n][d->curbuf], d->input_freqcbuf[n]);
        }
    }
}

static void
filter_part(struct drv *d, int t)
{
    const int P = d->P;
    const unsigned int bc = d->blockcounter;
    void *static_evalbuf = d->crossfadebuf0[t];     /* bfrun.c:1251-1258: shares storage */
    void *xf0 = d->crossfadebuf0[t], *xf1 = d->crossfadebuf1[t];
    int n, i, j;

    for (n = 0; n < d->n_filters; n++) {
        struct drv_filter *f = &d->filters[n];
        void *mixin[BFCUDA_MAXCHANNELS + 1];
        double scales[BFCUDA_MAXCHANNELS + BFCUDA_MAXFILTERS + 1];
        int coeff, delay, cblocks, prevcblocks, curblock, nin;

        if (f->thread != t) {
            continue;
        }
        if (f->procblocks < P) {
            f->procblocks++;        /* bfrun.c:1567-1571 */
        }
        coeff = f->coeff;
        delay = f->delayblocks;
        if (delay < 0) {
            delay = 0;
        } else if (delay > P - 1) {
            delay = P - 1;
        }
        /* bfrun.c:1585-1598 */
        cblocks = (coeff < 0 || d->coeff_n_blocks[coeff] > P - delay) ? P - delay
                                                                     : d->coeff_n_blocks[coeff];
        prevcblocks = (f->prevcoeff < 0 || d->coeff_n_blocks[f->prevcoeff] > P - delay)
                          ? P - delay : d->coeff_n_blocks[f->prevcoeff];
        curblock = (int)((bc + (unsigned int)delay) % (unsigned int)P);

        /* bfrun.c:1603-1681: input mix into the delay-line slot */
        nin = f->n_ch[IN];
        for (i = 0; i < nin; i++) {
            scales[i] = f->scale[IN][i] * d->bf[IN][f->ch[IN][i]].sf.scale;
            mixin[i] = d->input_freqcbuf[f->ch[IN][i]];
        }
        if (f->n_fin > 0) {
            void *fbufs[BFCUDA_MAXFILTERS];
            for (i = 0; i < f->n_fin; i++) {
                fbufs[i] = d->filters[f->fin[i]].ocbuf;
            }
            CV(mixnscale)(fbufs, static_evalbuf, f->fscale, f->n_fin, CV_MIXMODE_OUTPUT);
            CV(convolve_eval)(static_evalbuf, f->evalbuf, static_evalbuf);
            scales[nin] = 1.0;
            mixin[nin] = static_evalbuf;
            nin++;
        }
        CV(mixnscale)(mixin, f->cbuf[curblock], scales, nin, CV_MIXMODE_INPUT);

        /* bfrun.c:1687-1837: convolve */
        curblock = (int)(bc % (unsigned int)P);
        {
            const int xfade = f->crossfade && f->prevcoeff != coeff;
            if (P == 1) {
                if (xfade) {
                    if (f->prevcoeff < 0) {
                        CV(dirac_convolve)(f->cbuf[0], xf0);
                    } else {
                        CV(convolve)(f->cbuf[0], d->coeffs[f->prevcoeff][0], xf0);
                    }
                }
                if (coeff >= 0) {
                    CV(convolve_inplace)(f->cbuf[0], d->coeffs[coeff][0]);
                } else {
                    CV(dirac_convolve_inplace)(f->cbuf[0]);
                }
                if (xfade) {
                    CV(crossfade_inplace)(f->cbuf[0], xf0, xf1);
                }
            } else {
                if (xfade) {
                    if (f->prevcoeff < 0) {
                        CV(dirac_convolve)(f->cbuf[curblock], xf0);
                    } else {
                        CV(convolve)(f->cbuf[curblock], d->coeffs[f->prevcoeff][0], xf0);
                    }
                }
                if (coeff >= 0) {
                    CV(convolve)(f->cbuf[curblock], d->coeffs[coeff][0], f->ocbuf);
                    for (i = 1; i < cblocks && i < f->procblocks; i++) {
                        j = (int)((bc - (unsigned int)i) % (unsigned int)P);
                        CV(convolve_add)(f->cbuf[j], d->coeffs[coeff][i], f->ocbuf);
                    }
                } else {
                    CV(dirac_convolve)(f->cbuf[curblock], f->ocbuf);
                }
                if (xfade) {
                    if (f->prevcoeff >= 0) {
                        for (i = 1; i < prevcblocks && i < f->procblocks; i++) {
                            j = (int)((bc - (unsigned int)i) % (unsigned int)P);
                            CV(convolve_add)(f->cbuf[j], d->coeffs[f->prevcoeff][i], xf0);
                        }
                    }
                    CV(crossfade_inplace)(f->ocbuf, xf0, xf1);
                }
            }
        }
        f->prevcoeff = coeff;       /* bfrun.c:1838 */
    }

    /* bfrun.c:1847-1868: output mix, in filter order */
    for (n = 0; n < d->n_ch[OUT]; n++) {
        void *bufs[BFCUDA_MAXFILTERS];
        double scales[BFCUDA_MAXFILTERS];
        int cnt = 0;
        if (d->out_mix_thread[n] != t) {
            continue;
        }
        for (i = 0; i < d->n_filters; i++) {
            for (j = 0; j < d->filters[i].n_ch[OUT]; j++) {
                if (d->filters[i].ch[OUT][j] == n) {
                    bufs[cnt] = d->filters[i].ocbuf;
                    scales[cnt] = d->filters[i].scale[OUT][j] / d->bf[OUT][n].sf.scale;
                    cnt++;
                    break;
                }
            }
        }
        if (cnt > 0) {
            CV(mixnscale)(bufs, d->output_freqcbuf[n], scales, cnt, CV_MIXMODE_OUTPUT);
        }
    }
}

static void
inverse_part(struct drv *d, int t, uint8_t *outbuf)
{
    int n;
    if (d->out_phys != NULL) {
        /* bfrun.c:1937-2002: the virtual outputs of one physical channel are added in the time domain, in channel
           order, muted ones left out, and quantised once; all of it on one worker (the reference keeps the outputs of a
           physical channel in one process as well) */
        int i, k, filled;
        if (t != 0) {
            return;
        }
        for (n = 0; n < d->n_ch[OUT]; n++) {
            int first = 1, last = n;
            for (k = 0; k < d->n_ch[OUT]; k++) {
                if (d->out_phys[k] == d->out_phys[n]) {
                    if (k < n) first = 0;
                    last = k;
                }
            }
            if (!first) {
                continue;       /* handled with the first member of its group */
            }
            filled = 0;
            for (k = n; k <= last; k++) {
                if (d->out_phys[k] != d->out_phys[n]) {
                    continue;
                }
                CV(freq2time)(d->output_freqcbuf[k], d->timebuf[0]);
                drv_subsample_update(d, OUT, k, d->timebuf[0]);
                memcpy(d->debug_time[k], d->timebuf[0], (size_t)d->L * d->rs);
                if (d->muted[OUT][k]) {
                    continue;
                }
                if (!filled) {
                    memcpy(d->mixbuf, d->timebuf[0], (size_t)d->L * d->rs);
                } else if (d->rs == 4) {
                    for (i = 0; i < d->L; i++) ((float *)d->mixbuf)[i] += ((float *)d->timebuf[0])[i];
                } else {
                    for (i = 0; i < d->L; i++) ((double *)d->mixbuf)[i] += ((double *)d->timebuf[0])[i];
                }
                filled = 1;
            }
            if (!filled) {
                memset(d->mixbuf, 0, (size_t)d->L * d->rs);
            }
            CV(cbuf2raw)(d->mixbuf, outbuf, &d->bf[OUT][last], d->dither_state[last] != NULL, d->dither_state[last],
                         &d->overflow[last]);
            for (k = n; k <= last; k++) {
                if (d->out_phys[k] == d->out_phys[n]) {
                    d->overflow[k] = d->overflow[last];     /* bfrun.c:1994-1998 */
                }
            }
        }
        return;
    }
    /* bfrun.c:1877-1936 */
    for (n = 0; n < d->n_ch[OUT]; n++) {
        if (d->out_fft_thread[n] != t) {
            continue;
        }
        CV(freq2time)(d->output_freqcbuf[n], d->timebuf[t]);
        drv_subsample_update(d, OUT, n, d->timebuf[t]);     /* bfrun.c:1918-1925 */
        if (d->muted[OUT][n]) {
            memset(d->timebuf[t], 0, (size_t)d->L * d->rs);    /* (the reference mutes 1:1 outputs in dai.c) */
        }
        memcpy(d->debug_time[n], d->timebuf[t], (size_t)d->L * d->rs);
        CV(cbuf2raw)(d->timebuf[t], outbuf, &d->bf[OUT][n], d->dither_state[n] != NULL, d->dither_state[n],
                     &d->overflow[n]);     /* bfrun.c:1930-1935 */
    }
}

static void *
worker(void *arg)
{
    struct drv *d = ((void **)arg)[0];
    const int t = (int)(intptr_t)((void **)arg)[1];
    int b;

    for (b = 0; b < d->run_blocks; b++) {
        forward_part(d, t, d->run_in + (size_t)b * d->in_stride);
        if (d->n_threads > 1) {
            pthread_barrier_wait(&d->barrier);
        }
        filter_part(d, t);
        if (d->n_threads > 1) {
            pthread_barrier_wait(&d->barrier);
        }
        inverse_part(d, t, d->run_out + (size_t)b * d->out_stride);
        if (d->n_threads > 1) {
            pthread_barrier_wait(&d->barrier);
        }
        if (t == 0) {
            d->curbuf = !d->curbuf;     /* bfrun.c:2031-2034 */
            d->blockcounter++;
        }
        if (d->n_threads > 1) {
            pthread_barrier_wait(&d->barrier);
        }
    }
    return NULL;
}

/* Process n_blocks consecutive blocks; block b reads raw_in + b*in_stride, writes raw_out + b*out_stride
 * (stride 0 = reuse the same buffer).  Returns wall seconds, or a negative value on failure. */
double
DRV(run)(struct drv *d, int n_blocks, const void *raw_in, size_t in_stride, void *raw_out,
         size_t out_stride)
{
    struct timespec t0, t1;
    int t;

    d->run_in = raw_in;
    d->run_out = raw_out;
    d->in_stride = in_stride;
    d->out_stride = out_stride;
    d->run_blocks = n_blocks;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (d->n_threads == 1) {
        void *arg[2] = { d, (void *)(intptr_t)0 };
        worker(arg);
    } else {
        pthread_t th[d->n_threads];
        void *args[d->n_threads][2];
        pthread_barrier_init(&d->barrier, NULL, d->n_threads);
        for (t = 0; t < d->n_threads; t++) {
            args[t][0] = d;
            args[t][1] = (void *)(intptr_t)t;
            pthread_create(&th[t], NULL, worker, args[t]);
        }
        for (t = 0; t < d->n_threads; t++) {
            pthread_join(th[t], NULL);
        }
        pthread_barrier_destroy(&d->barrier);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (g_failed) {
        return -1.0;
    }
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

int
DRV(process_block)(struct drv *d, const void *raw_in, void *raw_out)
{
    return DRV(run)(d, 1, raw_in, 0, raw_out, 0) < 0.0 ? -1 : 0;
}

/* Thin pass-throughs so tests can drive single convolver.h calls of the same library. */
int DRV(cv_init)(int length, int realsize)
{
#ifdef DRV_REFERENCE
    bfconf->quiet = 1;
    return CV(init)("/dev/null", length, realsize);
#else
    orc_set_fail_handler(drv_fail_handler);
    return CV(init)(NULL, length, realsize);
#endif
}
void DRV(cv_set_safety_limit)(double limit)
{
#ifdef DRV_REFERENCE
    bfconf->safety_limit = limit;
#else
    orc_set_safety_limit(limit);
#endif
}
int DRV(cv_failed)(int reset)
{
    int f = g_failed;
    if (reset) {
        g_failed = 0;
    }
    return f;
}
int DRV(cv_cbufsize)(void) { return CV(cbufsize)(); }
void DRV(cv_raw2cbuf)(void *raw, void *cbuf, void *next, const struct bfcuda_buffer_format *bf)
{
    cv_buffer_format f;
    f.sf.isfloat = bf->sf.isfloat; f.sf.swap = bf->sf.swap; f.sf.bytes = bf->sf.bytes;
    f.sf.sbytes = bf->sf.sbytes; f.sf.scale = bf->sf.scale; f.sf.format = bf->sf.format;
    f.sample_spacing = bf->sample_spacing; f.byte_offset = bf->byte_offset;
    CV(raw2cbuf)(raw, cbuf, next, &f, NULL, NULL);
}
/* per-call dither: cv_dither_init builds the table and n states (dither_init, dither.c:75-139), cv_cbuf2raw_dither
 * is convolver_cbuf2raw with apply_dither = true on state `index` */
#ifdef DRV_REFERENCE
static struct dither_state *g_cv_dither[BFCUDA_MAXCHANNELS];
#else
static struct orc_dither_state g_cv_dither[BFCUDA_MAXCHANNELS];
#endif

int DRV(cv_dither_init)(int n_channels, int sample_rate, int realsize, int max_size, int max_samples_per_loop)
{
    if (n_channels < 1 || n_channels > BFCUDA_MAXCHANNELS) {
        return 0;
    }
#ifdef DRV_REFERENCE
    return dither_init(n_channels, sample_rate, realsize, max_size, max_samples_per_loop, g_cv_dither) ? 1 : 0;
#else
    return orc_dither_init(n_channels, sample_rate, realsize, max_size, max_samples_per_loop, g_cv_dither);
#endif
}

void DRV(cv_cbuf2raw_dither)(void *cbuf, void *out, const struct bfcuda_buffer_format *bf,
                             struct bfcuda_overflow *of, int index)
{
    cv_buffer_format f;
    cv_overflow o;
    f.sf.isfloat = bf->sf.isfloat; f.sf.swap = bf->sf.swap; f.sf.bytes = bf->sf.bytes;
    f.sf.sbytes = bf->sf.sbytes; f.sf.scale = bf->sf.scale; f.sf.format = bf->sf.format;
    f.sample_spacing = bf->sample_spacing; f.byte_offset = bf->byte_offset;
    o.n_overflows = of->n_overflows; o.intlargest = of->intlargest; o.largest = of->largest;
    o.max = of->max;
#ifdef DRV_REFERENCE
    CV(cbuf2raw)(cbuf, out, &f, 1, g_cv_dither[index], &o);
#else
    CV(cbuf2raw)(cbuf, out, &f, 1, &g_cv_dither[index], &o);
#endif
    of->n_overflows = o.n_overflows; of->intlargest = o.intlargest; of->largest = o.largest;
    of->max = o.max;
}

void DRV(cv_cbuf2raw)(void *cbuf, void *out, const struct bfcuda_buffer_format *bf,
                      struct bfcuda_overflow *of)
{
    cv_buffer_format f;
    cv_overflow o;
    f.sf.isfloat = bf->sf.isfloat; f.sf.swap = bf->sf.swap; f.sf.bytes = bf->sf.bytes;
    f.sf.sbytes = bf->sf.sbytes; f.sf.scale = bf->sf.scale; f.sf.format = bf->sf.format;
    f.sample_spacing = bf->sample_spacing; f.byte_offset = bf->byte_offset;
    o.n_overflows = of->n_overflows; o.intlargest = of->intlargest; o.largest = of->largest;
    o.max = of->max;
    CV(cbuf2raw)(cbuf, out, &f, 0, NULL, &o);
    of->n_overflows = o.n_overflows; of->intlargest = o.intlargest; of->largest = o.largest;
    of->max = o.max;
}
void DRV(cv_time2freq)(void *in, void *out) { CV(time2freq)(in, out); }
void DRV(cv_freq2time)(void *in, void *out) { CV(freq2time)(in, out); }
void DRV(cv_mixnscale)(void *in[], void *out, double scales[], int n, int mode)
{
    CV(mixnscale)(in, out, scales, n, mode);
}
void DRV(cv_convolve)(void *in, void *c, void *out) { CV(convolve)(in, c, out); }
void DRV(cv_convolve_add)(void *in, void *c, void *out) { CV(convolve_add)(in, c, out); }
void DRV(cv_convolve_inplace)(void *b, void *c) { CV(convolve_inplace)(b, c); }
void DRV(cv_dirac_convolve)(void *in, void *out) { CV(dirac_convolve)(in, out); }
void DRV(cv_dirac_convolve_inplace)(void *b) { CV(dirac_convolve_inplace)(b); }
void DRV(cv_crossfade_inplace)(void *in, void *xf, void *buf) { CV(crossfade_inplace)(in, xf, buf); }
void DRV(cv_convolve_eval)(void *in, void *buf, void *out) { CV(convolve_eval)(in, buf, out); }
int DRV(cv_coeffs2cbuf)(void *taps, int n, double scale, void *dest)
{
    return CV(coeffs2cbuf)(taps, n, scale, dest) != NULL;
}
void DRV(cv_runtime_coeffs2cbuf)(void *src, void *dest) { CV(runtime_coeffs2cbuf)(src, dest); }

/* convolver_td_* (convolver.h:134-146): the small ordered-layout convolver of the sub-sample delay */
int DRV(cv_td_block_length)(int n_coeffs) { return CV(td_block_length)(n_coeffs); }
void *DRV(cv_td_new)(void *coeffs, int n_coeffs) { return CV(td_new)(coeffs, n_coeffs); }
void DRV(cv_td_convolve)(void *tdc, void *overlap_block) { CV(td_convolve)(tdc, overlap_block); }

/* dither table and per-channel state of the cv_* surface, for tests that hand the SAME table and state to another
 * implementation: the table (dither_randtab / dither_randtab_size, dither.c:22-24) and a channel's struct dither_state */
const int8_t *DRV(cv_dither_table)(int *size)
{
#ifdef DRV_REFERENCE
    *size = dither_randtab_size;
    return dither_randtab;
#else
    return orc_dither_table(size);
#endif
}
void *DRV(cv_dither_state)(int index)
{
#ifdef DRV_REFERENCE
    return g_cv_dither[index];
#else
    return &g_cv_dither[index];
#endif
}

/*
 * bfcuda_convolver.h -- the reference's per-call convolver interface, served by the CUDA kernels.
 *
 * Same symbols, argument meaning, ownership and error behaviour as /root/reference/convolver.h:16-152
 * (implemented there by fftw_convolver.c + convolver_xmm.c).  Every buffer is a HOST pointer owned by
 * the caller, exactly as bfrun.c / bfconf.c pass them; each call copies its operands
 * to the device, runs the kernel and copies the result back, so an unmodified bfrun.o / bfconf.o links
 * and runs against libbfcuda.so.  Per-call granularity is hostile to a GPU -- the block-level interface
 * in bfcuda.h is the one the host should call per audio block; this surface exists for link
 * compatibility, for start-up work (coefficient preprocessing) and for buffer-for-buffer parity tests.
 *
 * Layouts are the reference's: time-domain cbufs of N = 2L reals, FFTW half-complex spectra, and the
 * blocked "4 real / 4 imaginary" layout between mixnscale(INPUT) and mixnscale(OUTPUT).
 *
 * struct bfcuda_buffer_format / struct bfcuda_overflow are layout-identical to struct buffer_format
 * (dai.h:30-34) and struct bfoverflow (bfmod.h:99-104).
 */
#ifndef BFCUDA_CONVOLVER_H
#define BFCUDA_CONVOLVER_H

#include "bfcuda.h"

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int bool_t;     /* defs.h:14 */

#define CONVOLVER_MIXMODE_INPUT 1       /* convolver.h:38-40 */
#define CONVOLVER_MIXMODE_INPUT_ADD 2
#define CONVOLVER_MIXMODE_OUTPUT 3

/* The host globals the reference's convolver reads: bf_exit (called on invalid mixmode /
 * sample size / safety limit; default exit(status)), bfconf->quiet, bfconf->safety_limit.  Non-finite
 * output calls the exit hook with status -5 where the reference abort()s (real2raw.h:27-31). */
void bfcuda_convolver_set_host(void (*bf_exit_hook)(int status), int quiet, double safety_limit);
const char *bfcuda_convolver_last_error(void);

/* Dither (convolver_cbuf2raw with apply_dither, fftw_convolver.c:489-515): the reference's convolver reads two more
 * host globals, the table dither_init() built -- dither_randtab / dither_randtab_size (dither.c:22-24, used by
 * dither_preloop_real2int_hp_tpdf, dither.h:28-38).  Hand them over once after dither_init(); the table stays the
 * host's (the preloop writes its slot 0 on a wrap, exactly as in the reference).  `dither_state` of
 * convolver_cbuf2raw is the host's struct dither_state, whose layout is: */
struct bfcuda_dither_state {    /* == struct dither_state, dither.h:17-22 */
    int randtab_ptr;
    int8_t *randtab;
    float sf[2];
    double sd[2];
};
void bfcuda_convolver_set_dither_table(int8_t *dither_randtab, int dither_randtab_size);

/* convolver.h:148-152 -- `config_filename` (FFTW wisdom) is ignored, as the header allows */
bool_t convolver_init(const char config_filename[], int length, int realsize);
int convolver_cbufsize(void);                                                   /* convolver.h:98-100 */
void convolver_raw2cbuf(void *rawbuf, void *cbuf, void *next_cbuf, struct bfcuda_buffer_format *bf,
                        void (*postprocess)(void *realbuf, int n_samples, void *arg), void *pp_arg);   /* :16-25 */
void convolver_time2freq(void *input_cbuf, void *output_cbuf);                  /* :27-30 */
void convolver_mixnscale(void *input_cbufs[], void *output_cbuf, double scales[], int n_bufs, int mixmode); /* :32-41 */
void convolver_convolve_inplace(void *cbuf, void *coeffs);                      /* :43-46 */
void convolver_convolve(void *input_cbuf, void *coeffs, void *output_cbuf);     /* :48-52 */
void convolver_crossfade_inplace(void *input_cbuf, void *crossfade_cbuf, void *buffer_cbuf);   /* :54-57 */
void convolver_convolve_add(void *input_cbuf, void *coeffs, void *output_cbuf); /* :59-63 */
void convolver_dirac_convolve(void *input_cbuf, void *output_cbuf);             /* :65-70 */
void convolver_dirac_convolve_inplace(void *cbuf);
void convolver_freq2time(void *input_cbuf, void *output_cbuf);                  /* :72-75 */
void convolver_convolve_eval(void *input_cbuf, void *buffer_cbuf, void *output_cbuf);  /* :77-86 */
void convolver_cbuf2raw(void *cbuf, void *outbuf, struct bfcuda_buffer_format *bf, bool_t apply_dither,
                        void *dither_state, struct bfcuda_overflow *overflow);  /* :88-95 */
void *convolver_coeffs2cbuf(void *coeffs, int n_coeffs, double scale, void *optional_dest);    /* :102-108 */
void convolver_runtime_coeffs2cbuf(void *src, void *dest);                      /* :110-113 */
bool_t convolver_verify_cbuf(void *cbufs[], int n_cbufs);                       /* :116-119 */
void convolver_debug_dump_cbuf(const char filename[], void *cbufs[], int n_cbufs);     /* :121-126 */

/* FFTW plans cannot be handed out: returns NULL and sets bfcuda_convolver_last_error() (there is no FFTW
 * behind this convolver; the only caller besides the convolver itself is convolver_td_new, below). */
void *convolver_fftplan(int order, int invert, int inplace);                    /* :128-132 */
/* The small ordered-layout convolver of the sub-sample delay (fftw_convolver.c:682-782, caller delay.c:415-506):
 * coefficients and blocks are HOST arrays of `realsize` reals as in the reference; a block is 2 * block_length
 * reals, convolved in place.  n_coeffs up to 16384 (float) / 8192 (double). */
typedef struct _td_conv_t_ td_conv_t;
int convolver_td_block_length(int n_coeffs);                                    /* :137-138 */
td_conv_t *convolver_td_new(void *coeffs, int n_coeffs);                        /* :140-142 */
void convolver_td_convolve(td_conv_t *tdc, void *overlap_block);                /* :144-146 */
/* extension (the reference never frees a td convolver): release its device memory */
void bfcuda_convolver_td_delete(td_conv_t *tdc);

#ifdef __cplusplus
}
#endif
#endif

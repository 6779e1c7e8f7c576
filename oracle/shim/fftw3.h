/*
 * oracle/shim/fftw3.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Stand-in for the FFTW3 header the reference includes (/root/reference/fftw_convolver.c:20).
 * FFTW3 is an un-vendored, un-pinned dependency of the reference (/root/reference/Makefile:21 links
 * -lfftw3 -lfftw3f; no version is named anywhere) and is not installed in this image, so the exact
 * subset of its API the convolver uses is provided here on top of oracle/shim/fft_shim.c:
 *   fftw{,f}_plan_r2r_1d (kinds FFTW_R2HC, FFTW_HC2R), fftw{,f}_execute_r2r (in == out allowed),
 *   fftw{,f}_import_wisdom_from_file, fftw{,f}_export_wisdom_to_file (no-ops).
 * Semantics follow FFTW's published definition of the r2r half-complex transforms (unnormalised).
 */
#ifndef BF_ORACLE_SHIM_FFTW3_H
#define BF_ORACLE_SHIM_FFTW3_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct shim_pland *fftw_plan;
typedef struct shim_planf *fftwf_plan;

typedef enum { FFTW_R2HC = 0, FFTW_HC2R = 1 } fftw_r2r_kind;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_r2r_1d(int n, double *in, double *out, fftw_r2r_kind kind, unsigned flags);
fftwf_plan fftwf_plan_r2r_1d(int n, float *in, float *out, fftw_r2r_kind kind, unsigned flags);
void fftw_execute_r2r(const fftw_plan p, double *in, double *out);
void fftwf_execute_r2r(const fftwf_plan p, float *in, float *out);
void fftw_destroy_plan(fftw_plan p);
void fftwf_destroy_plan(fftwf_plan p);
int fftw_import_wisdom_from_file(FILE *f);
int fftwf_import_wisdom_from_file(FILE *f);
void fftw_export_wisdom_to_file(FILE *f);
void fftwf_export_wisdom_to_file(FILE *f);

#ifdef __cplusplus
}
#endif
#endif

/*
 * oracle/shim/fft_shim.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * From-scratch CPU real FFT behind the FFTW3 subset declared in oracle/shim/fftw3.h.  The float API
 * runs a float core (so CPU timings of the reference path are honest), the double API a double core.
 * Plans are immutable after creation and execution uses thread-local scratch, so one plan may be
 * executed from many threads at once (the CPU baseline runs one worker per core).
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "fftw3.h"

#define REAL float
#define SFX f
#include "fft_core.inc"
#undef REAL
#undef SFX

#define REAL double
#define SFX d
#include "fft_core.inc"
#undef REAL
#undef SFX

fftwf_plan
fftwf_plan_r2r_1d(int n, float *in, float *out, fftw_r2r_kind kind, unsigned flags)
{
    (void)in; (void)out; (void)flags;
    return shim_plan_newf(n, (int)kind);
}

fftw_plan
fftw_plan_r2r_1d(int n, double *in, double *out, fftw_r2r_kind kind, unsigned flags)
{
    (void)in; (void)out; (void)flags;
    return shim_plan_newd(n, (int)kind);
}

void
fftwf_execute_r2r(const fftwf_plan p, float *in, float *out)
{
    if (p->kind == FFTW_R2HC) {
        shim_r2hcf(p, in, out);
    } else {
        shim_hc2rf(p, in, out);
    }
}

void
fftw_execute_r2r(const fftw_plan p, double *in, double *out)
{
    if (p->kind == FFTW_R2HC) {
        shim_r2hcd(p, in, out);
    } else {
        shim_hc2rd(p, in, out);
    }
}

void
fftwf_destroy_plan(fftwf_plan p)
{
    int i;
    if (p == NULL) return;
    for (i = 0; i < p->n_pass; i++) free(p->tw[i]);
    free(p->wn);
    free(p);
}

void
fftw_destroy_plan(fftw_plan p)
{
    int i;
    if (p == NULL) return;
    for (i = 0; i < p->n_pass; i++) free(p->tw[i]);
    free(p->wn);
    free(p);
}

int fftw_import_wisdom_from_file(FILE *f) { (void)f; return 0; }
int fftwf_import_wisdom_from_file(FILE *f) { (void)f; return 0; }
void fftw_export_wisdom_to_file(FILE *f) { (void)f; }
void fftwf_export_wisdom_to_file(FILE *f) { (void)f; }

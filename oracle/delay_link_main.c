/*
 * delay_link_main.c -- TEST INFRASTRUCTURE (not part of the product).
 *
 * Drives the reference's UNMODIFIED sub-sample delay code -- delay.c (delay_subsample_init / delay_subsample_update,
 * /root/reference/delay.c:415-510) with firwindow.c and emalloc.c, compiled where they lie under /root/reference -- on
 * top of a convolver that is chosen at LINK time:
 *
 *   oracle/_ref/delay_link_gpu   links brutefir_b200/libbfcuda.so: the reference object code calls convolver_init,
 *                                convolver_td_block_length / _new / _convolve (convolver.h:128-152) of the CUDA library.
 *                                This is boundary path A of INTEGRATION.md exercised for real: reference objects, no
 *                                source change, our shared library.
 *   oracle/_ref/delay_link_ref   links oracle/_ref/libbfref.so: the reference's own fftw_convolver.c.
 *
 * Both write the delayed fragments to stdout as raw reals; tests/test_gpu_link_reference.py compares them.
 * bfconf and bf_exit are the two host globals delay.c and the convolver expect from the main program
 * (pinfo.h:12, bfrun.h).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bfconf.h"
#include "convolver.h"
#include "delay.h"

#ifndef DELAY_LINK_NO_GLOBALS
static struct bfconf bfconf_storage;
struct bfconf *bfconf = &bfconf_storage;
void
bf_exit(int status)
{
    fprintf(stderr, "delay_link: bf_exit(%d)\n", status);
    exit(status);
}
#else
extern struct bfconf *bfconf;
#endif

static uint32_t lcg = 12345u;
static double
noise(void)
{
    lcg = lcg * 1664525u + 1013904223u;
    return ((double)(lcg >> 8) / (double)(1u << 24) - 0.5) * 0.5;
}

int
main(int argc, char *argv[])
{
    const int realsize = argc > 1 ? atoi(argv[1]) : 4;
    const int fragment = argc > 2 ? atoi(argv[2]) : 1024;
    const int half_len = argc > 3 ? atoi(argv[3]) : 15;       /* sdf_length: 31 taps -> 32-sample td blocks */
    const int steps = BF_SAMPLE_SLOTS;
    static const int schedule[] = { 0, 7, -13, 99, -99, 31, 1, -1 };
    int frag, n, blocksize;
    void *buf, *rest;

    bfconf->quiet = 1;
    if (!convolver_init("/dev/null", fragment, realsize)) {
        fprintf(stderr, "delay_link: convolver_init failed\n");
        return 2;
    }
    if (!delay_subsample_init(steps, half_len, 9.0, fragment, realsize)) {
        fprintf(stderr, "delay_link: delay_subsample_init failed\n");
        return 3;
    }
    blocksize = delay_subsample_filterblocksize();
    buf = calloc((size_t)fragment, (size_t)realsize);
    rest = calloc((size_t)blocksize, (size_t)realsize);
    for (frag = 0; frag < (int)(sizeof(schedule) / sizeof(schedule[0])); frag++) {
        for (n = 0; n < fragment; n++) {
            if (realsize == 4) {
                ((float *)buf)[n] = (float)noise();
            } else {
                ((double *)buf)[n] = noise();
            }
        }
        delay_subsample_update(buf, rest, schedule[frag] % steps);
        if (fwrite(buf, (size_t)realsize, (size_t)fragment, stdout) != (size_t)fragment) {
            return 4;
        }
    }
    fprintf(stderr, "delay_link: %d fragments of %d, td block %d, realsize %d\n", frag, fragment, blocksize, realsize);
    return 0;
}

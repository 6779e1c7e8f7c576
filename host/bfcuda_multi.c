/*
 * bfcuda_multi.c -- the multi-GPU host in C: BruteFIR's "one filter process per CPU" rule applied to GPUs, written
 * against the C ABI of include/bfcuda.h only.
 *
 * load_balance_filters() (/root/reference/bfconf.c:2227-2318) groups filters that mix into the same output (or are
 * chained) and deals the groups round robin over the processes.  Here a "process" is one engine on one GPU, all driven
 * from this one host thread: engine g holds the coefficient spectra and delay lines of its filters only, is handed an
 * interleaved block of just the input channels those filters read, and returns a block of just the output channels they
 * feed (one dai device per GPU, dai.c:537-576).  The host fans the frames of the input file out into the per-GPU blocks
 * and gathers the per-GPU output blocks into the frames of the output file; nothing is exchanged between the GPUs
 * (diagonal graphs and per-output groups of a matrix need no collective).
 *
 *   bfcuda_multi -g 4 -n 64 -L 8192 -P 128 [-m] [-c taps.f32|dirac] [-B 8] [-r 32|64] [-i fmt] [-o fmt] in.raw out.raw
 *     -g gpus : engines; engine k runs on CUDA device k modulo the number of devices present (spread over every
 *               (devices / gpus)-th device when the box has at least twice as many)
 *     -m      : n x n matrix (output o = sum over inputs i of filter o*n+i, each scaled 1/n) instead of the diagonal
 *     -B      : audio blocks per call (file-to-file mode); four calls are kept in flight per engine
 *   Output is byte-identical to bfcuda_run's for the same arguments (tests/test_gpu_host.py).
 */
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

#include "bfcuda.h"

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } while (0)
#define CHECK(call) do { int rc__ = (call); if (rc__ != 0) DIE("%s failed (%d): %s", #call, rc__, bfcuda_strerror()); } while (0)
#define MAXGPU 16
#define DEPTH 4

static int
parse_format(const char *s, struct bfcuda_sample_format *sf)
{
    /* the formats of bfconf.c:377-472 this host reads and writes */
    static const struct { const char *name; int isfloat, bytes, sbytes, id; } F[] = {
        { "S16_LE", 0, 2, 2, 2 }, { "S24_LE", 0, 3, 3, 6 }, { "S24_4LE", 0, 4, 3, 8 }, { "S32_LE", 0, 4, 4, 10 },
        { "FLOAT_LE", 1, 4, 4, 12 }, { "FLOAT64_LE", 1, 8, 8, 14 } };
    size_t i;
    for (i = 0; i < sizeof(F) / sizeof(F[0]); i++) {
        if (strcasecmp(s, F[i].name) == 0) {
            sf->isfloat = F[i].isfloat;
            sf->bytes = F[i].bytes;
            sf->sbytes = F[i].sbytes;
            sf->swap = 0;
            sf->format = F[i].id;
            sf->scale = sf->isfloat ? 1.0 : 1.0 / (double)((uint64_t)1 << ((sf->sbytes << 3) - 1));
            return 0;
        }
    }
    return -1;
}

static int
interleaved(struct bfcuda_buffer_format *bf, int n, const struct bfcuda_sample_format *sf, int fragsize)
{
    int c, n_bytes = n * sf->bytes * fragsize;      /* calc_buffer_format, dai.c:537-576 */
    for (c = 0; c < n; c++) {
        bf[c].sf = *sf;
        bf[c].sample_spacing = n;
        bf[c].byte_offset = c * sf->bytes;
    }
    if (n_bytes % 32 != 0) {
        n_bytes += 32 - n_bytes % 32;
    }
    return n_bytes;
}

static double
now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

struct shard {
    bfcuda_engine *eng;
    int n_in, n_out, n_filters;
    int *in_ch, *out_ch;            /* global channel of every local one, ascending */
    int *filt;                      /* global filter of every local one */
    size_t in_bytes, out_bytes;     /* of one local block */
    void *raw_in[DEPTH], *raw_out[DEPTH];
};

/* frames of the full interleaved block <-> frames of a shard's block: sample s of global channel ch[j] is local j */
static void
fan_out(const struct shard *sh, const unsigned char *full, int n, int bytes, int frames, unsigned char *local)
{
    int fr, j;
    for (fr = 0; fr < frames; fr++) {
        const unsigned char *src = full + (size_t)fr * n * bytes;
        unsigned char *dst = local + (size_t)fr * sh->n_in * bytes;
        for (j = 0; j < sh->n_in; j++) {
            memcpy(dst + (size_t)j * bytes, src + (size_t)sh->in_ch[j] * bytes, (size_t)bytes);
        }
    }
}

static void
gather(const struct shard *sh, const unsigned char *local, int n, int bytes, int frames, unsigned char *full)
{
    int fr, j;
    for (fr = 0; fr < frames; fr++) {
        const unsigned char *src = local + (size_t)fr * sh->n_out * bytes;
        unsigned char *dst = full + (size_t)fr * n * bytes;
        for (j = 0; j < sh->n_out; j++) {
            memcpy(dst + (size_t)sh->out_ch[j] * bytes, src + (size_t)j * bytes, (size_t)bytes);
        }
    }
}

int
main(int argc, char *argv[])
{
    int n = 8, L = 1024, P = 8, realbits = 32, matrix = 0, gpus = 2, batch = 1, a, g, f, k, rs, n_filters, n_dev;
    const char *fin = "S24_4LE", *fout = "S24_4LE", *coeff_path = "dirac", *in_path = NULL, *out_path = NULL;
    struct bfcuda_sample_format sf_in, sf_out;
    struct shard sh[MAXGPU];
    FILE *in, *out, *cf = NULL;
    unsigned char *full_in, *full_out;
    size_t full_in_bytes, full_out_bytes;
    int nblk[DEPTH];
    long blocks = 0;
    double t0, t1, one;

    for (a = 1; a < argc; a++) {
        if (!strcmp(argv[a], "-n") && a + 1 < argc) n = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-L") && a + 1 < argc) L = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-P") && a + 1 < argc) P = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-r") && a + 1 < argc) realbits = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-g") && a + 1 < argc) gpus = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-B") && a + 1 < argc) batch = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-i") && a + 1 < argc) fin = argv[++a];
        else if (!strcmp(argv[a], "-o") && a + 1 < argc) fout = argv[++a];
        else if (!strcmp(argv[a], "-c") && a + 1 < argc) coeff_path = argv[++a];
        else if (!strcmp(argv[a], "-m")) matrix = 1;
        else if (in_path == NULL) in_path = argv[a];
        else if (out_path == NULL) out_path = argv[a];
        else DIE("usage: %s -g gpus [-n ch] [-L len] [-P blocks] [-m] [-c taps|dirac] [-B blocks] [-r 32|64] [-i fmt] [-o fmt] in out", argv[0]);
    }
    if (gpus < 1 || gpus > MAXGPU || gpus > n) DIE("-g must be 1..%d and at most the channel count", MAXGPU);
    if (parse_format(fin, &sf_in) != 0 || parse_format(fout, &sf_out) != 0) DIE("Unknown sample format.");
    n_dev = bfcuda_device_count();
    if (n_dev < 1) DIE("no CUDA device (there is no CPU fallback)");
    rs = realbits / 8;
    n_filters = matrix ? n * n : n;
    one = matrix ? 1.0 / n : 1.0;
    if (strcmp(coeff_path, "dirac") != 0 && (cf = fopen(coeff_path, "rb")) == NULL) DIE("Could not open \"%s\".", coeff_path);

    /* Filter groups: filters that feed the same output stay together (bfconf.c:2234-2298).  Diagonal: group = filter;
     * matrix: group o = the n filters of output o.  Group k goes to engine k % gpus (bfconf.c:2300-2316). */
    for (g = 0; g < gpus; g++) {
        struct shard *s = &sh[g];
        struct bfcuda_config cfg;
        struct bfcuda_buffer_format *bf_in, *bf_out;
        struct bfcuda_filter *filters;
        int *chan, *cb, o, i, *used;
        memset(s, 0, sizeof(*s));
        s->in_ch = calloc((size_t)n, sizeof(int));
        s->out_ch = calloc((size_t)n, sizeof(int));
        s->filt = calloc((size_t)n_filters, sizeof(int));
        used = calloc((size_t)n, sizeof(int));
        for (o = 0; o < n; o++) {
            if (o % gpus != g) {
                continue;
            }
            s->out_ch[s->n_out++] = o;
            if (matrix) {
                for (i = 0; i < n; i++) {
                    s->filt[s->n_filters++] = o * n + i;
                    used[i] = 1;
                }
            } else {
                s->filt[s->n_filters++] = o;
                used[o] = 1;
            }
        }
        for (i = 0; i < n; i++) {
            if (used[i]) {
                used[i] = s->n_in;          /* global -> local input */
                s->in_ch[s->n_in++] = i;
            }
        }
        bf_in = calloc((size_t)s->n_in, sizeof(*bf_in));
        bf_out = calloc((size_t)s->n_out, sizeof(*bf_out));
        filters = calloc((size_t)s->n_filters, sizeof(*filters));
        chan = calloc(2 * (size_t)s->n_filters, sizeof(int));
        cb = calloc((size_t)s->n_filters, sizeof(int));
        memset(&cfg, 0, sizeof(cfg));
        cfg.filter_length = L;
        cfg.n_blocks = P;
        cfg.realsize = rs;
        cfg.n_channels[BFCUDA_IN] = s->n_in;
        cfg.n_channels[BFCUDA_OUT] = s->n_out;
        cfg.n_bytes[BFCUDA_IN] = interleaved(bf_in, s->n_in, &sf_in, L);
        cfg.n_bytes[BFCUDA_OUT] = interleaved(bf_out, s->n_out, &sf_out, L);
        cfg.formats[BFCUDA_IN] = bf_in;
        cfg.formats[BFCUDA_OUT] = bf_out;
        for (f = 0; f < s->n_filters; f++) {
            const int gf = s->filt[f], gi = matrix ? gf % n : gf, go = matrix ? gf / n : gf;
            for (o = 0; o < s->n_out && s->out_ch[o] != go; o++) {
            }
            chan[2 * f] = used[gi];
            chan[2 * f + 1] = o;
            filters[f].n_channels[BFCUDA_IN] = filters[f].n_channels[BFCUDA_OUT] = 1;
            filters[f].channels[BFCUDA_IN] = &chan[2 * f];
            filters[f].channels[BFCUDA_OUT] = &chan[2 * f + 1];
            filters[f].scale[BFCUDA_IN] = filters[f].scale[BFCUDA_OUT] = &one;
            filters[f].coeff = f;           /* the engine holds only its own filters' coefficient sets */
            cb[f] = P;
        }
        cfg.n_filters = s->n_filters;
        cfg.filters = filters;
        cfg.n_coeffs = s->n_filters;
        cfg.coeff_n_blocks = cb;
        /* The GPUs of an HGX box share PCIe switch uplinks in pairs (profiles/r2_copy_skew_n8.txt): with fewer engines than
         * devices take every (n_dev / gpus)-th device, one uplink per engine; otherwise engine k runs on device k modulo the
         * number of devices. */
        cfg.device = (gpus > 1 && n_dev >= 2 * gpus && n_dev % gpus == 0) ? g * (n_dev / gpus) : g % n_dev;
        cfg.max_batch = batch;
        CHECK(bfcuda_create(&cfg, &s->eng));
        s->in_bytes = (size_t)cfg.n_bytes[BFCUDA_IN];
        s->out_bytes = (size_t)cfg.n_bytes[BFCUDA_OUT];
        for (k = 0; k < DEPTH; k++) {
            s->raw_in[k] = bfcuda_host_alloc_near(cfg.device, s->in_bytes * (size_t)batch);
            s->raw_out[k] = bfcuda_host_alloc_near(cfg.device, s->out_bytes * (size_t)batch);
            if (s->raw_in[k] == NULL || s->raw_out[k] == NULL) DIE("%s", bfcuda_strerror());
        }
        free(used);
    }
    /* coefficients: taps of global filter f at offset f * L * P of the file, handed to the engine that owns it */
    {
        const size_t taps = (size_t)L * P;
        void *h = calloc(taps, (size_t)rs);
        for (g = 0; g < gpus; g++) {
            for (f = 0; f < sh[g].n_filters; f++) {
                memset(h, 0, taps * rs);
                if (cf == NULL) {
                    if (rs == 4) ((float *)h)[0] = 1.0f; else ((double *)h)[0] = 1.0;
                } else if (fseek(cf, (long)((size_t)sh[g].filt[f] * taps * rs), SEEK_SET) != 0 ||
                           fread(h, (size_t)rs, taps, cf) == 0) {
                    DIE("\"%s\" holds fewer than %d filters.", coeff_path, n_filters);
                }
                CHECK(bfcuda_coeff_from_taps(sh[g].eng, f, h, (int)taps, 1.0));
            }
        }
        free(h);
        if (cf != NULL) fclose(cf);
    }

    in = in_path == NULL || !strcmp(in_path, "-") ? stdin : fopen(in_path, "rb");
    out = out_path == NULL || !strcmp(out_path, "-") ? stdout : fopen(out_path, "wb");
    if (in == NULL || out == NULL) DIE("Could not open input or output: %s", strerror(errno));
    full_in_bytes = (size_t)n * sf_in.bytes * L;        /* the files hold plain frames, no device padding */
    full_out_bytes = (size_t)n * sf_out.bytes * L;
    full_in = malloc(full_in_bytes * (size_t)batch);
    full_out = malloc(full_out_bytes * (size_t)batch);
    fprintf(stderr, "bfcuda_multi: %d filters x %d taps over %d engine(s) on %d device(s), %d block(s) per call\n", n_filters,
            L * P, gpus, n_dev < gpus ? n_dev : gpus, batch);

    /* call k is submitted on every engine, then the outputs of call k-2 (complete by now: the engines run forward(k),
     * MAC(k-1) and inverse(k-2) side by side) are gathered and written */
    t0 = now();
    for (k = 0;; k++) {
        const int s = k % DEPTH;
        size_t got = fread(full_in, 1, full_in_bytes * (size_t)batch, in);
        if (got == 0) {
            break;
        }
        nblk[s] = (int)((got + full_in_bytes - 1) / full_in_bytes);
        if (got < (size_t)nblk[s] * full_in_bytes) {
            memset(full_in + got, 0, (size_t)nblk[s] * full_in_bytes - got);       /* dai.c:1312-1332 */
        }
        for (g = 0; g < gpus; g++) {
            int b;
            for (b = 0; b < nblk[s]; b++) {
                fan_out(&sh[g], full_in + (size_t)b * full_in_bytes, n, sf_in.bytes, L,
                        (unsigned char *)sh[g].raw_in[s] + (size_t)b * sh[g].in_bytes);
            }
            CHECK(bfcuda_process_blocks_async(sh[g].eng, nblk[s], sh[g].raw_in[s], sh[g].raw_out[s]));
        }
        if (k > 1) {
            const int p = (k - 2) % DEPTH;
            int b;
            for (g = 0; g < gpus; g++) {
                CHECK(bfcuda_wait_previous(sh[g].eng, 2));
                for (b = 0; b < nblk[p]; b++) {
                    gather(&sh[g], (unsigned char *)sh[g].raw_out[p] + (size_t)b * sh[g].out_bytes, n, sf_out.bytes, L,
                           full_out + (size_t)b * full_out_bytes);
                }
            }
            if (fwrite(full_out, 1, full_out_bytes * (size_t)nblk[p], out) != full_out_bytes * (size_t)nblk[p]) DIE("write failed");
        }
        blocks += nblk[s];
    }
    if (k > 0) {
        int q, b;
        for (g = 0; g < gpus; g++) {
            CHECK(bfcuda_synchronize(sh[g].eng));
        }
        for (q = k > 1 ? k - 2 : 0; q < k; q++) {
            const int p = q % DEPTH;
            for (g = 0; g < gpus; g++) {
                for (b = 0; b < nblk[p]; b++) {
                    gather(&sh[g], (unsigned char *)sh[g].raw_out[p] + (size_t)b * sh[g].out_bytes, n, sf_out.bytes, L,
                           full_out + (size_t)b * full_out_bytes);
                }
            }
            if (fwrite(full_out, 1, full_out_bytes * (size_t)nblk[p], out) != full_out_bytes * (size_t)nblk[p]) DIE("write failed");
        }
    }
    t1 = now();
    if (out != stdout) fclose(out);
    fprintf(stderr, "bfcuda_multi: %ld blocks in %.3f s\n", blocks, t1 - t0);
    for (g = 0; g < gpus; g++) {
        bfcuda_destroy(sh[g].eng);
    }
    return 0;
}

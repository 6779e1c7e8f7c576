/*
 * bfcuda_run.c -- a C host for the GPU convolution engine: the part of BruteFIR's filter process that stays
 * on the CPU, written against the C ABI of include/bfcuda.h only.
 *
 * It does what filter_process() + bfio_file do for the benchmark configurations of the reference
 * (/root/reference/bfrun.c:1420-2083, bfio_file.c:418-451, 569-586): read raw interleaved PCM blocks of
 * filter_length frames from a file (or stdin), run them through the engine, write raw PCM blocks to a file
 * (or stdout), zero-filling the last partial block the way dai_input does at end of input
 * (dai.c:1312-1332), and print the reference's per-stage benchmark table (bfrun.c:2035-2078) and realtime
 * index (bfrun.c:649-677).  The graph is "n channels, filter i: input i -> output i", optionally a full
 * n x n matrix (-m), which covers bench2/3/5-style and massive_config-style set-ups; bfconf's parser is not
 * duplicated here (INTEGRATION.md shows where the same three calls go in an unmodified bfrun.c).
 *
 *   cc -O2 -Iinclude host/bfcuda_run.c -o host/bfcuda_run -Lbrutefir_b200 -lbfcuda -Wl,-rpath,'$ORIGIN/../brutefir_b200' -lm
 *
 *   bfcuda_run -n 2 -L 4096 -P 16 -i S24_4LE -o S24_4LE -c taps.f32 [-r 32] [-s rate] [-m] [-b] [-l] [-t] [-R] [-B blocks] in.raw out.raw
 *     -B blocks: hand the engine up to `blocks` (<= 16) audio blocks per call and keep two calls in flight while
 *               the files are read and written (offline mode; bit-identical output, several times the throughput)
 *     -t      : text files on both sides, bfio_file's `text: true` (bfio_file.c:153-185, 308-420, 509-565): white-space
 *               separated numbers in, one line per sample frame out ("%+.16e", tab separated); the sample format is
 *               then FLOAT64_LE as the reference requires
 *     -R      : the per-block pattern of filter_process() as INTEGRATION.md section B shows it: control snapshot of
 *               every filter, ONE synchronous bfcuda_process_block(), the peak-meter read of every output -- timed
 *               per block (the engine is left to finish its ahead-of-time work between two blocks, as the block
 *               period of a sound card would); prints the median and the largest call latency
 *     -l      : the real-time schedule (BFCUDA_FLAG_LOW_LATENCY): block by block, partitions 1 .. P-1 of the next
 *               block summed ahead of time (half the call latency; partition sums within tolerance, not bit-identical)
 *     taps.f32: raw little-endian float32 (float64 with -r 64) taps, one filter after the other, L*P each
 *               ("dirac" = unit pulses, the reference's "dirac pulse" coefficient, bfconf.c:1905-1913)
 *     -f fmt  : format of the coefficient file, as the `format:` field of a coeff section (bfconf.c:783-812,
 *               load_coeff bfconf.c:1867-2030): "text" (one number per line, real_read bfconf.c:1725-1766) or a
 *               sample format name (raw_read bfconf.c:1780-1821: converted like raw2real and multiplied by the
 *               format's scale); default FLOAT_LE / FLOAT64_LE by -r
 *     -a dB   : `attenuation:` of the coeff sections (scale = 10^(-dB/20), bfconf.c:778-781)
 *     -k bytes: `skip:` bytes at the start of the coefficient file (bfconf.c:1886-1893)
 *     -K blocks: `blocks:` of the coeff sections (bfconf.c:823-826, 2827-2831): the coefficient sets are K <= P partitions long
 *     -f processed: the file holds the sets already in the convolver's processed layout, K blocks of
 *               convolver_cbufsize() bytes each per filter (COEFF_FORMAT_PROCESSED, bfconf.c:1924-1957)
 *     -c shm:ID/OFFSET/BLOCKS[,ID/OFFSET/BLOCKS...]: processed blocks in System V shared memory segments, as
 *               `filename: ID/OFFSET/BLOCKS` with `shared_mem: true` (bfconf.c:795-815, get_sharedmem bfconf.c:1825-1865);
 *               the parts are concatenated and must hold n_filters * K blocks
 */
#include <sys/ipc.h>
#include <sys/shm.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

#include "bfcuda.h"

struct fmt_entry {
    const char *name;
    int isfloat, bytes, sbytes, big_endian, id;
};
/* bfconf.c:377-472 (parse_sample_format); ids bfmod.h:33-48 */
static const struct fmt_entry FORMATS[] = {
    { "S8", 0, 1, 1, 0, 1 },          { "S16_LE", 0, 2, 2, 0, 2 },     { "S16_BE", 0, 2, 2, 1, 3 },
    { "S24_LE", 0, 3, 3, 0, 6 },      { "S24_3LE", 0, 3, 3, 0, 6 },    { "S24_BE", 0, 3, 3, 1, 7 },
    { "S24_3BE", 0, 3, 3, 1, 7 },     { "S24_4LE", 0, 4, 3, 0, 8 },    { "S24_4BE", 0, 4, 3, 1, 9 },
    { "S32_LE", 0, 4, 4, 0, 10 },     { "S32_BE", 0, 4, 4, 1, 11 },    { "FLOAT_LE", 1, 4, 4, 0, 12 },
    { "FLOAT_BE", 1, 4, 4, 1, 13 },   { "FLOAT64_LE", 1, 8, 8, 0, 14 }, { "FLOAT64_BE", 1, 8, 8, 1, 15 },
};

static int
parse_format(const char *s, struct bfcuda_sample_format *sf)
{
    size_t i;
    for (i = 0; i < sizeof(FORMATS) / sizeof(FORMATS[0]); i++) {
        if (strcasecmp(s, FORMATS[i].name) == 0) {
            sf->isfloat = FORMATS[i].isfloat;
            sf->bytes = FORMATS[i].bytes;
            sf->sbytes = FORMATS[i].sbytes;
            sf->swap = FORMATS[i].big_endian;   /* little-endian host */
            sf->format = FORMATS[i].id;
            sf->scale = sf->isfloat ? 1.0 : 1.0 / (double)((uint64_t)1 << ((sf->sbytes << 3) - 1));
            return 0;
        }
    }
    return -1;
}

/* calc_buffer_format for one interleaved device, dai.c:537-576 */
static int
interleaved(struct bfcuda_buffer_format *bf, int n, const struct bfcuda_sample_format *sf, int fragsize)
{
    int c, n_bytes = n * sf->bytes * fragsize;
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

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } while (0)
#define CHECK(call) do { int rc__ = (call); if (rc__ != 0) DIE("%s failed (%d): %s", #call, rc__, bfcuda_strerror()); } while (0)

/* bfio_file's text mode (bfio_file.c:308-420): white-space separated numbers, empty lines skipped, one double each.
 * Returns the bytes filled (a multiple of 8); a token that is not a number is an error as in the reference. */
static size_t read_text(FILE *in, void *buf, size_t bytes)
{
    double *a = buf;
    size_t i = 0, count = bytes / 8;
    while (i < count) {
        int r = fscanf(in, "%lf", &a[i]);
        if (r == EOF) {
            break;
        }
        if (r != 1) {
            DIE("File I/O: Read failed: bad text format.");
        }
        i++;
    }
    return i * 8;
}

/* bfio_file.c:28, 509-565: "%+.16e" per sample, tabs between the channels of a frame, one frame per line */
static int write_text(FILE *out, const void *buf, size_t bytes, int channels)
{
    const double *a = buf;
    size_t i, count = bytes / 8;
    for (i = 0; i < count; i++) {
        if (fprintf(out, "%+.16e%c", a[i], (i + 1) % (size_t)channels == 0 ? '\n' : '\t') < 0) {
            return -1;
        }
    }
    return 0;
}

int
main(int argc, char *argv[])
{
    int n = 2, L = 4096, P = 16, realbits = 32, rate = 48000, matrix = 0, bench = 0, low_latency = 0, text_io = 0, fin_set = 0, fout_set = 0, realtime = 0, device = 0, batch = 1, a;
    const char *fin = "S24_4LE", *fout = "S24_4LE", *coeff_path = "dirac", *in_path = NULL, *out_path = NULL;
    const char *coeff_fmt = NULL;
    double attenuation_db = 0.0;
    long coeff_skip = 0;
    int coeff_nblocks = 0;
    struct bfcuda_sample_format sf_in, sf_out;
    struct bfcuda_buffer_format *bf_in, *bf_out;
    struct bfcuda_filter *filters;
    struct bfcuda_config cfg;
    struct bfcuda_info info;
    bfcuda_engine *eng = NULL;
    int n_filters, *chan, *coeff_blocks, f, c, rs;
    double *ones, t0, t1, stage[BFCUDA_N_STAGES];
    long blocks = 0, stage_blocks = 0, launches = 0;
    void *raw_in[4], *raw_out[4];
    size_t in_bytes, out_bytes;
    int nblk[4], k;
    FILE *in, *out;

    for (a = 1; a < argc; a++) {
        if (!strcmp(argv[a], "-n") && a + 1 < argc) n = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-L") && a + 1 < argc) L = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-P") && a + 1 < argc) P = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-r") && a + 1 < argc) realbits = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-s") && a + 1 < argc) rate = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-d") && a + 1 < argc) device = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-i") && a + 1 < argc) { fin = argv[++a]; fin_set = 1; }
        else if (!strcmp(argv[a], "-o") && a + 1 < argc) { fout = argv[++a]; fout_set = 1; }
        else if (!strcmp(argv[a], "-c") && a + 1 < argc) coeff_path = argv[++a];
        else if (!strcmp(argv[a], "-f") && a + 1 < argc) coeff_fmt = argv[++a];
        else if (!strcmp(argv[a], "-a") && a + 1 < argc) attenuation_db = atof(argv[++a]);
        else if (!strcmp(argv[a], "-k") && a + 1 < argc) coeff_skip = atol(argv[++a]);
        else if (!strcmp(argv[a], "-K") && a + 1 < argc) coeff_nblocks = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-m")) matrix = 1;
        else if (!strcmp(argv[a], "-b")) bench = 1;
        else if (!strcmp(argv[a], "-l")) low_latency = 1;
        else if (!strcmp(argv[a], "-t")) text_io = 1;
        else if (!strcmp(argv[a], "-R")) realtime = 1;
        else if (!strcmp(argv[a], "-B") && a + 1 < argc) batch = atoi(argv[++a]);
        else if (in_path == NULL) in_path = argv[a];
        else if (out_path == NULL) out_path = argv[a];
        else DIE("usage: %s [-n ch] [-L len] [-P blocks] [-r 32|64] [-s rate] [-i fmt] [-o fmt] [-c taps|dirac] [-m] [-b] [-l] [-t] [-R] [-B blocks] [in [out]]", argv[0]);
    }
    if (text_io) {
        /* bfio_file.c:165-185: text conversion exists for the native FLOAT64 format only (AUTO selects it) */
        if ((fin_set && strcmp(fin, "FLOAT64_LE") != 0) || (fout_set && strcmp(fout, "FLOAT64_LE") != 0)) {
            DIE("File I/O: No support for text conversion of given sample format.");
        }
        fin = fout = "FLOAT64_LE";
    }
    if (parse_format(fin, &sf_in) != 0 || parse_format(fout, &sf_out) != 0) DIE("Unknown sample format.");
    rs = realbits / 8;
    n_filters = matrix ? n * n : n;
    bf_in = calloc(n, sizeof(*bf_in));
    bf_out = calloc(n, sizeof(*bf_out));
    filters = calloc(n_filters, sizeof(*filters));
    chan = calloc(2 * n_filters, sizeof(int));
    coeff_blocks = calloc(n_filters, sizeof(int));
    ones = calloc(1, sizeof(double));
    ones[0] = matrix ? 1.0 / n : 1.0;

    memset(&cfg, 0, sizeof(cfg));
    cfg.filter_length = L;
    cfg.n_blocks = P;
    cfg.realsize = rs;
    cfg.n_channels[BFCUDA_IN] = cfg.n_channels[BFCUDA_OUT] = n;
    cfg.n_bytes[BFCUDA_IN] = interleaved(bf_in, n, &sf_in, L);
    cfg.n_bytes[BFCUDA_OUT] = interleaved(bf_out, n, &sf_out, L);
    cfg.formats[BFCUDA_IN] = bf_in;
    cfg.formats[BFCUDA_OUT] = bf_out;
    for (f = 0; f < n_filters; f++) {
        chan[2 * f] = matrix ? f % n : f;       /* from_inputs */
        chan[2 * f + 1] = matrix ? f / n : f;   /* to_outputs */
        filters[f].n_channels[BFCUDA_IN] = filters[f].n_channels[BFCUDA_OUT] = 1;
        filters[f].channels[BFCUDA_IN] = &chan[2 * f];
        filters[f].channels[BFCUDA_OUT] = &chan[2 * f + 1];
        filters[f].scale[BFCUDA_IN] = filters[f].scale[BFCUDA_OUT] = ones;
        filters[f].coeff = f;
        coeff_blocks[f] = coeff_nblocks > 0 ? coeff_nblocks : P;        /* bfconf.c:2827-2831 */
    }
    if (coeff_nblocks > P) DIE("Too many blocks in coeff 0.");
    cfg.n_filters = n_filters;
    cfg.filters = filters;
    cfg.n_coeffs = n_filters;
    cfg.coeff_n_blocks = coeff_blocks;
    cfg.device = device;
    cfg.flags = (bench ? BFCUDA_FLAG_STAGE_TIMING : 0) | (low_latency ? BFCUDA_FLAG_LOW_LATENCY : 0);
    cfg.max_batch = batch;
    CHECK(bfcuda_create(&cfg, &eng));
    CHECK(bfcuda_get_info(eng, &info));

    /* coefficients already in the convolver's processed layout: a file of blocks, or System V shared memory
       (COEFF_FORMAT_PROCESSED, bfconf.c:1924-1957; get_sharedmem, bfconf.c:1825-1865) */
    if ((coeff_fmt != NULL && strcasecmp(coeff_fmt, "processed") == 0) || strncmp(coeff_path, "shm:", 4) == 0) {
        const int K = coeff_blocks[0];
        const size_t cbufsize = (size_t)2 * L * rs;        /* convolver_cbufsize(), fftw_convolver.c:520-524 */
        const size_t want = (size_t)n_filters * K;
        size_t have = 0;
        if (strncmp(coeff_path, "shm:", 4) == 0) {
            const char *p = coeff_path + 4;
            while (*p != '\0') {
                int id, off, nb, used = 0;
                unsigned char *seg;
                if (sscanf(p, "%d/%d/%d%n", &id, &off, &nb, &used) != 3) DIE("bad shared memory spec \"%s\"", coeff_path);
                if ((seg = shmat(id, NULL, SHM_RDONLY)) == (void *)-1) {
                    DIE("Failed to attach to shared memory with id %d: %s.", id, strerror(errno));
                }
                if ((off & 31) != 0) DIE("Shared memory pointer with id %d and offset %d is not aligned at a 32 byte boundary.", id, off);
                for (f = 0; f < nb && have < want; f++, have++) {
                    CHECK(bfcuda_coeff_set_block(eng, (int)(have / K), (int)(have % K), seg + off + (size_t)f * cbufsize));
                }
                shmdt(seg);
                p += used;
                if (*p == ',') p++;
            }
        } else {
            FILE *pf = fopen(coeff_path, "rb");
            void *blk = malloc(cbufsize);
            if (pf == NULL) DIE("Could not open \"%s\" for reading.", coeff_path);
            if (coeff_skip > 0) fseek(pf, coeff_skip, SEEK_SET);
            while (have < want && fread(blk, 1, cbufsize, pf) == cbufsize) {
                CHECK(bfcuda_coeff_set_block(eng, (int)(have / K), (int)(have % K), blk));
                have++;
            }
            fclose(pf);
            free(blk);
        }
        if (have != want) DIE("Shared memory block count mismatch in coeff %d.", (int)(have / K));
    } else
    /* coefficients: load_coeff (bfconf.c:1867-2030) for "dirac pulse", text and raw sample formats */
    {
        size_t taps = (size_t)L * coeff_blocks[0];
        void *h = calloc(taps, rs);
        FILE *cf = NULL;
        const int is_text = coeff_fmt != NULL && strcasecmp(coeff_fmt, "text") == 0;
        const double scale = pow(10.0, -attenuation_db / 20.0);      /* FROM_DB(-attenuation), bfconf.c:781 */
        struct bfcuda_sample_format csf;
        if (!is_text && parse_format(coeff_fmt != NULL ? coeff_fmt : (rs == 4 ? "FLOAT_LE" : "FLOAT64_LE"), &csf) != 0) {
            DIE("Unknown coefficient format \"%s\".", coeff_fmt);
        }
        if (strcmp(coeff_path, "dirac") != 0) {
            if ((cf = fopen(coeff_path, is_text ? "rt" : "rb")) == NULL) {
                DIE("Could not open \"%s\" for reading.", coeff_path);
            }
            if (coeff_skip > 0 && (fseek(cf, coeff_skip, SEEK_SET) != 0 || ftell(cf) != coeff_skip)) {
                DIE("Failed to skip %ld bytes of file \"%s\".", coeff_skip, coeff_path);
            }
        }
        for (f = 0; f < n_filters; f++) {
            size_t got = 0;
            memset(h, 0, taps * rs);
            if (cf == NULL) {
                if (rs == 4) ((float *)h)[0] = 1.0f; else ((double *)h)[0] = 1.0;
                got = taps;
            } else if (is_text) {
                char line[1024];
                while (got < taps && fgets(line, sizeof(line) - 1, cf) != NULL) {      /* real_read */
                    char *s0 = line, *e0;
                    double v;
                    while (*s0 == ' ' || *s0 == '\t') s0++;
                    if (*s0 == '\n' || *s0 == '\0') continue;
                    v = strtod(s0, &e0);
                    if (e0 == s0) DIE("Parse error in file %s: invalid floating point number.", coeff_path);
                    if (rs == 4) ((float *)h)[got] = (float)v; else ((double *)h)[got] = v;
                    got++;
                }
            } else {
                unsigned char raw[8];
                while (got < taps && fread(raw, csf.bytes, 1, cf) == 1) {               /* raw_read + raw2real */
                    double v;
                    int i;
                    unsigned char le[8];
                    for (i = 0; i < csf.bytes; i++) le[i] = csf.swap ? raw[csf.bytes - 1 - i] : raw[i];
                    if (csf.isfloat) {
                        if (csf.bytes == 4) { float t; memcpy(&t, le, 4); v = (double)t; }
                        else { memcpy(&v, le, 8); }
                    } else {
                        int32_t q = 0;
                        for (i = 0; i < csf.bytes; i++) q |= (int32_t)((uint32_t)le[i] << (8 * (i + 4 - csf.bytes)));
                        v = (double)(q >> (8 * (4 - csf.bytes)));       /* sign-extending shift, raw2real.h:106-142 */
                    }
                    /* the integer is converted unscaled into the real type, then multiplied by sf.scale in the real
                     * type (bfconf.c:1807-1818) */
                    if (rs == 4) ((float *)h)[got] = (float)v * (float)csf.scale;
                    else ((double *)h)[got] = v * csf.scale;
                    got++;
                }
            }
            if (got == 0) DIE("\"%s\" holds fewer than %d filters.", coeff_path, n_filters);
            CHECK(bfcuda_coeff_from_taps(eng, f, h, (int)taps, scale));
        }
        if (cf != NULL) fclose(cf);
        free(h);
    }

    in = in_path == NULL || !strcmp(in_path, "-") ? stdin : fopen(in_path, "rb");
    out = out_path == NULL || !strcmp(out_path, "-") ? stdout : fopen(out_path, "wb");
    if (in == NULL || out == NULL) DIE("Could not open input or output: %s", strerror(errno));
    in_bytes = (size_t)cfg.n_bytes[BFCUDA_IN];
    out_bytes = (size_t)cfg.n_bytes[BFCUDA_OUT];
    for (k = 0; k < 4; k++) {
        raw_in[k] = bfcuda_host_alloc(in_bytes * (size_t)batch);
        raw_out[k] = bfcuda_host_alloc(out_bytes * (size_t)batch);
        if (raw_in[k] == NULL || raw_out[k] == NULL) DIE("%s", bfcuda_strerror());
    }

    fprintf(stderr, "bfcuda_run: %d filters x %d taps (%d x %d) on %s, MAC %.1f MB/block, %d block(s) per call\n", n_filters,
            L * P, L, P, info.device_name, (double)info.mac_bytes_per_block / 1e6, batch);
    t0 = now();
    /* Call k is submitted, then call k-1's output (complete while call k runs) is written: the file I/O of the
     * reference's input / output processes overlapped with the filter process, with three buffer sets instead of
     * the dai double buffers. */
    if (realtime) {
        /* filter_process()'s per-block sequence, bfrun.c:1462-1478 (snapshot), 1494-2006 (the block), 1929-1936 (meter) */
        double *lat = NULL, sum = 0.0;
        long n_lat = 0, cap = 0, i, j;
        for (;;) {
            struct bfcuda_overflow of;
            double c0, c1;
            size_t got = text_io ? read_text(in, raw_in[0], in_bytes) : fread(raw_in[0], 1, in_bytes, in);
            if (got == 0) {
                break;
            }
            if (got < in_bytes) {
                memset((char *)raw_in[0] + got, 0, in_bytes - got);
            }
            c0 = now();
            for (f = 0; f < n_filters; f++) {
                struct bfcuda_filter_control fc;
                memset(&fc, 0, sizeof(fc));
                fc.coeff = filters[f].coeff;
                fc.delayblocks = filters[f].delayblocks;
                fc.scale[BFCUDA_IN] = filters[f].scale[BFCUDA_IN];
                fc.scale[BFCUDA_OUT] = filters[f].scale[BFCUDA_OUT];
                CHECK(bfcuda_set_control(eng, f, &fc));
            }
            CHECK(bfcuda_process_block(eng, raw_in[0], raw_out[0]));
            for (c = 0; c < n; c++) {
                CHECK(bfcuda_get_overflow(eng, c, &of));
            }
            c1 = now();
            if (n_lat == cap) {
                cap = cap ? 2 * cap : 1024;
                lat = realloc(lat, (size_t)cap * sizeof(*lat));
                if (lat == NULL) DIE("out of memory");
            }
            lat[n_lat++] = c1 - c0;
            if (text_io ? write_text(out, raw_out[0], out_bytes, n) != 0 : fwrite(raw_out[0], 1, out_bytes, out) != out_bytes) {
                DIE("write failed: %s", strerror(errno));
            }
            CHECK(bfcuda_synchronize(eng));     /* the block period passes */
            blocks++;
        }
        for (i = 1; i < n_lat; i++) {           /* insertion sort: a few thousand entries at most */
            double v = lat[i];
            for (j = i - 1; j >= 0 && lat[j] > v; j--) lat[j + 1] = lat[j];
            lat[j + 1] = v;
        }
        for (i = 0; i < n_lat; i++) sum += lat[i];
        if (n_lat > 0) {
            fprintf(stderr, "per-block call latency (snapshot + block + meter): median %.3f ms, mean %.3f ms, max %.3f ms over %ld blocks\n",
                    1e3 * lat[n_lat / 2], 1e3 * sum / (double)n_lat, 1e3 * lat[n_lat - 1], n_lat);
        }
        free(lat);
        k = 0;
    } else
    for (k = 0;; k++) {
        /* file to file: the engine keeps up to three calls in flight (its step graphs run forward(k), MAC(k-1) and
         * inverse(k-2) side by side); call k is submitted, then the output of call k-2 is written while k runs */
        const int s = k % 4;
        size_t got = text_io ? read_text(in, raw_in[s], in_bytes * (size_t)batch)
                             : fread(raw_in[s], 1, in_bytes * (size_t)batch, in);
        if (got == 0) {
            break;
        }
        nblk[s] = (int)((got + in_bytes - 1) / in_bytes);
        if (got < (size_t)nblk[s] * in_bytes) {
            memset((char *)raw_in[s] + got, 0, (size_t)nblk[s] * in_bytes - got);       /* dai.c:1312-1332 */
        }
        CHECK(bfcuda_process_blocks_async(eng, nblk[s], raw_in[s], raw_out[s]));
        if (k > 1) {
            const int p = (k - 2) % 4;
            CHECK(bfcuda_wait_previous(eng, 2));
            if (text_io ? write_text(out, raw_out[p], out_bytes * (size_t)nblk[p], n) != 0
                        : fwrite(raw_out[p], 1, out_bytes * (size_t)nblk[p], out) != out_bytes * (size_t)nblk[p]) {
                DIE("write failed: %s", strerror(errno));
            }
        }
        blocks += nblk[s];
    }
    if (k > 0) {
        int q;
        CHECK(bfcuda_synchronize(eng));
        for (q = k > 1 ? k - 2 : 0; q < k; q++) {
            const int p = q % 4;
            if (text_io ? write_text(out, raw_out[p], out_bytes * (size_t)nblk[p], n) != 0
                        : fwrite(raw_out[p], 1, out_bytes * (size_t)nblk[p], out) != out_bytes * (size_t)nblk[p]) {
                DIE("write failed: %s", strerror(errno));
            }
        }
    }
    t1 = now();
    if (out != stdout) fclose(out);

    for (c = 0; c < n; c++) {
        struct bfcuda_overflow of;
        CHECK(bfcuda_get_overflow(eng, c, &of));
        if (of.n_overflows > 0) {
            fprintf(stderr, "output %d: %u overflows, peak %.2f dB\n", c, of.n_overflows,
                    20.0 * log10(of.largest / of.max));     /* bfrun.c:555-618 */
        }
    }
    if (blocks > 0) {
        const double per_block = (t1 - t0) / (double)blocks, block_s = (double)L / (double)rate;
        fprintf(stderr, "%ld blocks, %.3f ms/block wall (file I/O included), rti %.4f, realtime multiple %.1f\n", blocks,
                per_block * 1e3, per_block / block_s, block_s / per_block);
        if (bench) {
            CHECK(bfcuda_stage_times(eng, stage, &stage_blocks, &launches));
            fprintf(stderr, "  device ms per period | raw2real+time2freq+mixscale1 | convolve | mixscale2+freq2time+real2raw\n"
                            "                       | %28.3f | %8.3f | %28.3f   (%ld kernel launches)\n",
                    stage[0], stage[1], stage[2], launches);
        }
    }
    for (k = 0; k < 4; k++) {
        bfcuda_host_free(raw_in[k]);
        bfcuda_host_free(raw_out[k]);
    }
    bfcuda_destroy(eng);
    return 0;
}

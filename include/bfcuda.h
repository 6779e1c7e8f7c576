/*
 * bfcuda.h -- C ABI of the B200-native BruteFIR convolution engine (libbfcuda.so).
 *
 * This is the drop-in boundary for ONE path of the reference: the uniformly partitioned
 * overlap-save convolver behind /root/reference/convolver.h plus the raw2real / real2raw sample
 * conversion feeding it.  Two surfaces are exported:
 *
 *  (1) bfcuda_convolver.h -- the reference's own per-call interface (convolver.h:16-152), same names,
 *      argument meaning and error behaviour, host pointers in and out, every call executed by the
 *      CUDA kernels.  It exists for link compatibility and for buffer-for-buffer parity tests.
 *
 *  (2) this header -- the block-level interface the (unchanged, C) host side calls once per audio
 *      block instead of the ~20 per-filter calls of filter_process() (bfrun.c:1420-2083).  One call
 *      covers bfrun.c:1494-2006: raw2cbuf + time2freq for every input, mixnscale(INPUT) into the
 *      frequency-domain delay line, convolve / convolve_add over all partitions (with dirac and
 *      crossfade variants), mixnscale(OUTPUT), freq2time and cbuf2raw for every output.
 *
 * Plain C types only; no CUDA or torch types cross the boundary.  All functions return 0 on success
 * and a negative BFCUDA_E* code on failure; bfcuda_strerror() gives the message of the calling
 * thread's last failure.  There is no CPU fallback: without a CUDA device every entry point that
 * computes fails with BFCUDA_ENODEV.
 *
 * One thread may call into one engine at a time (the reference's filter processes are single
 * threaded).  CUDA is initialised lazily by bfcuda_create(), i.e. after any fork()
 * the host performs (bfconf_init runs before bfrun forks, bfconf.c:2786 / bfrun.c:2312).
 */
#ifndef BFCUDA_H
#define BFCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFCUDA_IN 0     /* BF_IN,  bfmod.h:30 */
#define BFCUDA_OUT 1    /* BF_OUT, bfmod.h:31 */

#define BFCUDA_MAXCHANNELS 256  /* BF_MAXCHANNELS, bfmod.h:22 */
#define BFCUDA_MAXFILTERS 256   /* BF_MAXFILTERS,  bfmod.h:23 */

/* error codes */
#define BFCUDA_OK 0
#define BFCUDA_EINVAL (-1)      /* bad argument / unsupported configuration */
#define BFCUDA_ENODEV (-2)      /* no CUDA device / driver */
#define BFCUDA_ECUDA (-3)       /* CUDA runtime error */
#define BFCUDA_ENOMEM (-4)
#define BFCUDA_ENONFINITE (-5)  /* NaN/Inf in output (reference: abort(), real2raw.h:27-31) or coefficients */
#define BFCUDA_ESAFETY (-6)     /* safety limit exceeded (reference: bf_exit, real2raw.h:32-41) */
#define BFCUDA_ENOTSUP (-7)     /* on the reference surface but outside the accelerated path */
#define BFCUDA_ECOMM (-8)       /* NCCL failure */

/* struct sample_format, dai.h:21-28 (same field order and meaning) */
struct bfcuda_sample_format {
    int isfloat;
    int swap;       /* byte order differs from the (little endian) host */
    int bytes;      /* storage bytes per sample: 1, 2, 3, 4, 8 */
    int sbytes;     /* significant bytes; bits = 8 * sbytes */
    double scale;   /* 2^-(8*sbytes-1) for integers, 1.0 for floats (bfconf.c:473-477) */
    int format;     /* BF_SAMPLE_FORMAT_*, informational */
};

/* struct buffer_format, dai.h:30-34: one channel inside an interleaved or planar raw block */
struct bfcuda_buffer_format {
    struct bfcuda_sample_format sf;
    int sample_spacing;     /* in samples */
    int byte_offset;        /* in bytes */
};

/* struct bfoverflow, bfmod.h:99-104 */
struct bfcuda_overflow {
    unsigned int n_overflows;
    int32_t intlargest;
    double largest;
    double max;
};

/* struct bffilter (bfmod.h:118-126) + its initial struct bffilter_control (bfmod.h:128-133) */
struct bfcuda_filter {
    int crossfade;
    int n_channels[2];          /* [BFCUDA_IN] inputs mixed into this filter, [BFCUDA_OUT] outputs fed */
    const int *channels[2];     /* virtual channel indices */
    const double *scale[2];     /* fctrl.scale[IN|OUT][i], linear multipliers */
    int n_filters_in;           /* filter->filter chaining (to_filters / from_filters, bfrun.c:1603-1660) */
    const int *filters_in;      /* source filters; filters are listed in processing order, producers first */
    const double *fscale;       /* fctrl.fscale[i], NULL = all 1.0 */
    int coeff;                  /* initial coefficient set, -1 = no coefficients (dirac) */
    int delayblocks;            /* initial delay in blocks */
};

/* run-time control of one filter: struct bffilter_control, bfmod.h:128-133; snapshotted per block
 * exactly like bfrun.c:1462-1478.  NULL scale pointers leave the current scales untouched. */
struct bfcuda_filter_control {
    int coeff;
    int delayblocks;
    const double *scale[2];
    const double *fscale;       /* multipliers of the source filters (from_filters), NULL = unchanged */
};

struct bfcuda_config {
    int filter_length;          /* L, power of two (bfconf.c:1495-1520) */
    int n_blocks;               /* P partitions */
    int realsize;               /* 4 (float_bits 32) or 8 (float_bits 64) */
    int n_channels[2];
    const struct bfcuda_buffer_format *formats[2];  /* per virtual channel (virtual:physical 1:1) */
    int n_bytes[2];             /* bytes of one raw input / output block (dai_buffer_format[IO]->n_bytes) */
    int n_filters;
    const struct bfcuda_filter *filters;    /* in processing order */
    int n_coeffs;
    const int *coeff_n_blocks;  /* bfcoeff.n_blocks per coefficient set (<= n_blocks) */
    double safety_limit;        /* bfconf->safety_limit, 0 = off */
    int device;                 /* CUDA device ordinal */
    unsigned int flags;         /* BFCUDA_FLAG_* */
    int mac_split;              /* 0 = automatic; 1 = never split the partition sum (reference summation order);
                                   S > 1 = split it S ways */
    const int *apply_dither;    /* per output channel: dither requested (`dither: true` of its device, bfconf.c:3173-3217);
                                   NULL = none.  Like bfconf, the engine silently drops the request for float formats,
                                   for more than 16 bit at realsize 4 and for 32 bit formats.  HP-TPDF dither with
                                   first-order error feedback, dither_funs.h:7-68: a sequential recurrence per channel. */
    int sampling_rate;          /* only sizes the dither table (dither.c:75-96); 0 = 44100 */
    int max_dither_table_size;  /* bfconf->max_dither_table_size, 0 = no limit */
    int max_batch;              /* 0/1 = block by block (the reference's schedule).  B > 1 (<= 16 at realsize 4, <= 8 at
                                   realsize 8) lets bfcuda_process_blocks* take up to B consecutive blocks per call:
                                   offline / file-to-file throughput mode.  Results are bit-identical to B single
                                   calls; the I/O delay grows by the batch (not for real-time use). */
    int powersave;              /* bfconf->powersave (bfrun.c:1541-1552, 1613-1700): a silent input frame skips its transform
                                   and leaves a zero delay-line slot, and the multiply-accumulate skips zero slots -- on
                                   this engine: does not READ them, nor the coefficient blocks they would meet.  0 = off */
    double analog_powersave;    /* bfconf->analog_powersave as a linear level: a frame whose peak (times the sample
                                   format's scale) is below it is made truly zero (bfrun.c:722-772).  >= 1.0 (or 0) =
                                   only frames of exact zeros count as silent */
    const int *out_physical;    /* per virtual output channel: its physical channel, bfconf->virt2phys[OUT].  Outputs with
                                   the same id are added sample by sample in the time domain, in channel order, and
                                   quantised once (bfrun.c:1937-2002); they must carry the same buffer format, and they
                                   share one overflow record.  NULL = every output has a physical channel of its own.
                                   (Several virtual INPUTS on one physical channel need nothing: give them the same
                                   buffer format.) */
};

#define BFCUDA_FLAG_STAGE_TIMING 1u     /* record CUDA events around each stage of every block */
#define BFCUDA_FLAG_NO_GRAPH 2u         /* (reserved) */
#define BFCUDA_FLAG_KEEP_INPUT_SPECTRA 4u   /* keep every input's unscaled spectrum for bfcuda_debug_read */
#define BFCUDA_FLAG_NO_STREAM_SHARING 16u  /* give every filter its own delay line even where several filters are fed by
                                           the same input with the same scale and delay (they normally share one) */
#define BFCUDA_FLAG_LOW_LATENCY 32u     /* real-time schedule for block-by-block calls: the partitions 1 .. P-1 of the NEXT
                                           block only need spectra that are already in the delay line, so their sum is
                                           computed right after a block's output has left (while the host waits for the
                                           next input); when that input arrives only partition 0 is multiplied and the
                                           two partial sums are added.  The partition sum then is head + tail instead of
                                           the reference's left-to-right order (like mac_split = 2; within the parity
                                           tolerances, not bit-identical), the call latency drops by the MAC's time.
                                           Control changes, crossfades and batches fall back to the full sum. */
#define BFCUDA_FLAG_SERIAL_STAGES 8u    /* do not overlap the stages of consecutive launches (the engine normally runs
                                           launch n+1's forward and launch n-1's inverse stage beside launch n's
                                           multiply-accumulate): stage timings then are each stage running alone */

typedef struct bfcuda_engine bfcuda_engine;

const char *bfcuda_strerror(void);
int bfcuda_device_count(void);

int bfcuda_create(const struct bfcuda_config *config, bfcuda_engine **engine);
void bfcuda_destroy(bfcuda_engine *engine);

/* ---- coefficients ---------------------------------------------------------------------------- */

/* load_coeff + convolver_coeffs2cbuf (bfconf.c:1992-2019, fftw_convolver.c:526-573): split n_taps
 * reals (realsize bytes each) into blocks of L, scale, zero-pad, transform on the device.
 * Fails with BFCUDA_ENONFINITE on NaN/Inf, like the reference returning NULL. */
int bfcuda_coeff_from_taps(bfcuda_engine *engine, int coeff, const void *taps, int n_taps, double scale);
/* upload one block already in the reference's processed layout (the "processed" coefficient format,
 * bfconf.c:1924-1957, and what bfaccess->coeffs_data exposes) */
int bfcuda_coeff_set_block(bfcuda_engine *engine, int coeff, int block, const void *cbuf);
/* read one block back in the reference's processed layout */
int bfcuda_coeff_get_block(bfcuda_engine *engine, int coeff, int block, void *cbuf);
/* convolver_runtime_coeffs2cbuf (fftw_convolver.c:575-596): L taps -> one block, while running */
int bfcuda_coeff_runtime_block(bfcuda_engine *engine, int coeff, int block, const void *taps_L);

/* ---- control --------------------------------------------------------------------------------- */

int bfcuda_set_control(bfcuda_engine *engine, int filter, const struct bfcuda_filter_control *control);
/* Sub-sample delay of one channel (io = BFCUDA_IN: the postprocess hook of convolver_raw2cbuf, bfrun.c:1503-1526;
 * BFCUDA_OUT: before mixing and quantisation, bfrun.c:1918-1925).  `taps` are the n_taps = 2 * sdf_length + 1 reals
 * (engine precision) of the windowed sinc of the channel's current delay step -- what the host's unchanged delay.c /
 * firwindow.c hand to convolver_td_new (delay.c:486-499); the engine applies them as the causal FIR the td convolver
 * computes.  taps == NULL switches the channel's sub-sample delay off.  Takes effect at the next block; filter history
 * is kept across changes like delay.c's `rest`.  n_taps <= 1025. */
int bfcuda_set_subdelay(bfcuda_engine *engine, int io, int channel, const void *taps, int n_taps);
/* Mute of a virtual channel (icomm->ismuted, bfrun.c:1510-1525, 1953): a muted input reads as silence, a muted output
 * contributes nothing to its physical channel. */
int bfcuda_set_mute(bfcuda_engine *engine, int io, int channel, int muted);
int bfcuda_get_overflow(bfcuda_engine *engine, int out_channel, struct bfcuda_overflow *overflow);
int bfcuda_reset_overflow(bfcuda_engine *engine);

/* ---- the block step -------------------------------------------------------------------------- */

/* One audio block, host buffers laid out as the dai buffers are (dai.c:537-576): raw_in holds
 * n_bytes[IN] bytes, raw_out receives n_bytes[OUT] bytes.  Copies in, runs the three stages, copies
 * out and waits.  raw_in / raw_out should be page-locked (bfcuda_host_alloc) for full PCIe speed. */
int bfcuda_process_block(bfcuda_engine *engine, const void *raw_in, void *raw_out);

/* Pipelined form for offline / throughput use: enqueue only; raw_out is complete after
 * bfcuda_synchronize().  The caller must keep both buffers valid and unmodified until then. */
int bfcuda_process_block_async(bfcuda_engine *engine, const void *raw_in, void *raw_out);
int bfcuda_synchronize(bfcuda_engine *engine);

/* Batched form (config.max_batch > 1): n_blocks (1..max_batch) consecutive blocks, stored back to back
 * (n_bytes[IN] / n_bytes[OUT] apart), in one call.  Each stage runs once for the whole batch and the
 * multiply-accumulate reuses every coefficient and delay-line spectrum across the batch in registers, so
 * its HBM traffic per block drops ~n_blocks-fold; every output block is still computed with the reference's
 * operation order (bit-identical to n_blocks single-block calls).  A pending bfcuda_set_control() or a
 * crossfade block is split off and processed on its own. */
int bfcuda_process_blocks(bfcuda_engine *engine, int n_blocks, const void *raw_in, void *raw_out);
int bfcuda_process_blocks_async(bfcuda_engine *engine, int n_blocks, const void *raw_in, void *raw_out);
/* Wait until the output of a recent asynchronous call is complete in host memory without draining the pipeline:
 * calls_back = 0 is the most recent bfcuda_process_block[s]_async call, 1 the one before it (the engine keeps two
 * calls in flight).  A file-to-file host submits call k, then waits for call k-1 and writes its output while call k
 * runs (host/bfcuda_run.c). */
int bfcuda_wait_previous(bfcuda_engine *engine, int calls_back);

/* Device-resident form: input already in the engine's device staging buffer (see bfcuda_device_io),
 * output left in the device output buffer; no host<->device copy.  Enqueue only. */
int bfcuda_process_block_device(bfcuda_engine *engine);
int bfcuda_process_blocks_device(bfcuda_engine *engine, int n_blocks);
int bfcuda_device_io(bfcuda_engine *engine, int io, void **device_ptr, size_t *n_bytes);
int bfcuda_upload_input(bfcuda_engine *engine, const void *raw_in);
int bfcuda_upload_inputs(bfcuda_engine *engine, int n_blocks, const void *raw_in);
int bfcuda_download_output(bfcuda_engine *engine, void *raw_out);
int bfcuda_download_outputs(bfcuda_engine *engine, int n_blocks, void *raw_out);

void *bfcuda_host_alloc(size_t n_bytes);    /* page-locked host memory */
/* the same, placed on the NUMA node the CUDA device `device` is attached to (multi-GPU hosts: one process per GPU) */
void *bfcuda_host_alloc_near(int device, size_t n_bytes);
void bfcuda_host_free(void *p);
/* Copy-only baseline of the host-buffer path (no kernels): `reps` rounds of "n_blocks input blocks host -> device" and
 * "n_blocks output blocks device -> host" on the engine's copy streams, both directions concurrently.  Measures what
 * the bus and the host side can carry for this engine's block sizes (bench.py runs it on all ranks at once). */
int bfcuda_copy_baseline(bfcuda_engine *engine, int n_blocks, const void *raw_in, void *raw_out, int reps,
                         double *ms_per_rep, double *h2d_gbs, double *d2h_gbs);

/* ---- measurement ----------------------------------------------------------------------------- */

#define BFCUDA_STAGE_FORWARD 0  /* raw2real + R2C FFT + input mix  (raw2real, time2freq, mixscale1 columns) */
#define BFCUDA_STAGE_MAC 1      /* delay-line multiply-accumulate   (convolve column) */
#define BFCUDA_STAGE_INVERSE 2  /* output mix + C2R FFT + real2raw  (mixscale2, freq2time, real2raw columns) */
#define BFCUDA_N_STAGES 3

/* CUDA-event stopwatch on the engine's stream (torch.cuda.Event cannot see this stream). */
int bfcuda_timer_start(bfcuda_engine *engine);
int bfcuda_timer_stop(bfcuda_engine *engine, double *elapsed_ms);   /* synchronises */
/* With BFCUDA_FLAG_STAGE_TIMING: mean device milliseconds per block of each stage and the number of
 * kernel launches since the last call; resets the accumulators. */
int bfcuda_stage_times(bfcuda_engine *engine, double mean_ms[BFCUDA_N_STAGES], long *n_blocks,
                       long *n_kernel_launches);
/* switch BFCUDA_FLAG_STAGE_TIMING at run time (synchronises).  The per-stage events cost a few percent of
 * throughput (they are recorded on three streams per launch): time the whole job without them. */
int bfcuda_set_stage_timing(bfcuda_engine *engine, int on);
/* switch BFCUDA_FLAG_SERIAL_STAGES at run time (takes effect at the next launch) */
int bfcuda_set_serial_stages(bfcuda_engine *engine, int on);
/* static facts about the engine for roofline arithmetic */
struct bfcuda_info {
    int n_fft;                  /* N = 2 L */
    int mac_split;              /* partition-sum split actually used */
    int n_streams;              /* distinct delay-line streams (U in SURVEY.md 8(d)) */
    int kernels_per_block;
    int uses_graph;
    int sm_count;
    size_t mac_bytes_per_block; /* algorithmic: rs * N * (P*F + P*U + F), current coefficient lengths */
    size_t device_bytes;        /* device memory held by the engine */
    char device_name[64];
    int max_batch;
    size_t mac_bytes_per_batch; /* compulsory bytes of one full batch: rs * N * (P*F + (P+B-1)*U + B*F) */
};
int bfcuda_get_info(bfcuda_engine *engine, struct bfcuda_info *info);

/* ---- introspection for parity tests (reference layouts on the host side) ------------------------ */

#define BFCUDA_DBG_INPUT_SPECTRUM 1     /* input_freqcbuf[ch], FFTW half-complex order (bfrun.c:1547) */
#define BFCUDA_DBG_DELAYLINE 2          /* cbuf[filter][slot], blocked layout (bfrun.c:1671) */
#define BFCUDA_DBG_FILTER_OUTPUT 3      /* ocbuf[filter], blocked layout (bfrun.c:1737-1754) */
#define BFCUDA_DBG_OUTPUT_TIME 4        /* first L reals of the inverse transform of output ch (bfrun.c:1887) */
int bfcuda_debug_read(bfcuda_engine *engine, int what, int index, int slot, void *dst_N_reals);

/* ---- multi-GPU (one engine per process and GPU) ---------------------------------- */

#define BFCUDA_COMM_ID_BYTES 128
/* rank 0 creates an id, the host broadcasts it (any transport), every rank calls comm_init. */
int bfcuda_comm_unique_id(void *id_128_bytes);
int bfcuda_comm_init(bfcuda_engine *engine, int rank, int n_ranks, const void *id_128_bytes);
/* Mark output channels whose feeding filters live on several ranks: their time-domain blocks are
 * summed over NVLink (ncclAllReduce, realsize floats x L) between the inverse FFT and quantisation. */
int bfcuda_comm_shared_outputs(bfcuda_engine *engine, int n_shared, const int *out_channels);

#ifdef __cplusplus
}
#endif
#endif

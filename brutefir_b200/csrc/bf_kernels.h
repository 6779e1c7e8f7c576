// bf_kernels.h -- host-callable launchers of the sm_100a kernels (bf_kernels.cu).
//
// Plain C++ declarations; every pointer is a DEVICE pointer unless named host_*.  `realsize` selects
// the float (4) or double (8) instantiation, mirroring the reference's realsize switch
// (fftw_convolver.c:42, 796-808).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "bf_common.cuh"

namespace bf {

// ---- descriptors living in device memory ----------------------------------------------------------

// One destination of an input channel's spectrum: delay-line stream `stream`, multiplied by `scale`
// (= (real_t)(fctrl.scale[IN] * sf.scale), bfrun.c:1664-1665 + fftw_convfuns.h:18-20), written to ring
// slot (t + delay) % P (bfrun.c:1600).
struct FwdDest {
    int stream;
    int delay;
    double scale;   // already rounded to the real type
};

// A delay-line stream that mixes several inputs (mixnscale INPUT with n_bufs > 1)
struct MixStream {
    int stream;
    int delay;
    int n_inputs;
    int first;      // index into the (channel, scale) term arrays
};
struct MixTerm {
    int index;      // input channel (forward mix) or y-slot (output mix)
    double scale;   // already rounded to the real type
};

// One multiply-accumulate job: out[slot] = sum_{i < n_parts} FDL[stream][(t - i) % P] (*) H[hbase + i]
// (bfrun.c:1737-1754), or the dirac short cut (bfrun.c:1809, fftw_convfuns.h:606-619) when hbase < 0.
struct MacJob {
    int stream;
    int hbase;      // first coefficient block (units of N reals), -1 = dirac
    int n_parts;
    int out;        // y-slot
};

// One output channel of the inverse stage: terms [first, first + n) of the "new" mix and, when a
// feeding filter is crossfading this block, terms [xf_first, xf_first + n) of the "old" mix.
struct OutChan {
    int first;
    int n;
    int xf_first;   // -1 = no crossfade this block
    int shared;     // bit 0: summed across ranks before quantisation; bit 1: dithered (k_dither quantises it);
                    // bit 2: a virtual output mixed into another one's physical channel (k_virt_mix): never packed;
                    // any bit: the inverse stage / k_pack stop after the time-domain store
};

struct FftPlan {
    int N;          // real length = 2 L
    int realsize;
    void *tw;       // device: N/2 complex roots e^{-2 pi i j / N}
    void *tw2;      // device: per-pass tables of the size-specialised transform (bf_fft2.cuh), or NULL
    // transforms longer than one thread block holds (bf_fft4.cu): M = big_m1 x big_m2, 0 = not in use
    int big_m1, big_m2;
    void *big_tw1, *big_tw2;        // root tables of the two sub-transform sizes
    void *big_scr0, *big_scr1;      // scratch, big_items x M complex each
    void *big_old;                  // big_items x L reals: the old-coefficient signal of a crossfade block
    int big_items;                  // transforms per chunk
};

// big_items_hint: the largest number of transforms one launch will ask for (sizes the four-step scratch)
cudaError_t fft_plan_create(FftPlan *plan, int N, int realsize, int big_items_hint = 64);
// bf_fft4.cu: the four-step transform for N beyond one block (L >= 32768 float / 16384 double)
bool fft_big_supported(int N, int realsize);
cudaError_t fft_big_plan_create(FftPlan *plan, int max_items);
void fft_big_plan_destroy(FftPlan *plan);
cudaError_t launch_forward_big(const FftPlan &plan, const struct ForwardArgs &a, cudaStream_t s);
cudaError_t launch_inverse_big(const FftPlan &plan, const struct InverseArgs &a, cudaStream_t s);
cudaError_t launch_coeff_fft_big(const FftPlan &plan, const void *taps, int n_blocks, double scale, void *H, int hbase,
                                 cudaStream_t s);
// the forward / inverse stages read unpacked planar samples (k_unpack before, k_pack after): true for the
// size-specialised and the four-step transforms
inline bool plan_unpacks_first(const FftPlan &plan) { return plan.tw2 != nullptr || plan.big_m1 != 0; }
// one block holds the transform in shared memory (the generic kernels and the per-call convolver.h surface)
bool fft_single_block_supported(int N, int realsize);
// bf_fft2_kernels.cu: the size-specialised forward / inverse stages (float, 1024 <= L <= 16384)
bool fft2_supported(int N, int realsize);
cudaError_t fft2_plan_create(FftPlan *plan);
cudaError_t launch_forward2(const FftPlan &plan, const struct ForwardArgs &a, cudaStream_t s);
cudaError_t launch_inverse2(const FftPlan &plan, const struct InverseArgs &a, cudaStream_t s);
void fft_plan_destroy(FftPlan *plan);
bool fft_size_supported(int N, int realsize);

// ---- engine kernels --------------------------------------------------------------------------------

// Batching: one launch may process `batch` consecutive audio blocks (grid y = block within the batch).
// The delay-line ring has `ring` = 2 P + 2 max_batch slots per stream (bf_engine.cu, bfcuda_create), so that the
// spectra of the later blocks of a batch -- and of the next launch -- do not overwrite slots the earlier blocks still
// read; `t` is the ring slot of the batch's first block (0 <= t < ring).
struct ForwardArgs {
    const uint8_t *raw_in;      // block b at raw_in + b * in_stride
    const SampleFormat *fmt;    // [n_in]
    const void *prev_in;        // [n_in][L] reals: the block before the batch's first one
    void *prev_out;             // [n_in][L] reals: receives the batch's last block (never aliases prev_in when batch > 1)
    void *fdl;                  // [U][ring][N] planar spectra
    void *xin;                  // [batch][n_in][N] unscaled planar spectra, or NULL
    const uint8_t *need_xin;    // [n_in]
    const int *dest_first;      // [n_in + 1]
    const FwdDest *dests;
    int n_in;
    int n_vin;                  // rows of one block of xin: the inputs plus the evaluated filter outputs (k_eval)
    int ring;
    int t;
    int batch;
    size_t in_stride;
    int fast_fmt;               // 1: every input is an aligned 4-byte LE integer, 2: FLOAT_LE, 3: packed S24_LE in whole 32-channel tiles, 0: per-sample generic decode
    // size-specialised path (bf_fft2_kernels.cu): the samples arrive unpacked
    const void *xt_cur;         // [batch][n_in][L] reals, this launch's blocks
    const void *xt_prev;        // [n_in][L] reals, the block before the first one
    int single_dest;            // every input feeds exactly one delay-line stream and no input spectrum is kept
    // powersave (bfrun.c:1541-1552, 722-772): a frame [previous block | this block] whose peak is zero (mode 1) or
    // below the channel's level (mode 2) leaves zero spectra, and its delay-line slots are flagged for the MAC
    int powersave;              // 0 = off
    const unsigned int *amax_cur;   // [batch][n_in] peak |sample| of this launch's blocks, float bits (k_unpack)
    const unsigned int *amax_prev;  // [n_in] the same for the block before the first one
    const float *ps_thr;        // [n_in] mode 2: silent iff peak < ps_thr (= analog_powersave / sf.scale)
    uint8_t *slot_zero;         // [U][ring] 1 = the slot holds zeros (the MAC neither reads it nor its coefficients)
};

struct UnpackArgs {
    const uint8_t *raw_in;      // block b at raw_in + b * in_stride
    const SampleFormat *fmt;    // [n_in]
    void *xt;                   // [batch][n_in][L] reals
    int n_in;
    int batch;
    int L;
    size_t in_stride;
    int fast_fmt;
    unsigned int *amax;         // [batch][n_in] running peak |sample| as float bits (zeroed before the launch), or NULL
    const uint8_t *muted;       // [n_in] muted inputs read as zeros (bfrun.c:1523-1525), or NULL
};
cudaError_t launch_unpack(const FftPlan &plan, const UnpackArgs &a, cudaStream_t s);
cudaError_t launch_forward(const FftPlan &plan, const ForwardArgs &a, cudaStream_t s);

struct StreamMixArgs {
    const void *xin;            // [batch][n_in][N]; n_in here counts the virtual inputs (ForwardArgs::n_vin)
    void *fdl;
    const MixStream *streams;
    const MixTerm *terms;
    int n_streams;
    int n_in;
    int ring;
    int t;
    int batch;
    uint8_t *slot_zero;         // mixed slots are marked "not zero" (conservative), or NULL
};
cudaError_t launch_stream_mix(const FftPlan &plan, const StreamMixArgs &a, cudaStream_t s);

// Filter -> filter chaining (to_filters / from_filters): the mixed outputs of the source filters are evaluated in
// the time domain and become one more input spectrum of the consuming filter -- bfrun.c:1603-1660,
// convolver_convolve_eval (fftw_convolver.c:411-433): HC2R, frame = [previous valid block | this valid block], R2HC.
struct EvalEntry {
    int first;      // terms [first, first + n): (Y slot, fscale) of the source filters
    int n;
    int xf_first;   // terms of the "old coefficient" mix while a source filter crossfades this block, else -1
    int vin;        // row of xin that receives the evaluated spectrum (>= n_in); keep row = vin - n_in
};
struct EvalArgs {
    const void *Y;              // [split][batch][n_slots][N]
    const EvalEntry *entries;
    const MixTerm *terms;
    void *keep;                 // [n_eval][L] reals: previous valid block of every evaluated mix
    void *xin;                  // [batch][n_vin][N]
    int n_entries;
    int n_in;
    int n_vin;
    int n_slots;
    int split;
    int batch;
};
cudaError_t launch_eval(const FftPlan &plan, const EvalArgs &a, cudaStream_t s);

struct MacArgs {
    const void *fdl;
    const void *H;
    void *Y;                    // [split][batch][n_slots][N]
    const MacJob *jobs;
    int n_jobs;
    int n_slots;
    int ring;
    int split;
    int t;
    int batch;
    int variant;                // 0 = direct streaming loads, 1 = bulk-copy (TMA) staged (batch 1 only)
    unsigned long long neg_zero2;   // filled in by the batched launcher: two packed -0.0f (bf_mac_batch.cu, BinPairAcc)
    // Uneven two-way split for the low-latency schedule (batch 1, variant 0, split == 2): head > 0 makes partial 0
    // the partitions [0, head) and partial 1 the rest; z_count > 0 launches only the partials z_first .. z_first +
    // z_count - 1 (the "rest" is computed one block ahead, bf_engine.cu).
    int head, z_first, z_count;
    const uint8_t *slot_zero;   // [U][ring] powersave: slots flagged 1 are not read (nor the coefficients they meet), or NULL
};
cudaError_t launch_mac(const FftPlan &plan, const MacArgs &a, cudaStream_t s);
// The host stub of the kernel the most recent launch_forward / launch_mac call of this thread launched: how the engine
// finds those two nodes in a captured step graph (their arguments carry the ring slot, which changes per launch).
extern thread_local const void *g_last_func;
// the batched MAC with block-cooperative, bulk-copy staged operands (bf_mac_tile.cu)
bool mac_tile_applicable(const FftPlan &plan, const MacArgs &a);
bool mac_coop_by_default(const FftPlan &plan, const MacArgs &a);
cudaError_t launch_mac_tile(const FftPlan &plan, const MacArgs &a, cudaStream_t s);
// bins per thread the batched kernel will use for such a launch (bf_mac_batch.cu)
int mac_batch_lanes(int realsize, int batch, int n_jobs, int N);
// With split > 1 the MAC leaves `split` partial sums per output: add them, in order, into partial 0 (the consumers
// -- inverse stage, chaining evaluation -- then read one complete spectrum per filter).
cudaError_t launch_split_reduce(const FftPlan &plan, const MacArgs &a, cudaStream_t s);

struct InverseArgs {
    const void *Y;              // [split][batch][n_slots][N]
    const OutChan *chans;       // [n_out]
    const MixTerm *terms;
    void *out_time;             // [batch][n_out][L] reals (time domain, LSB units)
    uint8_t *raw_out;           // block b at raw_out + b * out_stride
    const SampleFormat *fmt;    // [n_out]
    Overflow *overflow;         // [n_out]
    unsigned int *status;       // BF_STATUS_* bits
    int n_out;
    int n_slots;
    int split;
    int batch;
    size_t out_stride;
    double safety_limit;
    int fast_fmt;               // as ForwardArgs::fast_fmt, for the outputs
    int simple_mix;             // every output is one filter's output, unsplit partition sum, no crossfade pending
    int any_xfade;              // some output of this launch crossfades (two transforms per output, bf_fft2_kernels.cu)
};
// mixnscale(OUTPUT) as a kernel of its own (fftw_convfuns.h:268-494, bfrun.c:1847-1868) for outputs fed by several
// filters: Z[o] = sum_j scale_j Y[slot_j], left to right, written to Y slot z_first + o (and the "old coefficient" mix
// of a crossfade block to z_first + n_out + o).  The inverse stage then reads ONE spectrum per output with scale 1.
struct OutMixArgs {
    void *Y;                    // [batch][n_slots][N] (partition split already reduced)
    const OutChan *mixes;       // [n_out]: first / n / xf_first of the filter terms; outputs with n <= 1 are skipped
    const MixTerm *terms;
    int n_out;
    int n_slots;
    int z_first;
    int batch;
};
cudaError_t launch_out_mix(const FftPlan &plan, const OutMixArgs &a, cudaStream_t s);

// size-specialised path: the inverse stage stops at out_time; this quantises / packs ALL channels from it
cudaError_t launch_pack(const FftPlan &plan, const InverseArgs &a, cudaStream_t s);
cudaError_t launch_inverse(const FftPlan &plan, const InverseArgs &a, cudaStream_t s);

// HP-TPDF dither with first-order error feedback (dither.c, dither.h:28-38, dither_funs.h:7-68): sequential per
// channel, one lane per dithered output walks its L samples (and the blocks of a batch) in order.
struct DitherChan {         // device-resident state of one dithered output (struct dither_state, dither.h:17-22)
    int out;                // output channel
    int randtab_ptr;
    double e0, e1;          // error feedback state sf[0..1] / sd[0..1], kept in the real type's precision
};
struct DitherArgs {
    DitherChan *chans;      // [n_dither]
    const int8_t *randtab;  // dither_randtab
    const void *randmap;    // 512 reals: map[d + 256] for d = randtab[n] - randtab[n-1]
    int randtab_size;
    int n_dither;
};
cudaError_t launch_dither(const FftPlan &plan, const InverseArgs &a, const DitherArgs &d, cudaStream_t s);

// Sub-sample delay (delay_subsample_update, delay.c:415-442, the postprocess hook of convolver_raw2cbuf, bfrun.c:1503-1526,
// and the output side, bfrun.c:1918-1925): the reference pushes the L new samples of a channel through its td convolver
// in overlap-save blocks -- a causal FIR over the channel's stream, y[n] = sum_k h[k] x[n - k], with the windowed-sinc taps
// of the channel's current delay step.  Here: that FIR in place on the unpacked samples / the inverse transform's output,
// one thread block per delayed channel (the blocks of a batch in order; `hist` carries the last n_taps - 1 inputs).
struct SubdelayChan {
    int ch;             // channel (row of the sample buffer)
    int tap_first;      // first tap in `taps`
    int n_taps;         // <= BF_SUBDELAY_MAX_TAPS
};
#define BF_SUBDELAY_MAX_TAPS 1025
struct SubdelayArgs {
    void *data;                 // [batch][n_ch][L] reals, filtered in place
    const SubdelayChan *chans;  // [n_chans]
    const void *taps;           // reals
    void *hist;                 // [n_ch][BF_SUBDELAY_MAX_TAPS - 1] reals, row = channel
    int n_chans, n_ch, batch, L;
};
cudaError_t launch_subdelay(const FftPlan &plan, const SubdelayArgs &a, cudaStream_t s);

// Several virtual outputs on one physical channel (bfrun.c:1937-2002): their time-domain blocks are added sample by
// sample, in channel order, muted ones left out, into the row of the group's last member, which alone is quantised;
// a group of one only serves the mute.
struct VirtGroup {
    int first, n;       // members[first .. first + n): virtual output channels, ascending
};
struct VirtMixArgs {
    void *out_time;             // [batch][n_out][L]
    const VirtGroup *groups;
    const int *members;
    const uint8_t *muted;       // [n_out]
    int n_groups, n_out, batch, L;
};
cudaError_t launch_virt_mix(const FftPlan &plan, const VirtMixArgs &a, cudaStream_t s);

// quantise + pack out_time rows of the channels with chans[o].shared == 1 (after the cross-rank sum)
cudaError_t launch_quantise_shared(const FftPlan &plan, const InverseArgs &a, cudaStream_t s);

// coefficient preprocessing (convolver_coeffs2cbuf, fftw_convolver.c:526-573): taps[b][L] -> H block
// hbase + b; blocks are scaled by `scale`, placed in the upper half of a zero frame, transformed and
// multiplied by 1/N.
cudaError_t launch_coeff_fft(const FftPlan &plan, const void *taps, int n_blocks, double scale, void *H,
                             int hbase, cudaStream_t s);

// ---- layout permutation and the per-call (convolver.h) kernels ----------------------------------------
enum PermuteMode { BLOCKED_TO_PLANAR = 0, PLANAR_TO_BLOCKED = 1, HC_TO_PLANAR = 2, PLANAR_TO_HC = 3 };
cudaError_t launch_permute(const FftPlan &plan, const void *src, void *dst, int n_spectra, int mode,
                           cudaStream_t s);

cudaError_t launch_r2hc(const FftPlan &plan, const void *in, void *out, int batch, cudaStream_t s);
cudaError_t launch_hc2r(const FftPlan &plan, const void *in, void *out, int batch, cudaStream_t s);
// mixnscale on the reference layouts: mode 1 = INPUT (hc -> blocked), 3 = OUTPUT (blocked -> hc)
cudaError_t launch_cv_mixnscale(const FftPlan &plan, const void *const *in_ptrs, const double *scales,
                                int n_bufs, void *out, int mode, cudaStream_t s);
// blocked-layout complex multiply: op 0 = convolve (d = b*c), 1 = convolve_add (d += b*c)
cudaError_t launch_cv_convolve(const FftPlan &plan, const void *b, const void *c, void *d, int op,
                               cudaStream_t s);
cudaError_t launch_cv_dirac(const FftPlan &plan, const void *b, void *d, cudaStream_t s);
// convolver_td_* (fftw_convolver.c:682-782): ordered half-complex buffers of n reals, n = 2 * blocklen of ANY power of
// two >= 2.  launch_td_scale: buf *= scale.  launch_td_mul: the in-place ordered product (convolve_inplace_ordered).
// launch_td_small: n < 8 has no shared-memory transform; forward (dir 0) or inverse (dir 1) DFT by definition.
cudaError_t launch_td_scale(int realsize, void *buf, int n, double scale, cudaStream_t s);
cudaError_t launch_td_mul(int realsize, void *buf, const void *coeffs, int n, cudaStream_t s);
cudaError_t launch_td_small(int realsize, const void *in, void *out, int n, int dir, cudaStream_t s);
// crossfade ramp over the first L samples (fftw_convolver.c:349-355): nw = old*(1-f n) + nw*f n
cudaError_t launch_cv_xfade_blend(const FftPlan &plan, const void *old_time, void *new_time, cudaStream_t s);
cudaError_t launch_cv_raw2real(const FftPlan &plan, const uint8_t *raw, SampleFormat fmt, void *dst,
                               cudaStream_t s);
cudaError_t launch_cv_real2raw(const FftPlan &plan, const void *src, uint8_t *raw, SampleFormat fmt,
                               Overflow *overflow, unsigned int *status, double safety_limit, cudaStream_t s);

}  // namespace bf

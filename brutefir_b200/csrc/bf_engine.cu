// bf_engine.cu -- the block-level engine behind include/bfcuda.h.
//
// Host-side state machine that replaces the per-block body of filter_process()
// (/root/reference/bfrun.c:1420-2083): it keeps the frequency-domain delay lines, coefficient spectra,
// the run-time control snapshot, dither state and the overflow counters in HBM, and turns each call (one audio
// block, or a batch of up to eight) into a short chain of kernel launches on three software-pipelined streams:
//   main   k_unpack -> k_forward2 (or the fused k_forward) [-> k_stream_mix]
//   s_mac  k_mac / k_mac_batch2 [-> k_split_reduce] [-> per chaining level: k_eval, k_stream_mix, k_mac ...]
//   s_inv  [k_out_mix ->] k_inverse2 (or k_inverse) [-> ncclAllReduce] -> k_pack [-> k_dither]
//
// HBM layout (all "planar" spectra, bf_common.cuh; B = max_batch, ring = 2 P + 2 B):
//   H        [sum of coeff blocks][N]       coefficient spectra, pre-scaled by 1/N  (bfconf->coeffs_data)
//   FDL      [F][ring][N]                   delay-line rings, slot (t + delay) % ring written per block
//                                           (cbuf[n][n_blocks], bfrun.c:1045, 1273-1287); filters with the same
//                                           input, scale and delay share a ring (update_streams)
//   Y        [2][split][B][2F + 2 n_out][N] filter outputs (ocbuf[n]), two generations; slots F.. hold the "old
//                                           coefficient" outputs of a crossfade block (crossfadebuf[0]), the last
//                                           2 n_out the mixes of outputs fed by several filters (k_out_mix)
//   xt       [2][B][n_in][L]                unpacked input blocks, two generations (input_timecbuf,
//                                           fftw_convolver.c:180-193); prev [n_in][L] on the fused path
//   xin      [B][n_in + n_eval][N]          unscaled spectra of inputs feeding a multi-input mix, and the evaluated
//                                           outputs of source filters (to_filters chaining)
//   out_time [B][n_out][L]                  inverse-transform output, LSB units
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bfcuda.h"
#include "bf_kernels.h"

namespace bf {
bool mac_tma_applicable(const FftPlan &plan);
}

using namespace bf;

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            return fail(BFCUDA_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                       \
        }                                                                                                \
    } while (0)

// ---- NCCL, bound lazily so that single-GPU use has no NCCL dependency --------------------------------
typedef struct { char internal[128]; } nccl_uid_t;
typedef void *nccl_comm_t;
struct NcclApi {
    void *handle;
    int (*GetUniqueId)(nccl_uid_t *);
    int (*CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int);
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    int (*CommDestroy)(nccl_comm_t);
    const char *(*GetErrorString)(int);
};
static NcclApi g_nccl;

static int nccl_load(void)
{
    if (g_nccl.handle != nullptr) {
        return 0;
    }
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) {
        h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    }
    if (h == nullptr) {
        return fail(BFCUDA_ECOMM, "cannot load libnccl.so.2: %s", dlerror());
    }
    g_nccl.GetUniqueId = (int (*)(nccl_uid_t *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(nccl_comm_t *, int, nccl_uid_t, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce =
        (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.GroupStart = (int (*)(void))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)(void))dlsym(h, "ncclGroupEnd");
    g_nccl.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.GroupStart ||
        !g_nccl.GroupEnd || !g_nccl.CommDestroy) {
        return fail(BFCUDA_ECOMM, "libnccl.so.2 lacks required symbols");
    }
    g_nccl.handle = h;
    return 0;
}

// ---- engine state ------------------------------------------------------------------------------------

struct FilterState {
    int crossfade;
    std::vector<int> ch[2];
    std::vector<double> scale[2];
    std::vector<int> fin;           // source filters (from_filters), all with a lower index
    std::vector<double> fscale;
    int level;                      // 0 = fed by inputs only; 1 + max level of the sources otherwise
    int eval;                       // index of its evaluated source mix (row n_in + eval of xin), -1 = none
    // Delay-line sharing: filters fed by the same single input with the same scale and block delay have
    // bit-identical delay lines, so they share one ring ("input spectra reused across filters").
    int stream;                     // ring this filter's delay line lives in (its own index, or a lower filter's)
    int key_ch, key_delay;          // the sharing key the ring was last written under: input, delay, rounded scale
    double key_scale;
    long key_since;                 // block count at which the key last changed (-1 = unchanged since creation)
    int coeff, prevcoeff, delayblocks;
    // Exact emulation of the reference's P-slot ring around run-time delay changes (begin_transitions):
    char *mirror;                   // P slots on the device = the reference's cbuf[filter][0..P) during a transition
    long trans_until;               // block count at which the ring is regular again, -1 = not in a transition
    std::vector<long> ref_id;       // [P]    block index whose spectrum each mirror slot holds (-1 = zeros)
    std::vector<long> eng_id;       // [ring] the same for the slots of the filter's (private) ring
};

#define TIMING_RING 64

// One captured step graph (graph_tick): the stages of THREE consecutive launches side by side -- forward(k), MAC(k-1),
// inverse(k-2) -- as independent branches of one CUDA graph.  `mask` says which stages it holds (4 forward, 2 MAC,
// 1 inverse; partial masks fill and drain the pipeline), `par` = k & 1 fixes every double-buffered operand.
struct StepGraph {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    cudaGraphNode_t n_fwd, n_mac;           // the two kernels whose arguments carry the ring slot
    cudaKernelNodeParams p_fwd, p_mac;      // as captured (their kernelParams point into `graph`)
    cudaGraphNode_t n_h2d, n_d2h, n_snap;   // mode 2: the copy nodes whose host side changes per call
    int n_kernels;
};
// a launch whose later stages have not been enqueued yet
struct PendingStage {
    bool valid;
    unsigned int step;          // launch number of the step
    int slot_t;                 // ring slot of its first block
    void *host_out;             // host-buffer call: where its output goes (null: device-resident call)
    unsigned int call;          // ... and which call that was (io_count at the time)
};

struct bfcuda_engine {
    int L, P, N, rs;
    int max_batch, fdl_ring;    // blocks per launch at most; delay-line slots per stream (see bfcuda_create)
    int slot_t;                 // ring slot of the next block (0 <= slot_t < ring)
    int prev_par;               // which of the two `prev` buffers holds the last input block
    int last_batch;             // blocks in the most recent launch (layout of Y for debug_read)
    int n_ch[2], n_bytes[2];
    int fast_fmt[2];            // uniform aligned 4-byte sample layout of all inputs / outputs (ForwardArgs::fast_fmt)
    int n_filters, n_coeffs;
    int device;
    unsigned int flags;
    double safety_limit;
    int split, mac_variant;
    int sm_count;
    char device_name[64];
    std::vector<bfcuda_buffer_format> fmt[2];
    std::vector<FilterState> filters;
    std::vector<int> coeff_n_blocks, coeff_hbase;
    int total_coeff_blocks;

    // Three stage streams, software-pipelined over consecutive launches: `stream` runs unpack + forward (and every
    // table / coefficient update), `s_mac` the multiply-accumulate, `s_inv` inverse + pack.  Launch n+1's forward
    // stage and launch n-1's inverse stage run beside launch n's MAC (filter outputs Y are double buffered, the
    // delay-line ring has a launch of slack); the MAC is the HBM-bound stage, the other two fill its gaps and tail.
    // `s_in` / `s_out` carry the host<->device copies of the streaming interface (double-buffered raw blocks).
    cudaStream_t stream, s_mac, s_inv, s_in, s_out;
    cudaEvent_t ev_h2d[2], ev_fwd[2], ev_inv;
    cudaEvent_t ev_call_done[4];    // output (and overflow snapshot) of host-buffer call c complete in host memory: [c & 3]
    // Step graphs (launch-bound shapes, small shards): see graph_tick
    bool graph_enabled;             // not BFCUDA_FLAG_NO_GRAPH
    bool graph_mode;                // the most recent launch went through graph_tick
    int graph_host;                 // ... 0: of a device-resident call; 1: host buffers, copies on the copy streams around
                                    // the graph; 2: host buffers, copies INSIDE the graph (small pinned blocks)
    bool graph_used;
    PendingStage pend_mac, pend_inv;
    StepGraph sg[3][8][4];          // [graph_host mode][stage mask][k & 3] (k & 1 unless mode 1)
    cudaEvent_t ev_cap[4];          // fork / join inside a capture
    cudaEvent_t ev_g_done[4];       // end of the graph of tick k, [k & 3]
    cudaEvent_t ev_out_read[4];     // last device -> host read of raw output buffer q (graph mode)
    cudaEvent_t ev_h2dq[4];         // input copy of tick k complete, [k & 3]
    cudaEvent_t ev_fwd_done[2], ev_mac_done[2], ev_inv_done[2], ev_join;  // per launch parity
    // BFCUDA_FLAG_LOW_LATENCY: the sum over partitions 1 .. P-1 of the next block, computed ahead (enqueue_batch)
    bool low_latency;
    bool tail_ready;            // Y partial 1 of the next launch's generation holds that sum, made with the current tables
    cudaEvent_t ev_tail_done;   // end of the most recent ahead-of-time launch (it reads the job table)
    unsigned int launch_no;     // launches enqueued so far
    size_t y_stride;            // bytes between the two generations of Y
    unsigned int io_count;      // blocks submitted through the host-buffer interface
    FftPlan plan;
    uint8_t *d_raw[2];          // raw blocks of the device-resident interface (and buffer 0 of the streaming one)
    uint8_t *d_raw2[2];         // second buffers of the streaming interface
    uint8_t *d_raw34[2][2];     // third and fourth: the step graphs keep four host-buffer calls in flight
    SampleFormat *d_fmt[2];
    void *d_prev[2], *d_fdl, *d_xin, *d_H, *d_Y, *d_out_time, *d_scratch;
    void *d_xt[2];              // size-specialised path: unpacked input blocks [max_batch][n_in][L], two generations
    int xt_par, xt_last_nb;     // generation holding the most recent launch's blocks, and how many it holds
    bool single_dest, simple_mix;   // ForwardArgs::single_dest / InverseArgs::simple_mix of the current tables
    Overflow *d_overflow;
    unsigned int *d_status;
    unsigned int *h_status;     // pinned: status word of the device-resident / download path
    // Host-buffer calls: the overflow records [n_out] and the status word as of the END of call k are snapshotted on
    // the device behind its last kernel (s_inv) and read out with its output block, double buffered by call parity --
    // call k+1's inverse stage may already be updating d_overflow / d_status while call k's read-out runs.
    char *d_snap[4];            // by call & 3
    char *h_snap[4];            // pinned
    size_t snap_bytes, snap_status_off;
    bool h_overflow_valid;      // false until such a call has been made (and after a reset / a device-resident call)
    unsigned int io_waited;     // value of io_count when the host last waited for the most recent call's read-out
    size_t device_bytes;

    // tables: host mirrors and device copies
    std::vector<FwdDest> h_dests;
    std::vector<int> h_dest_first;
    std::vector<uint8_t> h_need_xin;
    std::vector<MixStream> h_mix_streams;
    std::vector<MixTerm> h_mix_terms;
    std::vector<MacJob> h_jobs;
    std::vector<OutChan> h_chans;       // what the inverse stage reads
    std::vector<OutChan> h_mixes;       // the filter terms of every output (k_out_mix)
    OutChan *d_mixes;
    bool any_out_mix;
    std::vector<MixTerm> h_out_terms;
    std::vector<int> shared_out;
    std::vector<std::pair<int, int>> delay_changes;     // (filter, old delay) of delay changes since the last block (begin_transitions)
    FwdDest *d_dests;
    int *d_dest_first;
    uint8_t *d_need_xin;
    MixStream *d_mix_streams;
    MixTerm *d_mix_terms;
    MacJob *d_jobs;
    OutChan *d_chans;
    MixTerm *d_out_terms;
    // filter -> filter chaining: per level, the ranges of MAC jobs, mix streams and evaluations
    int n_eval, n_vin, n_levels;
    long merge_check_at;        // block count at which update_streams() should look for mergeable rings again, -1 = never
    int n_rings;                // delay-line rings in use (distinct streams; < n_filters when filters share)
    void *d_keep;
    EvalEntry *d_eval_entries;
    MixTerm *d_eval_terms;
    std::vector<EvalEntry> h_eval_entries;
    std::vector<MixTerm> h_eval_terms;
    std::vector<int> level_job_first, level_mix_first, level_eval_first;   // [n_levels + 1]
    // HP-TPDF dither
    std::vector<int> dither_of_out;     // per output: index into the dithered channels, -1 = not dithered
    DitherArgs dither;                  // device pointers; n_dither = 0 when nothing is dithered
    // powersave (bfrun.c:1541-1552, 1613-1700, 722-772)
    int powersave;                      // 0 off, 1 exact zeros only, 2 analog level
    unsigned int *d_amax[2];            // [max_batch][n_in] block peaks, by the generation of d_xt
    float *d_ps_thr;                    // [n_in]
    uint8_t *d_slot_zero;               // [F][ring]
    long ps_hold_until;                 // block count until which the MAC ignores the flags (after a delay transition)
    // virtual -> physical outputs, mute, sub-sample delay (bfrun.c:1503-1526, 1918-2002)
    std::vector<int> out_rep;           // per output: the member of its physical channel that is quantised (itself: alone)
    std::vector<VirtGroup> h_groups;
    std::vector<int> h_members;
    VirtGroup *d_groups;
    int *d_members;
    bool any_group;                     // some physical channel carries several virtual outputs
    std::vector<uint8_t> h_muted[2];
    uint8_t *d_muted[2];
    bool any_muted[2];
    std::vector<SubdelayChan> h_sd[2];  // channels with a sub-sample delay filter
    SubdelayChan *d_sd_chans[2];
    void *d_sd_taps[2], *d_sd_hist[2];  // [n_ch][BF_SUBDELAY_MAX_TAPS] taps, [n_ch][BF_SUBDELAY_MAX_TAPS - 1] history
    bool dirty, xfade_active;
    size_t mac_bytes;           // algorithmic MAC bytes of one block launched alone
    size_t mac_bytes_batch;     // compulsory MAC bytes of one full batch of max_batch blocks

    unsigned int t;
    // measurement
    cudaEvent_t timer[2];
    cudaEvent_t ring[TIMING_RING][6];
    int ring_fill;
    int ring_blocks[TIMING_RING];
    double stage_ms[BFCUDA_N_STAGES];
    long stage_blocks, launches;
    // multi-GPU
    nccl_comm_t comm;
    int n_ranks;
};

static size_t rs_bytes(const bfcuda_engine *e, size_t n_reals) { return n_reals * (size_t)e->rs; }

template <typename T>
static int dev_alloc(bfcuda_engine *e, T **p, size_t bytes, bool zero = true)
{
    *p = nullptr;
    if (bytes == 0) {
        bytes = 16;
    }
    cudaError_t err = cudaMalloc((void **)p, bytes);
    if (err != cudaSuccess) {
        return fail(BFCUDA_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(err));
    }
    e->device_bytes += bytes;
    if (zero) {
        err = cudaMemset(*p, 0, bytes);
        if (err != cudaSuccess) {
            return fail(BFCUDA_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(err));
        }
    }
    return 0;
}

static SampleFormat to_dev_format(const bfcuda_buffer_format &b)
{
    SampleFormat f;
    f.isfloat = b.sf.isfloat;
    f.swap = b.sf.swap;
    f.bytes = b.sf.bytes;
    f.sbytes = b.sf.sbytes;
    f.sample_spacing = b.sample_spacing;
    f.byte_offset = b.byte_offset;
    return f;
}

static double round_to_real(const bfcuda_engine *e, double v)
{
    return e->rs == 4 ? (double)(float)v : v;   // (real_t)scale, fftw_convfuns.h:18-20
}

static int clamp_delay(const bfcuda_engine *e, int delay)
{
    if (delay < 0) return 0;                    // bfrun.c:1579-1584
    if (delay > e->P - 1) return e->P - 1;
    return delay;
}

static int coeff_blocks_used(const bfcuda_engine *e, int coeff, int delay)
{
    // bfrun.c:1585-1598
    if (coeff < 0 || e->coeff_n_blocks[coeff] > e->P - delay) {
        return e->P - delay;
    }
    return e->coeff_n_blocks[coeff];
}

static char *ring_ptr(const bfcuda_engine *e, int stream)
{
    return (char *)e->d_fdl + rs_bytes(e, (size_t)stream * e->fdl_ring * e->N);
}

static bool shareable(const bfcuda_engine *e, const FilterState &fs)
{
    return !(e->flags & BFCUDA_FLAG_NO_STREAM_SHARING) && fs.eval < 0 && fs.ch[0].size() == 1;
}

// Keep the ring assignment consistent with the control snapshot, at a block boundary (before the tables are built
// and before any delay transition starts).  A filter whose key (input, scale, delay) changed leaves the ring it shared and
// takes a copy of it -- its history up to now IS that ring; if it was the ring's owner the remaining users move to a
// copy of their own.  Filters whose keys are equal and have been for longer than any slot lives (or since creation)
// hold identical delay lines and are merged again, no copy needed.
static int update_streams(bfcuda_engine *e)
{
    const int F = e->n_filters;
    const size_t ring_bytes = rs_bytes(e, (size_t)e->fdl_ring * e->N);
    for (int f = 0; f < F; f++) {
        FilterState &fs = e->filters[f];
        if (fs.ch[0].size() != 1 || fs.eval >= 0) {
            continue;
        }
        const int ch = fs.ch[0][0], delay = clamp_delay(e, fs.delayblocks);
        const double sc = round_to_real(e, fs.scale[0][0] * e->fmt[0][ch].sf.scale);
        if (ch == fs.key_ch && delay == fs.key_delay && sc == fs.key_scale) {
            continue;
        }
        const bool first_snapshot = fs.key_ch < 0;
        // key changed: separate f from the ring it shares
        std::vector<int> users;
        for (int u = 0; u < F; u++) {
            if (u != f && e->filters[u].stream == fs.stream) {
                users.push_back(u);
            }
        }
        if (!users.empty()) {
            if (fs.stream != f) {
                CU(cudaMemcpyAsync(ring_ptr(e, f), ring_ptr(e, fs.stream), ring_bytes, cudaMemcpyDeviceToDevice, e->stream));
                if (e->d_slot_zero != nullptr) {        // the powersave flags travel with the ring
                    CU(cudaMemcpyAsync(e->d_slot_zero + (size_t)f * e->fdl_ring, e->d_slot_zero + (size_t)fs.stream * e->fdl_ring,
                                       (size_t)e->fdl_ring, cudaMemcpyDeviceToDevice, e->stream));
                }
                fs.stream = f;
            } else {
                const int g = users[0];     // lowest index: the new owner
                CU(cudaMemcpyAsync(ring_ptr(e, g), ring_ptr(e, f), ring_bytes, cudaMemcpyDeviceToDevice, e->stream));
                if (e->d_slot_zero != nullptr) {
                    CU(cudaMemcpyAsync(e->d_slot_zero + (size_t)g * e->fdl_ring, e->d_slot_zero + (size_t)f * e->fdl_ring,
                                       (size_t)e->fdl_ring, cudaMemcpyDeviceToDevice, e->stream));
                }
                for (int u : users) {
                    e->filters[u].stream = g;
                }
            }
        }
        fs.key_ch = ch;
        fs.key_delay = delay;
        fs.key_scale = sc;
        fs.key_since = first_snapshot ? -1 : (long)e->t;
    }
    // merge rings that have become identical
    const long horizon = 2L * e->P + 2L * e->max_batch;     // no slot written before t - horizon is ever read again
    for (int f = 0; f < F; f++) {
        FilterState &fs = e->filters[f];
        if (!shareable(e, fs) || fs.stream != f) {
            continue;       // only ring owners look for an older twin; their followers move with them
        }
        const bool settled_f = fs.key_since < 0 || (long)e->t - fs.key_since >= horizon;
        for (int g = 0; g < f && settled_f; g++) {
            const FilterState &gs = e->filters[g];
            const bool settled_g = gs.key_since < 0 || (long)e->t - gs.key_since >= horizon;
            if (shareable(e, gs) && gs.stream == g && settled_g && gs.key_ch == fs.key_ch &&
                gs.key_delay == fs.key_delay && gs.key_scale == fs.key_scale) {
                for (int u = 0; u < F; u++) {
                    if (e->filters[u].stream == f) {
                        e->filters[u].stream = g;
                    }
                }
                break;
            }
        }
    }
    // a ring that is not settled yet may become mergeable later: look again then
    e->merge_check_at = -1;
    for (int f = 0; f < F; f++) {
        const FilterState &fs = e->filters[f];
        if (shareable(e, fs) && fs.key_since >= 0 && (long)e->t - fs.key_since < horizon) {
            const long at = fs.key_since + horizon;
            e->merge_check_at = e->merge_check_at < 0 ? at : std::min(e->merge_check_at, at);
        }
    }
    return 0;
}

// Rebuild every per-block table from the control snapshot: the analogue of bfrun.c:1460-1484 plus the
// scale / slot / coefficient bookkeeping spread over bfrun.c:1566-1600, 1663-1675, 1726-1777, 1847-1854.
static void invalidate_graphs(bfcuda_engine *e);
extern "C" {
static int sync_all(bfcuda_engine *e);      // drains the step-graph pipeline, then waits for every stream
static int wait_call(bfcuda_engine *e, unsigned int call);
static int graph_drain(bfcuda_engine *e);
}

static void build_tables(bfcuda_engine *e)
{
    const int F = e->n_filters;
    std::vector<std::vector<FwdDest>> per_ch(e->n_ch[0]);
    e->h_mix_streams.clear();
    e->h_mix_terms.clear();
    e->h_jobs.clear();
    std::fill(e->h_need_xin.begin(), e->h_need_xin.end(), (e->flags & BFCUDA_FLAG_KEEP_INPUT_SPECTRA) ? 1 : 0);
    e->xfade_active = false;
    size_t blocks_h = 0, blocks_x = 0;
    std::vector<int> stream_parts(std::max(1, F), 0);   // delay-line blocks read per ring (shared rings count once)

    e->h_eval_entries.clear();
    e->h_eval_terms.clear();
    e->level_job_first.assign(e->n_levels + 1, 0);
    e->level_mix_first.assign(e->n_levels + 1, 0);
    e->level_eval_first.assign(e->n_levels + 1, 0);
    for (int level = 0; level < e->n_levels; level++) {
      e->level_job_first[level] = (int)e->h_jobs.size();
      e->level_mix_first[level] = (int)e->h_mix_streams.size();
      e->level_eval_first[level] = (int)e->h_eval_entries.size();
      for (int f = 0; f < F; f++) {
        const FilterState &fs = e->filters[f];
        if (fs.level != level) {
            continue;
        }
        const int delay = clamp_delay(e, fs.delayblocks);
        const int nin = (int)fs.ch[0].size();
        if (fs.eval >= 0) {
            // bfrun.c:1603-1660: the scaled mix of the source filters' outputs, evaluated in the time domain, joins
            // the channel inputs as one more spectrum with scale 1.0 ("unecessary scale multiply", bfrun.c:1646-1647)
            EvalEntry en;
            en.first = (int)e->h_eval_terms.size();
            en.n = (int)fs.fin.size();
            en.xf_first = -1;
            en.vin = e->n_ch[0] + fs.eval;
            bool any_xf = false;
            for (size_t i = 0; i < fs.fin.size(); i++) {
                MixTerm tm;
                tm.index = fs.fin[i];
                tm.scale = round_to_real(e, fs.fscale[i]);
                e->h_eval_terms.push_back(tm);
                const FilterState &src = e->filters[fs.fin[i]];
                any_xf |= src.crossfade && src.prevcoeff != src.coeff;
            }
            if (any_xf) {
                en.xf_first = (int)e->h_eval_terms.size();
                for (size_t i = 0; i < fs.fin.size(); i++) {
                    MixTerm tm = e->h_eval_terms[en.first + i];
                    const FilterState &src = e->filters[fs.fin[i]];
                    if (src.crossfade && src.prevcoeff != src.coeff) {
                        tm.index = F + tm.index;
                    }
                    e->h_eval_terms.push_back(tm);
                }
            }
            e->h_eval_entries.push_back(en);
            MixStream ms;
            ms.stream = f;
            ms.delay = delay;
            ms.n_inputs = nin + 1;
            ms.first = (int)e->h_mix_terms.size();
            for (int i = 0; i < nin; i++) {
                MixTerm tm;
                tm.index = fs.ch[0][i];
                tm.scale = round_to_real(e, fs.scale[0][i] * e->fmt[0][fs.ch[0][i]].sf.scale);
                e->h_mix_terms.push_back(tm);
                e->h_need_xin[fs.ch[0][i]] = 1;
            }
            MixTerm te;
            te.index = en.vin;
            te.scale = 1.0;
            e->h_mix_terms.push_back(te);
            e->h_mix_streams.push_back(ms);
        } else if (nin == 1) {
            if (fs.stream == f) {       // the ring's owner writes it; the filters sharing it only read
                FwdDest d;
                d.stream = f;
                d.delay = delay;
                d.scale = round_to_real(e, fs.scale[0][0] * e->fmt[0][fs.ch[0][0]].sf.scale);
                per_ch[fs.ch[0][0]].push_back(d);
            }
        } else if (nin > 1) {
            MixStream ms;
            ms.stream = f;
            ms.delay = delay;
            ms.n_inputs = nin;
            ms.first = (int)e->h_mix_terms.size();
            for (int i = 0; i < nin; i++) {
                MixTerm tm;
                tm.index = fs.ch[0][i];
                tm.scale = round_to_real(e, fs.scale[0][i] * e->fmt[0][fs.ch[0][i]].sf.scale);
                e->h_mix_terms.push_back(tm);
                e->h_need_xin[fs.ch[0][i]] = 1;
            }
            e->h_mix_streams.push_back(ms);
        }
        MacJob jb;
        jb.stream = fs.stream;
        jb.hbase = fs.coeff < 0 ? -1 : e->coeff_hbase[fs.coeff];
        jb.n_parts = fs.coeff < 0 ? 1 : coeff_blocks_used(e, fs.coeff, delay);
        jb.out = f;
        e->h_jobs.push_back(jb);
        blocks_h += fs.coeff < 0 ? 0 : jb.n_parts;
        stream_parts[fs.stream] = std::max(stream_parts[fs.stream], jb.n_parts);
        if (fs.crossfade && fs.prevcoeff != fs.coeff) {
            // bfrun.c:1726-1736, 1755-1769: the same delay line through the previous coefficients
            MacJob old = jb;
            old.hbase = fs.prevcoeff < 0 ? -1 : e->coeff_hbase[fs.prevcoeff];
            old.n_parts = fs.prevcoeff < 0 ? 1 : coeff_blocks_used(e, fs.prevcoeff, delay);
            old.out = F + f;
            e->h_jobs.push_back(old);
            blocks_h += fs.prevcoeff < 0 ? 0 : old.n_parts;
            stream_parts[fs.stream] = std::max(stream_parts[fs.stream], old.n_parts);
            e->xfade_active = true;
        }
      }
    }
    e->n_rings = 0;
    for (int f = 0; f < F; f++) {
        blocks_x += stream_parts[f];
        e->n_rings += stream_parts[f] > 0;
    }
    // jobs that read the same ring sit next to each other within their level: their blocks run at the same time and
    // the second reader of a delay-line block finds it in L2
    for (int level = 0; level < e->n_levels; level++) {
        std::stable_sort(e->h_jobs.begin() + e->level_job_first[level],
                         e->h_jobs.begin() + (level + 1 < e->n_levels ? e->level_job_first[level + 1] : (int)e->h_jobs.size()),
                         [](const MacJob &a, const MacJob &b) { return a.stream < b.stream; });
    }
    e->level_job_first[e->n_levels] = (int)e->h_jobs.size();
    e->level_mix_first[e->n_levels] = (int)e->h_mix_streams.size();
    e->level_eval_first[e->n_levels] = (int)e->h_eval_entries.size();
    e->h_dests.clear();
    for (int c = 0; c < e->n_ch[0]; c++) {
        e->h_dest_first[c] = (int)e->h_dests.size();
        e->h_dests.insert(e->h_dests.end(), per_ch[c].begin(), per_ch[c].end());
    }
    e->h_dest_first[e->n_ch[0]] = (int)e->h_dests.size();

    // output mixes, filters in index order (bfrun.c:1351-1365, 1847-1868)
    e->h_out_terms.clear();
    for (int o = 0; o < e->n_ch[1]; o++) {
        OutChan oc;
        oc.first = (int)e->h_out_terms.size();
        oc.n = 0;
        oc.xf_first = -1;
        oc.shared = std::find(e->shared_out.begin(), e->shared_out.end(), o) != e->shared_out.end() ? 1 : 0;
        if (e->dither_of_out[o] >= 0) {
            oc.shared |= 2;
        }
        if (!e->out_rep.empty() && e->out_rep[(size_t)o] != o) {
            oc.shared |= 4;     // mixed into another virtual output's physical channel (k_virt_mix), never packed
        }
        bool any_xf = false;
        for (int f = 0; f < F; f++) {
            const FilterState &fs = e->filters[f];
            for (size_t j = 0; j < fs.ch[1].size(); j++) {
                if (fs.ch[1][j] == o) {
                    MixTerm tm;
                    tm.index = f;
                    tm.scale = round_to_real(e, fs.scale[1][j] / e->fmt[1][o].sf.scale);
                    e->h_out_terms.push_back(tm);
                    oc.n++;
                    any_xf |= fs.crossfade && fs.prevcoeff != fs.coeff;
                    break;      // "output exists only once per filter", bfrun.c:1360
                }
            }
        }
        if (any_xf) {
            oc.xf_first = (int)e->h_out_terms.size();
            for (int j = 0; j < oc.n; j++) {
                MixTerm tm = e->h_out_terms[oc.first + j];
                const FilterState &fs = e->filters[tm.index];
                if (fs.crossfade && fs.prevcoeff != fs.coeff) {
                    tm.index = F + tm.index;
                }
                e->h_out_terms.push_back(tm);
            }
        }
        e->h_mixes[o] = oc;
        e->h_chans[o] = oc;
    }
    // Outputs fed by several filters are mixed by k_out_mix into a spectrum of their own (Y slots 2F + o, and
    // 2F + n_out + o for the old-coefficient mix of a crossfade block); the inverse stage then sees ONE term with scale 1
    // for them, like for an output with a single filter.
    e->any_out_mix = false;
    for (int o = 0; o < e->n_ch[1]; o++) {
        const OutChan mx = e->h_mixes[o];
        if (mx.n <= 1) {
            continue;
        }
        e->any_out_mix = true;
        OutChan oc = mx;
        MixTerm tm;
        tm.scale = 1.0;
        tm.index = 2 * std::max(1, F) + o;
        oc.first = (int)e->h_out_terms.size();
        oc.n = 1;
        e->h_out_terms.push_back(tm);
        if (mx.xf_first >= 0) {
            tm.index = 2 * std::max(1, F) + e->n_ch[1] + o;
            oc.xf_first = (int)e->h_out_terms.size();
            e->h_out_terms.push_back(tm);
        }
        e->h_chans[o] = oc;
    }
    // algorithmic MAC traffic, SURVEY.md 8(d): rs * N * (coefficient blocks + delay-line blocks + outputs)
    e->mac_bytes = (size_t)e->rs * e->N * (blocks_h + blocks_x + e->h_jobs.size());
    // a batch of B blocks reads every coefficient block once, a window of n_parts + B - 1 delay-line blocks per
    // job, and writes B outputs per job
    const size_t Bm = (size_t)e->max_batch;
    e->mac_bytes_batch = (size_t)e->rs * e->N * (blocks_h + blocks_x + (size_t)e->n_rings * (Bm - 1) + e->h_jobs.size() * Bm);
    e->single_dest = !(e->flags & BFCUDA_FLAG_KEEP_INPUT_SPECTRA) && e->h_mix_streams.empty();
    for (int c = 0; c < e->n_ch[0]; c++) {
        e->single_dest = e->single_dest && per_ch[c].size() == 1;
    }
    e->simple_mix = !e->xfade_active;   // (split partition sums and multi-filter mixes are reduced by their own kernels)
    for (int o = 0; o < e->n_ch[1]; o++) {
        e->simple_mix = e->simple_mix && e->h_chans[o].n == 1;
    }
    e->dirty = false;
    invalidate_graphs(e);       // job counts and kernel variants are baked into the captured launches
}

template <typename T>
static cudaError_t upload_vec(T *dst, const std::vector<T> &v, cudaStream_t s)
{
    if (v.empty()) {
        return cudaSuccess;
    }
    return cudaMemcpyAsync(dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

static int upload_tables(bfcuda_engine *e)
{
    CU(upload_vec(e->d_dests, e->h_dests, e->stream));
    CU(upload_vec(e->d_dest_first, e->h_dest_first, e->stream));
    CU(upload_vec(e->d_need_xin, e->h_need_xin, e->stream));
    CU(upload_vec(e->d_mix_streams, e->h_mix_streams, e->stream));
    CU(upload_vec(e->d_mix_terms, e->h_mix_terms, e->stream));
    CU(upload_vec(e->d_jobs, e->h_jobs, e->stream));
    CU(upload_vec(e->d_chans, e->h_chans, e->stream));
    CU(upload_vec(e->d_mixes, e->h_mixes, e->stream));
    CU(upload_vec(e->d_out_terms, e->h_out_terms, e->stream));
    CU(upload_vec(e->d_eval_entries, e->h_eval_entries, e->stream));
    CU(upload_vec(e->d_eval_terms, e->h_eval_terms, e->stream));
    return 0;
}

static int choose_split(const bfcuda_engine *e, int requested)
{
    if (requested < 1) {
        const char *env = getenv("BFCUDA_MAC_SPLIT");       // experiments: force the split of the automatic rule
        if (env != nullptr && atoi(env) >= 1) {
            requested = atoi(env);
        }
    }
    if (requested >= 1) {
        return std::min(requested, std::max(1, e->P));
    }
    // Automatic: keep the reference's summation order (split 1) whenever the (filter, bin) space alone
    // fills the machine; otherwise split the partition sum so that ~512 threads per SM have work.
    // (The batched kernel gives a thread fewer bins but B blocks of independent work: one 256-thread block per SM
    // is enough there.)
    int lanes = 16 / e->rs;
    if (e->max_batch > 1) {
        lanes = mac_batch_lanes(e->rs, e->max_batch, std::max(1, e->n_filters), e->N);     // the launcher's own table
    }
    const long threads = (long)std::max(1, e->n_filters) * (e->N / 2 / lanes);
    long target = (long)e->sm_count * (e->max_batch > 1 ? 256 : 512);
    if (e->max_batch > 1 && threads * 2 >= target) {
        // at least four warps per SM of the batched kernel: keep the whole partition sum in one thread.  (The 8-filter
        // shard of the headline job -- 32768 threads for 37888 slots -- was split three ways by the rule below: 45 us per
        // launch against 37-39 us unsplit, profiles/r2_macsweep_w2split1.txt; and unsplit is the reference's summation order.)
        return 1;
    }
    if (e->max_batch > 1 && threads < target) {
        // once the sum has to be split anyway, split it far enough for ~4 warps per scheduler: the batched kernel's
        // throughput at low occupancy follows the warp count (32 x 256 bins x 1024 partitions: 62 us per 8 blocks with
        // 5 partials)
        target = (long)e->sm_count * 512;
    }
    long s = (target + threads - 1) / threads;
    s = std::min<long>(s, std::max(1, e->P / 8));
    return (int)std::max<long>(1, s);
}

// ---- dither table (dither.c:37-139) -------------------------------------------------------------------------
// "maximally equidistributed combined Tausworthe generator" (GSL taus) with the default seed, one int8 per sample;
// channel j starts `spacing` bytes after channel j-1 (10 s of samples, at least 1 s and at least one block).
static uint32_t tausrand(uint32_t state[3])
{
#define BF_TAUSWORTHE(s, a, b, c, d) ((s & c) << d) ^ (((s << a) ^ s) >> b)
    state[0] = BF_TAUSWORTHE(state[0], 13, 19, (uint32_t)4294967294U, 12);
    state[1] = BF_TAUSWORTHE(state[1], 2, 25, (uint32_t)4294967288U, 4);
    state[2] = BF_TAUSWORTHE(state[2], 3, 11, (uint32_t)4294967280U, 17);
    return state[0] ^ state[1] ^ state[2];
}

static int setup_dither(bfcuda_engine *e, const struct bfcuda_config *c)
{
    e->dither_of_out.assign(std::max(1, e->n_ch[1]), -1);
    if (c->apply_dither == nullptr) {
        return 0;
    }
    std::vector<DitherChan> chans;
    for (int o = 0; o < e->n_ch[1]; o++) {
        const bfcuda_sample_format &sf = e->fmt[1][o].sf;
        // bfconf.c:3173-3217: no dither on float formats, on more than 16 bit at float_bits 32, on 32 bit formats
        if (!c->apply_dither[o] || sf.isfloat || (e->rs == 4 && sf.sbytes > 2) || sf.sbytes >= 4) {
            continue;
        }
        e->dither_of_out[o] = (int)chans.size();
        DitherChan dc;
        dc.out = o;
        dc.randtab_ptr = 0;
        dc.e0 = dc.e1 = 0.0;
        chans.push_back(dc);
    }
    const int n = (int)chans.size();
    if (n == 0) {
        return 0;
    }
    const int rate = c->sampling_rate > 0 ? c->sampling_rate : 44100;
    int spacing = 10 * rate;
    const int minspacing = std::max(rate, e->L);
    spacing = std::max(spacing, minspacing);
    if (c->max_dither_table_size > 0 && (long)n * spacing > c->max_dither_table_size) {
        spacing = c->max_dither_table_size / n;
    }
    if (spacing < minspacing) {
        return fail(BFCUDA_EINVAL, "Maximum dither table size %d bytes is too small, must at least be %d bytes.",
                    c->max_dither_table_size, n * rate * minspacing);
    }
    const size_t size = (size_t)n * spacing + 1;
    std::vector<int8_t> tab(size);
    uint32_t st[3];
    {
        uint32_t seed = 1;      // tausinit(state, 0): "default seed is 1"
#define BF_LCG(v) ((69069 * (v)) & 0xFFFFFFFFU)
        st[0] = BF_LCG(seed);
        st[1] = BF_LCG(st[0]);
        st[2] = BF_LCG(st[1]);
        for (int i = 0; i < 6; i++) {
            tausrand(st);
        }
    }
    for (size_t i = 0; i < size; i++) {
        tab[i] = (int8_t)(tausrand(st) & 0x000000FF);
    }
    for (int j = 0; j < n; j++) {
        chans[j].randtab_ptr = j * spacing + 1;
    }
    // dither.c:113-131: integer difference -> dither in (-1, +1) plus the +0.5 that makes truncation a mid-tread
    // requantiser.  Entry 511 (difference +255, which the reference's 511-entry map does not cover) continues the line.
    std::vector<double> mapd(512);
    std::vector<float> mapf(512);
    mapf[0] = -0.5f;
    mapd[0] = -0.5;
    for (int k = -255; k < 254; k++) {
        mapf[k + 256] = (float)(0.5 + 1.0 / 255.0 + 1.0 / 255.0 * (float)k);
        mapd[k + 256] = 0.5 + 1.0 / 255.0 + 1.0 / 255.0 * (double)k;
    }
    mapf[510] = 1.5f;
    mapd[510] = 1.5;
    mapf[511] = (float)(1.5 + 1.0 / 255.0);
    mapd[511] = 1.5 + 1.0 / 255.0;
    int rc;
    int8_t *d_tab = nullptr;
    void *d_map = nullptr;
    if ((rc = dev_alloc(e, &d_tab, size, false)) != 0) return rc;
    e->dither.randtab = d_tab;
    if ((rc = dev_alloc(e, &d_map, 512 * (size_t)e->rs, false)) != 0) return rc;
    e->dither.randmap = d_map;
    if ((rc = dev_alloc(e, &e->dither.chans, sizeof(DitherChan) * n, false)) != 0) return rc;
    CU(cudaMemcpy(d_tab, tab.data(), size, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_map, e->rs == 4 ? (const void *)mapf.data() : (const void *)mapd.data(), 512 * (size_t)e->rs,
                  cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->dither.chans, chans.data(), sizeof(DitherChan) * n, cudaMemcpyHostToDevice));
    e->dither.randtab_size = (int)size;
    e->dither.n_dither = n;
    return 0;
}

// ======================================================================================================
// C ABI
// ======================================================================================================

static void invalidate_graphs(bfcuda_engine *e)
{
    for (int h = 0; h < 3; h++) {
        for (int m = 0; m < 8; m++) {
            for (int p = 0; p < 4; p++) {
                StepGraph &g = e->sg[h][m][p];
                if (g.exec != nullptr) cudaGraphExecDestroy(g.exec);
                if (g.graph != nullptr) cudaGraphDestroy(g.graph);
                memset(&g, 0, sizeof(g));
            }
        }
    }
}

extern "C" {

const char *bfcuda_strerror(void)
{
    return g_err;
}

int bfcuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void bfcuda_destroy(bfcuda_engine *e)
{
    if (e == nullptr) {
        return;
    }
    cudaSetDevice(e->device);
    for (cudaStream_t st : { e->stream, e->s_mac, e->s_inv, e->s_in, e->s_out }) {
        if (st) cudaStreamSynchronize(st);
    }
    if (e->comm != nullptr && g_nccl.handle != nullptr) {
        g_nccl.CommDestroy(e->comm);
    }
    void *ptrs[] = { e->d_groups, e->d_members, e->d_muted[0], e->d_muted[1], e->d_sd_chans[0], e->d_sd_chans[1], e->d_sd_taps[0],
                     e->d_sd_taps[1], e->d_sd_hist[0], e->d_sd_hist[1], e->d_amax[0], e->d_amax[1], e->d_ps_thr, e->d_slot_zero, e->d_raw34[0][0], e->d_raw34[0][1], e->d_raw34[1][0], e->d_raw34[1][1], e->d_xt[0], e->d_xt[1], e->d_raw[0], e->d_raw[1], e->d_raw2[0], e->d_raw2[1], e->d_fmt[0], e->d_fmt[1], e->d_prev[0], e->d_prev[1], e->d_fdl, e->d_xin, e->d_H,
                     e->d_Y, e->d_out_time, e->d_scratch, e->d_overflow, e->d_dests, e->d_dest_first,
                     e->d_need_xin, e->d_mix_streams, e->d_mix_terms, e->d_jobs, e->d_chans, e->d_out_terms,
                     e->d_keep, e->d_eval_entries, e->d_eval_terms, e->d_mixes, e->dither.chans, (void *)e->dither.randtab,
                     (void *)e->dither.randmap };
    for (void *p : ptrs) {
        if (p != nullptr) {
            cudaFree(p);
        }
    }
    for (FilterState &fs : e->filters) {
        if (fs.mirror != nullptr) {
            cudaFree(fs.mirror);
        }
    }
    if (e->h_status) cudaFreeHost(e->h_status);
    for (int i = 0; i < 4; i++) {
        if (e->h_snap[i]) cudaFreeHost(e->h_snap[i]);
        if (e->d_snap[i]) cudaFree(e->d_snap[i]);
    }
    invalidate_graphs(e);
    fft_plan_destroy(&e->plan);
    for (int i = 0; i < 2; i++) {
        if (e->timer[i]) cudaEventDestroy(e->timer[i]);
    }
    for (int i = 0; i < TIMING_RING; i++) {
        for (int j = 0; j < 6; j++) {
            if (e->ring[i][j]) cudaEventDestroy(e->ring[i][j]);
        }
    }
    for (cudaEvent_t ev : { e->ev_h2d[0], e->ev_h2d[1], e->ev_fwd[0], e->ev_fwd[1], e->ev_call_done[0], e->ev_call_done[1],
                            e->ev_call_done[2], e->ev_call_done[3], e->ev_cap[0], e->ev_cap[1], e->ev_cap[2], e->ev_cap[3],
                            e->ev_g_done[0], e->ev_g_done[1], e->ev_g_done[2], e->ev_g_done[3], e->ev_out_read[0],
                            e->ev_out_read[1], e->ev_out_read[2], e->ev_out_read[3], e->ev_h2dq[0], e->ev_h2dq[1],
                            e->ev_h2dq[2], e->ev_h2dq[3], e->ev_inv, e->ev_fwd_done[0], e->ev_fwd_done[1], e->ev_mac_done[0],
                            e->ev_mac_done[1], e->ev_inv_done[0], e->ev_inv_done[1], e->ev_join, e->ev_tail_done }) {
        if (ev) cudaEventDestroy(ev);
    }
    for (cudaStream_t st : { e->stream, e->s_mac, e->s_inv, e->s_in, e->s_out }) {
        if (st) cudaStreamDestroy(st);
    }
    delete e;
}

int bfcuda_create(const struct bfcuda_config *c, bfcuda_engine **out)
{
    if (c == nullptr || out == nullptr) {
        return fail(BFCUDA_EINVAL, "null argument");
    }
    *out = nullptr;
    // convolver_init's checks (fftw_convolver.c:796-803) + bfconf's limits (bfconf.c:1495-1520, bfmod.h:22-23)
    if (c->realsize != 4 && c->realsize != 8) {
        return fail(BFCUDA_EINVAL, "Invalid real size %d.", c->realsize);
    }
    if (c->filter_length < 4 || (c->filter_length & (c->filter_length - 1)) != 0) {
        return fail(BFCUDA_EINVAL, "Invalid length %d.", c->filter_length);
    }
    if (c->n_blocks < 1) {
        return fail(BFCUDA_EINVAL, "Invalid number of blocks %d.", c->n_blocks);
    }
    if (c->max_batch > (c->realsize == 4 ? 16 : 8)) {
        return fail(BFCUDA_EINVAL, "max_batch %d exceeds %d", c->max_batch, c->realsize == 4 ? 16 : 8);
    }
    if (c->n_channels[0] < 0 || c->n_channels[0] > BFCUDA_MAXCHANNELS || c->n_channels[1] < 0 ||
        c->n_channels[1] > BFCUDA_MAXCHANNELS || c->n_filters < 0 || c->n_filters > BFCUDA_MAXFILTERS) {
        return fail(BFCUDA_EINVAL, "channel or filter count out of range");
    }
    if (!fft_size_supported(2 * c->filter_length, c->realsize)) {
        return fail(BFCUDA_ENOTSUP, "filter_length %d at realsize %d is beyond the four-step transform (max %d)",
                    c->filter_length, c->realsize, 1 << 22);
    }
    if (!fft_single_block_supported(2 * c->filter_length, c->realsize)) {
        for (int f = 0; f < c->n_filters; f++) {
            if (c->filters[f].n_filters_in > 0) {
                return fail(BFCUDA_ENOTSUP, "filter -> filter chaining needs partitions of at most %d samples",
                            c->realsize == 4 ? 16384 : 8192);
            }
        }
    }
    for (int io = 0; io < 2; io++) {
        if (c->n_bytes[io] < 0 || (c->n_channels[io] > 0 && c->formats[io] == nullptr)) {
            return fail(BFCUDA_EINVAL, "%s: negative block size or missing buffer formats", io ? "output" : "input");
        }
        for (int n = 0; n < c->n_channels[io]; n++) {
            const bfcuda_buffer_format &b = c->formats[io][n];
            // integer formats: 1 <= sbytes <= bytes <= 4 (bfconf.c:377-472); the clip limits are 2^(8 sbytes - 1)
            const bool okf = b.sf.isfloat ? (b.sf.bytes == 4 || b.sf.bytes == 8)
                                          : (b.sf.bytes >= 1 && b.sf.bytes <= 4 && b.sf.sbytes >= 1 && b.sf.sbytes <= b.sf.bytes);
            if (!okf || b.sample_spacing < 1 || b.byte_offset < 0 ||
                (long)b.byte_offset + ((long)(c->filter_length - 1) * b.sample_spacing + 1) * b.sf.bytes >
                    c->n_bytes[io]) {
                // raw2real.h:154-158 / real2raw.h:245-249 ("Sample byte size %d is not supported.")
                return fail(BFCUDA_EINVAL, "%s channel %d: unsupported sample format or layout (bytes %d)",
                            io ? "output" : "input", n, b.sf.bytes);
            }
        }
    }
    if ((c->n_filters > 0 && c->filters == nullptr) || (c->n_coeffs > 0 && c->coeff_n_blocks == nullptr) || c->n_coeffs < 0) {
        return fail(BFCUDA_EINVAL, "missing filter or coefficient table");
    }
    for (int f = 0; f < c->n_filters; f++) {
        const bfcuda_filter &s = c->filters[f];
        for (int i = 0; i < s.n_filters_in; i++) {
            // processing order is topological, producers first (bfconf.c:2933-2964)
            if (s.filters_in == nullptr || s.filters_in[i] < 0 || s.filters_in[i] >= f) {
                return fail(BFCUDA_EINVAL, "filter %d: source filter index must be in 0..%d (filters in processing "
                                           "order, producers first)", f, f - 1);
            }
        }
        if (s.coeff >= c->n_coeffs) {
            return fail(BFCUDA_EINVAL, "filter %d: coefficient index %d out of range", f, s.coeff);
        }
        for (int io = 0; io < 2; io++) {
            if (s.n_channels[io] < 0 || (s.n_channels[io] > 0 && (s.channels[io] == nullptr || s.scale[io] == nullptr))) {
                return fail(BFCUDA_EINVAL, "filter %d: missing channel or scale array", f);
            }
            for (int i = 0; i < s.n_channels[io]; i++) {
                if (s.channels[io][i] < 0 || s.channels[io][i] >= c->n_channels[io]) {
                    return fail(BFCUDA_EINVAL, "filter %d: channel index out of range", f);
                }
            }
        }
    }
    for (int n = 0; n < c->n_coeffs; n++) {
        if (c->coeff_n_blocks[n] < 1 || c->coeff_n_blocks[n] > c->n_blocks) {
            return fail(BFCUDA_EINVAL, "coefficient set %d: %d blocks (must be 1..%d)", n, c->coeff_n_blocks[n],
                        c->n_blocks);
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BFCUDA_ENODEV, "no CUDA device available (and there is no CPU fallback)");
    }
    if (c->device < 0 || c->device >= ndev) {
        return fail(BFCUDA_EINVAL, "device %d out of range (%d devices)", c->device, ndev);
    }
    CU(cudaSetDevice(c->device));

    bfcuda_engine *e = new bfcuda_engine();
    memset(&e->plan, 0, sizeof(e->plan));
    e->L = c->filter_length;
    e->P = c->n_blocks;
    e->N = 2 * e->L;
    e->rs = c->realsize;
    e->max_batch = c->max_batch < 1 ? 1 : c->max_batch;
    // Ring slots per delay line.  The reference's ring has exactly P (bfrun.c:1045, 1600).  Here: + 2 B - 1 so that
    // launch n+1's forward stage can run beside launch n's MAC, + P - 1 so that a block written "ahead" under the
    // largest block delay (P - 1) never lands on a slot a running MAC still reads, and every slot a block can read
    // (P back) stays distinct from the slots being written (what the reference's P-slot ring aliases after a delay
    // change is reproduced from an exact mirror, begin_transitions).
    e->fdl_ring = 2 * e->P + 2 * e->max_batch;
    e->slot_t = 0;
    e->d_xt[0] = e->d_xt[1] = nullptr;
    memset(&e->dither, 0, sizeof(e->dither));
    e->d_keep = nullptr;
    e->d_mixes = nullptr;
    e->any_out_mix = false;
    e->d_eval_entries = nullptr;
    e->d_eval_terms = nullptr;
    e->n_eval = 0;
    e->n_levels = 1;
    e->n_rings = 0;
    e->merge_check_at = -1;
    e->xt_par = 0;
    e->xt_last_nb = 1;
    e->single_dest = e->simple_mix = false;
    e->prev_par = 0;
    e->last_batch = 1;
    e->device = c->device;
    e->flags = c->flags;
    e->safety_limit = c->safety_limit;
    e->n_filters = c->n_filters;
    e->n_coeffs = c->n_coeffs;
    e->stream = e->s_mac = e->s_inv = e->s_in = e->s_out = nullptr;
    for (int i = 0; i < 2; i++) {
        e->ev_fwd_done[i] = e->ev_mac_done[i] = e->ev_inv_done[i] = nullptr;
    }
    e->ev_join = nullptr;
    e->ev_tail_done = nullptr;
    e->low_latency = false;
    e->tail_ready = false;
    e->launch_no = 0;
    e->y_stride = 0;
    e->ev_h2d[0] = e->ev_h2d[1] = e->ev_fwd[0] = e->ev_fwd[1] = nullptr;
    e->ev_inv = nullptr;
    e->io_count = 0;
    e->comm = nullptr;
    e->n_ranks = 1;
    e->device_bytes = 0;
    e->t = 0;
    e->ring_fill = 0;
    e->stage_blocks = e->launches = 0;
    e->h_status = nullptr;
    for (int i = 0; i < 4; i++) {
        e->h_snap[i] = e->d_snap[i] = nullptr;
        e->ev_call_done[i] = e->ev_cap[i] = nullptr;
    }
    for (int i = 0; i < 4; i++) {
        e->ev_g_done[i] = e->ev_out_read[i] = e->ev_h2dq[i] = nullptr;
    }
    e->d_raw34[0][0] = e->d_raw34[0][1] = e->d_raw34[1][0] = e->d_raw34[1][1] = nullptr;
    memset(e->sg, 0, sizeof(e->sg));
    e->graph_enabled = !(c->flags & BFCUDA_FLAG_NO_GRAPH) && getenv("BFCUDA_NO_GRAPH") == nullptr;
    e->graph_mode = e->graph_used = false;
    e->graph_host = 0;
    e->d_groups = nullptr;
    e->d_members = nullptr;
    e->any_group = false;
    for (int io = 0; io < 2; io++) {
        e->d_muted[io] = nullptr;
        e->any_muted[io] = false;
        e->d_sd_chans[io] = nullptr;
        e->d_sd_taps[io] = e->d_sd_hist[io] = nullptr;
    }
    e->powersave = 0;
    e->d_amax[0] = e->d_amax[1] = nullptr;
    e->d_ps_thr = nullptr;
    e->d_slot_zero = nullptr;
    e->ps_hold_until = -1;
    e->pend_mac.valid = e->pend_inv.valid = false;
    e->h_overflow_valid = false;
    e->io_waited = 0;
    memset(e->stage_ms, 0, sizeof(e->stage_ms));
    memset(e->timer, 0, sizeof(e->timer));
    memset(e->ring, 0, sizeof(e->ring));
    void **zero[] = { (void **)&e->d_raw[0], (void **)&e->d_raw[1], (void **)&e->d_raw2[0], (void **)&e->d_raw2[1],
                      (void **)&e->d_fmt[0], (void **)&e->d_fmt[1],
                      &e->d_prev[0], &e->d_prev[1], &e->d_fdl, &e->d_xin, &e->d_H, &e->d_Y, &e->d_out_time, &e->d_scratch,
                      (void **)&e->d_overflow, (void **)&e->d_status, (void **)&e->d_dests, (void **)&e->d_dest_first,
                      (void **)&e->d_need_xin, (void **)&e->d_mix_streams, (void **)&e->d_mix_terms,
                      (void **)&e->d_jobs, (void **)&e->d_chans, (void **)&e->d_out_terms };
    for (void **p : zero) {
        *p = nullptr;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess) {
        e->sm_count = prop.multiProcessorCount;
        snprintf(e->device_name, sizeof(e->device_name), "%s", prop.name);
    } else {
        e->sm_count = 148;
        e->device_name[0] = '\0';
    }
    for (int io = 0; io < 2; io++) {
        e->n_ch[io] = c->n_channels[io];
        e->n_bytes[io] = c->n_bytes[io];
        e->fmt[io].assign(c->formats[io], c->formats[io] + c->n_channels[io]);
        // the layouts every shipped config uses get compile-time paths in the FFT stages: all channels aligned
        // 4-byte little-endian integers (S32_LE, S24_4LE) or all FLOAT_LE
        int fast = e->n_ch[io] > 0 && e->n_bytes[io] % 4 == 0 ? (e->fmt[io][0].sf.isfloat ? 2 : 1) : 0;
        for (const bfcuda_buffer_format &b : e->fmt[io]) {
            if (b.sf.bytes != 4 || b.sf.swap || (b.byte_offset & 3) != 0 || (b.sf.isfloat ? 2 : 1) != fast ||
                (!b.sf.isfloat && b.sf.sbytes != 3 && b.sf.sbytes != 4)) {
                fast = 0;
            }
        }
        if (fast == 0 && e->n_ch[io] > 0 && e->n_ch[io] % 4 == 0) {
            // packed 24-bit little-endian, all channels interleaved in channel order (massive_config's / xtc_config's
            // "S24_LE"): tiles of up to 32 channels, 3 nc contiguous and 4-byte aligned bytes per sample time (nc % 4 == 0)
            bool packed = e->n_bytes[io] % 4 == 0;
            for (int ch = 0; ch < e->n_ch[io]; ch++) {
                const bfcuda_buffer_format &b = e->fmt[io][(size_t)ch];
                if (b.sf.isfloat || b.sf.swap || b.sf.bytes != 3 || b.sf.sbytes != 3 || b.sample_spacing != e->n_ch[io] ||
                    b.byte_offset != 3 * ch) {
                    packed = false;
                }
            }
            if (packed) {
                fast = 3;
            }
        }
        e->fast_fmt[io] = fast;
    }
    e->filters.resize(e->n_filters);
    for (int f = 0; f < e->n_filters; f++) {
        const bfcuda_filter &s = c->filters[f];
        FilterState &fs = e->filters[f];
        fs.crossfade = s.crossfade;
        for (int io = 0; io < 2; io++) {
            fs.ch[io].assign(s.channels[io], s.channels[io] + s.n_channels[io]);
            fs.scale[io].assign(s.scale[io], s.scale[io] + s.n_channels[io]);
        }
        fs.coeff = fs.prevcoeff = s.coeff;      // bfrun.c:1326
        fs.delayblocks = s.delayblocks;
        fs.level = 0;
        fs.eval = -1;
        fs.stream = f;
        fs.key_ch = -1;         // no key yet: update_streams() takes the first snapshot at block 0
        fs.key_delay = 0;
        fs.key_scale = 0.0;
        fs.key_since = -1;
        fs.mirror = nullptr;
        fs.trans_until = -1;
        if (s.n_filters_in > 0) {
            fs.fin.assign(s.filters_in, s.filters_in + s.n_filters_in);
            if (s.fscale != nullptr) {
                fs.fscale.assign(s.fscale, s.fscale + s.n_filters_in);
            } else {
                fs.fscale.assign(s.n_filters_in, 1.0);
            }
            for (int src : fs.fin) {
                fs.level = std::max(fs.level, e->filters[src].level + 1);
            }
            fs.eval = e->n_eval++;
            e->n_levels = std::max(e->n_levels, fs.level + 1);
        }
    }
    e->n_vin = e->n_ch[0] + e->n_eval;
    e->coeff_n_blocks.assign(c->coeff_n_blocks, c->coeff_n_blocks + c->n_coeffs);
    e->coeff_hbase.resize(e->n_coeffs);
    e->total_coeff_blocks = 0;
    for (int n = 0; n < e->n_coeffs; n++) {
        e->coeff_hbase[n] = e->total_coeff_blocks;
        e->total_coeff_blocks += e->coeff_n_blocks[n];
    }
    e->split = choose_split(e, c->mac_split);
    const char *variant = getenv("BFCUDA_MAC_VARIANT");
    e->mac_variant = variant != nullptr ? atoi(variant) : 0;
    if ((e->flags & BFCUDA_FLAG_LOW_LATENCY) && e->n_levels == 1 && e->P > 1) {
        // head = partition 0, tail = the others, one block ahead: a two-way split of uneven halves
        e->low_latency = true;
        e->split = 2;
        e->mac_variant = 0;
    }

    int rc = 0;
#define TRY(x)                    \
    do {                          \
        rc = (x);                 \
        if (rc != 0) goto error;  \
    } while (0)
#define TRYCU(call)                                                                                \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            rc = fail(BFCUDA_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__));               \
            goto error;                                                                            \
        }                                                                                          \
    } while (0)
    {
        const size_t N = e->N, L = e->L, F = std::max(1, e->n_filters);
        {
            // Stream priorities.  Round 1 gave the latency-bound FFT stages the higher priority (their blocks took the SMs
            // the MAC's retiring blocks freed instead of queueing behind its whole grid).  With the batched MAC that is the
            // wrong way round on small shards: a 128-block MAC launch that finds 64 + 64 transform blocks already holding
            // whole SMs starts a third of its blocks late, and the step is the MAC plus that delay (8-filter shard:
            // 57 us per step against 43 us with the MAC first; 16 / 32 filters 75-84 -> 72, 146 -> 133-135; the full job
            // 270 -> 259-266, block by block 170 -> 164; profiles/r2_prio*.txt).  BFCUDA_MAC_PRIO=0 restores the old order.
            int lo = 0, hi = 0;
            TRYCU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            const char *pe = getenv("BFCUDA_MAC_PRIO");
            const bool mac_first = pe == nullptr || atoi(pe) != 0;
            TRYCU(cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, mac_first ? lo : hi));
            TRYCU(cudaStreamCreateWithPriority(&e->s_mac, cudaStreamNonBlocking, mac_first ? hi : lo));
            TRYCU(cudaStreamCreateWithPriority(&e->s_inv, cudaStreamNonBlocking, mac_first ? lo : hi));
        }
        TRYCU(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
        TRYCU(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
        for (cudaEvent_t *ev : { &e->ev_h2d[0], &e->ev_h2d[1], &e->ev_fwd[0], &e->ev_fwd[1], &e->ev_call_done[0],
                                 &e->ev_call_done[1], &e->ev_call_done[2], &e->ev_call_done[3], &e->ev_cap[0],
                                 &e->ev_cap[1], &e->ev_cap[2], &e->ev_cap[3], &e->ev_g_done[0], &e->ev_g_done[1],
                                 &e->ev_g_done[2], &e->ev_g_done[3], &e->ev_out_read[0], &e->ev_out_read[1],
                                 &e->ev_out_read[2], &e->ev_out_read[3], &e->ev_h2dq[0], &e->ev_h2dq[1], &e->ev_h2dq[2],
                                 &e->ev_h2dq[3], &e->ev_inv, &e->ev_fwd_done[0], &e->ev_fwd_done[1],
                                 &e->ev_mac_done[0], &e->ev_mac_done[1], &e->ev_inv_done[0], &e->ev_inv_done[1],
                                 &e->ev_join, &e->ev_tail_done }) {
            TRYCU(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
        }
        TRYCU(fft_plan_create(&e->plan, e->N, e->rs, e->max_batch * std::max(1, std::max(e->n_ch[0], e->n_ch[1]))));
        TRYCU(cudaEventCreate(&e->timer[0]));
        TRYCU(cudaEventCreate(&e->timer[1]));
        if (e->flags & BFCUDA_FLAG_STAGE_TIMING) {
            for (int i = 0; i < TIMING_RING; i++) {
                for (int j = 0; j < 6; j++) {
                    TRYCU(cudaEventCreate(&e->ring[i][j]));
                }
            }
        }
        TRYCU(cudaMallocHost((void **)&e->h_status, sizeof(unsigned int)));
        *e->h_status = 0;
        e->snap_status_off = sizeof(Overflow) * (size_t)std::max(1, e->n_ch[1]);
        e->snap_bytes = e->snap_status_off + 16;
        for (int i = 0; i < 4; i++) {
            TRYCU(cudaMallocHost((void **)&e->h_snap[i], e->snap_bytes));
            memset(e->h_snap[i], 0, e->snap_bytes);
            TRY(dev_alloc(e, &e->d_snap[i], e->snap_bytes));
        }
        const size_t B = (size_t)e->max_batch;
        TRY(dev_alloc(e, &e->d_raw[0], B * e->n_bytes[0]));
        TRY(dev_alloc(e, &e->d_raw[1], B * e->n_bytes[1]));
        TRY(dev_alloc(e, &e->d_raw2[0], B * e->n_bytes[0]));
        TRY(dev_alloc(e, &e->d_raw2[1], B * e->n_bytes[1]));
        for (int io = 0; io < 2; io++) {
            TRY(dev_alloc(e, &e->d_raw34[io][0], B * e->n_bytes[io]));
            TRY(dev_alloc(e, &e->d_raw34[io][1], B * e->n_bytes[io]));
        }
        TRY(dev_alloc(e, &e->d_fmt[0], sizeof(SampleFormat) * std::max(1, e->n_ch[0])));
        TRY(dev_alloc(e, &e->d_fmt[1], sizeof(SampleFormat) * std::max(1, e->n_ch[1])));
        TRY(dev_alloc(e, &e->d_prev[0], rs_bytes(e, (size_t)e->n_ch[0] * L)));
        TRY(dev_alloc(e, &e->d_prev[1], rs_bytes(e, (size_t)e->n_ch[0] * L)));
        if (plan_unpacks_first(e->plan)) {
            TRY(dev_alloc(e, &e->d_xt[0], rs_bytes(e, B * (size_t)std::max(1, e->n_ch[0]) * L)));
            TRY(dev_alloc(e, &e->d_xt[1], rs_bytes(e, B * (size_t)std::max(1, e->n_ch[0]) * L)));
        }
        if (c->powersave && plan_unpacks_first(e->plan) && e->n_ch[0] > 0) {
            // (partitions shorter than 64 samples run the fused generic kernels, which have no separate unpack pass to
            // take the peaks in: there powersave is left off -- with exact-zero detection it changes no result anyway)
            const double level = c->analog_powersave <= 0.0 ? 1.0 : c->analog_powersave;
            e->powersave = level >= 1.0 ? 1 : 2;
            TRY(dev_alloc(e, &e->d_amax[0], sizeof(unsigned int) * B * (size_t)e->n_ch[0]));
            TRY(dev_alloc(e, &e->d_amax[1], sizeof(unsigned int) * B * (size_t)e->n_ch[0]));
            TRY(dev_alloc(e, &e->d_ps_thr, sizeof(float) * (size_t)e->n_ch[0]));
            TRY(dev_alloc(e, &e->d_slot_zero, F * (size_t)e->fdl_ring, false));
            TRYCU(cudaMemset(e->d_slot_zero, 1, F * (size_t)e->fdl_ring));      // nothing written yet: every slot is zeros
            std::vector<float> thr((size_t)e->n_ch[0]);
            for (int n = 0; n < e->n_ch[0]; n++) {
                // silent iff scale * peak < level (bfrun.c:767); the peak is kept in float
                thr[(size_t)n] = (float)(level / e->fmt[0][n].sf.scale);
            }
            TRYCU(cudaMemcpy(e->d_ps_thr, thr.data(), sizeof(float) * thr.size(), cudaMemcpyHostToDevice));
        } else if (c->powersave && c->analog_powersave > 0.0 && c->analog_powersave < 1.0) {
            rc = fail(BFCUDA_ENOTSUP, "analog powersave needs partitions of at least 64 samples");
            goto error;
        }
        TRY(dev_alloc(e, &e->d_fdl, rs_bytes(e, F * (size_t)e->fdl_ring * N)));
        TRY(dev_alloc(e, &e->d_xin, rs_bytes(e, B * (size_t)std::max(1, e->n_vin) * N)));
        TRY(dev_alloc(e, &e->d_keep, rs_bytes(e, (size_t)std::max(1, e->n_eval) * L)));
        TRY(dev_alloc(e, &e->d_eval_entries, sizeof(EvalEntry) * std::max(1, e->n_eval)));
        {
            size_t terms = 1;
            for (const FilterState &fs : e->filters) {
                terms += 2 * fs.fin.size();
            }
            TRY(dev_alloc(e, &e->d_eval_terms, sizeof(MixTerm) * terms));
        }
        TRY(dev_alloc(e, &e->d_H, rs_bytes(e, (size_t)std::max(1, e->total_coeff_blocks) * N)));
        e->y_stride = rs_bytes(e, (size_t)e->split * B * (2 * F + 2 * (size_t)e->n_ch[1]) * N);
        TRY(dev_alloc(e, &e->d_Y, 2 * e->y_stride));
        TRY(dev_alloc(e, &e->d_out_time, rs_bytes(e, B * (size_t)std::max(1, e->n_ch[1]) * L)));
        TRY(dev_alloc(e, &e->d_scratch, rs_bytes(e, 4 * N)));
        {
            // the overflow records and the status word share one allocation: one copy snapshots both
            char *blob = nullptr;
            TRY(dev_alloc(e, &blob, e->snap_bytes));
            e->d_overflow = reinterpret_cast<Overflow *>(blob);
            e->d_status = reinterpret_cast<unsigned int *>(blob + e->snap_status_off);
        }
        TRY(dev_alloc(e, &e->d_dests, sizeof(FwdDest) * F));
        TRY(dev_alloc(e, &e->d_dest_first, sizeof(int) * (e->n_ch[0] + 1)));
        TRY(dev_alloc(e, &e->d_need_xin, (size_t)std::max(1, e->n_ch[0])));
        TRY(dev_alloc(e, &e->d_mix_streams, sizeof(MixStream) * F));
        {
            size_t terms = 1;
            for (const FilterState &fs : e->filters) {
                terms += fs.ch[0].size() + 1;
            }
            TRY(dev_alloc(e, &e->d_mix_terms, sizeof(MixTerm) * terms));
            terms = 1;
            for (const FilterState &fs : e->filters) {
                terms += 2 * fs.ch[1].size();
            }
            terms += 2 * (size_t)e->n_ch[1];
            TRY(dev_alloc(e, &e->d_out_terms, sizeof(MixTerm) * terms));
        }
        TRY(dev_alloc(e, &e->d_jobs, sizeof(MacJob) * 2 * F));
        TRY(dev_alloc(e, &e->d_chans, sizeof(OutChan) * std::max(1, e->n_ch[1])));
        TRY(dev_alloc(e, &e->d_mixes, sizeof(OutChan) * std::max(1, e->n_ch[1])));
        e->h_dest_first.assign(e->n_ch[0] + 1, 0);
        e->h_need_xin.assign(std::max(1, e->n_ch[0]), 0);
        e->h_chans.assign(e->n_ch[1], OutChan());
        e->h_mixes.assign(e->n_ch[1], OutChan());

        for (int io = 0; io < 2; io++) {
            std::vector<SampleFormat> f(e->n_ch[io]);
            for (int n = 0; n < e->n_ch[io]; n++) {
                f[n] = to_dev_format(e->fmt[io][n]);
            }
            if (!f.empty()) {
                TRYCU(cudaMemcpy(e->d_fmt[io], f.data(), sizeof(SampleFormat) * f.size(), cudaMemcpyHostToDevice));
            }
        }
        // virtual -> physical outputs (bfconf->virt2phys[OUT], n_virtperphys): group the outputs by physical id
        e->out_rep.resize((size_t)std::max(1, e->n_ch[1]));
        for (int io = 0; io < 2; io++) {
            e->h_muted[io].assign((size_t)std::max(1, e->n_ch[io]), 0);
            TRY(dev_alloc(e, &e->d_muted[io], (size_t)std::max(1, e->n_ch[io])));
        }
        {
            std::vector<char> seen((size_t)std::max(1, e->n_ch[1]), 0);
            for (int o = 0; o < e->n_ch[1]; o++) {
                if (seen[(size_t)o]) {
                    continue;
                }
                e->out_rep[(size_t)o] = o;
                VirtGroup g;
                g.first = (int)e->h_members.size();
                g.n = 0;
                for (int q = o; q < e->n_ch[1]; q++) {
                    if (q == o || (c->out_physical != nullptr && c->out_physical[q] == c->out_physical[o])) {
                        const bfcuda_buffer_format &a = e->fmt[1][o], &b = e->fmt[1][q];
                        if (a.byte_offset != b.byte_offset || a.sample_spacing != b.sample_spacing || a.sf.bytes != b.sf.bytes ||
                            a.sf.sbytes != b.sf.sbytes || a.sf.isfloat != b.sf.isfloat || a.sf.swap != b.sf.swap) {
                            rc = fail(BFCUDA_EINVAL, "outputs %d and %d share a physical channel but not a buffer format", o, q);
                            goto error;
                        }
                        seen[(size_t)q] = 1;
                        e->h_members.push_back(q);
                        g.n++;
                    }
                }
                for (int j = 0; j < g.n; j++) {
                    e->out_rep[(size_t)e->h_members[(size_t)(g.first + j)]] = e->h_members[(size_t)(g.first + g.n - 1)];
                }
                e->any_group = e->any_group || g.n > 1;
                e->h_groups.push_back(g);
            }
            if (e->any_group) {
                if (!plan_unpacks_first(e->plan)) {
                    rc = fail(BFCUDA_ENOTSUP, "several outputs on one physical channel need partitions of at least 64 samples");
                    goto error;
                }
                for (int o = 0; o < e->n_ch[1]; o++) {
                    if (c->apply_dither != nullptr && c->apply_dither[o] && e->out_rep[(size_t)o] != o) {
                        rc = fail(BFCUDA_ENOTSUP, "dither on a physical channel that carries several virtual outputs");
                        goto error;
                    }
                }
            }
            TRY(dev_alloc(e, &e->d_groups, sizeof(VirtGroup) * std::max<size_t>(1, e->h_groups.size())));
            TRY(dev_alloc(e, &e->d_members, sizeof(int) * std::max<size_t>(1, e->h_members.size())));
            if (!e->h_groups.empty()) {
                TRYCU(cudaMemcpy(e->d_groups, e->h_groups.data(), sizeof(VirtGroup) * e->h_groups.size(), cudaMemcpyHostToDevice));
                TRYCU(cudaMemcpy(e->d_members, e->h_members.data(), sizeof(int) * e->h_members.size(), cudaMemcpyHostToDevice));
            }
        }
        TRY(setup_dither(e, c));
        TRY(bfcuda_reset_overflow(e));
        e->dirty = true;
        e->xfade_active = false;
        build_tables(e);
        e->dirty = true;    // tables still have to be uploaded by the first block
    }
    *out = e;
    return 0;
error:
    bfcuda_destroy(e);
    return rc;
#undef TRY
#undef TRYCU
}

int bfcuda_reset_overflow(bfcuda_engine *e)
{
    if (e == nullptr) return fail(BFCUDA_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    std::vector<Overflow> of(e->n_ch[1]);
    for (int n = 0; n < e->n_ch[1]; n++) {
        of[n].n_overflows = 0;
        of[n].intlargest = 0;
        of[n].largest = 0.0;
        // bfrun.c:2264-2279
        if (e->fmt[1][n].sf.isfloat) {
            of[n].max = 1.0;
        } else {
            of[n].max = (double)((uint64_t)1 << ((e->fmt[1][n].sf.sbytes << 3) - 1)) - 1;
        }
    }
    if (e->stream != nullptr) {
        int rc = sync_all(e);
        if (rc != 0) return rc;
    }
    if (!of.empty()) {
        CU(cudaMemcpy(e->d_overflow, of.data(), sizeof(Overflow) * of.size(), cudaMemcpyHostToDevice));
    }
    CU(cudaMemset(e->d_status, 0, sizeof(unsigned int)));
    e->h_overflow_valid = false;
    return 0;
}

int bfcuda_get_overflow(bfcuda_engine *e, int out_channel, struct bfcuda_overflow *overflow)
{
    if (e == nullptr || overflow == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    if (out_channel < 0 || out_channel >= e->n_ch[1]) return fail(BFCUDA_EINVAL, "output channel out of range");
    // the virtual outputs of one physical channel share its record (bfrun.c:1994-1998)
    out_channel = e->out_rep.empty() ? out_channel : e->out_rep[(size_t)out_channel];
    Overflow of;
    if (e->h_overflow_valid && e->io_count > 0) {
        // host-buffer interface: the records came back with the most recent call's output -- wait for that read-out
        // only (no other stream: the next block's ahead-of-time work keeps running), no copy; after a synchronous
        // call not even that (the per-output peak-meter loop of bfrun.c:1929-1936 costs nothing)
        if (e->io_waited != e->io_count) {
            CU(cudaSetDevice(e->device));
            int rc = wait_call(e, e->io_count - 1u);
            if (rc != 0) return rc;
            e->io_waited = e->io_count;
        }
        of = reinterpret_cast<const Overflow *>(e->h_snap[(e->io_count - 1u) & 3u])[out_channel];
    } else {
        CU(cudaSetDevice(e->device));
        int rc = sync_all(e);
        if (rc != 0) return rc;
        CU(cudaMemcpy(&of, e->d_overflow + out_channel, sizeof(of), cudaMemcpyDeviceToHost));
    }
    overflow->n_overflows = of.n_overflows;
    overflow->intlargest = of.intlargest;
    overflow->largest = of.largest;
    overflow->max = of.max;
    return 0;
}

// ---- coefficients ------------------------------------------------------------------------------------

// The multiply-accumulate of the most recent launch (s_mac), and in the low-latency schedule the ahead-of-time sum,
// may still be reading d_H when a run-time coefficient update arrives: order the update (on e->stream) behind them so
// that a block sees a coefficient set either wholly old or wholly new (bfrun.c:1462-1478 snapshots per block).
static int wait_coeff_readers(bfcuda_engine *e)
{
    if (e->launch_no >= 1) {
        CU(cudaStreamWaitEvent(e->stream, e->ev_mac_done[(e->launch_no - 1) & 1u], 0));
    }
    if (e->low_latency) {
        CU(cudaStreamWaitEvent(e->stream, e->ev_tail_done, 0));
    }
    return 0;
}

static int check_coeff(bfcuda_engine *e, int coeff, int block)
{
    if (e == nullptr) return fail(BFCUDA_EINVAL, "null engine");
    e->tail_ready = false;      // every coefficient setter comes through here: a sum made ahead of time is stale
    if (coeff < 0 || coeff >= e->n_coeffs) return fail(BFCUDA_EINVAL, "coefficient index %d out of range", coeff);
    if (block < 0 || block >= e->coeff_n_blocks[coeff]) {
        return fail(BFCUDA_EINVAL, "coefficient block %d out of range", block);
    }
    return 0;
}

int bfcuda_coeff_from_taps(bfcuda_engine *e, int coeff, const void *taps, int n_taps, double scale)
{
    int rc = check_coeff(e, coeff, 0);
    if (rc != 0) return rc;
    if (taps == nullptr || n_taps < 0) return fail(BFCUDA_EINVAL, "bad taps");
    CU(cudaSetDevice(e->device));
    const int nb = e->coeff_n_blocks[coeff];
    const size_t total = (size_t)nb * e->L;
    const size_t n = std::min<size_t>((size_t)n_taps, total);
    // NaN/Inf check on the scaled taps in the real type (fftw_convolver.c:540-556); zero-extension to whole
    // blocks as load_coeff does (bfconf.c:1982-2019)
    std::vector<unsigned char> padded(total * e->rs, 0);
    if (e->rs == 4) {
        const float *src = (const float *)taps;
        const float s = (float)scale;
        for (size_t i = 0; i < n; i++) {
            if (!std::isfinite((double)(src[i] * s))) {
                return fail(BFCUDA_ENONFINITE, "NaN or Inf value among coefficients.");
            }
        }
    } else {
        const double *src = (const double *)taps;
        for (size_t i = 0; i < n; i++) {
            if (!std::isfinite(src[i] * scale)) {
                return fail(BFCUDA_ENONFINITE, "NaN or Inf value among coefficients.");
            }
        }
    }
    memcpy(padded.data(), taps, n * e->rs);
    void *d_taps = nullptr;
    CU(cudaMalloc(&d_taps, padded.size()));
    rc = wait_coeff_readers(e);
    if (rc != 0) {
        cudaFree(d_taps);
        return rc;
    }
    cudaError_t err = cudaMemcpyAsync(d_taps, padded.data(), padded.size(), cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) {
        err = launch_coeff_fft(e->plan, d_taps, nb, scale, e->d_H, e->coeff_hbase[coeff], e->stream);
    }
    if (err == cudaSuccess) {
        err = cudaStreamSynchronize(e->stream);
    }
    cudaFree(d_taps);
    if (err != cudaSuccess) {
        return fail(BFCUDA_ECUDA, "coefficient preprocessing failed: %s", cudaGetErrorString(err));
    }
    return 0;
}

int bfcuda_coeff_runtime_block(bfcuda_engine *e, int coeff, int block, const void *taps_L)
{
    int rc = check_coeff(e, coeff, block);
    if (rc != 0) return rc;
    CU(cudaSetDevice(e->device));
    // convolver_runtime_coeffs2cbuf (fftw_convolver.c:575-596): no scale, no NaN check.  Ordered behind the
    // multiply-accumulate in flight, so a running engine picks the new block up at a block boundary.
    rc = wait_coeff_readers(e);
    if (rc != 0) return rc;
    CU(cudaMemcpyAsync(e->d_scratch, taps_L, rs_bytes(e, e->L), cudaMemcpyHostToDevice, e->stream));
    CU(launch_coeff_fft(e->plan, e->d_scratch, 1, 1.0, e->d_H, e->coeff_hbase[coeff] + block, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

int bfcuda_coeff_set_block(bfcuda_engine *e, int coeff, int block, const void *cbuf)
{
    int rc = check_coeff(e, coeff, block);
    if (rc != 0) return rc;
    CU(cudaSetDevice(e->device));
    char *dst = (char *)e->d_H + rs_bytes(e, (size_t)(e->coeff_hbase[coeff] + block) * e->N);
    rc = wait_coeff_readers(e);
    if (rc != 0) return rc;
    CU(cudaMemcpyAsync(e->d_scratch, cbuf, rs_bytes(e, e->N), cudaMemcpyHostToDevice, e->stream));
    CU(launch_permute(e->plan, e->d_scratch, dst, 1, BLOCKED_TO_PLANAR, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

int bfcuda_coeff_get_block(bfcuda_engine *e, int coeff, int block, void *cbuf)
{
    int rc = check_coeff(e, coeff, block);
    if (rc != 0) return rc;
    CU(cudaSetDevice(e->device));
    const char *src = (const char *)e->d_H + rs_bytes(e, (size_t)(e->coeff_hbase[coeff] + block) * e->N);
    CU(launch_permute(e->plan, src, e->d_scratch, 1, PLANAR_TO_BLOCKED, e->stream));
    CU(cudaMemcpyAsync(cbuf, e->d_scratch, rs_bytes(e, e->N), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

// ---- control -----------------------------------------------------------------------------------------

int bfcuda_set_control(bfcuda_engine *e, int filter, const struct bfcuda_filter_control *c)
{
    if (e == nullptr || c == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    if (filter < 0 || filter >= e->n_filters) return fail(BFCUDA_EINVAL, "filter index out of range");
    if (c->coeff >= e->n_coeffs) return fail(BFCUDA_EINVAL, "coefficient index %d out of range", c->coeff);
    FilterState &fs = e->filters[filter];
    // The host takes this snapshot for every filter before every block (bfrun.c:1462-1478); mostly nothing has changed,
    // and then nothing may happen here either: a rebuild of the tables waits for the previous launch's stages and
    // discards a partition sum made ahead of time.
    const int new_coeff = c->coeff < 0 ? -1 : c->coeff;
    bool changed = new_coeff != fs.coeff || c->delayblocks != fs.delayblocks;
    for (int io = 0; io < 2 && !changed; io++) {
        if (c->scale[io] != nullptr) {
            changed = !std::equal(fs.scale[io].begin(), fs.scale[io].end(), c->scale[io]);
        }
    }
    if (!changed && c->fscale != nullptr) {
        changed = !std::equal(fs.fscale.begin(), fs.fscale.end(), c->fscale);
    }
    if (!changed) {
        return 0;
    }
    fs.coeff = new_coeff;
    if (clamp_delay(e, c->delayblocks) != clamp_delay(e, fs.delayblocks)) {
        // remember the delay the ring was last written under (first change since the last block wins)
        bool known = false;
        for (const std::pair<int, int> &fx : e->delay_changes) {
            known = known || fx.first == filter;
        }
        if (!known) {
            e->delay_changes.push_back(std::make_pair(filter, clamp_delay(e, fs.delayblocks)));
        }
    }
    fs.delayblocks = c->delayblocks;
    for (int io = 0; io < 2; io++) {
        if (c->scale[io] != nullptr) {
            fs.scale[io].assign(c->scale[io], c->scale[io] + fs.ch[io].size());
        }
    }
    if (c->fscale != nullptr) {
        fs.fscale.assign(c->fscale, c->fscale + fs.fin.size());
    }
    e->dirty = true;    // takes effect at the next block, like the snapshot at bfrun.c:1462-1478
    return 0;
}

int bfcuda_set_mute(bfcuda_engine *e, int io, int channel, int muted)
{
    if (e == nullptr || (io != 0 && io != 1)) return fail(BFCUDA_EINVAL, "bad argument");
    if (channel < 0 || channel >= e->n_ch[io]) return fail(BFCUDA_EINVAL, "channel out of range");
    if (!plan_unpacks_first(e->plan)) return fail(BFCUDA_ENOTSUP, "mute needs partitions of at least 64 samples");
    if ((e->h_muted[io][(size_t)channel] != 0) == (muted != 0)) {
        return 0;
    }
    CU(cudaSetDevice(e->device));
    int rc = sync_all(e);       // the flags are read by launches in flight: change them between blocks
    if (rc != 0) return rc;
    e->h_muted[io][(size_t)channel] = muted ? 1 : 0;
    e->any_muted[io] = false;
    for (uint8_t m : e->h_muted[io]) {
        e->any_muted[io] = e->any_muted[io] || m != 0;
    }
    CU(cudaMemcpy(e->d_muted[io], e->h_muted[io].data(), e->h_muted[io].size(), cudaMemcpyHostToDevice));
    return 0;
}

int bfcuda_set_subdelay(bfcuda_engine *e, int io, int channel, const void *taps, int n_taps)
{
    if (e == nullptr || (io != 0 && io != 1)) return fail(BFCUDA_EINVAL, "bad argument");
    if (channel < 0 || channel >= e->n_ch[io]) return fail(BFCUDA_EINVAL, "channel out of range");
    if (taps != nullptr && (n_taps < 1 || n_taps > BF_SUBDELAY_MAX_TAPS)) {
        return fail(BFCUDA_EINVAL, "sub-sample delay filter of %d taps (1..%d)", n_taps, BF_SUBDELAY_MAX_TAPS);
    }
    if (!plan_unpacks_first(e->plan)) return fail(BFCUDA_ENOTSUP, "sub-sample delay needs partitions of at least 64 samples");
    CU(cudaSetDevice(e->device));
    int rc = sync_all(e);
    if (rc != 0) return rc;
    const size_t nch = (size_t)std::max(1, e->n_ch[io]);
    if (e->d_sd_taps[io] == nullptr) {
        if ((rc = dev_alloc(e, &e->d_sd_taps[io], rs_bytes(e, nch * BF_SUBDELAY_MAX_TAPS))) != 0) return rc;
        if ((rc = dev_alloc(e, &e->d_sd_hist[io], rs_bytes(e, nch * (BF_SUBDELAY_MAX_TAPS - 1)))) != 0) return rc;
        if ((rc = dev_alloc(e, &e->d_sd_chans[io], sizeof(SubdelayChan) * nch)) != 0) return rc;
    }
    std::vector<SubdelayChan> &list = e->h_sd[io];
    for (size_t i = 0; i < list.size(); i++) {
        if (list[i].ch == channel) {
            list.erase(list.begin() + (long)i);
            break;
        }
    }
    if (taps != nullptr) {
        SubdelayChan sc;
        sc.ch = channel;
        sc.tap_first = channel * BF_SUBDELAY_MAX_TAPS;
        sc.n_taps = n_taps;
        list.push_back(sc);
        CU(cudaMemcpy((char *)e->d_sd_taps[io] + rs_bytes(e, (size_t)sc.tap_first), taps, rs_bytes(e, (size_t)n_taps),
                      cudaMemcpyHostToDevice));
    }
    if (!list.empty()) {
        CU(cudaMemcpy(e->d_sd_chans[io], list.data(), sizeof(SubdelayChan) * list.size(), cudaMemcpyHostToDevice));
    }
    return 0;
}

// ---- the block step ------------------------------------------------------------------------------------

static int flush_timing_ring(bfcuda_engine *e)
{
    if (e->ring_fill == 0) {
        return 0;
    }
    CU(cudaEventSynchronize(e->ring[e->ring_fill - 1][5]));
    for (int i = 0; i < e->ring_fill; i++) {
        // forward: [0,1], MAC: [2,3] on the main stream; inverse: [4,5] on its own stream
        static const int from[BFCUDA_N_STAGES] = { 0, 2, 4 }, to[BFCUDA_N_STAGES] = { 1, 3, 5 };
        for (int s = 0; s < BFCUDA_N_STAGES; s++) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, e->ring[i][from[s]], e->ring[i][to[s]]));
            e->stage_ms[s] += ms;
        }
    }
    for (int i = 0; i < e->ring_fill; i++) {
        e->stage_blocks += e->ring_blocks[i];
    }
    e->ring_fill = 0;
    return 0;
}

// ---- kernel argument blocks of one launch (shared by enqueue_batch and the step graphs) --------------------------
static bool in_transition(const bfcuda_engine *e);
static ForwardArgs make_forward_args(const bfcuda_engine *e, int nb, int slot_t, const uint8_t *raw_in)
{
    ForwardArgs fa;
    fa.raw_in = raw_in;
    fa.fmt = e->d_fmt[0];
    fa.prev_in = e->d_prev[e->prev_par];
    fa.prev_out = e->d_prev[e->prev_par ^ 1];
    fa.fdl = e->d_fdl;
    fa.xin = e->d_xin;
    fa.need_xin = e->d_need_xin;
    fa.dest_first = e->d_dest_first;
    fa.dests = e->d_dests;
    fa.n_in = e->n_ch[0];
    fa.n_vin = e->n_vin;
    fa.ring = e->fdl_ring;
    fa.t = slot_t;
    fa.batch = nb;
    fa.in_stride = (size_t)e->n_bytes[0];
    fa.fast_fmt = e->fast_fmt[0];
    fa.xt_cur = fa.xt_prev = nullptr;
    fa.single_dest = e->single_dest ? 1 : 0;
    fa.powersave = 0;           // set_xt_generation() fills these in on the unpack-first paths
    fa.amax_cur = fa.amax_prev = nullptr;
    fa.ps_thr = e->d_ps_thr;
    fa.slot_zero = e->d_slot_zero;
    return fa;
}

// the operands that follow the generation of the unpacked-sample buffers: this launch's blocks in d_xt[gen], the block
// before them the last one of the other generation (and the same for the powersave peaks)
static void set_xt_generation(const bfcuda_engine *e, ForwardArgs &fa, int gen, int prev_nb)
{
    fa.xt_cur = e->d_xt[gen];
    fa.xt_prev = (const char *)e->d_xt[gen ^ 1] + rs_bytes(e, (size_t)(prev_nb - 1) * e->n_ch[0] * e->L);
    if (e->powersave) {
        fa.powersave = e->powersave;
        fa.amax_cur = e->d_amax[gen];
        fa.amax_prev = e->d_amax[gen ^ 1] + (size_t)(prev_nb - 1) * e->n_ch[0];
    }
}

static UnpackArgs make_unpack_args(const bfcuda_engine *e, int nb, const uint8_t *raw_in, int xt_gen)
{
    UnpackArgs ua;
    ua.raw_in = raw_in;
    ua.fmt = e->d_fmt[0];
    ua.xt = e->d_xt[xt_gen];
    ua.n_in = e->n_ch[0];
    ua.batch = nb;
    ua.L = e->L;
    ua.in_stride = (size_t)e->n_bytes[0];
    ua.fast_fmt = e->fast_fmt[0];
    ua.amax = e->powersave ? e->d_amax[xt_gen] : nullptr;
    ua.muted = e->any_muted[0] ? e->d_muted[0] : nullptr;
    return ua;
}

static MacArgs make_mac_args(const bfcuda_engine *e, int nb, int slot_t, int y_gen)
{
    MacArgs ma;
    ma.fdl = e->d_fdl;
    ma.H = e->d_H;
    ma.Y = (char *)e->d_Y + (size_t)y_gen * e->y_stride;
    ma.jobs = e->d_jobs;
    ma.n_jobs = e->level_job_first[1];
    ma.n_slots = 2 * std::max(1, e->n_filters) + 2 * e->n_ch[1];
    ma.ring = e->fdl_ring;
    ma.split = e->split;
    ma.t = slot_t;
    ma.batch = nb;
    ma.variant = (e->mac_variant == 1 && !mac_tma_applicable(e->plan)) ? 0 : e->mac_variant;
    ma.head = ma.z_first = ma.z_count = 0;
    // powersave: zero slots are skipped -- unless a delay transition has been rewriting slots behind the flags' back
    ma.slot_zero = (e->powersave && !in_transition(e) && (long)e->t >= e->ps_hold_until) ? e->d_slot_zero : nullptr;
    ma.neg_zero2 = 0x8000000080000000ull;   // two packed -0.0f (bf_mac_batch.cu, BinPairAcc): the launcher sets the same
    return ma;
}

static InverseArgs make_inverse_args(const bfcuda_engine *e, int nb, int y_gen, uint8_t *raw_out)
{
    InverseArgs ia;
    ia.Y = (char *)e->d_Y + (size_t)y_gen * e->y_stride;
    ia.chans = e->d_chans;
    ia.terms = e->d_out_terms;
    ia.out_time = e->d_out_time;
    ia.raw_out = raw_out;
    ia.fmt = e->d_fmt[1];
    ia.overflow = e->d_overflow;
    ia.status = e->d_status;
    ia.n_out = e->n_ch[1];
    ia.n_slots = 2 * std::max(1, e->n_filters) + 2 * e->n_ch[1];
    ia.split = 1;               // already reduced (launch_split_reduce)
    ia.batch = nb;
    ia.out_stride = (size_t)e->n_bytes[1];
    ia.safety_limit = e->safety_limit;
    ia.fast_fmt = e->fast_fmt[1];
    if (ia.fast_fmt == 3) {
        // the packed path shuffles across all 32 lanes of a tile: not with outputs k_pack leaves out (dither, virtual groups)
        bool whole = true;
        for (size_t o = 0; o < e->out_rep.size(); o++) {
            whole = whole && e->out_rep[o] == (int)o;
        }
        for (int d : e->dither_of_out) {
            whole = whole && d < 0;
        }
        if (!whole) {
            ia.fast_fmt = 0;
        }
    }
    ia.simple_mix = e->simple_mix ? 1 : 0;
    ia.any_xfade = e->xfade_active ? 1 : 0;
    return ia;
}

static SubdelayArgs make_subdelay_args(const bfcuda_engine *e, int io, void *data, int nb)
{
    SubdelayArgs sa;
    sa.data = data;
    sa.chans = e->d_sd_chans[io];
    sa.taps = e->d_sd_taps[io];
    sa.hist = e->d_sd_hist[io];
    sa.n_chans = (int)e->h_sd[io].size();
    sa.n_ch = e->n_ch[io];
    sa.batch = nb;
    sa.L = e->L;
    return sa;
}

static OutMixArgs make_out_mix_args(const bfcuda_engine *e, int nb, const InverseArgs &ia)
{
    OutMixArgs oa;
    oa.Y = const_cast<void *>(ia.Y);
    oa.mixes = e->d_mixes;
    oa.terms = e->d_out_terms;
    oa.n_out = e->n_ch[1];
    oa.n_slots = ia.n_slots;
    oa.z_first = 2 * std::max(1, e->n_filters);
    oa.batch = nb;
    return oa;
}

// Enqueue the kernels of `nb` consecutive blocks as ONE launch per stage (nb <= max_batch; callers make sure
// no control change or crossfade falls inside): unpack + forward on the main stream, MAC on s_mac, inverse + pack on
// s_inv, ordered by per-parity events (launch n: p = n & 1):
//   forward(n)  after MAC(n-2)      -- the ring slots it overwrites were last read there (ring = 2P + 2B)
//   MAC(n)      after forward(n) and inverse(n-2)   -- Y[p] is free again
//   inverse(n)  after MAC(n)
//   raw_in / raw_out : device raw blocks, block b at + b * n_bytes
//   in_ready         : event the forward stage must wait for (input copy), or null
//   out_free         : event the inverse stage must wait for (previous read-out of raw_out), or null
//   fwd_done         : recorded after the forward stage (raw_in may be overwritten afterwards), or null
// On return e->ev_inv marks the end of the inverse stage.
static bool in_transition(const bfcuda_engine *e);
static int transition_repair(bfcuda_engine *e, int level, cudaStream_t stream);

static int enqueue_batch(bfcuda_engine *e, int nb, uint8_t *raw_in, uint8_t *raw_out, cudaEvent_t in_ready,
                         cudaEvent_t out_free, cudaEvent_t fwd_done)
{
    const bool timing = (e->flags & BFCUDA_FLAG_STAGE_TIMING) != 0;
    cudaEvent_t *ev = nullptr;
    if (timing) {
        if (e->ring_fill == TIMING_RING) {
            int rc = flush_timing_ring(e);
            if (rc != 0) return rc;
        }
        ev = e->ring[e->ring_fill];
        e->ring_blocks[e->ring_fill] = nb;
    }
    const int par = (int)(e->launch_no & 1u);
    const bool have_prev2 = e->launch_no >= 2;
    if (in_ready != nullptr) {
        CU(cudaStreamWaitEvent(e->stream, in_ready, 0));
    }
    if (have_prev2) {
        CU(cudaStreamWaitEvent(e->stream, e->ev_mac_done[par], 0));
    }
    if (e->n_levels > 1 && e->launch_no >= 1) {
        // chained filters: the previous launch's later levels still read the input spectra this forward stage rewrites
        CU(cudaStreamWaitEvent(e->stream, e->ev_mac_done[par ^ 1], 0));
    }
    if ((e->flags & BFCUDA_FLAG_SERIAL_STAGES) && e->launch_no >= 1) {
        // measurement aid: no overlap between launches, so the stage events bracket each stage running alone
        CU(cudaStreamWaitEvent(e->stream, e->ev_inv_done[par ^ 1], 0));
    }
    if (timing) CU(cudaEventRecord(ev[0], e->stream));

    ForwardArgs fa = make_forward_args(e, nb, e->slot_t, raw_in);
    if (plan_unpacks_first(e->plan)) {
        // size-specialised / four-step path: unpack the raw blocks into planar reals first (raw2real), transforms read those
        e->xt_par ^= 1;
        UnpackArgs ua = make_unpack_args(e, nb, raw_in, e->xt_par);
        if (ua.amax != nullptr) {
            CU(cudaMemsetAsync(ua.amax, 0, sizeof(unsigned int) * (size_t)nb * e->n_ch[0], e->stream));
        }
        CU(launch_unpack(e->plan, ua, e->stream));
        e->launches += e->n_ch[0] > 0;
        if (!e->h_sd[0].empty()) {
            // the postprocess hook of convolver_raw2cbuf: delay_subsample_update on the new samples (bfrun.c:1503-1526)
            CU(launch_subdelay(e->plan, make_subdelay_args(e, 0, ua.xt, nb), e->stream));
            e->launches++;
        }
        set_xt_generation(e, fa, e->xt_par, e->xt_last_nb);
        e->xt_last_nb = nb;
    }
    CU(launch_forward(e->plan, fa, e->stream));
    e->prev_par ^= 1;
    e->launches += e->n_ch[0] > 0;
    StreamMixArgs sa;
    sa.xin = e->d_xin;
    sa.fdl = e->d_fdl;
    sa.terms = e->d_mix_terms;
    sa.n_in = e->n_vin;
    sa.ring = e->fdl_ring;
    sa.t = e->slot_t;
    sa.batch = nb;
    sa.slot_zero = e->d_slot_zero;
    if (e->level_mix_first[1] > 0) {
        // mixes of input channels only (level 0); the mixes that contain an evaluated filter output follow their
        // source filters' MAC below
        sa.streams = e->d_mix_streams;
        sa.n_streams = e->level_mix_first[1];
        CU(launch_stream_mix(e->plan, sa, e->stream));
        e->launches++;
    }
    const bool transition = in_transition(e);      // then nb == 1 (enqueue_blocks)
    if (transition) {
        int rc = transition_repair(e, 0, e->stream);
        if (rc != 0) return rc;
    }
    if (timing) CU(cudaEventRecord(ev[1], e->stream));
    if (fwd_done != nullptr) {
        CU(cudaEventRecord(fwd_done, e->stream));
    }
    CU(cudaEventRecord(e->ev_fwd_done[par], e->stream));
    CU(cudaStreamWaitEvent(e->s_mac, e->ev_fwd_done[par], 0));
    if (have_prev2) {
        CU(cudaStreamWaitEvent(e->s_mac, e->ev_inv_done[par], 0));     // inverse(n-2) has consumed Y[par]
    }
    if (timing) CU(cudaEventRecord(ev[2], e->s_mac));

    MacArgs ma = make_mac_args(e, nb, e->slot_t, par);
    const bool ll = e->low_latency && nb == 1;
    if (ll) {
        ma.head = 1;
        if (e->tail_ready) {
            ma.z_count = 1;     // partial 1 was computed ahead of time: only partition 0 is left
        }
    }
    e->tail_ready = false;
    CU(launch_mac(e->plan, ma, e->s_mac));
    e->launches += ma.n_jobs > 0;
    if (e->split > 1) {
        CU(launch_split_reduce(e->plan, ma, e->s_mac));
        e->launches += ma.n_jobs > 0;
    }
    for (int level = 1; level < e->n_levels; level++) {
        // filter -> filter chaining (bfrun.c:1603-1660): evaluate the finished source outputs, mix them with the
        // channel inputs into the consumers' delay lines, then run the consumers' partitions
        EvalArgs ea;
        ea.Y = ma.Y;
        ea.entries = e->d_eval_entries + e->level_eval_first[level];
        ea.terms = e->d_eval_terms;
        ea.keep = e->d_keep;
        ea.xin = e->d_xin;
        ea.n_entries = e->level_eval_first[level + 1] - e->level_eval_first[level];
        ea.n_in = e->n_ch[0];
        ea.n_vin = e->n_vin;
        ea.n_slots = ma.n_slots;
        ea.split = 1;           // already reduced
        ea.batch = nb;
        CU(launch_eval(e->plan, ea, e->s_mac));
        sa.streams = e->d_mix_streams + e->level_mix_first[level];
        sa.n_streams = e->level_mix_first[level + 1] - e->level_mix_first[level];
        CU(launch_stream_mix(e->plan, sa, e->s_mac));
        if (transition) {
            int rc = transition_repair(e, level, e->s_mac);
            if (rc != 0) return rc;
        }
        ma.jobs = e->d_jobs + e->level_job_first[level];
        ma.n_jobs = e->level_job_first[level + 1] - e->level_job_first[level];
        CU(launch_mac(e->plan, ma, e->s_mac));
        e->launches += 3;
        if (e->split > 1) {
            CU(launch_split_reduce(e->plan, ma, e->s_mac));
            e->launches++;
        }
    }
    if (timing) CU(cudaEventRecord(ev[3], e->s_mac));
    CU(cudaEventRecord(e->ev_mac_done[par], e->s_mac));
    CU(cudaStreamWaitEvent(e->s_inv, e->ev_mac_done[par], 0));
    if (out_free != nullptr) {
        CU(cudaStreamWaitEvent(e->s_inv, out_free, 0));
    }
    if (timing) CU(cudaEventRecord(ev[4], e->s_inv));

    InverseArgs ia = make_inverse_args(e, nb, par, raw_out);
    if (e->any_out_mix) {
        OutMixArgs oa = make_out_mix_args(e, nb, ia);
        CU(launch_out_mix(e->plan, oa, e->s_inv));
        e->launches++;
    }
    CU(launch_inverse(e->plan, ia, e->s_inv));
    e->launches += e->n_ch[1] > 0;
    if (!e->h_sd[1].empty()) {
        // bfrun.c:1918-1925: delay_subsample_update on the output block, before mixing and quantisation
        CU(launch_subdelay(e->plan, make_subdelay_args(e, 1, e->d_out_time, nb), e->s_inv));
        e->launches++;
    }
    if (e->any_group || e->any_muted[1]) {
        VirtMixArgs va;
        va.out_time = e->d_out_time;
        va.groups = e->d_groups;
        va.members = e->d_members;
        va.muted = e->any_muted[1] ? e->d_muted[1] : nullptr;
        va.n_groups = (int)e->h_groups.size();
        va.n_out = e->n_ch[1];
        va.batch = nb;
        va.L = e->L;
        CU(launch_virt_mix(e->plan, va, e->s_inv));
        e->launches++;
    }
    const bool pack_all = plan_unpacks_first(e->plan);   // size-specialised / four-step path: real2raw is a kernel of its own
    if (!e->shared_out.empty() || pack_all) {
        // outputs fed from several ranks: sum the L valid time-domain samples over NVLink, then quantise
        // (after the inverse FFT, before real2raw)
        if (e->comm != nullptr && !e->shared_out.empty()) {
            const int dtype = e->rs == 4 ? 7 /* ncclFloat32 */ : 8 /* ncclFloat64 */;
            g_nccl.GroupStart();
            for (int b = 0; b < nb; b++) {
                for (int o : e->shared_out) {
                    char *row = (char *)e->d_out_time + rs_bytes(e, ((size_t)b * e->n_ch[1] + o) * e->L);
                    int r = g_nccl.AllReduce(row, row, (size_t)e->L, dtype, 0 /* ncclSum */, e->comm, e->s_inv);
                    if (r != 0) {
                        g_nccl.GroupEnd();
                        return fail(BFCUDA_ECOMM, "ncclAllReduce failed: %s",
                                    g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
                    }
                }
            }
            g_nccl.GroupEnd();
        }
        if (pack_all) {
            CU(launch_pack(e->plan, ia, e->s_inv));
        } else {
            CU(launch_quantise_shared(e->plan, ia, e->s_inv));
        }
        e->launches++;
    }
    if (e->dither.n_dither > 0) {
        CU(launch_dither(e->plan, ia, e->dither, e->s_inv));
        e->launches++;
    }
    if (timing) {
        CU(cudaEventRecord(ev[5], e->s_inv));
        e->ring_fill++;
    }
    CU(cudaEventRecord(e->ev_inv_done[par], e->s_inv));
    CU(cudaEventRecord(e->ev_inv, e->s_inv));
    if (ll && !e->dirty && !e->xfade_active && !transition && ma.n_jobs > 0) {
        // The next block's partitions 1 .. P-1 read ring slots that are all written by now: start their sum behind this
        // block's (short) MAC, into the other generation of Y, which the inverse stage of the previous launch must
        // have left.  A control change before the next block discards it (enqueue_blocks).
        MacArgs mt = ma;
        mt.Y = (char *)e->d_Y + (size_t)(par ^ 1) * e->y_stride;
        mt.t = (e->slot_t + 1) % e->fdl_ring;
        mt.z_first = 1;
        mt.z_count = 1;
        // ... and not before THIS block's inverse stage is through: the long-running blocks of the sum would hold the
        // SMs the (short, latency-critical) inverse and packing kernels are waiting for
        CU(cudaStreamWaitEvent(e->s_mac, e->ev_inv_done[par], 0));
        CU(launch_mac(e->plan, mt, e->s_mac));
        CU(cudaEventRecord(e->ev_tail_done, e->s_mac));
        e->launches++;
        e->tail_ready = true;
    }
    e->launch_no++;

    // bfrun.c:1838, 2034
    for (FilterState &fs : e->filters) {
        fs.prevcoeff = fs.coeff;
    }
    e->t += (unsigned int)nb;
    e->slot_t = (e->slot_t + nb) % e->fdl_ring;
    e->last_batch = nb;
    return 0;
}

// A run-time change of a filter's block delay (cfd, bfrun.c:1579-1600).  The reference's ring has exactly P slots,
// written at (w + delay) % P and read at (t - i) % P: after a change it reads slots that alias other blocks -- blocks
// written "ahead" under a larger old delay, blocks of a ring turn earlier where a larger new delay skips slots -- and
// changes that follow each other within P blocks stack these effects.  The engine's ring is longer (virtual slot u
// lives at u % ring, distinct physical slots for what the reference aliases), so for 2 P blocks after a change the
// filter runs in a TRANSITION: a mirror of P slots is kept that IS the reference's ring (every block the filter's
// newly written spectrum is copied to mirror[(t + delay) % P]), and before each multiply-accumulate every ring slot the
// block reads (virtual t - i, i < P - delay, i <= t: the reference stops at the number of blocks processed so far,
// procblocks, bfrun.c:1567-1571, 1745) is made equal to mirror[(t - i) % P] where the two differ.  Which block a slot
// holds is tracked on the host (ref_id / eng_id), so only aliasing slots are copied.  After P blocks of constant delay
// every residue has been rewritten and the mirror equals the plain ring again; the transition is kept for 2 P blocks
// so that everything a LATER change can still read (P slots back) is regular when its snapshot is taken.  Blocks in
// a transition are launched one at a time.
static bool in_transition(const bfcuda_engine *e)
{
    for (const FilterState &fs : e->filters) {
        if (fs.trans_until >= 0) {
            return true;
        }
    }
    return false;
}

static int begin_transitions(bfcuda_engine *e)
{
    const size_t nb = rs_bytes(e, e->N);
    const int R = e->fdl_ring, P = e->P;
    const long T = (long)e->t;
    for (FilterState &fs : e->filters) {
        if (fs.trans_until >= 0 && T >= fs.trans_until) {
            fs.trans_until = -1;        // regular again
        }
    }
    for (const std::pair<int, int> &fx : e->delay_changes) {
        const int f = fx.first, d_old = fx.second;
        FilterState &fs = e->filters[f];
        if (clamp_delay(e, fs.delayblocks) == d_old) {
            continue;       // changed and changed back between two blocks
        }
        if (fs.trans_until < 0) {
            // snapshot: the last P writes went to virtual slots w + d_old (regular), the newest is T - 1 + d_old
            if (fs.mirror == nullptr) {
                int rc = dev_alloc(e, &fs.mirror, nb * (size_t)P, false);
                if (rc != 0) return rc;
            }
            const char *ring = ring_ptr(e, fs.stream);     // private by now: update_streams() ran first
            auto phys = [&](long u) { return (int)((((long)e->slot_t + (u - T)) % R + R) % R); };
            const long umax = T - 1 + d_old;
            fs.eng_id.assign((size_t)R, -1);
            fs.ref_id.assign((size_t)P, -1);
            for (long u = umax; u > umax - R; u--) {
                fs.eng_id[(size_t)phys(u)] = u - d_old >= 0 ? u - d_old : -1;
            }
            for (long u = umax; u > umax - P; u--) {
                if (u - d_old < 0) {
                    continue;
                }
                const size_t s = (size_t)(((u % P) + P) % P);
                fs.ref_id[s] = u - d_old;
                CU(cudaMemcpyAsync(fs.mirror + nb * s, ring + nb * (size_t)phys(u), nb, cudaMemcpyDeviceToDevice, e->stream));
            }
        }
        fs.trans_until = T + 2L * P;
        // the repairs copy whole slots without their powersave flags: read everything until the last repaired slot is
        // out of every filter's reach
        e->ps_hold_until = std::max(e->ps_hold_until, T + 3L * P + 2L * e->max_batch);
    }
    e->delay_changes.clear();
    return 0;
}

// One block (nb = 1) of the filters of `level` that are in a transition, after their delay-line slot of this block has
// been written and before their partitions are multiplied; `stream` is the stream both of those run on.
static int transition_repair(bfcuda_engine *e, int level, cudaStream_t stream)
{
    const size_t nb = rs_bytes(e, e->N);
    const int R = e->fdl_ring, P = e->P;
    const long t = (long)e->t;
    for (size_t f = 0; f < e->filters.size(); f++) {
        FilterState &fs = e->filters[f];
        if (fs.trans_until < 0 || fs.level != level) {
            continue;
        }
        char *ring = ring_ptr(e, fs.stream);
        auto phys = [&](long u) { return (int)((((long)e->slot_t + (u - t)) % R + R) % R); };
        const int d = clamp_delay(e, fs.delayblocks);
        // this block's spectrum: virtual slot t + d, the reference's slot (t + d) % P
        const size_t sw = (size_t)((t + d) % P);
        CU(cudaMemcpyAsync(fs.mirror + nb * sw, ring + nb * (size_t)phys(t + d), nb, cudaMemcpyDeviceToDevice, stream));
        fs.ref_id[sw] = t;
        fs.eng_id[(size_t)phys(t + d)] = t;
        for (long i = 0; i < P - d && i <= t; i++) {
            const long u = t - i;
            const size_t s = (size_t)(u % P), q = (size_t)phys(u);
            if (fs.eng_id[q] == fs.ref_id[s]) {
                continue;
            }
            if (fs.ref_id[s] < 0) {
                CU(cudaMemsetAsync(ring + nb * q, 0, nb, stream));
            } else {
                CU(cudaMemcpyAsync(ring + nb * q, fs.mirror + nb * s, nb, cudaMemcpyDeviceToDevice, stream));
            }
            fs.eng_id[q] = fs.ref_id[s];
        }
    }
    return 0;
}

// Process n blocks that already sit in device memory: split them into launches of at most max_batch blocks;
// a pending control change or a crossfade block is processed on its own (its tables differ from its
// neighbours', bfrun.c:1462-1478, 1726-1777).
static int enqueue_blocks(bfcuda_engine *e, int n, uint8_t *raw_in, uint8_t *raw_out, cudaEvent_t in_ready,
                          cudaEvent_t out_free, cudaEvent_t fwd_done)
{
    int done = 0;
    while (done < n) {
        int nb = std::min(n - done, e->max_batch);
        if (e->merge_check_at >= 0 && (long)e->t >= e->merge_check_at) {
            e->dirty = true;
        }
        const bool transition = in_transition(e);
        if (e->dirty || e->xfade_active || transition) {
            // the previous launches' MAC and inverse stages (other streams) still read the job / output-mix tables,
            // and a delay transition rewrites ring slots the previous MAC reads
            if (e->launch_no >= 1) {
                const int prev = (int)((e->launch_no - 1) & 1u);
                CU(cudaStreamWaitEvent(e->stream, e->ev_mac_done[prev], 0));
                CU(cudaStreamWaitEvent(e->stream, e->ev_inv_done[prev], 0));
            }
            if (e->low_latency) {
                // an ahead-of-time launch may still read the job table (no-op while the event was never recorded)
                CU(cudaStreamWaitEvent(e->stream, e->ev_tail_done, 0));
                e->tail_ready = false;
            }
            int frc = update_streams(e);
            if (frc != 0) return frc;
            frc = begin_transitions(e);
            if (frc != 0) return frc;
            build_tables(e);
            int rc = upload_tables(e);
            if (rc != 0) return rc;
            if (e->xfade_active || in_transition(e)) {
                nb = 1;
            }
        }
        if (nb != 1) {
            e->tail_ready = false;      // Y is laid out per batch size
        }
        const bool lastpart = done + nb == n;
        int rc = enqueue_batch(e, nb, raw_in + (size_t)done * e->n_bytes[0], raw_out + (size_t)done * e->n_bytes[1],
                               done == 0 ? in_ready : nullptr, done == 0 ? out_free : nullptr,
                               lastpart ? fwd_done : nullptr);
        if (rc != 0) return rc;
        done += nb;
    }
    return 0;
}

static int check_status_word(unsigned int st)
{
    if (st & 1u) {
        return fail(BFCUDA_ENONFINITE, "NaN or Inf values in the output! Bad output.");
    }
    if (st & 2u) {
        return fail(BFCUDA_ESAFETY, "Safety limit exceeded on output.");
    }
    return 0;
}

static int check_status(bfcuda_engine *e)       // the device-resident / download path's word
{
    return check_status_word(*e->h_status);
}

static unsigned int snap_status(const bfcuda_engine *e, unsigned int call)
{
    return *reinterpret_cast<const unsigned int *>(e->h_snap[call & 3u] + e->snap_status_off);
}

// ======================================================================================================
// Step graphs
//
// A step of a small shard or a short partition is a handful of kernels of a few microseconds each: what bounds it is
// the host's enqueue cost (five launches plus a dozen event operations on three streams) and, inside the GPU, the gaps
// between dependent launches.  For calls that need no synchronous answer (the asynchronous and device-resident entry
// points) the engine therefore runs a SKEWED pipeline as ONE graph launch per call: tick k runs forward(k), MAC(k-1)
// and inverse(k-2) as three independent branches of a captured CUDA graph -- the same overlap the three stage streams
// give, without the events, and one cudaGraphLaunch instead of ~15 API calls.  The ring slot is the only per-launch
// quantity: it travels through cudaGraphExecKernelNodeSetParams on the two kernels that use it.  Output k arrives with
// tick k+2; everything that needs a complete state (synchronize, wait_previous, a control change, the synchronous
// call) drains the pipeline with the partial-mask graphs.  Anything the plain case does not cover -- pending control
// changes, crossfades, delay transitions, chained filters, the cross-rank sum, dither, stage timing -- takes the
// stream path (enqueue_blocks) after a drain.
// ======================================================================================================

// Where the step graphs pay: the stream path costs the host ~30 us per step (five launches and a dozen event
// operations) and leaves ~1 us bubbles between dependent kernels, which only matters when the step's GPU time is in
// that range.  Measured (profiles/r2_graph_ab.txt): 2 x 64 K taps 27 -> 11 us per block, 32 x 256 K taps 34 -> 28 us,
// but an 8-filter shard of the headline job (45 us of MAC per step) is 7 % faster on the streams, whose stages flow
// from step to step without the join at the end of every graph.  Automatic rule: graphs when the step's MAC traffic
// is below what HBM moves in ~21 us.  BFCUDA_GRAPH=0 / 1 forces it.
static bool graph_preferred(const bfcuda_engine *e)
{
    static const char *env = getenv("BFCUDA_GRAPH");
    if (env != nullptr) {
        return atoi(env) != 0;
    }
    // block by block the graphs win at every size measured since the MAC stream has the higher priority (the full job
    // 163.6 -> 159.3 us per block, shards of 2 / 4 / 8 ranks 87.8 -> 86.2, 55.3 -> 53.4, 34.9 -> 32.9; host-buffer figures
    // within +-1-3 %; profiles/r2_graph_b1.txt); batched calls keep the size rule
    if (e->max_batch == 1) {
        return true;
    }
    return e->mac_bytes_batch <= (size_t)136500000;
}

static bool graph_eligible(const bfcuda_engine *e, int n_blocks)
{
    return e->graph_enabled && graph_preferred(e) && e->plan.tw2 != nullptr && e->n_levels == 1 && n_blocks == e->max_batch && !e->dirty &&
           !e->xfade_active && !in_transition(e) && !e->low_latency && e->comm == nullptr && e->shared_out.empty() &&
           e->dither.n_dither == 0 && !(e->flags & (BFCUDA_FLAG_STAGE_TIMING | BFCUDA_FLAG_SERIAL_STAGES)) &&
           e->single_dest && e->launch_no >= 2 && e->mac_variant == 0 && e->n_ch[0] > 0 && e->n_ch[1] > 0 &&
           e->level_job_first[1] > 0 && !(e->merge_check_at >= 0 && (long)e->t >= e->merge_check_at) &&
           e->h_sd[0].empty() && e->h_sd[1].empty() && !e->any_group && !e->any_muted[0] && !e->any_muted[1];
}

// every stream waits for everything enqueued so far on every other one (device side only): taken when the engine
// switches between the stream path and the graph path, whose bookkeeping of buffer reuse differs
static int join_streams(bfcuda_engine *e)
{
    cudaStream_t all[5] = { e->stream, e->s_mac, e->s_inv, e->s_in, e->s_out };
    for (int i = 0; i < 5; i++) {
        CU(cudaEventRecord(e->ev_join, all[i]));
        for (int j = 0; j < 5; j++) {
            if (j != i) CU(cudaStreamWaitEvent(all[j], e->ev_join, 0));
        }
    }
    return 0;
}

static uint8_t *raw_queue(const bfcuda_engine *e, int io, int q)
{
    return q == 0 ? e->d_raw[io] : (q == 1 ? e->d_raw2[io] : e->d_raw34[io][q - 2]);
}

// host_in / host_out / host_snap: mode 2 only (copies inside the graph), else null
static int capture_step_graph(bfcuda_engine *e, StepGraph &g, int mask, const UnpackArgs &ua, const ForwardArgs &fa,
                              const MacArgs &ma, const InverseArgs &ia, const void *host_in, void *host_out, void *host_snap)
{
    const void *f_fwd = nullptr, *f_mac = nullptr;
    const int nb = e->max_batch;
    CU(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeRelaxed));
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (err == cudaSuccess) err = r; return err == cudaSuccess; };
    ok(cudaEventRecord(e->ev_cap[0], e->stream));
    if (mask & 4) {
        if (host_in != nullptr) {
            ok(cudaMemcpyAsync(const_cast<uint8_t *>(ua.raw_in), host_in, (size_t)nb * e->n_bytes[0], cudaMemcpyHostToDevice, e->stream));
        }
        if (ua.amax != nullptr) {
            ok(cudaMemsetAsync(ua.amax, 0, sizeof(unsigned int) * (size_t)nb * e->n_ch[0], e->stream));
        }
        if (ok(launch_unpack(e->plan, ua, e->stream)) && ok(launch_forward(e->plan, fa, e->stream))) {
            f_fwd = g_last_func;
        }
    }
    if (mask & 2) {
        ok(cudaStreamWaitEvent(e->s_mac, e->ev_cap[0], 0));
        if (ok(launch_mac(e->plan, ma, e->s_mac))) {
            f_mac = g_last_func;
        }
        if (e->split > 1) ok(launch_split_reduce(e->plan, ma, e->s_mac));
        ok(cudaEventRecord(e->ev_cap[1], e->s_mac));
        ok(cudaStreamWaitEvent(e->stream, e->ev_cap[1], 0));
    }
    if (mask & 1) {
        ok(cudaStreamWaitEvent(e->s_inv, e->ev_cap[0], 0));
        if (e->any_out_mix) ok(launch_out_mix(e->plan, make_out_mix_args(e, nb, ia), e->s_inv));
        ok(launch_inverse(e->plan, ia, e->s_inv));
        ok(launch_pack(e->plan, ia, e->s_inv));
        if (host_out != nullptr) {
            ok(cudaMemcpyAsync(host_out, ia.raw_out, (size_t)nb * e->n_bytes[1], cudaMemcpyDeviceToHost, e->s_inv));
            ok(cudaMemcpyAsync(host_snap, e->d_overflow, e->snap_bytes, cudaMemcpyDeviceToHost, e->s_inv));
        }
        ok(cudaEventRecord(e->ev_cap[2], e->s_inv));
        ok(cudaStreamWaitEvent(e->stream, e->ev_cap[2], 0));
    }
    cudaGraph_t graph = nullptr;
    cudaError_t end = cudaStreamEndCapture(e->stream, &graph);
    if (err != cudaSuccess || end != cudaSuccess || graph == nullptr) {
        if (graph != nullptr) cudaGraphDestroy(graph);
        cudaGetLastError();
        return fail(BFCUDA_ECUDA, "step graph capture failed: %s", cudaGetErrorString(err != cudaSuccess ? err : end));
    }
    memset(&g, 0, sizeof(g));
    g.graph = graph;
    size_t n = 0;
    CU(cudaGraphGetNodes(graph, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    CU(cudaGraphGetNodes(graph, nodes.data(), &n));
    for (size_t i = 0; i < n; i++) {
        cudaGraphNodeType ty;
        CU(cudaGraphNodeGetType(nodes[i], &ty));
        if (ty == cudaGraphNodeTypeMemcpy) {
            cudaMemcpy3DParms mp;
            CU(cudaGraphMemcpyNodeGetParams(nodes[i], &mp));
            if (host_in != nullptr && mp.dstPtr.ptr == (void *)ua.raw_in) {
                g.n_h2d = nodes[i];
            } else if (host_out != nullptr && mp.srcPtr.ptr == (void *)ia.raw_out) {
                g.n_d2h = nodes[i];
            } else if (host_out != nullptr && mp.srcPtr.ptr == (void *)e->d_overflow) {
                g.n_snap = nodes[i];
            }
            continue;
        }
        if (ty != cudaGraphNodeTypeKernel) {
            continue;
        }
        g.n_kernels++;
        cudaKernelNodeParams kp;
        CU(cudaGraphKernelNodeGetParams(nodes[i], &kp));
        if ((mask & 4) && kp.func == f_fwd && g.n_fwd == nullptr) {
            g.n_fwd = nodes[i];
            g.p_fwd = kp;
        } else if ((mask & 2) && kp.func == f_mac && g.n_mac == nullptr) {
            g.n_mac = nodes[i];
            g.p_mac = kp;
        }
    }
    if (((mask & 4) && g.n_fwd == nullptr) || ((mask & 2) && g.n_mac == nullptr) ||
        ((mask & 4) && host_in != nullptr && g.n_h2d == nullptr) ||
        ((mask & 1) && host_out != nullptr && (g.n_d2h == nullptr || g.n_snap == nullptr))) {
        cudaGraphDestroy(graph);
        memset(&g, 0, sizeof(g));
        return fail(BFCUDA_ECUDA, "step graph: could not identify the forward / MAC kernel nodes");
    }
    cudaError_t ierr = cudaGraphInstantiate(&g.exec, graph, 0);
    if (ierr != cudaSuccess) {
        cudaGraphDestroy(graph);
        memset(&g, 0, sizeof(g));
        return fail(BFCUDA_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ierr));
    }
    return 0;
}

// One tick of the skewed pipeline.  have_f: a new step enters (its raw input is in the device input buffer of this
// tick's parity); the pending MAC / inverse stages of the two steps before it ride along.  Without have_f the tick only
// drains.  host_out / call: destination and number of the host-buffer call that brought the new step.
static int graph_tick(bfcuda_engine *e, bool have_f, void *host_out, unsigned int call, const void *host_in = nullptr)
{
    const int nb = e->max_batch;
    const int mask = (have_f ? 4 : 0) | (e->pend_mac.valid ? 2 : 0) | (e->pend_inv.valid ? 1 : 0);
    if (mask == 0) {
        return 0;
    }
    const unsigned int k = have_f ? e->launch_no : (e->pend_mac.valid ? e->pend_mac.step + 1u : e->pend_inv.step + 2u);
    const bool host = e->graph_host == 1;       // copies on the copy streams, around the graph
    const bool inside = e->graph_host == 2;     // copies are nodes of the graph: one raw buffer per direction is enough
    // the double-buffered operands (Y, unpacked samples) follow k & 1; a host-buffer pipeline with outside copies keeps
    // FOUR raw blocks per direction in flight (the copies of a 16 MiB step take longer than its kernels): k & 3
    const int par = (int)(k & (host ? 3u : 1u));
    uint8_t *d_in = host ? raw_queue(e, 0, par) : e->d_raw[0];
    uint8_t *d_out = host ? raw_queue(e, 1, par) : e->d_raw[1];
    UnpackArgs ua;
    ForwardArgs fa;
    MacArgs ma;
    InverseArgs ia;
    memset(&ua, 0, sizeof(ua));
    memset(&fa, 0, sizeof(fa));
    memset(&ma, 0, sizeof(ma));
    memset(&ia, 0, sizeof(ia));
    if (have_f) {
        e->xt_par ^= 1;             // == (k + 1) & 1: the stream path toggles it once per launch too
        ua = make_unpack_args(e, nb, d_in, e->xt_par);
        fa = make_forward_args(e, nb, e->slot_t, d_in);
        set_xt_generation(e, fa, e->xt_par, e->xt_last_nb);
        e->xt_last_nb = nb;
    }
    if (mask & 2) {
        ma = make_mac_args(e, nb, e->pend_mac.slot_t, (int)(e->pend_mac.step & 1u));
    }
    if (mask & 1) {
        ia = make_inverse_args(e, nb, (int)(e->pend_inv.step & 1u), d_out);
    }
    StepGraph &g = e->sg[e->graph_host][mask][par];
    const bool copy_out = inside && (mask & 1) && e->pend_inv.host_out != nullptr;
    void *snap_dst = copy_out ? (void *)e->h_snap[e->pend_inv.call & 3u] : nullptr;
    if (g.exec == nullptr) {
        int rc = capture_step_graph(e, g, mask, ua, fa, ma, ia, inside && have_f ? host_in : nullptr,
                                    copy_out ? e->pend_inv.host_out : nullptr, snap_dst);
        if (rc != 0) return rc;
    } else {
        if (inside && have_f) {
            CU(cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.n_h2d, d_in, host_in, (size_t)nb * e->n_bytes[0], cudaMemcpyHostToDevice));
        }
        if (copy_out) {
            CU(cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.n_d2h, e->pend_inv.host_out, d_out, (size_t)nb * e->n_bytes[1],
                                                  cudaMemcpyDeviceToHost));
            CU(cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.n_snap, snap_dst, e->d_overflow, e->snap_bytes, cudaMemcpyDeviceToHost));
        }
        if (mask & 4) {
            cudaKernelNodeParams kp = g.p_fwd;
            void *args[2] = { &fa, g.p_fwd.kernelParams[1] };
            kp.kernelParams = args;
            CU(cudaGraphExecKernelNodeSetParams(g.exec, g.n_fwd, &kp));
        }
        if (mask & 2) {
            cudaKernelNodeParams kp = g.p_mac;
            void *args[2] = { &ma, g.p_mac.kernelParams[1] };
            kp.kernelParams = args;
            CU(cudaGraphExecKernelNodeSetParams(g.exec, g.n_mac, &kp));
        }
    }
    const bool out_to_host = (mask & 1) && host && e->pend_inv.host_out != nullptr;
    if (out_to_host) {
        // the inverse stage overwrites raw output buffer `par`: its last read-out (two ticks ago) must be through
        CU(cudaStreamWaitEvent(e->stream, e->ev_out_read[par], 0));
    }
    CU(cudaGraphLaunch(g.exec, e->stream));
    e->launches += g.n_kernels;
    e->graph_used = true;
    if (out_to_host) {
        const unsigned int c = e->pend_inv.call;
        CU(cudaMemcpyAsync(e->d_snap[c & 3u], e->d_overflow, e->snap_bytes, cudaMemcpyDeviceToDevice, e->stream));
        CU(cudaEventRecord(e->ev_g_done[par], e->stream));
        CU(cudaStreamWaitEvent(e->s_out, e->ev_g_done[par], 0));
        CU(cudaMemcpyAsync(e->pend_inv.host_out, d_out, (size_t)nb * e->n_bytes[1], cudaMemcpyDeviceToHost, e->s_out));
        CU(cudaMemcpyAsync(e->h_snap[c & 3u], e->d_snap[c & 3u], e->snap_bytes, cudaMemcpyDeviceToHost, e->s_out));
        CU(cudaEventRecord(e->ev_call_done[c & 3u], e->s_out));
        CU(cudaEventRecord(e->ev_out_read[par], e->s_out));
        e->h_overflow_valid = e->n_ch[1] > 0;
    } else if (copy_out) {
        // the copies ran inside the graph: the call's output is complete when the graph is
        CU(cudaEventRecord(e->ev_call_done[e->pend_inv.call & 3u], e->stream));
        e->h_overflow_valid = e->n_ch[1] > 0;
    } else if (!inside) {
        CU(cudaEventRecord(e->ev_g_done[par], e->stream));
    }
    if (!inside) {
        CU(cudaEventRecord(e->ev_inv, e->stream));
    }
    // advance the pipeline
    e->pend_inv = e->pend_mac;
    e->pend_mac.valid = false;
    if (have_f) {
        e->pend_mac.valid = true;
        e->pend_mac.step = e->launch_no;
        e->pend_mac.slot_t = e->slot_t;
        e->pend_mac.host_out = host_out;
        e->pend_mac.call = call;
        e->launch_no++;
        for (FilterState &fs : e->filters) {
            fs.prevcoeff = fs.coeff;        // bfrun.c:1838, 2034
        }
        e->t += (unsigned int)nb;
        e->slot_t = (e->slot_t + nb) % e->fdl_ring;
        e->last_batch = nb;
    }
    return 0;
}

// complete every stage that is still pending in the skewed pipeline
static int graph_drain(bfcuda_engine *e)
{
    while (e->pend_mac.valid || e->pend_inv.valid) {
        int rc = graph_tick(e, false, nullptr, 0);
        if (rc != 0) return rc;
    }
    return 0;
}

static int leave_graph_mode(bfcuda_engine *e)
{
    if (!e->graph_mode) {
        return 0;
    }
    int rc = graph_drain(e);
    if (rc != 0) return rc;
    e->graph_mode = false;
    return join_streams(e);
}

static int enter_graph_mode(bfcuda_engine *e, int host)
{
    if (e->graph_mode && e->graph_host == host) {
        return 0;
    }
    int rc = leave_graph_mode(e);       // device-resident <-> host-buffer: different raw buffers
    if (rc != 0) return rc;
    rc = join_streams(e);
    if (rc != 0) return rc;
    e->graph_mode = true;
    e->graph_host = host;
    return 0;
}

static int sync_all(bfcuda_engine *e)
{
    int rc = graph_drain(e);
    if (rc != 0) return rc;
    CU(cudaStreamSynchronize(e->s_in));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaStreamSynchronize(e->s_mac));
    CU(cudaStreamSynchronize(e->s_inv));
    CU(cudaStreamSynchronize(e->s_out));
    e->io_waited = e->io_count;
    return 0;
}

// the output of host-buffer call `call` is complete in host memory
static int wait_call(bfcuda_engine *e, unsigned int call)
{
    if ((e->pend_mac.valid && e->pend_mac.host_out != nullptr && e->pend_mac.call <= call) ||
        (e->pend_inv.valid && e->pend_inv.host_out != nullptr && e->pend_inv.call <= call)) {
        int rc = graph_drain(e);        // its later stages are still waiting for the next ticks
        if (rc != 0) return rc;
    }
    CU(cudaEventSynchronize(e->ev_call_done[call & 3u]));
    return 0;
}

// Copies as nodes of the step graph (mode 2): one launch per call and no event traffic around it.  Measured SLOWER than
// the copies on their own streams (profiles/r2_graph_copies.txt: 2 x 64 K taps block by block x3600 against x4850
// through host buffers, 8 blocks per call x22 700 against x32 200; only 32 x 256 K at 8 blocks gains 5 %): inside the
// graph a copy's latency sits in its branch's dependency chain every tick, outside it overlaps the neighbouring ticks.
// Kept opt-in (BFCUDA_GRAPH_COPIES=1); needs page-locked host memory and blocks up to 1 MiB per direction and call.
static bool copies_fit_in_graph(bfcuda_engine *e, int n_blocks, const void *raw_in, void *raw_out)
{
    static const char *env = getenv("BFCUDA_GRAPH_COPIES");
    if (env == nullptr || atoi(env) == 0) {
        return false;
    }
    if ((size_t)n_blocks * e->n_bytes[0] > ((size_t)1 << 20) || (size_t)n_blocks * e->n_bytes[1] > ((size_t)1 << 20)) {
        return false;
    }
    for (const void *p : { raw_in, (const void *)raw_out }) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        if (at.type != cudaMemoryTypeHost) {
            return false;
        }
    }
    return true;
}

static int process_blocks_host(bfcuda_engine *e, int n_blocks, const void *raw_in, void *raw_out, bool allow_graph)
{
    if (e == nullptr || raw_in == nullptr || raw_out == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    if (n_blocks < 1 || n_blocks > e->max_batch) {
        return fail(BFCUDA_EINVAL, "n_blocks %d outside 1..max_batch (%d)", n_blocks, e->max_batch);
    }
    CU(cudaSetDevice(e->device));
    const unsigned int call = e->io_count;
    if (allow_graph && graph_eligible(e, n_blocks) && copies_fit_in_graph(e, n_blocks, raw_in, raw_out)) {
        // small pinned blocks: the copies become nodes of the step graph (one launch per call and nothing around it)
        int rc = enter_graph_mode(e, 2);
        if (rc != 0) return rc;
        rc = graph_tick(e, true, raw_out, call, raw_in);
        if (rc != 0) return rc;
        e->io_count++;
        return 0;
    }
    if (allow_graph && graph_eligible(e, n_blocks)) {
        int rc = enter_graph_mode(e, 1);
        if (rc != 0) return rc;
        const int q = (int)(e->launch_no & 3u);
        uint8_t *d_in = raw_queue(e, 0, q);
        CU(cudaStreamWaitEvent(e->s_in, e->ev_g_done[q], 0));       // the forward stage four ticks ago has consumed d_in
        CU(cudaMemcpyAsync(d_in, raw_in, (size_t)n_blocks * e->n_bytes[0], cudaMemcpyHostToDevice, e->s_in));
        CU(cudaEventRecord(e->ev_h2dq[q], e->s_in));
        CU(cudaStreamWaitEvent(e->stream, e->ev_h2dq[q], 0));
        rc = graph_tick(e, true, raw_out, call);
        if (rc != 0) return rc;
        e->io_count++;
        return 0;
    }
    int rc = leave_graph_mode(e);
    if (rc != 0) return rc;
    // double-buffered raw blocks: copy-in of call k+1 and copy-out of call k-1 overlap the kernels of call k
    const int b = (int)(call & 1u);
    uint8_t *d_in = b ? e->d_raw2[0] : e->d_raw[0];
    uint8_t *d_out = b ? e->d_raw2[1] : e->d_raw[1];
    const bool reuse = call >= 2;
    if (reuse) {
        CU(cudaStreamWaitEvent(e->s_in, e->ev_fwd[b], 0));      // forward of call k-2 has consumed d_in
    }
    CU(cudaMemcpyAsync(d_in, raw_in, (size_t)n_blocks * e->n_bytes[0], cudaMemcpyHostToDevice, e->s_in));
    CU(cudaEventRecord(e->ev_h2d[b], e->s_in));
    rc = enqueue_blocks(e, n_blocks, d_in, d_out, e->ev_h2d[b], reuse ? e->ev_call_done[(call - 2u) & 3u] : nullptr,
                        e->ev_fwd[b]);
    if (rc != 0) return rc;
    // the peak-meter records and the status word travel with the block (bfrun.c:1929-1936 reads them after every
    // block): snapshot them behind this call's last kernel, read the snapshot out with the output
    CU(cudaMemcpyAsync(e->d_snap[call & 3u], e->d_overflow, e->snap_bytes, cudaMemcpyDeviceToDevice, e->s_inv));
    CU(cudaEventRecord(e->ev_inv, e->s_inv));
    CU(cudaStreamWaitEvent(e->s_out, e->ev_inv, 0));
    CU(cudaMemcpyAsync(raw_out, d_out, (size_t)n_blocks * e->n_bytes[1], cudaMemcpyDeviceToHost, e->s_out));
    CU(cudaMemcpyAsync(e->h_snap[call & 3u], e->d_snap[call & 3u], e->snap_bytes, cudaMemcpyDeviceToHost, e->s_out));
    e->h_overflow_valid = e->n_ch[1] > 0;
    CU(cudaEventRecord(e->ev_call_done[call & 3u], e->s_out));
    e->io_count++;
    return 0;
}

int bfcuda_process_blocks_async(bfcuda_engine *e, int n_blocks, const void *raw_in, void *raw_out)
{
    return process_blocks_host(e, n_blocks, raw_in, raw_out, true);
}

int bfcuda_process_block_async(bfcuda_engine *e, const void *raw_in, void *raw_out)
{
    return bfcuda_process_blocks_async(e, 1, raw_in, raw_out);
}

int bfcuda_wait_previous(bfcuda_engine *e, int calls_back)
{
    if (e == nullptr) return fail(BFCUDA_EINVAL, "null engine");
    if (calls_back < 0 || calls_back > 3 || (unsigned int)calls_back >= e->io_count) {
        return fail(BFCUDA_EINVAL, "only the four most recent asynchronous calls can be waited for");
    }
    CU(cudaSetDevice(e->device));
    // per-call events and snapshots are kept for the four most recent calls
    const unsigned int call = e->io_count - 1u - (unsigned int)calls_back;
    int rc = wait_call(e, call);
    if (rc != 0) return rc;
    if (calls_back == 0) {
        e->io_waited = e->io_count;
    }
    return check_status_word(snap_status(e, call));
}

int bfcuda_synchronize(bfcuda_engine *e)
{
    if (e == nullptr) return fail(BFCUDA_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    int rc = sync_all(e);
    if (rc != 0) return rc;
    if (e->io_count > 0) {
        rc = check_status_word(snap_status(e, e->io_count - 1u));
        if (rc != 0) return rc;
    }
    return check_status(e);
}

int bfcuda_process_block(bfcuda_engine *e, const void *raw_in, void *raw_out)
{
    return bfcuda_process_blocks(e, 1, raw_in, raw_out);
}

int bfcuda_process_blocks(bfcuda_engine *e, int n_blocks, const void *raw_in, void *raw_out)
{
    // the answer is wanted now: the stream path (a skewed pipeline would have to be drained right away)
    int rc = process_blocks_host(e, n_blocks, raw_in, raw_out, false);
    if (rc != 0) return rc;
    // the call's own read-out, not every stream: in the low-latency schedule the next block's partial sum keeps running
    return bfcuda_wait_previous(e, 0);
}

int bfcuda_process_blocks_device(bfcuda_engine *e, int n_blocks)
{
    if (e == nullptr) return fail(BFCUDA_EINVAL, "null engine");
    if (n_blocks < 1 || n_blocks > e->max_batch) {
        return fail(BFCUDA_EINVAL, "n_blocks %d outside 1..max_batch (%d)", n_blocks, e->max_batch);
    }
    CU(cudaSetDevice(e->device));
    e->h_overflow_valid = false;        // no read-out follows a device-resident call
    if (graph_eligible(e, n_blocks)) {
        int rc = enter_graph_mode(e, 0);
        if (rc != 0) return rc;
        return graph_tick(e, true, nullptr, 0);
    }
    int rc = leave_graph_mode(e);
    if (rc != 0) return rc;
    return enqueue_blocks(e, n_blocks, e->d_raw[0], e->d_raw[1], nullptr, nullptr, nullptr);
}

int bfcuda_process_block_device(bfcuda_engine *e)
{
    return bfcuda_process_blocks_device(e, 1);
}

int bfcuda_device_io(bfcuda_engine *e, int io, void **device_ptr, size_t *n_bytes)
{
    if (e == nullptr || (io != 0 && io != 1)) return fail(BFCUDA_EINVAL, "bad argument");
    if (device_ptr) *device_ptr = e->d_raw[io];
    if (n_bytes) *n_bytes = (size_t)e->n_bytes[io] * e->max_batch;
    return 0;
}

int bfcuda_upload_inputs(bfcuda_engine *e, int n_blocks, const void *raw_in)
{
    if (e == nullptr || raw_in == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    if (n_blocks < 1 || n_blocks > e->max_batch) return fail(BFCUDA_EINVAL, "n_blocks outside 1..max_batch");
    CU(cudaSetDevice(e->device));
    int rc = sync_all(e);
    if (rc != 0) return rc;
    CU(cudaMemcpyAsync(e->d_raw[0], raw_in, (size_t)n_blocks * e->n_bytes[0], cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

int bfcuda_upload_input(bfcuda_engine *e, const void *raw_in)
{
    return bfcuda_upload_inputs(e, 1, raw_in);
}

int bfcuda_download_output(bfcuda_engine *e, void *raw_out)
{
    return bfcuda_download_outputs(e, 1, raw_out);
}

int bfcuda_download_outputs(bfcuda_engine *e, int n_blocks, void *raw_out)
{
    if (e == nullptr || raw_out == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    if (n_blocks < 1 || n_blocks > e->max_batch) return fail(BFCUDA_EINVAL, "n_blocks outside 1..max_batch");
    CU(cudaSetDevice(e->device));
    int rc = sync_all(e);
    if (rc != 0) return rc;
    CU(cudaMemcpyAsync(raw_out, e->d_raw[1], (size_t)n_blocks * e->n_bytes[1], cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(e->h_status, e->d_status, sizeof(unsigned int), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return check_status(e);
}

void *bfcuda_host_alloc(size_t n_bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, n_bytes ? n_bytes : 16) != cudaSuccess) {
        cudaGetLastError();
        fail(BFCUDA_ENOMEM, "cudaMallocHost(%zu) failed", n_bytes);
        return nullptr;
    }
    memset(p, 0, n_bytes);
    return p;
}

// Page-locked host memory on the NUMA node the device hangs off (sysfs: /sys/bus/pci/devices/<id>/numa_node): with
// one process per GPU on a two-socket host, buffers that land on the other socket cross the inter-socket link on every
// copy.  The node is applied as a PREFERRED memory policy around the allocation (set_mempolicy by syscall number: no
// libnuma in the image); anything that fails leaves the default placement.
void *bfcuda_host_alloc_near(int device, size_t n_bytes)
{
    int node = -1;
    char busid[32] = "";
    if (cudaDeviceGetPCIBusId(busid, (int)sizeof(busid), device) == cudaSuccess) {
        for (char *c = busid; *c; c++) {
            if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
        }
        char path[96];
        snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", busid);
        FILE *f = fopen(path, "r");
        if (f != nullptr) {
            if (fscanf(f, "%d", &node) != 1) node = -1;
            fclose(f);
        }
    } else {
        cudaGetLastError();
    }
    bool bound = false;
#ifdef SYS_set_mempolicy
    if (node >= 0 && node < 64) {
        unsigned long mask = 1ul << node;
        bound = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, 65ul) == 0;
    }
#endif
    void *p = bfcuda_host_alloc(n_bytes);
#ifdef SYS_set_mempolicy
    if (bound) {
        syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
    }
#endif
    return p;
}

// Copy-only baseline of the host-buffer path: `reps` rounds of the H2D copy of n_blocks input blocks and the D2H copy of
// n_blocks output blocks, on the engine's own copy streams, both directions at once as in the pipelined calls, no
// kernels.  What the bus / the host side can carry for this engine's block sizes -- with every rank calling it at the
// same time, what the box can carry.  Reports the wall time of the whole exchange (CUDA events) and the two rates.
int bfcuda_copy_baseline(bfcuda_engine *e, int n_blocks, const void *raw_in, void *raw_out, int reps, double *ms_per_rep,
                         double *h2d_gbs, double *d2h_gbs)
{
    if (e == nullptr || raw_in == nullptr || raw_out == nullptr || reps < 1) return fail(BFCUDA_EINVAL, "bad argument");
    if (n_blocks < 1 || n_blocks > e->max_batch) return fail(BFCUDA_EINVAL, "n_blocks outside 1..max_batch");
    CU(cudaSetDevice(e->device));
    int rc = sync_all(e);
    if (rc != 0) return rc;
    const size_t bi = (size_t)n_blocks * e->n_bytes[0], bo = (size_t)n_blocks * e->n_bytes[1];
    cudaEvent_t ev[4];
    for (cudaEvent_t &x : ev) CU(cudaEventCreate(&x));
    CU(cudaEventRecord(ev[0], e->s_in));
    CU(cudaEventRecord(ev[2], e->s_out));
    for (int r = 0; r < reps; r++) {
        CU(cudaMemcpyAsync(r & 1 ? e->d_raw2[0] : e->d_raw[0], raw_in, bi, cudaMemcpyHostToDevice, e->s_in));
        CU(cudaMemcpyAsync(raw_out, r & 1 ? e->d_raw2[1] : e->d_raw[1], bo, cudaMemcpyDeviceToHost, e->s_out));
    }
    CU(cudaEventRecord(ev[1], e->s_in));
    CU(cudaEventRecord(ev[3], e->s_out));
    CU(cudaEventSynchronize(ev[1]));
    CU(cudaEventSynchronize(ev[3]));
    float m_in = 0.f, m_out = 0.f;
    CU(cudaEventElapsedTime(&m_in, ev[0], ev[1]));
    CU(cudaEventElapsedTime(&m_out, ev[2], ev[3]));
    for (cudaEvent_t &x : ev) cudaEventDestroy(x);
    if (ms_per_rep) *ms_per_rep = (double)std::max(m_in, m_out) / reps;
    if (h2d_gbs) *h2d_gbs = m_in > 0.f ? (double)bi * reps / (m_in * 1e-3) / 1e9 : 0.0;
    if (d2h_gbs) *d2h_gbs = m_out > 0.f ? (double)bo * reps / (m_out * 1e-3) / 1e9 : 0.0;
    return 0;
}

void bfcuda_host_free(void *p)
{
    if (p != nullptr) {
        cudaFreeHost(p);
    }
}

// ---- measurement ---------------------------------------------------------------------------------------

int bfcuda_timer_start(bfcuda_engine *e)
{
    if (e == nullptr) return fail(BFCUDA_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    int rc = sync_all(e);
    if (rc != 0) return rc;
    CU(cudaEventRecord(e->timer[0], e->stream));
    return 0;
}

int bfcuda_timer_stop(bfcuda_engine *e, double *elapsed_ms)
{
    if (e == nullptr || elapsed_ms == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    CU(cudaSetDevice(e->device));
    // the stop event must come after everything enqueued on any of the engine's streams -- including the stages of
    // the last two steps that the skewed step graphs have not run yet
    {
        int rc = graph_drain(e);
        if (rc != 0) return rc;
    }
    for (cudaStream_t st : { e->s_mac, e->s_inv, e->s_out }) {
        CU(cudaEventRecord(e->ev_join, st));
        CU(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
    }
    CU(cudaEventRecord(e->timer[1], e->stream));
    CU(cudaEventSynchronize(e->timer[1]));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e->timer[0], e->timer[1]));
    *elapsed_ms = (double)ms;
    return 0;
}

int bfcuda_set_stage_timing(bfcuda_engine *e, int on)
{
    if (e == nullptr) return fail(BFCUDA_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    int rc = sync_all(e);
    if (rc != 0) return rc;
    rc = flush_timing_ring(e);
    if (rc != 0) return rc;
    if (on) {
        for (int i = 0; i < TIMING_RING; i++) {
            for (int j = 0; j < 6; j++) {
                if (e->ring[i][j] == nullptr) {
                    CU(cudaEventCreate(&e->ring[i][j]));
                }
            }
        }
        e->flags |= BFCUDA_FLAG_STAGE_TIMING;
    } else {
        e->flags &= ~BFCUDA_FLAG_STAGE_TIMING;
    }
    return 0;
}

int bfcuda_set_serial_stages(bfcuda_engine *e, int on)
{
    if (e == nullptr) return fail(BFCUDA_EINVAL, "null engine");
    if (on) {
        e->flags |= BFCUDA_FLAG_SERIAL_STAGES;
    } else {
        e->flags &= ~BFCUDA_FLAG_SERIAL_STAGES;
    }
    return 0;
}

int bfcuda_stage_times(bfcuda_engine *e, double mean_ms[BFCUDA_N_STAGES], long *n_blocks, long *n_kernel_launches)
{
    if (e == nullptr) return fail(BFCUDA_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    int rc = flush_timing_ring(e);
    if (rc != 0) return rc;
    for (int s = 0; s < BFCUDA_N_STAGES; s++) {
        if (mean_ms) mean_ms[s] = e->stage_blocks > 0 ? e->stage_ms[s] / (double)e->stage_blocks : 0.0;
        e->stage_ms[s] = 0.0;
    }
    if (n_blocks) *n_blocks = e->stage_blocks;
    if (n_kernel_launches) *n_kernel_launches = e->launches;
    e->stage_blocks = 0;
    e->launches = 0;
    return 0;
}

int bfcuda_get_info(bfcuda_engine *e, struct bfcuda_info *info)
{
    if (e == nullptr || info == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    if (e->dirty) {
        build_tables(e);
        e->dirty = true;
    }
    memset(info, 0, sizeof(*info));
    info->n_fft = e->N;
    info->mac_split = e->split;
    info->n_streams = e->n_rings;
    info->kernels_per_block = 3 + (e->level_mix_first.size() > 1 && e->level_mix_first[1] > 0 ? 1 : 0) +
                              3 * (e->n_levels - 1) +
                              (plan_unpacks_first(e->plan) ? 2 : (e->shared_out.empty() ? 0 : 1));
    info->uses_graph = e->graph_used ? 1 : 0;
    info->max_batch = e->max_batch;
    info->mac_bytes_per_batch = e->mac_bytes_batch;
    info->sm_count = e->sm_count;
    info->mac_bytes_per_block = e->mac_bytes;
    info->device_bytes = e->device_bytes;
    snprintf(info->device_name, sizeof(info->device_name), "%s", e->device_name);
    return 0;
}

// ---- introspection ---------------------------------------------------------------------------------------

int bfcuda_debug_read(bfcuda_engine *e, int what, int index, int slot, void *dst)
{
    if (e == nullptr || dst == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    CU(cudaSetDevice(e->device));
    {
        int rc = sync_all(e);
        if (rc != 0) return rc;
    }
    const size_t nb = rs_bytes(e, e->N);
    switch (what) {
    case BFCUDA_DBG_INPUT_SPECTRUM: {
        if (!(e->flags & BFCUDA_FLAG_KEEP_INPUT_SPECTRA)) return fail(BFCUDA_EINVAL, "input spectra are only kept with flag 4 (debug)");
        if (index < 0 || index >= e->n_ch[0]) return fail(BFCUDA_EINVAL, "input channel out of range");
        const char *src = (const char *)e->d_xin + nb * index;
        CU(launch_permute(e->plan, src, e->d_scratch, 1, PLANAR_TO_HC, e->stream));
        CU(cudaMemcpyAsync(dst, e->d_scratch, nb, cudaMemcpyDeviceToHost, e->stream));
        break;
    }
    case BFCUDA_DBG_DELAYLINE: {
        if (index < 0 || index >= e->n_filters || slot < 0 || slot >= e->P) {
            return fail(BFCUDA_EINVAL, "filter or slot out of range");
        }
        // `slot` is the reference's numbering, cbuf[filter][(t + delay) % P] (bfrun.c:1600): the reference's slot s
        // holds the block written at the latest time tau <= t_last with (tau + delay) % P == s; the engine's ring is
        // longer (see bfcuda_create) and keeps that block at (tau + delay) % ring.
        const FilterState &dfs = e->filters[index];
        if (dfs.trans_until >= 0) {
            // in a delay transition the mirror IS the reference's ring (begin_transitions)
            if (dfs.ref_id[(size_t)slot] < 0) {
                memset(dst, 0, nb);
                break;
            }
            CU(launch_permute(e->plan, dfs.mirror + nb * (size_t)slot, e->d_scratch, 1, PLANAR_TO_BLOCKED, e->stream));
            CU(cudaMemcpyAsync(dst, e->d_scratch, nb, cudaMemcpyDeviceToHost, e->stream));
            break;
        }
        const int d = clamp_delay(e, e->filters[index].delayblocks);
        const long t_last = (long)e->t - 1;
        const long back = (((t_last + d - slot) % e->P) + e->P) % e->P;
        if (t_last - back < 0) {
            memset(dst, 0, nb);     // never written: still the zeros it was allocated with
            break;
        }
        const int R = e->fdl_ring;
        const int phys = (int)((((long)e->slot_t - 1 - back + d) % R + R) % R);
        const char *src = ring_ptr(e, e->filters[index].stream) + nb * (size_t)phys;
        CU(launch_permute(e->plan, src, e->d_scratch, 1, PLANAR_TO_BLOCKED, e->stream));
        CU(cudaMemcpyAsync(dst, e->d_scratch, nb, cudaMemcpyDeviceToHost, e->stream));
        break;
    }
    case BFCUDA_DBG_FILTER_OUTPUT: {
        if (index < 0 || index >= 2 * e->n_filters) return fail(BFCUDA_EINVAL, "filter out of range");
        const size_t n_slots = 2 * (size_t)std::max(1, e->n_filters) + 2 * (size_t)e->n_ch[1];
        std::vector<unsigned char> part(nb);
        for (int z = 0; z < 1; z++) {       // partial 0 holds the complete sum (launch_split_reduce)
            // the last block of the most recent launch
            const char *src = (const char *)e->d_Y + (size_t)((e->launch_no + 1) & 1u) * e->y_stride +
                              nb * (((size_t)z * e->last_batch + (e->last_batch - 1)) * n_slots + index);
            CU(launch_permute(e->plan, src, e->d_scratch, 1, PLANAR_TO_BLOCKED, e->stream));
            CU(cudaMemcpyAsync(z == 0 ? dst : (void *)part.data(), e->d_scratch, nb, cudaMemcpyDeviceToHost,
                               e->stream));
            CU(cudaStreamSynchronize(e->stream));
            if (z > 0) {
                // same order as the inverse kernel's partial sum
                if (e->rs == 4) {
                    float *a = (float *)dst;
                    const float *b = (const float *)part.data();
                    for (int i = 0; i < e->N; i++) a[i] = a[i] + b[i];
                } else {
                    double *a = (double *)dst;
                    const double *b = (const double *)part.data();
                    for (int i = 0; i < e->N; i++) a[i] = a[i] + b[i];
                }
            }
        }
        break;
    }
    case BFCUDA_DBG_OUTPUT_TIME: {
        if (index < 0 || index >= e->n_ch[1]) return fail(BFCUDA_EINVAL, "output channel out of range");
        const char *src = (const char *)e->d_out_time +
                          rs_bytes(e, ((size_t)(e->last_batch - 1) * e->n_ch[1] + index) * e->L);
        CU(cudaMemcpyAsync(dst, src, rs_bytes(e, e->L), cudaMemcpyDeviceToHost, e->stream));
        break;
    }
    default:
        return fail(BFCUDA_EINVAL, "unknown debug item %d", what);
    }
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

// ---- multi-GPU --------------------------------------------------------------------------------------------

int bfcuda_comm_unique_id(void *id_128_bytes)
{
    if (id_128_bytes == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    int rc = nccl_load();
    if (rc != 0) return rc;
    nccl_uid_t id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != 0) return fail(BFCUDA_ECOMM, "ncclGetUniqueId failed (%d)", r);
    memcpy(id_128_bytes, &id, sizeof(id));
    return 0;
}

int bfcuda_comm_init(bfcuda_engine *e, int rank, int n_ranks, const void *id_128_bytes)
{
    if (e == nullptr || id_128_bytes == nullptr) return fail(BFCUDA_EINVAL, "null argument");
    int rc = nccl_load();
    if (rc != 0) return rc;
    CU(cudaSetDevice(e->device));
    nccl_uid_t id;
    memcpy(&id, id_128_bytes, sizeof(id));
    int r = g_nccl.CommInitRank(&e->comm, n_ranks, id, rank);
    if (r != 0) {
        e->comm = nullptr;
        return fail(BFCUDA_ECOMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    }
    e->n_ranks = n_ranks;
    return 0;
}

int bfcuda_comm_shared_outputs(bfcuda_engine *e, int n_shared, const int *out_channels)
{
    if (e == nullptr || (n_shared > 0 && out_channels == nullptr)) return fail(BFCUDA_EINVAL, "null argument");
    for (int i = 0; i < n_shared; i++) {
        if (out_channels[i] < 0 || out_channels[i] >= e->n_ch[1]) {
            return fail(BFCUDA_EINVAL, "shared output channel out of range");
        }
    }
    e->shared_out.assign(out_channels, out_channels + n_shared);
    e->dirty = true;
    return 0;
}

}  // extern "C"

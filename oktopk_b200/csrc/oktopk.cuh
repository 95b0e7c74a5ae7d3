// Shared declarations of the fused sparse-allreduce kernels: device-resident per-bucket state,
// launch parameters, symmetric-block layout.  Included by the .cu files (nvcc) and bindings.cpp (g++).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#ifndef OKT_MAXP
#define OKT_MAXP 16
#endif

namespace okt {

constexpr int kThreads = 512;             // threads per CTA of the persistent kernels
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 1024;              // (idx,val) entries per pulled chunk (4 KB + 4 KB)
constexpr int kHistBins = 2048;           // 11-bit radix digit
constexpr int kMaxWarpsTotal = 16384;     // per-warp counters for the quantile cuts
constexpr int kGuardMax = 64;            // rungs of the over-selection ladder (fine guard rungs + coarse cap rungs)
constexpr int kGuardFineMax = 15;         // most fine (reference guard) rungs
constexpr int kCtasPerSm = 1;           // persistent kernels: one 512-thread CTA per SM (<= 128 registers/thread)
constexpr int kPackTile = 2;            // 128-bit vectors per thread per tile of the streaming pass
constexpr int kPackStages = 4;          // TMA ring depth of the streaming pass
constexpr int kPackSmemBytes = kPackStages * 2 * kPackTile * kThreads * 16;   // (grad + residual) tiles: 128 KB
constexpr int kScanStages = 2 * kPackStages;   // the region scan reuses the whole ring as 8 single-array stages

// one record per call, written by block 0 at the end of the fused kernel (observability: --trace, PROFILING)
constexpr int kTraceLen = 256;
struct TraceRec {
    uint32_t epoch;
    int local_count, global_count, recv_total, gather_total, overflow_send, overflow_gather, redo;
    float local_thr, global_thr;
    // phase durations in microseconds (globaltimer, block 0): exact-threshold/re-partition work, pack pass, wait for
    // the peers' reduce-scatter flags (= how far the slowest peer is behind), reduce, global selection, wait for the
    // allgather flags, final phase
    float us_local, us_pack, us_wait_rs, us_reduce, us_gselect, us_wait_ag, us_final;
    unsigned long long t_begin;                                   // globaltimer at kernel entry
};

// ---- device-resident, zero-initialised, one per bucket ---------------------------------------
struct OktState {
    unsigned long long bar;               // grid-barrier ticket counter
    unsigned int tick[4];                 // last-CTA-done tickets (pack / gselect / final / spare), self-resetting
    float local_thr;                      // threshold carried into the next call (after adaptation)
    float local_thr_used;                 // threshold actually applied in the last call (after guard)
    float global_thr;
    uint32_t epoch;                       // completed calls; the running call uses epoch + 1
    int edges[OKT_MAXP + 1];              // region edges, edges[0] = 0, edges[P] = n
    int send_cursor[OKT_MAXP];            // per-destination slot cursors (reset in-kernel)
    int gather_cursor;
    int cand_cursor;                      // first-touch candidate list of the reduce phase (reset in-kernel)
    int guard_counts[kGuardMax];          // selected elements whose HIGHEST passed ladder rung is j (suffix sums = #(|acc| > T_j))
    uint32_t sel_prefix;                  // radix-select running prefix / remaining rank
    uint32_t sel_krem;
    int cuts[OKT_MAXP];
    // statistics of the last call (read lazily by the host; never on the hot path)
    int stat_local_count;
    int stat_global_count;
    int stat_recv_total;                  // entries pulled in the reduce phase
    int stat_gather_total;                // entries pulled in the final phase
    int stat_overflow_send;               // entries dropped in the LAST call because a send slot was full
    int stat_overflow_gather;             // same for the allgather slot
    int stat_redo;                        // pack passes repeated in the last call (overflow policy: raise threshold + redo)
    int stat_dense_fallback;              // 1 if the last call took TopkDSA's dense allgather path
    int stat_mode;
    int fault;                            // FaultCode of the first bounded wait that timed out (0 = healthy)
    float pack_thr;                       // threshold the pack pass finally selected with (after overflow redos)
    unsigned long long cum_overflow_send;     // cumulative (64-bit: a diverging run must not wrap the counter)
    unsigned long long cum_overflow_gather;
    unsigned long long cum_redo;
    unsigned long long snap_overflow_send;    // cumulative values at the end of the previous call
    unsigned long long snap_overflow_gather;
    unsigned long long snap_redo;
    unsigned long long t_phase[8];        // globaltimer stamps (block 0): 0 after local phase, 1 pack done, 2 reduce done,
                                          // 3 gselect done, 4 end, 5 kernel entry, 6 rs flags in, 7 ag flags in
    double gs_sum, gs_sumsq;              // Gaussiank moments
    double clip_sumsq;                    // norm_clip: sum of squares of the incoming gradient (reset at the end of the call)
    uint32_t hist[kHistBins];
    int wcounts[kMaxWarpsTotal];
    TraceRec trace[kTraceLen];            // per-call history ring (index = epoch % kTraceLen), see read_trace
};

// ---- byte offsets inside every rank's symmetric block (identical on all ranks) -----------------
//
// Send slots.  LOSSLESS layout (cap == 0, the default): ONE buffer of ~n entries in which destination d's slot
// starts at  slot_off(d) = align4(edges[d]) + 4 d  -- its capacity is at least the length of region d, i.e. at
// least the number of elements that can possibly be selected for d, so the exchange can never overflow, whatever
// the threshold (the reference gets the same guarantee from host-side Alltoall count handshakes,
// VGG/allreducer.py:708-726).  Sender and receiver derive the offsets from the region edges both already hold.
// BOUNDED layout (cap > 0): P slots of `cap` entries; overflow is handled by the in-kernel policy (raise the
// threshold and redo the pack pass; classic-residual schemes keep the unsent entries in the residual).
// The send slots are single-buffered: a peer has finished pulling call e's slots before it publishes its
// allgather flag of call e, which every rank waits for before it leaves call e.  The gather slots are
// double-buffered by call parity (the gather-type kernels have no second handshake).
struct SymmLayout {
    size_t rs_mbox;      // uint64 [2][MAXP]       reduce-scatter mailbox: (epoch<<32 | count) from src
    size_t rs_thr;       // float  [2][MAXP]       src's final local threshold (receiver-side filter)
    size_t ag_mbox;      // uint64 [2][MAXP]       allgather mailbox
    size_t cut_mbox;     // uint64 [2][MAXP]
    size_t cut_data;     // int32  [2][MAXP][MAXP]
    size_t done_mbox;    // uint64 [2][MAXP]       "finished reading your memory" flags (dense fallback, tree schemes)
    size_t tree_mbox;    // uint64 [2][MAXP]       gTopk: per-round list-ready flags (((epoch<<5)|round) << 32 | count)
    size_t send_idx;     // int32  [scap]          my selections, bucketed by destination region
    size_t send_val;     // float  [scap]
    size_t gat_idx;      // int32  [2][gcap]       my region's globally selected entries
    size_t gat_val;      // float  [2][gcap]
    size_t total;
    int cap;             // bounded per-destination capacity; 0 = lossless layout
    int scap;            // entries in the send buffer
    int gcap;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline SymmLayout make_layout(int P, int n, int cap, int gcap) {
    SymmLayout L;
    size_t o = 0;
    L.rs_mbox = o;  o += sizeof(uint64_t) * 2 * OKT_MAXP;
    L.rs_thr = o;   o += sizeof(float) * 2 * OKT_MAXP;
    L.ag_mbox = o;  o += sizeof(uint64_t) * 2 * OKT_MAXP;
    L.cut_mbox = o; o += sizeof(uint64_t) * 2 * OKT_MAXP;
    L.cut_data = o; o += sizeof(int) * 2 * OKT_MAXP * OKT_MAXP;
    L.done_mbox = o; o += sizeof(uint64_t) * 2 * OKT_MAXP;
    L.tree_mbox = o; o += sizeof(uint64_t) * 2 * OKT_MAXP;
    o = align_up(o, 1024);
    const size_t scap = cap > 0 ? (size_t)P * cap : align_up((size_t)n, 4) + 4 * (size_t)OKT_MAXP + 1024;
    L.send_idx = o; o += sizeof(int) * scap;   o = align_up(o, 1024);
    L.send_val = o; o += sizeof(float) * scap; o = align_up(o, 1024);
    L.gat_idx = o;  o += sizeof(int) * 2 * (size_t)gcap;      o = align_up(o, 1024);
    L.gat_val = o;  o += sizeof(float) * 2 * (size_t)gcap;    o = align_up(o, 1024);
    L.total = o;
    L.cap = cap;
    L.scap = (int)scap;
    L.gcap = gcap;
    return L;
}

enum Phase : int {
    PH_LOCAL = 0,      // two-pass iterations: acc -> residual, exact k-th |acc| or guard, re-partition
    PH_PACK = 1,       // (accumulate +) select + pack per destination region
    PH_PUBLISH_RS = 2, // finalise threshold, publish counts to the region owners
    PH_REDUCE = 3,     // pull every source's slot for my region, scatter-add
    PH_GSELECT = 4,    // global selection on my region, pack the allgather slot
    PH_PUBLISH_AG = 5,
    PH_FINAL = 6,      // pull all slots, (exact top-k,) scatter result, clear residual, adapt
    PH_END = 7
};

enum ResidualMode : int { RES_OKTOPK = 0, RES_LOCAL_GT = 1, RES_LOCAL_GE = 2 };
enum GlobalMode : int { GLB_THRESHOLD = 0, GLB_EXACT_TOPK = 1, GLB_ALL_NONZERO = 2 };

struct OktParams {
    float* g;            // gradient bucket (result written in place)
    float* res;          // residual / accumulator
    OktState* st;
    int* cand;           // local scratch: region-local indices that received a contribution (capacity ccap)
    int ccap;
    int cand_mode;       // 1: first-touch candidate list (low density), 0: region scan (high density)
    float prefilter;     // exact iterations: radix-select only elements above prefilter * carried threshold (0 = all)
    char* peers[OKT_MAXP];   // every rank's symmetric block as mapped into this process
    float* peer_g[OKT_MAXP]; // every rank's gradient bucket (null when g is not the symmetric bucket): dense fallback
    SymmLayout L;
    int n, P, rank, k;
    int exact_local, repartition, uniform_regions;
    int residual_mode, global_mode;
    int deterministic, pull_tma;
    int phase_begin, phase_end;
    // Over-selection ladder T_0 = thr, T_j = T_{j-1} * (j <= guard_loops ? guard_factor : cap_factor).  The reference's
    // guard (VGG/compression.py:392-404) climbs the first guard_loops rungs while the count exceeds guard_limit; the cap
    // (cap_limit > 0, a B200-side extension: the counts of ALL rungs come out of the same streaming pass for free) keeps
    // climbing the coarse rungs while the count exceeds cap_limit, so a stale threshold can never ship more than
    // cap_limit entries per rank.
    int guard_loops, guard_limit;
    float guard_factor;
    int cap_limit, cap_rungs;
    float cap_factor;
    double l_low_cnt, l_high_cnt;       // local adaptation bounds, already multiplied by k
    float l_factor;
    double g_low_cnt, g_high_cnt;
    float g_inc, g_dec;
    unsigned long long timeout_ns;      // bound of every cross-GPU wait (0 = unbounded)
    int max_redo;                       // bounded slots: how many times the pack pass may be repeated with a raised threshold
    float redo_factor;                  // first raise; squared after every further attempt
    int dense_nnz_limit;                // GLB_ALL_NONZERO (TopkDSA): total gathered nnz >= this => dense allgather path (0 = never)
    int* host_fault;                    // mapped pinned int: fault code mirrored to the host without a sync (may be null)
    int trace;                          // 1: write a TraceRec per call
};

// ---- gather-type schemes (TopkAopt / Gaussiank / TopkA): select -> own slot -> everyone adds all ---
enum GatherSelect : int { GS_THRESHOLD_REUSE = 0, GS_GAUSSIAN = 1, GS_EXACT_TOPK = 2 };

struct GatherParams {
    float* g;
    float* res;
    OktState* st;
    char* peers[OKT_MAXP];
    SymmLayout L;
    int n, P, rank, k;
    int select_mode;          // GatherSelect
    int exact_now;            // GS_THRESHOLD_REUSE: recompute the exact threshold in this call
    int gauss_mode;           // 0 vgg, 1 lstm, 2 bert
    int gauss_loops;
    float gauss_factor;
    float density;
    int pull_tma;
    unsigned long long timeout_ns;
    int reselect;             // TopkA2: global top-k re-selection of the gathered union + put-back of the losers
    float clip_max_norm;      // > 0: scale the incoming gradient so that its L2 norm is at most this (norm_clip)
    unsigned* bitmap;         // n/32 words, all-zero between calls (exact first-touch detection for the union list)
    int* cand;                // union candidate list (capacity ccap)
    int ccap;
    int* host_fault;
};

// ---- gTopk: log2(P) rounds of pairwise list merges toward rank 0, then a broadcast ------------------
struct TreeParams {
    float* g;
    float* res;
    OktState* st;
    char* peers[OKT_MAXP];
    SymmLayout L;
    int n, P, rank, k;
    int pull_tma;
    unsigned long long timeout_ns;
    float clip_max_norm;
    unsigned* bitmap;         // n/32 words, all-zero between calls
    int* cand;                // two union lists of ccap/2 entries each (ping-pong across rounds)
    int ccap;
    int* sel_idx;             // my original picks (for the put-back of the non-survivors), capacity selcap
    float* sel_val;
    int selcap;
    int* host_fault;
};

// ---- dense allreduce over peer memory -------------------------------------------------------------
struct DenseParams {
    float* bufs[OKT_MAXP];    // every rank's gradient bucket (symmetric allocation)
    uint64_t* flags[OKT_MAXP];  // every rank's flag block: uint64 [2][grid][MAXP]
    unsigned long long* epoch;  // local, one counter per CTA
    int n, P, rank;
    float scale;              // 1/P
    int* fault;               // bucket fault word (OktState::fault)
    unsigned long long timeout_ns;
    float* mc;                // multicast (NVLS) mapping of the bucket, or null: switch-side reduction with multimem.*
    int* host_fault;
};

// ---- host-callable launchers (implemented in the .cu files) ----------------------------------------
int okt_max_coop_grid(int device);
cudaError_t launch_oktopk(const OktParams& p, int grid, cudaStream_t stream);
cudaError_t launch_gather_scheme(const GatherParams& p, int grid, cudaStream_t stream);
cudaError_t launch_gtopk(const TreeParams& p, int grid, cudaStream_t stream);
int gtopk_max_coop_grid(int device);
int gather_max_coop_grid(int device);
cudaError_t launch_dense_allreduce(const DenseParams& p, int grid, cudaStream_t stream);
cudaError_t launch_kth_abs(const float* x, int n, int k, OktState* st, float* out_thr, int grid, cudaStream_t stream);
cudaError_t launch_fused_sgd(float* p, float* g, float* mom, int n, float lr, float momentum, float dampening,
                             float weight_decay, int nesterov, int first_step, int zero_grad, float grad_scale,
                             const float* lr_ptr, const int* fault, cudaStream_t stream);
cudaError_t launch_fused_bert_adam(float* p, float* g, float* m, float* v, int n, float lr, float b1, float b2,
                                   float eps, float weight_decay, int zero_grad, const float* lr_ptr,
                                   const int* fault, cudaStream_t stream);
// multi-tensor gradient landing: copy up to kLandMax autograd-produced gradient tensors into the flat bucket in ONE launch
constexpr int kLandMax = 96;
struct LandParams {
    const float* src[kLandMax];
    long long dst_off[kLandMax];      // element offset inside the bucket
    int numel[kLandMax];
    int blk_begin[kLandMax + 1];      // first CTA of tensor t (CTAs are dealt proportionally to size)
    int count;
};
cudaError_t launch_land(const LandParams& lp, float* bucket, cudaStream_t stream);
// fused (conv-bias +) BatchNorm + ReLU, channels_last fp32, training mode (csrc/bnrelu.cu)
int bn_num_blocks(int M, int C);
cudaError_t launch_bn_forward(const float* x, float* y, float* partial, const float* gamma, const float* beta, const float* cbias,
                              float* save_mean, float* save_invstd, float* rmean, float* rvar, long long* nbt, float momentum,
                              float eps, int relu, int M, int C, cudaStream_t stream);
cudaError_t launch_bn_backward(const float* x, const float* dy, float* dx, float* partial, const float* gamma, const float* beta,
                               const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, int relu, int M,
                               int C, cudaStream_t stream);
cudaError_t launch_maxpool2_fwd(const float* x, float* y, unsigned char* arg, int N, int H, int W, int C, cudaStream_t stream);
cudaError_t launch_maxpool2_bwd(const float* dy, const unsigned char* arg, float* dx, int N, int H, int W, int C,
                                cudaStream_t stream);
cudaError_t launch_momentum_correct(float* g, float* buf, int n, float momentum, cudaStream_t stream);
cudaError_t launch_l2norm_sq(const float* x, int n, float* out, cudaStream_t stream);
cudaError_t launch_scale(float* x, int n, const float* norm_sq, float max_norm, cudaStream_t stream);

}  // namespace okt

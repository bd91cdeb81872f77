// Device-side state of one search context (one process / one B200).
#pragma once
#include <cuda_runtime.h>

#include <chrono>

#include "common.h"

namespace so {

#define SO_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            so::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, \
                          cudaGetErrorString(e_));                                             \
            return SO_ENODEV;                                                                  \
        }                                                                                      \
    } while (0)

// growable device buffer
template <class T>
struct DBuf {
    T *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return SO_OK;
        if (p) cudaFree(p);
        p = nullptr;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e != cudaSuccess) {
            cap = 0;
            set_error("cudaMalloc of %zu bytes failed: %s", want * sizeof(T), cudaGetErrorString(e));
            return SO_ENOMEM;
        }
        cap = want;
        return SO_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// One target chunk index resident in HBM (K3).
struct ChunkIndex {
    i64 c0 = 0, c1 = 0;          // target ordinals [c0, c1)
    uint32_t total = 0;          // residues in the chunk
    uint32_t n_seeds = 0;        // len(locus)
    i64 n_used = 0;              // non-empty buckets
    i64 threshold = 0;
    uint32_t max_tlen = 0;
    uint32_t max_bucket = 0;     // largest bucket (bounds the seed hits of a query: threshold * len + max_bucket)
    uint32_t *d_start = nullptr; // [NC + 1] start[b] = #entries with bucket < b
    uint32_t *d_locus = nullptr; // [n_seeds] reverse insertion order inside a bucket
    uint32_t *d_soas = nullptr;  // [M + 1] residue offsets inside the chunk
    uint2 *d_hdsst = nullptr;    // [n_seeds] (sequence + 1, position) of every locus entry (bisect quirk applied)
    double build_ms = 0;
};

// Candidates of a query block against one chunk, in the reference's order, as packed 64-bit values
// (target ordinal << 40 | score << 20 | diagonal + kCandDiagBias) in a reusable PINNED host buffer.
static const int kCandDiagBias = 1 << 19;
struct PackedCands {
    uint64_t *vals = nullptr;
    size_t n = 0, cap = 0;
    std::vector<uint64_t> offsets;  // [nq + 1]
    int reserve(size_t want) {
        if (want <= cap) return SO_OK;
        size_t ncap = want + want / 2 + 4096;
        uint64_t *nv = nullptr;
        if (cudaMallocHost((void **)&nv, ncap * 8) != cudaSuccess) {
            set_error("cudaMallocHost of %zu bytes failed", ncap * 8);
            return SO_ENOMEM;
        }
        if (n) memcpy(nv, vals, n * 8);
        if (vals) cudaFreeHost(vals);
        vals = nv;
        cap = ncap;
        return SO_OK;
    }
    void release() {
        if (vals) cudaFreeHost(vals);
        vals = nullptr;
        n = cap = 0;
    }
};
static inline so_cand unpack_cand(uint64_t v) {
    so_cand cd;
    const int diag = (int)(v & 0xfffffu) - kCandDiagBias;
    cd.target = (uint32_t)(v >> 40);
    cd.score = (uint32_t)((v >> 20) & 0xfffffu);
    // guess_start (fsearch.py:2544-2553): d = sst - qst = -diag -> head of the diagonal
    if (diag < 0)
        cd.qi = 0, cd.qj = (uint32_t)(-diag);
    else
        cd.qi = (uint32_t)diag, cd.qj = 0;
    return cd;
}

// Candidates of one query block on the device (H3, select.cu): per-query lists with a fixed capacity (one entry per
// target at most), filled chunk by chunk in the reference's concatenation order, then sorted / cut by k_h3_select.
struct BlockStore {
    DBuf<uint64_t> vals;    // [nq][capq] packed candidates (target << 40 | score << 20 | diagonal + bias)
    DBuf<uint32_t> count;   // [nq] candidates per query, [nq] work counter, [nq + 1] error flag
    DBuf<uint64_t> keys;    // per resident CTA: (0xffffffff - score) << 32 | index
    DBuf<uint32_t> la, lb;  // per resident CTA: stopper lists of the partition in flight
    DBuf<uint64_t> sel;     // [nq][vmax] the first min(vmax, n) candidates in the reference's sorted order
    DBuf<uint32_t> sel_n;   // [nq] n = len(hits)
    i64 nq = 0;
    size_t capq = 0;
    int vmax = 0, grid = 0;
    int prepare(so_ctx *c, i64 nq, size_t capq, int vmax, cudaStream_t st);
    int append(const uint64_t *d_cv, const uint32_t *d_bounds, int n, int q0, cudaStream_t st);
    int select(cudaStream_t st);
    void release();
};

struct HostProfile {  // wall-clock breakdown of the host side (SO_PROFILE=1 prints it)
    double cand_ms = 0, d2h_ms = 0, order_ms = 0, rounds_ms = 0, align_ms = 0, replay_ms = 0, final_ms = 0, total_ms = 0;
};

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

}  // namespace so

struct so_ctx {
    int device = 0;
    so::Params P;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaStream_t stream_aln = nullptr;        // alignment rounds run on their own stream (pipeline worker thread)
    cudaEvent_t ev_aln[4] = {};

    // targets (raw bytes) and queries (seg-masked bytes), packed, resident in HBM
    so::i64 n_t = 0, n_q = 0;
    std::vector<uint64_t> t_off, q_off;       // host copies of the offsets
    std::vector<uint8_t> q_masked;            // host copy of the masked queries
    const uint8_t *t_host = nullptr;          // caller-owned (valid during so_set_targets only)
    uint8_t *d_tres = nullptr, *d_qres = nullptr;
    uint8_t *d_tcls = nullptr, *d_qcls = nullptr;  // 5-bit BLOSUM62 residue classes of the same buffers
    // X-drop view of the targets (search.cu, k_single_ungap): classes with position 0 of every
    // sequence replaced by a terminator, forward copy at tung_off[0] and reversed copy at tung_off[1]
    uint8_t *d_tung = nullptr;
    uint32_t tung_off[2] = {0, 0};
    uint64_t *d_toff = nullptr, *d_qoff = nullptr;
    uint32_t *d_perm = nullptr;               // per query: positions in the reference quicksort order of -kscs (S3)
    size_t q_cap_bytes = 0, q_cap_seqs = 0;   // capacity of the query buffers (grow-only: so_set_queries runs per block)
    std::vector<uint32_t> q_perm_host;        // reused host staging of d_perm
    so::i64 sub_block = 0;                    // >0: fixed number of queries per seeding sub-block (tests)
    uint32_t max_qlen = 0, max_tlen = 0;

    std::vector<so::ChunkIndex> chunks;

    // scratch (grown on demand, reused)
    so::DBuf<uint8_t> scratch[64];
    // second candidate-production lane (so_search runs two producer threads on alternating query blocks so
    // that one lane's host synchronisations and D2H copies overlap the other lane's kernels)
    // (up to kMaxLanes lanes: lane 0 = scratch / stream / ev above, lanes 1.. = the *_x arrays; with more lanes than
    // two, kernels of different stages -- issue-bound X-drop, latency-bound cell passes -- share the SMs more often)
    enum { kMaxLanes = 4 };
    so::DBuf<uint8_t> scratch_x[kMaxLanes - 1][64];
    cudaStream_t stream_x[kMaxLanes - 1] = {};
    cudaEvent_t ev_x[kMaxLanes - 1][8] = {};
    so::DBuf<uint8_t> *lane_scratch(int lane) { return lane ? scratch_x[lane - 1] : scratch; }
    // shared_stream: every lane enqueues on ONE stream (kernels of different query blocks never run side by side; a lane
    // waits for its own block through an event, so its host work still overlaps the other lanes' kernels)
    bool shared_stream = false;
    cudaEvent_t ev_sync[kMaxLanes] = {};
    unsigned long long h_ctl[kMaxLanes][11] = {};  // control block of the lane's last sync-free block (copied before the wait)
    cudaStream_t lane_stream(int lane) const { return (lane && !shared_stream) ? stream_x[lane - 1] : stream; }
    // wait until everything this lane has enqueued so far is done
    cudaError_t lane_wait(int lane) {
        cudaError_t e = cudaEventRecord(ev_sync[lane], lane_stream(lane));
        return e != cudaSuccess ? e : cudaEventSynchronize(ev_sync[lane]);
    }
    cudaEvent_t *lane_ev(int lane) { return lane ? ev_x[lane - 1] : ev; }
    std::vector<cudaEvent_t> ev_pool[kMaxLanes];
    size_t ev_used[kMaxLanes] = {};  // per lane: stage events of the sync-free path, read at the end of a block
    so_stats stats_lane[kMaxLanes] = {};
    int n_lanes = 1;                 // measured: one production lane + the alignment worker beats two lanes (126 vs 129 ms / 4096 queries)
    double d2h_ms_lane[kMaxLanes] = {};
    so::DBuf<uint64_t> trace;
    void *h_pinned = nullptr;
    size_t h_pinned_cap = 0;

    so_stats stats = {};
    so_stats stats_aln = {};                  // written by the alignment side only; merged by merge_align_stats
    so::HostProfile prof;
    std::vector<so::PackedCands> cand_pool;   // one pinned buffer per chunk, reused across query blocks
    so::BlockStore bstore[kMaxLanes];         // per production lane
    enum { kMaxSlots = 2 * kMaxLanes };
    uint64_t *h_sel[kMaxSlots] = {};          // pinned, per pipeline slot: selected candidates of a block
    uint32_t *h_sel_n[kMaxSlots] = {};
    size_t h_sel_cap[kMaxSlots] = {};
};

namespace so {
// implemented in the .cu files
int upload_tables();
int align_pairs(so_ctx *c, const so_pair *pairs, i64 n, so_aln *out);
int build_chunk_index(so_ctx *c, ChunkIndex &ix);
void free_chunk_index(ChunkIndex &ix);
// candidates of queries [q_begin, q_end) against one chunk: copied to `out` (pinned host memory) or, when `bs` is
// given, appended on the device to the block's per-query lists (query q_begin = list bs_q0)
int chunk_candidates(so_ctx *c, const ChunkIndex &ix, i64 q_begin, i64 q_end, PackedCands &out, int lane = 0,
                     BlockStore *bs = nullptr, int bs_q0 = 0);
int qsort_prefix_device(so_ctx *c, const uint32_t *keys, i64 n, i64 need, uint32_t *perm_out);
int upload_search_config(so_ctx *c);
// sync-free candidate production of a whole query block (one pattern, one alphabet, sequences < 8192): every chunk's
// candidates are appended to `bs` with no host synchronisation; `eligible` = false when the block needs the general
// path (nothing was launched); finish_fast_block reads the flags / counters after the block's stream sync
int block_candidates_fast(so_ctx *c, i64 b0, i64 b1, int lane, BlockStore &bs, bool &eligible, int only_chunk = -1);
int enqueue_fast_ctl(so_ctx *c, int lane);  // D2H of the block's flags / counters, enqueued behind its kernels
int finish_fast_block(so_ctx *c, int lane, bool &redo);
void merge_lane_stats(so_ctx *c);
int ensure_pinned(so_ctx *c, size_t bytes);
void merge_align_stats(so_ctx *c);
int classify_residues(so_ctx *c, const uint8_t *d_in, uint8_t *d_out, size_t n);
int build_ungap_targets(so_ctx *c);
}  // namespace so

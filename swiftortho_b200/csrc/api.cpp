// C ABI entry points (include/swiftortho_b200.h) and the host side of the search:
// context set-up, query preparation (H1 seg mask, S3 position order), the candidate merge / sort /
// stop rule of blastp (H3, lib/fsearch.py:3039-3106) driving the alignment kernels in rounds, the
// e-value filter and the final per-query order (F, lib/fsearch.py:3071-3072, 3108-3110), and the
// 16-column text writer (lib/fsearch.py:3233-3243).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <cmath>
#include <thread>

#include "context.h"

struct so_fasta;  // fasta.cpp

namespace so {

template <class F>
static void parallel_for(i64 n, F f) {
    // host threads of one process: all cores by default, cores / WORLD_SIZE under a multi-rank launch (one process per
    // GPU shares the host), SO_HOST_THREADS overrides
    static const unsigned cap = []() {
        unsigned nt = std::thread::hardware_concurrency();
        if (nt == 0) nt = 1;
        if (const char *w = getenv("LOCAL_WORLD_SIZE") ? getenv("LOCAL_WORLD_SIZE") : getenv("WORLD_SIZE")) {
            const int ws = atoi(w);
            if (ws > 1) nt = std::max(1u, nt / (unsigned)ws);
        }
        if (const char *e = getenv("SO_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(e));
        return std::min(nt, 32u);
    }();
    unsigned nt = cap;
    if ((i64)nt > n) nt = (unsigned)std::max<i64>(n, 1);
    if (n < 4 || nt == 1) {
        for (i64 i = 0; i < n; i++) f(i);
        return;
    }
    std::atomic<i64> next(0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&]() {
            for (;;) {
                const i64 step = n >= 1024 ? 16 : 1;
                i64 i0 = next.fetch_add(step);
                if (i0 >= n) return;
                for (i64 i = i0; i < std::min<i64>(n, i0 + step); i++) f(i);
            }
        });
    for (auto &t : th) t.join();
}

int ensure_pinned(so_ctx *c, size_t bytes) {
    if (bytes <= c->h_pinned_cap) return SO_OK;
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    c->h_pinned = nullptr;
    c->h_pinned_cap = 0;
    SO_CUDA(cudaMallocHost(&c->h_pinned, bytes + bytes / 4));
    c->h_pinned_cap = bytes + bytes / 4;
    return SO_OK;
}

static int check_offsets(const uint64_t *off, i64 n, uint32_t &maxlen, const char *what) {
    maxlen = 0;
    for (i64 i = 0; i < n; i++) {
        if (off[i + 1] < off[i]) {
            set_error("%s offsets are not monotone", what);
            return SO_EINVAL;
        }
        uint64_t L = off[i + 1] - off[i];
        if (L >= 65536) {
            set_error("%s %lld has %llu residues; the limit is 65535", what, (long long)i, (unsigned long long)L);
            return SO_ELIMIT;
        }
        maxlen = std::max<uint32_t>(maxlen, (uint32_t)L);
    }
    return SO_OK;
}

// ---------------------------------------------------------------------------------------------
// H3 + F: one query's candidate list -> alignment requests -> rows
// ---------------------------------------------------------------------------------------------
struct QueryState {
    std::vector<uint64_t> order;  // candidates in the reference's sorted order (prefix of length `limit`);
                                  // low 32 bits = index into the chunk-major concatenation of the candidates
    i64 qord = 0;                 // global query ordinal
    i64 limit = 0;                // min(vmax, len(hits))
    i64 next = 0;                 // next candidate (in sorted order) to align
    i64 lead = 0;                 // candidates up to the last one whose UNGAPPED score already passes the e-value
    i64 last_hits = 0;            // candidates of the last round that gave a row
    int rounds = 0;               // alignment rounds this query took part in
    double mmiss = 0;
    i64 unmch = 0, bv = 0;
    bool done = false;
    std::vector<so_cand> sel;     // the first `limit` candidates in sorted order
    std::vector<so_hit> rows;
    // requests of the current round
    i64 req_first = 0, req_count = 0;  // candidates covered
};

// Alignment rounds.  The stop rule (fsearch.py:3103) is sequential per query: it ends after ceil(mmiss) consecutive
// misses.  The first round of a query submits exactly the candidates the rule is certain to reach: the candidates
// up to the last one whose ungapped diagonal score alone passes the e-value (they are hits unless the banded
// alignment scores lower, plus those above a diagonal score random pairs rarely reach) + ceil(mmiss).  Most queries
// end there; the others continue with exactly ceil(mmiss - unmch) more, then with rounds of 64, 128, ... candidates.
static const i64 kRoundMax = 4096;

}  // namespace so

using namespace so;

extern "C" {

int so_abi_version(void) { return SO_ABI_VERSION; }
const char *so_last_error(void) { return so::get_error(); }

int so_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int so_seg(const uint8_t *seq, int64_t n, uint8_t *out) {
    if (n < 0 || (n > 0 && (!seq || !out))) {
        set_error("so_seg: bad argument");
        return SO_EINVAL;
    }
    so::seg_mask(seq, n, out);
    return SO_OK;
}

int so_qsort_perm(const int64_t *keys, int64_t n, int32_t *perm) {
    if (n < 0 || (n > 0 && (!keys || !perm))) {
        set_error("so_qsort_perm: bad argument");
        return SO_EINVAL;
    }
    so::qsort_perm((const so::i64 *)keys, n, perm);
    return SO_OK;
}

int so_qsort_prefix_device(so_ctx *c, const uint32_t *keys, int64_t n, int64_t need, uint32_t *perm) {
    if (!c || n < 0 || need < 0 || (n > 0 && (!keys || !perm))) {
        set_error("so_qsort_prefix_device: bad argument");
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(c->device));
    return so::qsort_prefix_device(c, keys, n, need, perm);
}

int64_t so_score2bit(int64_t raw) { return so::score2bit(raw); }
double so_bit2e(int64_t D, int64_t ql, int64_t tl, int64_t bit) { return so::bit2e(D, ql, tl, bit); }
int so_f2s(double e, char *out, int cap) {
    if (!out || cap < 2) return SO_EINVAL;
    std::string s = so::f2s(e);
    snprintf(out, (size_t)cap, "%s", s.c_str());
    return SO_OK;
}

void so_free(void *p) { free(p); }

int so_ctx_create(int device, const so_params *p, so_ctx **out) {
    if (!out) {
        set_error("so_ctx_create: null out");
        return SO_EINVAL;
    }
    Params P;
    int rc = parse_params(p, P);
    if (rc != SO_OK) return rc;
    int ndev = so_device_count();
    if (ndev <= 0) {
        set_error("no CUDA device visible: swiftortho_b200 has no CPU fallback");
        return SO_ENODEV;
    }
    if (device < 0 || device >= ndev) {
        set_error("device %d out of range (%d visible)", device, ndev);
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(device));
    so_ctx *c = new so_ctx();
    c->device = device;
    c->P = P;
    SO_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SO_CUDA(cudaStreamCreateWithFlags(&c->stream_aln, cudaStreamNonBlocking));
    for (auto &st : c->stream_x) SO_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto &e : c->ev_sync) SO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &evs : c->ev_x)
        for (auto &e : evs) SO_CUDA(cudaEventCreate(&e));
    for (auto &e : c->ev) SO_CUDA(cudaEventCreate(&e));
    for (auto &e : c->ev_aln) SO_CUDA(cudaEventCreate(&e));
    if ((rc = so::upload_tables()) != SO_OK) {
        delete c;
        return rc;
    }
    *out = c;
    return SO_OK;
}

void so_ctx_destroy(so_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto &ix : c->chunks) so::free_chunk_index(ix);
    for (auto &s : c->scratch) s.release();
    for (auto &sx : c->scratch_x)
        for (auto &s : sx) s.release();
    for (auto &evs : c->ev_x)
        for (auto &e : evs)
            if (e) cudaEventDestroy(e);
    for (auto &st : c->stream_x)
        if (st) cudaStreamDestroy(st);
    for (auto &e : c->ev_sync)
        if (e) cudaEventDestroy(e);
    c->trace.release();
    if (c->d_tres) cudaFree(c->d_tres);
    if (c->d_qres) cudaFree(c->d_qres);
    if (c->d_toff) cudaFree(c->d_toff);
    if (c->d_qoff) cudaFree(c->d_qoff);
    if (c->d_perm) cudaFree(c->d_perm);
    if (c->d_tcls) cudaFree(c->d_tcls);
    if (c->d_qcls) cudaFree(c->d_qcls);
    if (c->d_tung) cudaFree(c->d_tung);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    for (auto &pc : c->cand_pool) pc.release();
    for (auto &b : c->bstore) b.release();
    for (int k = 0; k < so_ctx::kMaxSlots; k++) {
        if (c->h_sel[k]) cudaFreeHost(c->h_sel[k]);
        if (c->h_sel_n[k]) cudaFreeHost(c->h_sel_n[k]);
    }
    for (auto &e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : c->ev_aln)
        if (e) cudaEventDestroy(e);
    if (c->stream_aln) cudaStreamDestroy(c->stream_aln);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int so_set_targets(so_ctx *c, const uint8_t *residues, const uint64_t *offsets, int64_t n) {
    if (!c || !offsets || n < 0 || (!residues && n > 0 && offsets[n] > 0)) {
        set_error("so_set_targets: bad argument");
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(c->device));
    int rc = check_offsets(offsets, n, c->max_tlen, "target");
    if (rc != SO_OK) return rc;
    if (n >= (1 << 24)) {
        set_error("at most 16777215 target sequences per context");
        return SO_ELIMIT;
    }
    for (auto &ix : c->chunks) so::free_chunk_index(ix);
    c->chunks.clear();
    if (c->d_tres) cudaFree(c->d_tres);
    if (c->d_toff) cudaFree(c->d_toff);
    if (c->d_tcls) cudaFree(c->d_tcls);
    c->d_tres = nullptr, c->d_toff = nullptr, c->d_tcls = nullptr;
    c->n_t = n;
    c->t_off.assign(offsets, offsets + n + 1);
    const uint64_t base = offsets[0];
    for (auto &v : c->t_off) v -= base;
    const size_t bytes = (size_t)c->t_off[(size_t)n];
    SO_CUDA(cudaMalloc((void **)&c->d_tres, bytes + 64));
    SO_CUDA(cudaMalloc((void **)&c->d_toff, ((size_t)n + 1) * 8));
    SO_CUDA(cudaMemcpyAsync(c->d_tres, residues + base, bytes, cudaMemcpyHostToDevice, c->stream));
    SO_CUDA(cudaMemcpyAsync(c->d_toff, c->t_off.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    SO_CUDA(cudaMalloc((void **)&c->d_tcls, bytes + 64));
    if ((rc = so::classify_residues(c, c->d_tres, c->d_tcls, bytes)) != SO_OK) return rc;
    if ((rc = so::build_ungap_targets(c)) != SO_OK) return rc;
    SO_CUDA(cudaStreamSynchronize(c->stream));
    c->stats.h2d_bytes += (i64)bytes + ((i64)n + 1) * 8;
    return SO_OK;
}

// host half of so_set_queries: seg masks and the S3 position order of a query set, computed without touching the device
// or the context's query state, so a caller can prepare block i + 1 on another thread while so_search runs on block i
struct so_qprep {
    int64_t n = 0;
    uint32_t max_qlen = 0;
    double t_host = 0;
    std::vector<uint64_t> q_off;
    std::vector<uint8_t> masked;
    std::vector<uint32_t> perm;
};

int so_queries_prepare(const so_ctx *c, const uint8_t *residues, const uint64_t *offsets, int64_t n, so_qprep **out) {
    if (!c || !out || !offsets || n < 0 || (!residues && n > 0 && offsets[n] > 0)) {
        set_error("so_queries_prepare: bad argument");
        return SO_EINVAL;
    }
    so_qprep *p = new so_qprep();
    int rc = check_offsets(offsets, n, p->max_qlen, "query");
    if (rc != SO_OK) {
        delete p;
        return rc;
    }
    Timer tm;
    p->n = n;
    p->q_off.assign(offsets, offsets + n + 1);
    const uint64_t base = offsets[0];
    for (auto &v : p->q_off) v -= base;
    const size_t bytes = (size_t)p->q_off[(size_t)n];
    p->masked.resize(bytes);
    std::vector<uint32_t> &perm = p->perm;
    perm.resize(bytes);
    const bool flt = c->P.flt;
    const int mink = c->P.mink;
    const uint8_t *src = residues + base;
    uint8_t *dst = p->masked.data();
    const uint64_t *qo = p->q_off.data();
    // H1 (seg) and the S3 position order: kscs = sliding BLOSUM62 self score over the shortest seed
    // span, positions sorted with the reference quicksort by -kscs (fsearch.py:2647-2656, 2668)
    so::parallel_for(n, [&](i64 q) {
        const i64 L = (i64)(qo[q + 1] - qo[q]);
        const uint8_t *s = src + qo[q];
        uint8_t *m = dst + qo[q];
        if (flt)
            so::seg_mask(s, L, m);
        else if (L > 0)
            memcpy(m, s, (size_t)L);
        const i64 P = L - mink + 1;
        if (P <= 0) return;
        std::vector<uint64_t> v((size_t)P);
        i64 sc = 0;
        for (int i = 0; i < mink; i++) sc += so::score_bytes(m[i], m[i]);
        for (i64 i = 0; i < P; i++) {
            if (i > 0) sc = sc - so::score_bytes(m[i - 1], m[i - 1]) + so::score_bytes(m[i - 1 + mink], m[i - 1 + mink]);
            // key = -kscs, biased into uint32 (order preserving)
            v[(size_t)i] = ((uint64_t)(uint32_t)(0x40000000 - sc) << 32) | (uint32_t)i;
        }
        so::qsort_prefix(v, P);
        uint32_t *pp = perm.data() + qo[q];
        for (i64 i = 0; i < P; i++) pp[i] = (uint32_t)v[(size_t)i];
    });
    p->t_host = tm.ms();
    *out = p;
    return SO_OK;
}

void so_qprep_free(so_qprep *p) { delete p; }

int so_set_queries(so_ctx *c, const uint8_t *residues, const uint64_t *offsets, int64_t n) {
    if (!c) {
        set_error("so_set_queries: bad argument");
        return SO_EINVAL;
    }
    so_qprep *p = nullptr;
    int rc = so_queries_prepare(c, residues, offsets, n, &p);
    if (rc != SO_OK) return rc;
    rc = so_set_queries_prepared(c, p);
    so_qprep_free(p);
    return rc;
}

// device half: the prepared set becomes the context's query set (H2D of the masked residues, offsets and position
// order, residue classes computed on the device); `p` is left empty
int so_set_queries_prepared(so_ctx *c, so_qprep *p) {
    if (!c || !p) {
        set_error("so_set_queries_prepared: bad argument");
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(c->device));
    int rc;
    Timer tm;
    const int64_t n = p->n;
    c->n_q = n;
    c->max_qlen = p->max_qlen;
    c->q_off.swap(p->q_off);
    c->q_masked.swap(p->masked);
    c->q_perm_host.swap(p->perm);
    const size_t bytes = (size_t)c->q_off[(size_t)n];
    std::vector<uint32_t> &perm = c->q_perm_host;
    uint8_t *dst = c->q_masked.data();
    const uint64_t *qo = c->q_off.data();
    const double t_host0 = p->t_host;
    const double t_host = t_host0;
    // device buffers grow only: so_set_queries is called once per query block on the end-to-end path
    if (bytes > c->q_cap_bytes || !c->d_qres) {
        if (c->d_qres) cudaFree(c->d_qres);
        if (c->d_perm) cudaFree(c->d_perm);
        if (c->d_qcls) cudaFree(c->d_qcls);
        c->d_qres = nullptr, c->d_perm = nullptr, c->d_qcls = nullptr;
        const size_t cap = bytes + bytes / 4 + 4096;
        SO_CUDA(cudaMalloc((void **)&c->d_qres, cap + 64));
        SO_CUDA(cudaMalloc((void **)&c->d_perm, (cap + 16) * 4));
        SO_CUDA(cudaMalloc((void **)&c->d_qcls, cap + 64));
        c->q_cap_bytes = cap;
    }
    if ((size_t)n > c->q_cap_seqs || !c->d_qoff) {
        if (c->d_qoff) cudaFree(c->d_qoff);
        c->d_qoff = nullptr;
        const size_t cap = (size_t)n + (size_t)n / 4 + 64;
        SO_CUDA(cudaMalloc((void **)&c->d_qoff, (cap + 1) * 8));
        c->q_cap_seqs = cap;
    }
    SO_CUDA(cudaMemcpyAsync(c->d_qres, dst, bytes, cudaMemcpyHostToDevice, c->stream));
    SO_CUDA(cudaMemcpyAsync(c->d_qoff, qo, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    SO_CUDA(cudaMemcpyAsync(c->d_perm, perm.data(), bytes * 4, cudaMemcpyHostToDevice, c->stream));
    if ((rc = so::classify_residues(c, c->d_qres, c->d_qcls, bytes)) != SO_OK) return rc;
    SO_CUDA(cudaStreamSynchronize(c->stream));
    c->stats.h2d_bytes += (i64)bytes * 5 + ((i64)n + 1) * 8;
    c->stats.ms_host += tm.ms();
    if (getenv("SO_PROFILE")) fprintf(stderr, "so_set_queries: host (seg + S3 order) %.1f ms, total %.1f ms\n", t_host, tm.ms());
    return SO_OK;
}

int so_index_build(so_ctx *c) {
    if (!c || !c->d_tres) {
        set_error("so_index_build: load targets first");
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(c->device));
    for (auto &ix : c->chunks) so::free_chunk_index(ix);
    c->chunks.clear();
    // makedb (fsearch.py:2283-2295): Start = 0 if -L == -1, End = N if -U == -1; chunks of -c sequences
    i64 Start = c->P.rst == -1 ? 0 : c->P.rst, End = c->P.red == -1 ? c->n_t : c->P.red;
    Start = std::min<i64>(std::max<i64>(Start, 0), c->n_t);
    End = std::min<i64>(std::max<i64>(End, 0), c->n_t);
    for (i64 s = Start; s < End; s += c->P.chunk) {
        so::ChunkIndex ix;
        ix.c0 = s;
        ix.c1 = std::min<i64>(s + c->P.chunk, End);
        int rc = so::build_chunk_index(c, ix);
        if (rc != SO_OK) {
            so::free_chunk_index(ix);
            return rc;
        }
        c->chunks.push_back(ix);
    }
    return SO_OK;
}

int64_t so_index_chunks(const so_ctx *c) { return c ? (int64_t)c->chunks.size() : 0; }

int so_index_info_get(const so_ctx *c, int64_t chunk, so_index_info *info) {
    if (!c || !info || chunk < 0 || chunk >= (int64_t)c->chunks.size()) {
        set_error("so_index_info_get: bad chunk");
        return SO_EINVAL;
    }
    const so::ChunkIndex &ix = c->chunks[(size_t)chunk];
    info->chunk_start = ix.c0, info->chunk_end = ix.c1, info->n_seeds = ix.n_seeds, info->n_buckets_used = ix.n_used;
    info->threshold = ix.threshold, info->build_ms = ix.build_ms;
    return SO_OK;
}

int so_index_export(so_ctx *c, int64_t chunk, uint32_t *start, uint32_t *locus) {
    if (!c || chunk < 0 || chunk >= (int64_t)c->chunks.size()) {
        set_error("so_index_export: bad chunk");
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(c->device));
    const so::ChunkIndex &ix = c->chunks[(size_t)chunk];
    if (start) SO_CUDA(cudaMemcpy(start, ix.d_start, ((size_t)c->P.nc + 1) * 4, cudaMemcpyDeviceToHost));
    if (locus && ix.n_seeds) SO_CUDA(cudaMemcpy(locus, ix.d_locus, (size_t)ix.n_seeds * 4, cudaMemcpyDeviceToHost));
    return SO_OK;
}

int so_candidates(so_ctx *c, int64_t chunk, int64_t q_begin, int64_t q_end, uint64_t **cand_offsets, so_cand **cands) {
    if (!c || !cand_offsets || !cands || chunk < 0 || chunk >= (int64_t)c->chunks.size() || q_begin < 0 ||
        q_end > c->n_q || q_begin > q_end) {
        set_error("so_candidates: bad argument");
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(c->device));
    so::PackedCands bc;
    int rc = so::upload_search_config(c);
    // the sync-free cell path when the configuration allows it (the path so_search takes), else the general path
    const so::ChunkIndex &cix = c->chunks[(size_t)chunk];
    const i64 nqc = q_end - q_begin;
    const size_t capc = (size_t)std::max<i64>(cix.c1 - cix.c0, 1);
    if (rc == SO_OK && nqc > 0 && (double)nqc * (double)capc * 8. < 3e9) {
        so::BlockStore &bs = c->bstore[0];
        bool fast = false, redo = false;
        rc = bs.prepare(c, nqc, capc, 1, c->stream);
        if (rc == SO_OK) rc = so::block_candidates_fast(c, q_begin, q_end, 0, bs, fast, (int)chunk);
        if (rc == SO_OK && fast) rc = so::enqueue_fast_ctl(c, 0);
        if (rc == SO_OK && fast && c->lane_wait(0) != cudaSuccess) {
            set_error("CUDA error in so_candidates");
            rc = SO_ENODEV;
        }
        if (rc == SO_OK && fast) rc = so::finish_fast_block(c, 0, redo);
        if (rc == SO_OK && fast && !redo) {
            std::vector<uint32_t> cnt((size_t)nqc);
            SO_CUDA(cudaMemcpy(cnt.data(), bs.count.p, (size_t)nqc * 4, cudaMemcpyDeviceToHost));
            uint64_t tot = 0;
            for (uint32_t v : cnt) tot += v;
            *cand_offsets = (uint64_t *)malloc(((size_t)nqc + 1) * 8);
            *cands = (so_cand *)malloc(std::max<size_t>(1, (size_t)tot) * sizeof(so_cand));
            if (!*cand_offsets || !*cands) {
                set_error("out of host memory");
                return SO_ENOMEM;
            }
            std::vector<uint64_t> tmp;
            uint64_t w = 0;
            (*cand_offsets)[0] = 0;
            for (i64 k = 0; k < nqc; k++) {
                tmp.resize(cnt[(size_t)k]);
                if (!tmp.empty())
                    SO_CUDA(cudaMemcpy(tmp.data(), bs.vals.p + (size_t)k * bs.capq, tmp.size() * 8, cudaMemcpyDeviceToHost));
                for (uint64_t v : tmp) (*cands)[w++] = so::unpack_cand(v);
                (*cand_offsets)[k + 1] = w;
            }
            so::merge_lane_stats(c);
            return SO_OK;
        }
    }
    if (rc == SO_OK) rc = so::chunk_candidates(c, c->chunks[(size_t)chunk], q_begin, q_end, bc);
    so::merge_lane_stats(c);
    if (rc != SO_OK) {
        bc.release();
        return rc;
    }
    *cand_offsets = (uint64_t *)malloc(bc.offsets.size() * 8);
    *cands = (so_cand *)malloc(std::max<size_t>(1, bc.n) * sizeof(so_cand));
    if (!*cand_offsets || !*cands) {
        bc.release();
        set_error("out of host memory");
        return SO_ENOMEM;
    }
    memcpy(*cand_offsets, bc.offsets.data(), bc.offsets.size() * 8);
    for (size_t k = 0; k < bc.n; k++) (*cands)[k] = so::unpack_cand(bc.vals[k]);
    bc.release();
    return SO_OK;
}

int so_align_batch(so_ctx *c, const so_pair *pairs, int64_t n, so_aln *out) {
    if (!c || n < 0 || (n > 0 && (!pairs || !out))) {
        set_error("so_align_batch: bad argument");
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(c->device));
    int rc = so::align_pairs(c, pairs, n, out);
    so::merge_align_stats(c);
    return rc;
}

int so_stats_get(const so_ctx *c, so_stats *s) {
    if (!c || !s) return SO_EINVAL;
    *s = c->stats;
    return SO_OK;
}

int so_stats_reset(so_ctx *c) {
    if (!c) return SO_EINVAL;
    memset(&c->stats, 0, sizeof c->stats);
    return SO_OK;
}

// test / tuning hook (not part of the reference surface): fixed seeding sub-block size
int so_set_sub_block(so_ctx *c, int64_t n) {
    if (!c) return SO_EINVAL;
    c->sub_block = n;
    return SO_OK;
}

// test / tuning hook: number of candidate-production lanes used by so_search (1 .. 4; default 2; 0 = measurement
// mode).  bench.py uses 0 to time the kernels of a step without another stream sharing the GPU.
int so_set_lanes(so_ctx *c, int n) {
    if (!c || n < 0 || n > so_ctx::kMaxLanes) return SO_EINVAL;
    c->n_lanes = n;
    return SO_OK;
}

int so_search(so_ctx *c, int64_t q_begin, int64_t q_end, so_hit **rows_out, int64_t *n_rows) {
    if (!c || !rows_out || !n_rows || q_begin < 0 || q_end > c->n_q || q_begin > q_end) {
        set_error("so_search: bad argument");
        return SO_EINVAL;
    }
    if (!c->d_qres || !c->d_tres) {
        set_error("so_search: load targets and queries first");
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(c->device));
    Timer total;
    const Params &P = c->P;
    const i64 D = c->n_t;
    const double max_miss = std::max(P.max_miss, 1e-3);                      // fsearch.py:2970
    const i64 vmax = (i64)std::max(100., std::max((double)(P.v + 100), (double)P.v * 1.1));  // fsearch.py:3059
    std::vector<so_hit> all_rows;
    const size_t nch = c->chunks.size();
    // query block size: the candidates of a block live in per-query device lists (select.cu)
    i64 QB = 592;  // 4 x 148: the kernels that take one CTA per query (cell passes, candidate sort, selection) run whole waves
    if (const char *e = getenv("SO_QUERY_BLOCK")) QB = std::max<i64>(16, atoll(e));  // tuning hook
    const int nprod = std::max(1, std::min<int>(c->n_lanes, so_ctx::kMaxLanes));
    c->shared_stream = true;  // SO_SHARED_STREAM=0: one stream per lane (kernels of two query blocks share the SMs)
    if (const char *e = getenv("SO_SHARED_STREAM")) c->shared_stream = atoi(e) != 0;
    const int kSlots = 2 * nprod;  // a lane produces block k + nprod while the worker still orders block k
    // at most one candidate per (query, target): fixed per-query capacity of the device lists
    i64 capq = 0;
    for (const auto &ix : c->chunks) capq += ix.c1 - ix.c0;
    capq = std::max<i64>(capq, 1);
    // keep one lane's lists within ~6 GB (512 queries against up to 1.4 M targets)
    QB = std::max<i64>(16, std::min<i64>(QB, (i64)(6000000000ll / (capq * 8))));
    if (c->cand_pool.size() < (size_t)so_ctx::kMaxLanes) c->cand_pool.resize((size_t)so_ctx::kMaxLanes);
    {
        int rc = so::upload_search_config(c);
        if (rc != SO_OK) return rc;
    }
    // Pipeline: two producer threads (one stream + scratch set each) produce the candidates of alternating
    // query blocks on the GPU, so one lane's host synchronisations and D2H copies overlap the other lane's
    // kernels; a worker thread sorts / selects / aligns (own stream) / filters the blocks in block order.
    struct Job {
        i64 b0, b1;
        int slot;
    };
    std::mutex mu;
    std::condition_variable cv;
    std::map<int, Job> jobs;  // by block number
    int next_blk = 0;         // next block the worker consumes
    // n_lanes == 0: measurement mode: one lane, and candidate production and alignment rounds never share the GPU, so
    // the CUDA-event durations of the kernels are not inflated by kernels of the other stage
    const bool serial = c->n_lanes == 0;
    std::mutex gpu_mu;
    int producers_left = nprod;
    bool abort_all = false;
    bool producer_done = false, slot_busy[so_ctx::kMaxSlots] = {};
    int worker_rc = SO_OK;
    std::string worker_err;
    so_stats wstats;
    memset(&wstats, 0, sizeof wstats);

    // queries whose candidates are selected but not aligned yet: alignment rounds run over several blocks
    // at once so one launch carries enough alignments to fill the GPU
    std::vector<QueryState> pending;

    const i64 selcap = std::max<i64>(1, std::min<i64>(vmax, capq));
    auto order_block = [&](i64 b0, i64 b1, int slot, const std::function<void()> &release_slot) -> int {
        const i64 nq = b1 - b0;
        std::vector<QueryState> qs((size_t)nq);
        // PASS 2 head (fsearch.py:3039-3059): the merge, the quicksort by -score and the [:vmax] cut ran on the
        // device (select.cu); h_sel holds the selected candidates of every query in sorted order
        Timer th;
        const uint64_t *hs = c->h_sel[slot];
        const uint32_t *hn = c->h_sel_n[slot];
        so::parallel_for(nq, [&](i64 k) {
            QueryState &s = qs[(size_t)k];
            s.qord = b0 + k;
            const i64 n = (i64)hn[k];
            s.limit = std::min<i64>(vmax, n);
            double mm = (double)n * max_miss + 1;
            mm = std::max(mm, 100. / mm);
            mm = std::min(std::max(mm, 10.), 120.);
            s.mmiss = mm;
            s.done = s.limit == 0;
            s.sel.resize((size_t)s.limit);
            const i64 li = (i64)(c->q_off[(size_t)s.qord + 1] - c->q_off[(size_t)s.qord]);
            for (i64 i = 0; i < s.limit; i++) {
                const so_cand cd = so::unpack_cand(hs[(size_t)k * (size_t)selcap + (size_t)i]);
                s.sel[(size_t)i] = cd;
                const i64 lj = (i64)(c->t_off[(size_t)cd.target + 1] - c->t_off[(size_t)cd.target]);
                // (random diagonals rarely chain above ~60: e^(-0.267 x) tail; a homolog's best diagonal usually does,
                // even when indels keep its ungapped score below the e-value cut)
                if (cd.score >= 60 || so::bit2e(D, li, lj, so::score2bit((i64)cd.score)) <= P.expect) s.lead = i + 1;
            }
        });
        wstats.ms_host += th.ms();
        c->prof.order_ms += th.ms();
        release_slot();  // the pinned selection of this block is no longer needed
        for (auto &q : qs) pending.push_back(std::move(q));
        return SO_OK;
    };

    // ---- rolling alignment pool.  `pending` holds the queries whose candidates are selected, in query order; the
    // worker issues ONE alignment launch per iteration that carries the next round of EVERY unfinished query (first
    // rounds of newly arrived blocks together with later rounds of the stragglers), replays the sequential stop rule
    // on the host and retires finished queries from the front in query order.
    std::deque<QueryState> pool;
    struct Req {
        int q;      // index into the pool
        int first;  // first pair of this candidate
        int count;  // pairs (tiles) of this candidate
        int cand;   // index into sel
    };
    std::vector<so_pair> pairs;
    std::vector<so_aln> alns;
    std::vector<Req> reqs;

    auto finalize_front = [&]() {
        // final per-query order: qsort_u by -bit, first v rows (fsearch.py:3108-3110); queries leave in query order
        Timer tf;
        size_t nfin = 0;
        while (nfin < pool.size() && pool[nfin].done) nfin++;
        if (nfin == 0) return;
        so::parallel_for((i64)nfin, [&](i64 k) {
            QueryState &s = pool[(size_t)k];
            const i64 n = (i64)s.rows.size();
            if (n == 0) return;
            std::vector<uint64_t> v((size_t)n);
            for (i64 i = 0; i < n; i++)
                v[(size_t)i] = ((uint64_t)(uint32_t)(0x7fffffff - (int32_t)s.rows[(size_t)i].bit) << 32) | (uint32_t)i;
            so::qsort_prefix(v, n);
            std::vector<so_hit> sorted;
            const i64 lim = std::min<i64>(std::max<i64>(0, P.v), n);
            sorted.reserve((size_t)lim);
            for (i64 i = 0; i < lim; i++) sorted.push_back(s.rows[(size_t)(uint32_t)v[(size_t)i]]);
            s.rows.swap(sorted);
        });
        for (size_t k = 0; k < nfin; k++) {
            all_rows.insert(all_rows.end(), pool[k].rows.begin(), pool[k].rows.end());
            wstats.queries++;
        }
        pool.erase(pool.begin(), pool.begin() + (std::ptrdiff_t)nfin);
        wstats.ms_host += tf.ms();
        c->prof.final_ms += tf.ms();
    };

    // one round over every unfinished query of the pool; returns SO_OK (and does nothing when all are finished)
    auto align_round = [&]() -> int {
        std::unique_lock<std::mutex> gl(gpu_mu, std::defer_lock);
        if (serial) gl.lock();
        Timer trd;
        pairs.clear();
        reqs.clear();
        const i64 nq = (i64)pool.size();
        for (i64 k = 0; k < nq; k++) {
            QueryState &s = pool[(size_t)k];
            if (s.done) continue;
            const i64 qi_ord = s.qord;
            const i64 li = (i64)(c->q_off[(size_t)qi_ord + 1] - c->q_off[(size_t)qi_ord]);
            // first round: the candidates the stop rule is certain to reach (see kRoundMax); a query that needs more
            // rounds (its hits go on beyond the ungapped prediction) then takes 64, 128, 256, ... candidates per round,
            // so every query finishes within a few launches (a round trip costs more than the alignments it saves)
            const i64 left = std::max<i64>((i64)std::ceil(s.mmiss - (double)s.unmch), 1);
            const i64 want = s.next == 0   ? s.lead + (i64)std::ceil(s.mmiss)
                             : s.rounds == 1 ? left
                                             : std::max<i64>(left, (i64)64 << std::min(s.rounds - 2, 6));
            s.rounds++;
            const i64 hi = std::min<i64>(s.limit, s.next + std::min<i64>(want, kRoundMax));
            s.last_hits = 0;
            for (i64 h = s.next; h < hi; h++) {
                const int ci = (int)h;
                const so_cand &cd = s.sel[(size_t)ci];
                const i64 lj = (i64)(c->t_off[(size_t)cd.target + 1] - c->t_off[(size_t)cd.target]);
                Req r;
                r.q = (int)k, r.first = (int)pairs.size(), r.cand = ci, r.count = 0;
                if (li < 4096 && lj < 4096) {
                    so_pair p;
                    p.query = qi_ord, p.target = cd.target, p.q_off = 0, p.q_len = (int32_t)li, p.t_off = 0;
                    p.t_len = (int32_t)lj, p.qst = (int32_t)cd.qi, p.sst = (int32_t)cd.qj;
                    pairs.push_back(p);
                    r.count = 1;
                } else {
                    // kswat_st_long (fsearch.py:1480-1498): 4096-tiles down the diagonal; a tile
                    // whose target slice is empty is skipped (undefined in the reference)
                    i64 j = cd.qj;
                    for (i64 i = cd.qi; i < li; i += 4096, j += 4096) {
                        if (j >= lj) continue;
                        so_pair p;
                        p.query = qi_ord, p.target = cd.target, p.q_off = (int32_t)i;
                        p.q_len = (int32_t)std::min<i64>(4096, li - i), p.t_off = (int32_t)j;
                        p.t_len = (int32_t)std::min<i64>(4096, lj - j), p.qst = 0, p.sst = 0;
                        pairs.push_back(p);
                        r.count++;
                    }
                }
                reqs.push_back(r);
            }
        }
        if (reqs.empty()) return SO_OK;
        alns.resize(pairs.size());
        Timer ta;
        int rc = so::align_pairs(c, pairs.data(), (i64)pairs.size(), alns.data());
        if (rc != SO_OK) return rc;
        c->prof.align_ms += ta.ms();
        Timer tr;
        // replay (fsearch.py:3062-3106)
        size_t r = 0;
        while (r < reqs.size()) {
            const int k = reqs[r].q;
            QueryState &s = pool[(size_t)k];
            const i64 qi_ord = s.qord;
            const i64 li = (i64)(c->q_off[(size_t)qi_ord + 1] - c->q_off[(size_t)qi_ord]);
            for (; r < reqs.size() && reqs[r].q == k; r++) {
                if (s.done) continue;  // computed but never reached by the sequential rule (counted as wasted)
                wstats.alignments_used += reqs[r].count;
                const Req &rq = reqs[r];
                const so_cand &cd = s.sel[(size_t)rq.cand];
                const i64 lj = (i64)(c->t_off[(size_t)cd.target + 1] - c->t_off[(size_t)cd.target]);
                bool any = false;
                for (int t = 0; t < rq.count; t++) {
                    const so_aln &a = alns[(size_t)(rq.first + t)];
                    const so_pair &p = pairs[(size_t)(rq.first + t)];
                    const i64 bit = so::score2bit(a.raw_score);
                    const double e = so::bit2e(D, li, lj, bit);
                    if (e <= P.expect) {
                        so_hit hrow;
                        memset(&hrow, 0, sizeof hrow);
                        hrow.query = qi_ord, hrow.target = cd.target, hrow.qlen = (int32_t)li, hrow.tlen = (int32_t)lj;
                        hrow.aln_len = a.aln_len, hrow.mismatch = a.mismatch, hrow.gaps = a.gaps;
                        hrow.qst = a.qst + p.q_off + 1, hrow.qed = a.qed + p.q_off;
                        hrow.sst = a.sst + p.t_off + 1, hrow.sed = a.sed + p.t_off;
                        hrow.raw_score = a.raw_score, hrow.n_ident = a.n_ident, hrow.bit = bit;
                        hrow.identity = a.aln_len ? (double)a.n_ident * (100. / (double)a.aln_len) : NAN;
                        hrow.evalue = e;
                        s.rows.push_back(hrow);
                        any = true;
                        s.bv++;
                    }
                }
                if (any)
                    s.unmch = 0, s.last_hits++;
                else
                    s.unmch++;
                s.next++;
                if ((double)s.unmch >= s.mmiss || (double)s.bv >= (double)P.v + s.mmiss) s.done = true;
            }
            if (s.next >= s.limit) s.done = true;
        }
        wstats.ms_host += tr.ms();
        c->prof.replay_ms += tr.ms();
        c->prof.rounds_ms += trd.ms();
        return SO_OK;
    };

    size_t kMinFresh = 1;  // new queries that trigger an alignment round (the worker takes every block that arrived
                           // while it was busy with the previous round, so rounds grow by themselves under load)
    if (const char *e = getenv("SO_ALIGN_BATCH")) kMinFresh = (size_t)std::max(1, atoi(e));  // tuning hook
    std::thread worker([&]() {
        cudaSetDevice(c->device);
        size_t fresh = 0;
        for (;;) {
            // absorb blocks (in block order) until a round is due
            for (;;) {
                Job j;
                bool have = false;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    // a round is worth launching once enough new queries are in (first rounds carry most of the
                    // alignments) or nothing more will come; until then wait for the next block
                    const bool go = fresh >= kMinFresh || (producer_done && !jobs.count(next_blk));
                    if (!go) cv.wait(lk, [&] { return jobs.count(next_blk) || producer_done; });
                    if (jobs.count(next_blk)) {
                        j = jobs[next_blk];
                        jobs.erase(next_blk);
                        next_blk++;
                        have = true;
                    }
                }
                if (!have) break;
                bool released = false;
                auto release = [&]() {
                    if (released) return;
                    released = true;
                    std::lock_guard<std::mutex> lk(mu);
                    slot_busy[j.slot] = false;
                    cv.notify_all();
                };
                if (worker_rc == SO_OK) {
                    int rc = order_block(j.b0, j.b1, j.slot, release);
                    if (rc != SO_OK) {
                        std::lock_guard<std::mutex> lk(mu);
                        worker_rc = rc;
                        worker_err = so::get_error();
                        cv.notify_all();
                    }
                }
                release();
                fresh += (size_t)(j.b1 - j.b0);
            }
            fresh = 0;
            for (auto &q : pending) pool.push_back(std::move(q));
            pending.clear();
            bool active = false;
            for (const auto &q : pool)
                if (!q.done) {
                    active = true;
                    break;
                }
            if (active && worker_rc == SO_OK) {
                int rc = align_round();
                if (rc != SO_OK) {
                    std::lock_guard<std::mutex> lk(mu);
                    worker_rc = rc;
                    worker_err = so::get_error();
                    cv.notify_all();
                }
            }
            if (worker_rc != SO_OK)
                for (auto &q : pool) q.done = true;  // drain: nothing more is aligned after an error
            finalize_front();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (pool.empty() && producer_done && !jobs.count(next_blk)) break;
            }
        }
    });
    int prod_rcs[so_ctx::kMaxLanes] = {};
    std::string prod_errs[so_ctx::kMaxLanes];
    auto produce = [&](int pid) {
        cudaSetDevice(c->device);
        int rc = SO_OK;
        for (int blk = pid;; blk += nprod) {
            const i64 b0 = q_begin + (i64)blk * QB;
            if (b0 >= q_end || rc != SO_OK) break;
            const i64 b1 = std::min<i64>(q_end, b0 + QB);
            const int slot = blk % kSlots;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !slot_busy[slot] || abort_all || worker_rc != SO_OK; });
                if (worker_rc != SO_OK || abort_all) break;
                slot_busy[slot] = true;
            }
            // PASS 1 (fsearch.py:2990-3016): candidates of every chunk, appended on the device to the block's
            // per-query lists in chunk order; then the device selection and the copy of the selected candidates
            Timer tc;
            std::unique_lock<std::mutex> gl(gpu_mu, std::defer_lock);
            if (serial) gl.lock();
            cudaStream_t st = c->lane_stream(pid);
            so::BlockStore &bs = c->bstore[pid];
            const i64 nqb = b1 - b0;
            rc = bs.prepare(c, nqb, (size_t)capq, (int)selcap, st);
            bool fast = false;
            if (rc == SO_OK) rc = so::block_candidates_fast(c, b0, b1, pid, bs, fast);
            auto general_path = [&]() {
                for (size_t ch = 0; ch < nch && rc == SO_OK; ch++)
                    rc = so::chunk_candidates(c, c->chunks[ch], b0, b1, c->cand_pool[(size_t)pid], pid, &bs, 0);
            };
            if (!fast) general_path();
            if (rc == SO_OK) rc = bs.select(st);
            if (rc == SO_OK) {
                const size_t need_sel = (size_t)nqb * (size_t)selcap;
                if (need_sel > c->h_sel_cap[slot]) {
                    if (c->h_sel[slot]) cudaFreeHost(c->h_sel[slot]);
                    if (c->h_sel_n[slot]) cudaFreeHost(c->h_sel_n[slot]);
                    c->h_sel[slot] = nullptr, c->h_sel_n[slot] = nullptr, c->h_sel_cap[slot] = 0;
                    const size_t cap = (size_t)std::max<i64>(QB, nqb) * (size_t)selcap;
                    if (cudaMallocHost((void **)&c->h_sel[slot], cap * 8) != cudaSuccess ||
                        cudaMallocHost((void **)&c->h_sel_n[slot], ((size_t)std::max<i64>(QB, nqb) + 8) * 4) != cudaSuccess) {
                        so::set_error("cudaMallocHost of the selection buffers failed");
                        rc = SO_ENOMEM;
                    } else
                        c->h_sel_cap[slot] = cap;
                }
            }
            if (rc == SO_OK) {
                Timer td;
                uint32_t flags[2] = {0, 0};
                cudaMemcpyAsync(c->h_sel[slot], bs.sel.p, (size_t)nqb * (size_t)selcap * 8, cudaMemcpyDeviceToHost, st);
                cudaMemcpyAsync(c->h_sel_n[slot], bs.sel_n.p, (size_t)nqb * 4, cudaMemcpyDeviceToHost, st);
                cudaMemcpyAsync(flags, bs.count.p + nqb, 8, cudaMemcpyDeviceToHost, st);
                if (fast) so::enqueue_fast_ctl(c, pid);
                cudaError_t e = c->lane_wait(pid);
                if (e != cudaSuccess) {
                    so::set_error("CUDA error in the candidate selection: %s", cudaGetErrorString(e));
                    rc = SO_ENODEV;
                } else if (flags[1]) {
                    so::set_error("device candidate selection failed (flag %u)", flags[1]);
                    rc = SO_ELIMIT;
                }
                c->stats_lane[pid].d2h_bytes += (i64)nqb * selcap * 8 + nqb * 4 + 8;
                c->stats_lane[pid].kernel_launches += 1;
                c->d2h_ms_lane[pid] += td.ms();
                if (rc == SO_OK && fast) {
                    bool redo = false;
                    rc = so::finish_fast_block(c, pid, redo);
                    if (rc == SO_OK && redo) {
                        // a cell too large for the shared-memory sort: the whole block goes through the general path
                        fast = false;
                        rc = bs.prepare(c, nqb, (size_t)capq, (int)selcap, st);
                        general_path();
                        if (rc == SO_OK) rc = bs.select(st);
                        if (rc == SO_OK) {
                            cudaMemcpyAsync(c->h_sel[slot], bs.sel.p, (size_t)nqb * (size_t)selcap * 8, cudaMemcpyDeviceToHost, st);
                            cudaMemcpyAsync(c->h_sel_n[slot], bs.sel_n.p, (size_t)nqb * 4, cudaMemcpyDeviceToHost, st);
                            cudaMemcpyAsync(flags, bs.count.p + nqb, 8, cudaMemcpyDeviceToHost, st);
                            e = c->lane_wait(pid);
                            if (e != cudaSuccess) {
                                so::set_error("CUDA error in the candidate selection: %s", cudaGetErrorString(e));
                                rc = SO_ENODEV;
                            } else if (flags[1]) {
                                so::set_error("device candidate selection failed (flag %u)", flags[1]);
                                rc = SO_ELIMIT;
                            }
                            c->stats_lane[pid].redo_blocks += 1;
                        }
                    }
                }
            }
            std::lock_guard<std::mutex> lk(mu);
            c->prof.cand_ms += tc.ms();
            if (rc == SO_OK)
                jobs[blk] = Job{b0, b1, slot};
            else
                slot_busy[slot] = false;
            cv.notify_all();
        }
        if (rc != SO_OK) prod_errs[pid] = so::get_error();
        prod_rcs[pid] = rc;
        std::lock_guard<std::mutex> lk(mu);
        if (rc != SO_OK) producer_done = abort_all = true;  // a missing block must not stall the worker
        if (--producers_left == 0) producer_done = true;
        cv.notify_all();
    };
    std::vector<std::thread> producers;
    for (int pid = 1; pid < nprod; pid++) producers.emplace_back(produce, pid);
    produce(0);
    for (auto &t : producers) t.join();
    int prod_rc = SO_OK;
    std::string prod_err;
    for (int pid = 0; pid < nprod && prod_rc == SO_OK; pid++)
        if (prod_rcs[pid] != SO_OK) prod_rc = prod_rcs[pid], prod_err = prod_errs[pid];
    worker.join();
    c->stats.ms_host += wstats.ms_host;
    c->stats.queries += wstats.queries;
    c->stats.alignments_used += wstats.alignments_used;
    so::merge_align_stats(c);
    so::merge_lane_stats(c);
    if (prod_rc != SO_OK) {
        set_error("%s", prod_err.c_str());
        return prod_rc;
    }
    if (worker_rc != SO_OK) {
        set_error("%s", worker_err.c_str());
        return worker_rc;
    }
    c->stats.rows += (i64)all_rows.size();
    *n_rows = (int64_t)all_rows.size();
    *rows_out = (so_hit *)malloc(std::max<size_t>(1, all_rows.size()) * sizeof(so_hit));
    if (!*rows_out) {
        set_error("out of host memory");
        return SO_ENOMEM;
    }
    if (!all_rows.empty()) memcpy(*rows_out, all_rows.data(), all_rows.size() * sizeof(so_hit));
    c->stats.ms_total += total.ms();
    c->prof.total_ms += total.ms();
    if (getenv("SO_PROFILE")) {
        const so::HostProfile &p = c->prof;
        fprintf(stderr, "so_search profile (cumulative ms): total %.1f | candidates %.1f (d2h %.1f) | order %.1f | rounds %.1f (align %.1f replay %.1f) | final %.1f\n",
                p.total_ms, p.cand_ms, p.d2h_ms, p.order_ms, p.rounds_ms, p.align_ms, p.replay_ms, p.final_ms);
    }
    return SO_OK;
}

int so_write_rows(const so_hit *rows, int64_t n, const so_fasta *queries, const so_fasta *targets, const char *path,
                  int append) {
    if (n < 0 || (n > 0 && !rows) || !queries || !targets || !path) {
        set_error("so_write_rows: bad argument");
        return SO_EINVAL;
    }
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) {
        set_error("cannot open %s for writing", path);
        return SO_EIO;
    }
    // rows are formatted by all host threads in contiguous slices and written in order
    const int64_t kSlice = 2048;
    const int64_t nslices = (n + kSlice - 1) / kSlice;
    std::vector<std::string> bufs((size_t)nslices);
    std::atomic<int> bad(0);
    so::parallel_for(nslices, [&](so::i64 sl) {
        std::string &buf = bufs[(size_t)sl];
        const int64_t k0 = sl * kSlice, k1 = std::min<int64_t>(n, (sl + 1) * kSlice);
        // one pass for the size (ids and headers vary), one pass writing into the buffer: no per-field allocation
        size_t cap = 0;
        for (int64_t k = k0; k < k1; k++) {
            const char *hq, *ht;
            int64_t lq, lt;
            if (so_fasta_header(queries, rows[k].query, &hq, &lq) != SO_OK || so_fasta_header(targets, rows[k].target, &ht, &lt) != SO_OK) {
                bad = 1;
                return;
            }
            cap += (size_t)lq + 2 * (size_t)lt + 13 * 24 + 900;
        }
        buf.resize(cap);
        char *p = &buf[0];
        auto put_int = [&](long long v) {
            char t[24];
            int m = 0;
            unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
            do t[m++] = (char)('0' + u % 10), u /= 10;
            while (u);
            if (v < 0) *p++ = '-';
            while (m) *p++ = t[--m];
        };
        for (int64_t k = k0; k < k1; k++) {
            const so_hit &r = rows[k];
            const char *hq, *ht;
            int64_t lq, lt;
            so_fasta_header(queries, r.query, &hq, &lq);
            so_fasta_header(targets, r.target, &ht, &lt);
            // ids = header up to the first space (fsearch.py:3066)
            int64_t iq = 0, it = 0;
            while (iq < lq && hq[iq] != ' ') iq++;
            while (it < lt && ht[it] != ' ') it++;
            memcpy(p, hq, (size_t)iq), p += iq, *p++ = '\t';
            memcpy(p, ht, (size_t)it), p += it, *p++ = '\t';
            p += so::fmt_identity_to(p, r.identity), *p++ = '\t';
            put_int(r.aln_len), *p++ = '\t';
            put_int(r.mismatch), *p++ = '\t';
            put_int(r.gaps), *p++ = '\t';
            put_int(r.qst), *p++ = '\t';
            put_int(r.qed), *p++ = '\t';
            put_int(r.sst), *p++ = '\t';
            put_int(r.sed), *p++ = '\t';
            p += so::f2s_to(p, r.evalue), *p++ = '\t';
            put_int((long long)r.bit), *p++ = '\t';
            put_int(r.qlen), *p++ = '\t';
            put_int(r.tlen), *p++ = '\t';
            put_int((long long)r.query), *p++ = '\t';
            memcpy(p, ht, (size_t)lt), p += lt, *p++ = '\n';
        }
        buf.resize((size_t)(p - &buf[0]));
    });
    if (bad) {
        fclose(f);
        return SO_EINVAL;
    }
    for (const std::string &buf : bufs)
        if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
            fclose(f);
            set_error("short write on %s", path);
            return SO_EIO;
        }
    fclose(f);
    return SO_OK;
}
}

// K1-K6: index build, seed lookup, diagonal grouping and chained X-drop scoring on sm_100a.
//
// Index of one target chunk (replaces Fasta.build_msav, lib/fsearch.py:2208-2280)
//   k_target_seeds   one thread per (residue offset, alphabet, pattern): whole-span x/X veto,
//                    FNV-1a-32 over the care positions + the pattern ordinal, % NC, cross-pattern
//                    dedup (spseeds_fnv, fsearch.py:519-556).  The key is written at the MIRRORED
//                    insertion ordinal, so a STABLE radix sort on the bucket alone reproduces the
//                    reference's reverse-insertion order inside a bucket (fsearch.py:2258-2266).
//   radix sort       cub::DeviceRadixSort (library) on ceil(log2(NC+1)) bits
//   k_bucket_starts  dense start[NC+1] table (start[b] = #entries with bucket < b) = the
//                    reference's `start` list after the fill pass
//   k_locus_decode   (sequence+1, position) of every locus entry with the reference's bisect quirk
//                    (fsearch.py:2638-2642 + 134-153: largest idx with soas[idx] < x, so position 0
//                    of sequence k is attributed to sequence k-1 at position len(k-1))
// Search of a query block against one chunk (replaces Fasta.find_msav_m, fsearch.py:2645-2724)
//   k_query_seeds    seeds of every query position -> bucket range [st, ed) with the reference's
//                    clipping (get_bin_mem, fsearch.py:2530-2541: L = len(locus)-1; bucket NC-1 empty)
//   k_filter         one warp per query: walk the positions in the reference's quicksort order of
//                    -kscs (order computed once on the host), keep while the running bucket total
//                    is <= threshold*len (fsearch.py:2667-2677)
//   exclusive scan   cub::DeviceScan (library)
//   grouping         one pattern + one alphabet: k_cell_pass<false/true> partition the hits into (query, target)
//                    cells, k_cell_small / k_cell_warp / k_cell_block order every cell by (diagonal, qst) -> the
//                    sorted 64-bit keys (query | target | diagonal | qst).  Otherwise: k_expand (key + the hit's
//                    ordinal in the reference's scan order = its "first appearance" rank) and
//                    cub::DeviceRadixSort (library) on the packed key
//   head flags       cub::DeviceScan over "first hit of a (query, target, diagonal) group" -> group index
//   k_group_desc     group heads, head keys, X-drop descriptors (two-seed chains carried inline)
//   k_xdrop          chained X-drop-30 extension of every group (ungap / get_ungap_scores,
//                    fsearch.py:2454-2509); k_group_ungap_generic for sequences >= 8192 residues
//   compaction       cub::DeviceSelect: groups with score >= 25, in order
//   k_pair_select    per (query, target): best diagonal with first-appearance tie break, candidate order
//                    key = rank of the first passing diagonal (fsearch.py:2696-2719)
//   radix sort       candidates by (query, first-passing rank) -> reference candidate order
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>

#include "context.h"

namespace so {

struct SeedCfg {
    int S, A, step;
    uint32_t nc;
    int k[16];
    uint32_t care[16];
};
__constant__ SeedCfg c_cfg;
__constant__ uint16_t c_alpha[4 * 256];
__constant__ int8_t c_score2[kClasses * kClasses];
__device__ int g_xdrop_tab[26 * 32];  // k_xdrop's step table, one entry per (target class, query class): upload_cfg
__constant__ uint8_t c_code2[256];

static int upload_cfg(const Params &P) {
    SeedCfg h;
    memset(&h, 0, sizeof h);
    h.S = (int)P.patterns.size();
    h.A = (int)P.alphabets.size();
    h.step = P.step;
    h.nc = P.nc;
    for (int s = 0; s < h.S; s++) {
        h.k[s] = (int)P.patterns[(size_t)s].size();
        uint32_t m = 0;
        for (int j = 0; j < h.k[s]; j++)
            if (P.patterns[(size_t)s][(size_t)j] != '0') m |= 1u << j;
        h.care[s] = m;
    }
    uint16_t al[4 * 256];
    memset(al, 0, sizeof al);
    for (int a = 0; a < h.A; a++)
        for (int i = 0; i < 256; i++) al[a * 256 + i] = P.alphabets[(size_t)a][(size_t)i];
    int8_t tbl[kClasses * kClasses];
    uint8_t code[256];
    make_score_table(tbl);
    make_code_table(code);
    SO_CUDA(cudaMemcpyToSymbol(c_cfg, &h, sizeof h));
    SO_CUDA(cudaMemcpyToSymbol(c_alpha, al, sizeof al));
    SO_CUDA(cudaMemcpyToSymbol(c_score2, tbl, sizeof tbl));
    SO_CUDA(cudaMemcpyToSymbol(c_code2, code, sizeof code));
    // k_xdrop's packed step entries (see the kernel's header comment): 1 - score * 8192, 2^29 for a terminator (class 24) or a
    // class outside the table, 0 for the skip class (25).  Built once here: every CTA then copies 832 words 32 times
    // instead of deriving each lane-private entry from a divergent constant-memory read.
    int xt[26 * 32];
    for (int e = 0; e < 26 * 32; e++) {
        const int ct = e >> 5, cq = e & 31;
        xt[e] = ct == 25 ? 0 : (ct < 24 && cq < 24) ? 1 - (int)tbl[cq * kClasses + ct] * 8192 : 1 << 29;
    }
    SO_CUDA(cudaMemcpyToSymbol(g_xdrop_tab, xt, sizeof xt));
    return SO_OK;
}

// hash of the seed of pattern s / alphabet a starting at seq[i]; false if the span leaves the
// sequence or holds x/X
__device__ __forceinline__ bool seed_hash(const uint8_t *seq, int L, int i, int a, int s, const uint16_t *s_alpha,
                                          uint32_t &bucket) {
    const int k = c_cfg.k[s];
    if (i + k > L) return false;
    const uint32_t care = c_cfg.care[s];
    uint32_t n = 0x811c9dc5u;
    for (int j = 0; j < k; j++) {
        const uint32_t ch = seq[i + j];
        if (ch == 'x' || ch == 'X') return false;
        if ((care >> j) & 1u) n = (n ^ (uint32_t)s_alpha[a * 256 + ch]) * 0x01000193u;
    }
    n = (n ^ (uint32_t)s) * 0x01000193u;
    bucket = n % c_cfg.nc;
    return true;
}

// emitted by spseeds_fnv? (valid and not a repeat of an earlier pattern's (bucket, i) in this alphabet)
__device__ __forceinline__ bool seed_emitted(const uint8_t *seq, int L, int i, int a, int s,
                                             const uint16_t *s_alpha, uint32_t &bucket) {
    if (!seed_hash(seq, L, i, a, s, s_alpha, bucket)) return false;
    for (int s2 = 0; s2 < s; s2++) {
        uint32_t b2;
        if (seed_hash(seq, L, i, a, s2, s_alpha, b2) && b2 == bucket) return false;
    }
    return true;
}

__device__ __forceinline__ void load_alpha(uint16_t *s_alpha) {
    for (int k = threadIdx.x; k < c_cfg.A * 256; k += blockDim.x) s_alpha[k] = c_alpha[k];
    __syncthreads();
}

// largest j with a[j] <= x  (a[0] = 0 <= x always)
__device__ __forceinline__ int seq_of(const uint32_t *__restrict__ a, int n, uint32_t x) {
    int lo = 0, hi = n;  // a has n+1 entries; answer in [0, n-1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= x)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------ index
__global__ void __launch_bounds__(256) k_target_seeds(const uint8_t *__restrict__ res, const uint32_t *__restrict__ soas,
                                                      int M, uint32_t total, uint32_t nslots,
                                                      uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    __shared__ uint16_t s_alpha[4 * 256];
    load_alpha(s_alpha);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nslots) return;
    const int AS = c_cfg.A * c_cfg.S;
    const uint32_t x = tid / AS;
    const int as = (int)(tid % AS);
    const int a = as / c_cfg.S, s = as % c_cfg.S;
    const int j = seq_of(soas, M, x);
    const uint32_t base = soas[j];
    const int L = (int)(soas[j + 1] - base);
    const int i = (int)(x - base);
    const uint32_t ord = base * AS + (uint32_t)as * (uint32_t)L + (uint32_t)i;  // insertion ordinal
    uint32_t bucket = c_cfg.nc;
    if (i % c_cfg.step == 0) {
        uint32_t b;
        if (seed_emitted(res + base, L, i, a, s, s_alpha, b)) bucket = b;
    }
    const uint32_t pos = nslots - 1 - ord;
    keys[pos] = bucket;
    vals[pos] = x;
    (void)total;
}

__global__ void __launch_bounds__(256) k_bucket_starts(const uint32_t *__restrict__ keys, uint32_t n, uint32_t nc,
                                                       uint32_t *__restrict__ start) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n) return;
    const long long prev = p == 0 ? -1 : (long long)min(keys[p - 1], nc);
    const long long cur = p == n ? (long long)nc : (long long)min(keys[p], nc);
    for (long long b = prev + 1; b <= cur; b++) start[b] = p;
}

__global__ void __launch_bounds__(256) k_locus_decode(const uint32_t *__restrict__ locus, uint32_t n,
                                                      const uint32_t *__restrict__ soas, int M,
                                                      uint2 *__restrict__ out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t x = locus[p];
    // bisect(soas, x): largest idx with soas[idx] < x, -1 if none
    int lo = -1, hi = M + 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (soas[mid] < x)
            lo = mid;
        else
            hi = mid;
    }
    uint2 r;
    r.x = (uint32_t)(lo + 1);                 // 0 = the reference's idx -1 (never scores; dropped)
    r.y = lo >= 0 ? x - soas[lo] : 0u;        // sst (may equal the sequence length)
    out[p] = r;
}

void free_chunk_index(ChunkIndex &ix) {
    if (ix.d_start) cudaFree(ix.d_start);
    if (ix.d_locus) cudaFree(ix.d_locus);
    if (ix.d_soas) cudaFree(ix.d_soas);
    if (ix.d_hdsst) cudaFree(ix.d_hdsst);
    ix.d_start = ix.d_locus = ix.d_soas = nullptr;
    ix.d_hdsst = nullptr;
}

int build_chunk_index(so_ctx *c, ChunkIndex &ix) {
    const Params &P = c->P;
    int rc;
    if ((rc = upload_cfg(P)) != SO_OK) return rc;
    const i64 M = ix.c1 - ix.c0;
    std::vector<uint32_t> soas((size_t)M + 1);
    const uint64_t base = c->t_off[(size_t)ix.c0];
    uint32_t maxlen = 0;
    for (i64 j = 0; j <= M; j++) {
        uint64_t v = c->t_off[(size_t)(ix.c0 + j)] - base;
        if (v > 0xfffffff0ull) {
            set_error("target chunk holds more than 4 Gi residues; lower -c");
            return SO_ELIMIT;
        }
        soas[(size_t)j] = (uint32_t)v;
        if (j > 0) maxlen = std::max<uint32_t>(maxlen, soas[(size_t)j] - soas[(size_t)j - 1]);
    }
    ix.total = soas[(size_t)M];
    ix.max_tlen = maxlen;
    const int AS = (int)(P.alphabets.size() * P.patterns.size());
    const uint64_t nslots64 = (uint64_t)ix.total * (uint64_t)AS;
    if (nslots64 > 0x7fffff00ull) {
        set_error("target chunk has too many seed slots (%llu); lower -c", (unsigned long long)nslots64);
        return SO_ELIMIT;
    }
    const uint32_t nslots = (uint32_t)nslots64;
    SO_CUDA(cudaMalloc((void **)&ix.d_soas, ((size_t)M + 1) * sizeof(uint32_t)));
    SO_CUDA(cudaMemcpyAsync(ix.d_soas, soas.data(), ((size_t)M + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice,
                            c->stream));
    SO_CUDA(cudaMalloc((void **)&ix.d_start, ((size_t)P.nc + 1) * sizeof(uint32_t)));
    SO_CUDA(cudaEventRecord(c->ev[0], c->stream));
    std::vector<uint32_t> h_counts;  // sizes of the non-empty buckets, in bucket order
    if (nslots > 0) {
        if ((rc = c->scratch[4].reserve((size_t)nslots * 4)) != SO_OK) return rc;
        if ((rc = c->scratch[5].reserve((size_t)nslots * 4)) != SO_OK) return rc;
        if ((rc = c->scratch[6].reserve((size_t)nslots * 4)) != SO_OK) return rc;
        if ((rc = c->scratch[7].reserve((size_t)nslots * 4)) != SO_OK) return rc;
        uint32_t *k_in = (uint32_t *)c->scratch[4].p, *v_in = (uint32_t *)c->scratch[5].p;
        uint32_t *k_out = (uint32_t *)c->scratch[6].p, *v_out = (uint32_t *)c->scratch[7].p;
        k_target_seeds<<<(nslots + 255) / 256, 256, 0, c->stream>>>(c->d_tres + base, ix.d_soas, (int)M, ix.total,
                                                                    nslots, k_in, v_in);
        int bits = 1;
        while (bits < 32 && (1ull << bits) <= (uint64_t)P.nc) bits++;
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (int)nslots, 0, bits, c->stream);
        if ((rc = c->scratch[11].reserve(tmp)) != SO_OK) return rc;
        SO_CUDA(cub::DeviceRadixSort::SortPairs(c->scratch[11].p, tmp, k_in, k_out, v_in, v_out, (int)nslots, 0, bits,
                                                c->stream));
        k_bucket_starts<<<(nslots + 1 + 255) / 256, 256, 0, c->stream>>>(k_out, nslots, P.nc, ix.d_start);
        c->stats.kernel_launches += 2;
        c->stats.lib_launches += 1;
        SO_CUDA(cudaGetLastError());
        // bucket sizes: run lengths of the sorted keys (library), so only the ~4*10^5 non-empty bucket counts travel to
        // the host instead of all keys; invalid slots carry key NC and form the last run
        uint32_t *d_uniq = k_in, *d_runs = v_in;                 // (free after the sort)
        uint32_t *d_nruns = (uint32_t *)c->scratch[11].p;
        {
            size_t t2 = 0;
            cub::DeviceRunLengthEncode::Encode(nullptr, t2, k_out, d_uniq, d_runs, d_nruns, (int)nslots, c->stream);
            if ((rc = c->scratch[10].reserve(t2 + 16)) != SO_OK) return rc;
            if ((rc = c->scratch[11].reserve(std::max<size_t>(tmp, 64))) != SO_OK) return rc;
            d_nruns = (uint32_t *)c->scratch[11].p;
            SO_CUDA(cub::DeviceRunLengthEncode::Encode(c->scratch[10].p, t2, k_out, d_uniq, d_runs, d_nruns, (int)nslots, c->stream));
            c->stats.lib_launches += 1;
        }
        uint32_t nruns = 0;
        SO_CUDA(cudaMemcpyAsync(&nruns, d_nruns, 4, cudaMemcpyDeviceToHost, c->stream));
        SO_CUDA(cudaStreamSynchronize(c->stream));
        h_counts.resize(nruns);
        uint32_t last_key = 0;
        if (nruns) {
            SO_CUDA(cudaMemcpyAsync(h_counts.data(), d_runs, (size_t)nruns * 4, cudaMemcpyDeviceToHost, c->stream));
            SO_CUDA(cudaMemcpyAsync(&last_key, d_uniq + (nruns - 1), 4, cudaMemcpyDeviceToHost, c->stream));
            SO_CUDA(cudaStreamSynchronize(c->stream));
        }
        c->stats.d2h_bytes += (i64)nruns * 4 + 8;
        uint32_t nseeds = nslots;
        if (nruns && last_key >= P.nc) {                         // the run of the invalid slots
            nseeds -= h_counts.back();
            h_counts.pop_back();
        }
        ix.n_seeds = nseeds;
        if (nseeds > 0) {
            SO_CUDA(cudaMalloc((void **)&ix.d_locus, (size_t)nseeds * sizeof(uint32_t)));
            SO_CUDA(cudaMalloc((void **)&ix.d_hdsst, (size_t)nseeds * sizeof(uint2)));
            SO_CUDA(cudaMemcpyAsync(ix.d_locus, v_out, (size_t)nseeds * 4, cudaMemcpyDeviceToDevice, c->stream));
            k_locus_decode<<<(nseeds + 255) / 256, 256, 0, c->stream>>>(ix.d_locus, nseeds, ix.d_soas, (int)M,
                                                                        ix.d_hdsst);
            c->stats.kernel_launches += 1;
        }
    } else {
        ix.n_seeds = 0;
        SO_CUDA(cudaMemsetAsync(ix.d_start, 0, ((size_t)P.nc + 1) * sizeof(uint32_t), c->stream));
    }
    SO_CUDA(cudaEventRecord(c->ev[1], c->stream));
    SO_CUDA(cudaStreamSynchronize(c->stream));
    SO_CUDA(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    ix.build_ms = ms;
    // get_mu_sd (fsearch.py:746-761) over the non-empty bucket counts in bucket order: N starts
    // at 1, sequential double accumulation; threshold = int(mu + 2 sd) (fsearch.py:2248-2250)
    {
        double N = 1, mu = 0.;
        const std::vector<uint32_t> &counts = h_counts;
        for (uint32_t v : counts) ix.max_bucket = std::max<uint32_t>(ix.max_bucket, v);
        for (uint32_t v : counts) mu += (double)v, N += 1;
        mu /= N;
        double sd = 0.;
        for (uint32_t v : counts) sd += std::pow((double)v - mu, 2);
        sd = std::sqrt(sd / N);
        ix.n_used = (i64)counts.size();
        ix.threshold = (i64)(mu + 2 * sd);
        // `thr < 1 and DB.threshold or thr` (fsearch.py:2992)
        if (!(P.thr < 1 && ix.threshold != 0)) ix.threshold = P.thr;
    }
    return SO_OK;
}

// ------------------------------------------------------------------------------------------ search
struct BlockGeom {
    int nq;            // queries in the sub-block
    int qb0;           // first query ordinal
    int qst_bits, diag_bits, hd_bits;
    int diag_bias;
    uint32_t L;        // len(locus) - 1
    uint32_t nc;
    long long thr_mul; // threshold
    int mink;
    int c0;            // chunk start (global target ordinal)
};

// slot layout of query q: slot_off[q] + as * Lq + i
__global__ void __launch_bounds__(256) k_query_seeds(const uint8_t *__restrict__ qres, const uint64_t *__restrict__ qoff,
                                                     const uint32_t *__restrict__ slot_off, BlockGeom g,
                                                     const uint32_t *__restrict__ start, uint32_t nslots,
                                                     uint32_t *__restrict__ slot_st, uint32_t *__restrict__ slot_cnt) {
    __shared__ uint16_t s_alpha[4 * 256];
    load_alpha(s_alpha);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nslots) return;
    const int ql = seq_of(slot_off, g.nq, tid);
    const uint64_t qbase = qoff[g.qb0 + ql];
    const int L = (int)(qoff[g.qb0 + ql + 1] - qbase);
    const uint32_t rel = tid - slot_off[ql];
    const int as = (int)(rel / (uint32_t)L);
    const int i = (int)(rel % (uint32_t)L);
    const int a = as / c_cfg.S, s = as % c_cfg.S;
    uint32_t st = 0, cnt = 0, b;
    if (seed_emitted(qres + qbase, L, i, a, s, s_alpha, b)) {
        // get_bin_mem: the last bucket is empty; end clipped to len(locus)-1
        if (b + 1 < g.nc) {
            st = start[b];
            uint32_t ed = min(start[b + 1], g.L);
            cnt = ed > st ? ed - st : 0;
        }
    }
    slot_st[tid] = st;
    slot_cnt[tid] = cnt;
}

// one warp per query
__global__ void __launch_bounds__(256) k_filter(const uint64_t *__restrict__ qoff, const uint32_t *__restrict__ slot_off,
                                                const uint32_t *__restrict__ perm, BlockGeom g,
                                                uint32_t *__restrict__ slot_cnt) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= g.nq) return;
    const uint64_t qbase = qoff[g.qb0 + w];
    const int L = (int)(qoff[g.qb0 + w + 1] - qbase);
    const int P = L - g.mink + 1;
    if (P <= 0) return;
    const int AS = c_cfg.A * c_cfg.S;
    const uint32_t so0 = slot_off[w];
    const long long thr = g.thr_mul * (long long)L;
    long long carry = 0;
    for (int r0 = 0; r0 < P; r0 += 32) {
        const int r = r0 + lane;
        int pos = -1;
        long long ct = 0;
        if (r < P) {
            pos = (int)perm[qbase + r];
            for (int as = 0; as < AS; as++) ct += slot_cnt[so0 + (uint32_t)as * (uint32_t)L + (uint32_t)pos];
        }
        long long inc = ct;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const long long before = carry + inc - ct;  // running total before this position
        if (r < P && before > thr)
            for (int as = 0; as < AS; as++) slot_cnt[so0 + (uint32_t)as * (uint32_t)L + (uint32_t)pos] = 0;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// one warp per slot
__global__ void __launch_bounds__(256) k_expand(const uint32_t *__restrict__ slot_off, BlockGeom g, uint32_t nslots,
                                                const uint64_t *__restrict__ qoff,
                                                const uint32_t *__restrict__ slot_st, const uint32_t *__restrict__ slot_cnt,
                                                const uint64_t *__restrict__ slot_out, const uint2 *__restrict__ hdsst,
                                                uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < nslots; slot += nwarps) {
        const uint32_t cnt = slot_cnt[slot];
        if (cnt == 0) continue;
        const int ql = seq_of(slot_off, g.nq, slot);
        const int L = (int)(qoff[g.qb0 + ql + 1] - qoff[g.qb0 + ql]);
        const uint32_t rel = slot - slot_off[ql];
        const int qst = (int)(rel % (uint32_t)L);
        const uint32_t st = slot_st[slot], out0 = (uint32_t)slot_out[slot];
        const uint64_t hi = (uint64_t)ql << (g.hd_bits + g.diag_bits + g.qst_bits);
        for (uint32_t k = lane; k < cnt; k += 32) {
            const uint2 e = hdsst[st + k];
            uint64_t key = ~0ull;  // hits attributed to "sequence -1" can never score: dropped
            if (e.x != 0) {
                const int diag = qst - (int)e.y + g.diag_bias;
                key = hi | ((uint64_t)e.x << (g.diag_bits + g.qst_bits)) | ((uint64_t)diag << g.qst_bits) | (uint64_t)qst;
            }
            keys[out0 + k] = key;
            if (vals) vals[out0 + k] = out0 + k;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Hit grouping without a global sort (keys-only path; SO_CUB_SORT=1 forces the device-wide radix sort).
//
// The seed hits of a query arrive as ~200 bucket lists, each already in descending (sequence, position)
// order.  Instead of a 5-pass device-wide radix sort of 64-bit keys they are partitioned straight into
// (query, target) cells -- a cell holds everything pair selection later folds together:
//   1. k_cell_pass<false>: one CTA per query, one shared-memory counter per target of the chunk
//      (4 B x (M + 2) <= 227 KB), one shared-memory atomic per hit;
//   2. cub::DeviceScan over the [query][target] counts -> cell bases;
//   3. k_cell_pass<true>: same traversal, the counters are now write cursors; the 64-bit key goes to its cell;
//   4. in-cell order (diagonal, qst): cells hold ~2.7 hits on average, so k_cell_small sorts cells of up to 8
//      hits with one THREAD per cell (sorting network on the 32-bit cell-local bits), k_cell_warp sorts cells
//      of up to kCellWarp hits with one warp per cell (rank sort out of shared memory), k_cell_block the rest
//      (cub::BlockRadixSort).  All in place.
// Each hit is read twice from the index (8 B) and its key is written once and touched once more, against
// 5 x (8 B + 8 B) for the radix sort; the hits the reference attributes to "sequence -1" (they can never
// score and were only sorted to the end before) are not materialised at all.
// A cell above kCellMax hits falls back to the device-wide sort for that sub-block.
// ---------------------------------------------------------------------------------------------
enum { kCellSmall = 8, kCellWarp = 192, kCellMax = 16384 };

// Count pass (SCATTER = false): per work item (query, target range) the per-target counters are scanned inside the
// CTA; cell_local[query][target] receives the exclusive prefix inside the item and utot[item] its total, so no
// device-wide scan over the cells is needed (k_unit_scan turns the <= 16 K item totals into item bases).  A cell
// above `cell_max` hits raises flags[0] (the block is then redone through the device-wide sort).
// Scatter pass (SCATTER = true): cursors start at ubase[item] + cell_local[..].
template <bool SCATTER>
__global__ void __launch_bounds__(1024) k_cell_pass(const uint32_t *__restrict__ slot_off, BlockGeom g,
                                                    const uint64_t *__restrict__ qoff,
                                                    const uint32_t *__restrict__ slot_st, const uint32_t *__restrict__ slot_cnt,
                                                    const uint2 *__restrict__ hdsst, uint32_t NB, uint32_t nsplit,
                                                    uint32_t *__restrict__ cell_local, uint32_t *__restrict__ utot,
                                                    const uint32_t *__restrict__ ubase, uint32_t *__restrict__ sub,
                                                    uint32_t cell_max, uint32_t *__restrict__ flags,
                                                    uint32_t *__restrict__ next_query) {
    if (SCATTER && flags[0]) return;
    __shared__ uint32_t s_part[32];
    // a work item is (query, one of `nsplit` target ranges): with two ranges the counters of a CTA need half the
    // shared memory, two CTAs (64 warps) fit an SM and hide the latency of this loop better; every CTA walks all
    // bucket lists of its query and keeps the hits of its range (lists are in descending target order, so the
    // upper ranges stop early)
    extern __shared__ uint32_t s_cell[];  // [NB / nsplit] per-target counters (count pass) or write cursors (scatter pass)
    const uint32_t NBh = (NB + nsplit - 1) / nsplit;
    __shared__ int s_qi;
    __shared__ uint32_t s_next_slot;
    const int lane = threadIdx.x & 31;
    constexpr int U = SCATTER ? 4 : 8;  // bucket rows in flight per warp (8 in the scatter pass measured the same)
    for (;;) {            // CTAs draw queries from a counter: no wave quantisation with one 200 KB CTA per SM
        if (threadIdx.x == 0) {
            s_qi = (int)atomicAdd(next_query, 1u);
            s_next_slot = 0;
        }
        __syncthreads();
        if ((uint32_t)s_qi >= (uint32_t)g.nq * nsplit) break;
        const int qi = (int)((uint32_t)s_qi / nsplit);
        const uint32_t lo = ((uint32_t)s_qi % nsplit) * NBh, span = min(NB, lo + NBh) - lo;  // targets [lo, lo + span)
        {
            const uint32_t ub = SCATTER ? ubase[s_qi] : 0u;
            for (uint32_t b = threadIdx.x; b < span; b += blockDim.x)
                s_cell[b] = SCATTER ? ub + cell_local[(size_t)qi * NB + lo + b] : 0u;
        }
        __syncthreads();
        const uint32_t so0 = slot_off[qi], nsl = slot_off[qi + 1] - so0;
        const int L = (int)(qoff[g.qb0 + qi + 1] - qoff[g.qb0 + qi]);
        for (;;) {  // warps draw slots one at a time: bucket lists differ a lot in length
            uint32_t sl = 0;
            if (lane == 0) sl = atomicAdd(&s_next_slot, 1u);
            sl = __shfl_sync(0xffffffffu, sl, 0);
            if (sl >= nsl) break;
            const uint32_t cnt = slot_cnt[so0 + sl];
            if (cnt == 0) continue;
            const uint32_t st = slot_st[so0 + sl];
            const int qst = (int)(sl % (uint32_t)L);
            for (uint32_t kb = 0; kb < cnt; kb += 32 * U) {  // warp-uniform trip count (the loop votes)
                const uint32_t k0 = kb + (uint32_t)lane;
                uint2 e[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t k = k0 + 32u * u;
                    e[u] = k < cnt ? hdsst[st + k] : make_uint2(0u, 0u);
                }
                bool below = false;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (e[u].x == 0) continue;  // past the end, or the reference's "sequence -1" (never scores): dropped
                    const uint32_t b = e[u].x - lo;
                    if (b >= span) {            // another range's target
                        below |= e[u].x < lo;
                        continue;
                    }
                    if (!SCATTER)
                        atomicAdd(&s_cell[b], 1u);
                    else {
                        // only the cell-local part (diagonal | qst) is stored: 4 B per hit keep the partially
                        // written sectors of the queries in flight inside L2
                        const uint32_t pos = atomicAdd(&s_cell[b], 1u);
                        const int diag = qst - (int)e[u].y + g.diag_bias;
                        sub[pos] = ((uint32_t)diag << g.qst_bits) | (uint32_t)qst;
                    }
                }
                if (__any_sync(0xffffffffu, below)) break;  // descending targets: nothing of this range is left
            }
        }
        __syncthreads();
        if (!SCATTER) {
            // exclusive scan of the item's counters in shared memory: every thread owns `per` consecutive counters
            // (per is odd: no bank conflicts), block scan over the per-thread sums
            const uint32_t per = ((span + blockDim.x - 1) / blockDim.x) | 1u;
            const uint32_t b0 = threadIdx.x * per, b1 = min(span, b0 + per);
            uint32_t sum = 0, mx = 0;
            for (uint32_t b = b0; b < b1; b++) {
                const uint32_t v = s_cell[b];
                sum += v;
                mx = max(mx, v);
            }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) s_part[threadIdx.x >> 5] = inc;
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0 && mx > cell_max) atomicExch(flags, 1u);
            __syncthreads();
            if (threadIdx.x < 32) {
                const uint32_t v = s_part[threadIdx.x];
                uint32_t w = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
                    if (lane >= o) w += u;
                }
                s_part[threadIdx.x] = w - v;
                if (threadIdx.x == 31) utot[s_qi] = w;
            }
            __syncthreads();
            uint32_t run = s_part[threadIdx.x >> 5] + inc - sum;
            for (uint32_t b = b0; b < b1; b++) {
                const uint32_t v = s_cell[b];
                s_cell[b] = run;
                run += v;
            }
            __syncthreads();
            for (uint32_t b = threadIdx.x; b < span; b += blockDim.x) cell_local[(size_t)qi * NB + lo + b] = s_cell[b];
            __syncthreads();
        }
    }
}

__device__ __forceinline__ void cex(uint32_t &a, uint32_t &b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo, b = hi;
}

// X-drop descriptor flags (uint2.y: query residue index inside the sub-block's view in the low 24 bits)
static const uint32_t kDescSingle = 1u << 24, kDescNoLeft = 1u << 25;
// chained group of exactly two seeds whose qst differ by 1..32 (the usual chain: two overlapping k-mers):
// bit 26 + (delta - 1) in bits 27..31, so k_xdrop never has to read the hits of such a group
static const uint32_t kDescPair = 1u << 26;
static const uint32_t kDescSkipX = 0xffffffffu;  // uint2.x of a hit that is not the head of its diagonal group
static const uint32_t kHeadBit = 0x80000000u;    // sorted cell-local key: first hit of a (query, target, diagonal) group

// item bases from the item totals (<= 16 K items): one CTA; also publishes the number of valid hits
__global__ void __launch_bounds__(1024) k_unit_scan(const uint32_t *__restrict__ utot, uint32_t U, uint32_t *__restrict__ ubase,
                                                    unsigned long long *__restrict__ counters, const uint32_t *__restrict__ flags) {
    __shared__ uint32_t s_part[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (uint32_t i0 = 0; i0 < U; i0 += 1024) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t v = i < U ? utot[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_part[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t t = s_part[threadIdx.x];
            uint32_t w = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            s_part[threadIdx.x] = w - t;
        }
        __syncthreads();
        const uint32_t ex = s_carry + s_part[threadIdx.x >> 5] + inc - v;
        if (i < U) ubase[i] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ubase[U] = s_carry;
        counters[0] = flags[0] ? 0ull : (unsigned long long)s_carry;  // hits to score (none when the block is to be redone)
        counters[2] = 0ull;                                            // group pool of k_xdrop
        counters[6] += (unsigned long long)s_carry;                    // statistic: seed hits
        uint32_t *t = reinterpret_cast<uint32_t *>(counters + 8);      // transient: list lengths, work counters
        t[0] = t[1] = t[2] = t[3] = 0u;
    }
}

// extent of cell c = (query qi, target t) in the hit arrays
__device__ __forceinline__ void cell_extent(uint32_t c, uint32_t NB, uint32_t NBh, uint32_t nsplit,
                                            const uint32_t *__restrict__ cell_local, const uint32_t *__restrict__ ubase,
                                            uint32_t &qi, uint32_t &t, uint32_t &off, uint32_t &n) {
    qi = c / NB;
    t = c - qi * NB;
    const uint32_t r = t / NBh, unit = qi * nsplit + r;
    const uint32_t ub = ubase[unit];
    off = ub + cell_local[c];
    const bool last = (t + 1 == NB) || ((t + 1) % NBh == 0);
    const uint32_t end = last ? ubase[unit + 1] : ub + cell_local[c + 1];
    n = end - off;
}

// sorted hit i of a cell: key + head bit, and the X-drop descriptor when it is the head of its diagonal group
// (kp / k1 / k2 = previous / next / second next sorted keys; has* = they exist inside the cell)
__device__ __forceinline__ void cell_emit(uint32_t p, uint32_t k, bool hasp, uint32_t kp, bool has1, uint32_t k1, bool has2,
                                          uint32_t k2, uint32_t tbase, uint32_t qrel, uint32_t cid, const BlockGeom &g,
                                          uint32_t *__restrict__ ssub, uint2 *__restrict__ desc, uint32_t *__restrict__ cellid) {
    const uint32_t qmask = (1u << g.qst_bits) - 1u;
    const uint32_t dg = k >> g.qst_bits;
    const bool head = !hasp || (kp >> g.qst_bits) != dg;
    ssub[p] = k | (head ? kHeadBit : 0u);
    if (!head) {
        desc[p] = make_uint2(kDescSkipX, 0u);
        return;
    }
    cellid[p] = cid;  // (query << 16 | target + 1): read back by k_xdrop for the groups that pass
    const int qst = (int)(k & qmask);
    const int diag = (int)dg - g.diag_bias;
    const int sst = qst - diag;
    const bool multi = has1 && (k1 >> g.qst_bits) == dg;
    uint32_t flags = (multi ? 0u : kDescSingle) | ((qst == 0 || sst == 0) ? kDescNoLeft : 0u);
    if (multi && !(has2 && (k2 >> g.qst_bits) == dg)) {
        const uint32_t delta = (k1 & qmask) - (uint32_t)qst;
        if (delta >= 1 && delta <= 32) flags |= kDescPair | ((delta - 1) << 27);
    }
    desc[p] = make_uint2(tbase + (uint32_t)sst, (qrel + (uint32_t)qst) | flags);
}

// one thread per cell: cells of 1..8 hits are sorted in registers on their cell-local bits (diagonal | qst) and
// written out as sorted keys + descriptors; larger cells are queued for k_cell_warp / k_cell_block
__global__ void __launch_bounds__(256) k_cell_small(const uint32_t *__restrict__ cell_local, const uint32_t *__restrict__ ubase,
                                                    uint32_t ncells, uint32_t NB, uint32_t NBh, uint32_t nsplit, BlockGeom g,
                                                    const uint64_t *__restrict__ qoff, const uint64_t *__restrict__ toff, uint64_t qa,
                                                    const uint32_t *__restrict__ sub, uint32_t *__restrict__ ssub,
                                                    uint2 *__restrict__ desc, uint32_t *__restrict__ cellid,
                                                    uint32_t *__restrict__ wlist,
                                                    uint32_t *__restrict__ blist, uint32_t *__restrict__ lcount,
                                                    const uint32_t *__restrict__ flags) {
    // 2-D launch: blockIdx.y = query, x = target (no division per thread; `flags` is not consulted here: a block that
    // is going to be redone only queues its large cells, which k_cell_block then skips)
    (void)flags;
    (void)ncells;
    const uint32_t qi = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = qi * NB + t;
    uint32_t off = 0, n = 0;
    if (t < NB) {
        const uint32_t unit = nsplit == 1 ? qi : qi * nsplit + t / NBh;
        const uint32_t ub = ubase[unit];
        off = ub + cell_local[c];
        const bool last = (t + 1 == NB) || (nsplit != 1 && (t + 1) % NBh == 0);
        n = (last ? ubase[unit + 1] : ub + cell_local[c + 1]) - off;
    }
    {   // larger cells are queued: one atomic per warp and list
        const int lane = threadIdx.x & 31;
        const bool qw = n > (uint32_t)kCellSmall && n <= (uint32_t)kCellWarp, qb = n > (uint32_t)kCellWarp;
        const unsigned mw = __ballot_sync(0xffffffffu, qw), mb = __ballot_sync(0xffffffffu, qb);
        if (mw) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(lcount, (uint32_t)__popc(mw));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (qw) wlist[base + __popc(mw & ((1u << lane) - 1u))] = c;
        }
        if (mb) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(lcount + 1, (uint32_t)__popc(mb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (qb) blist[base + __popc(mb & ((1u << lane) - 1u))] = c;
        }
    }
    if (n == 0 || n > (uint32_t)kCellSmall) return;
    uint32_t k[kCellSmall];
#pragma unroll
    for (int i = 0; i < kCellSmall; i++) k[i] = (uint32_t)i < n ? sub[off + i] : 0xffffffffu;
    if (n > 4) {  // 19-comparator network on 8
        cex(k[0], k[1]), cex(k[2], k[3]), cex(k[4], k[5]), cex(k[6], k[7]);
        cex(k[0], k[2]), cex(k[1], k[3]), cex(k[4], k[6]), cex(k[5], k[7]);
        cex(k[1], k[2]), cex(k[5], k[6]), cex(k[0], k[4]), cex(k[3], k[7]);
        cex(k[1], k[5]), cex(k[2], k[6]);
        cex(k[1], k[4]), cex(k[3], k[6]);
        cex(k[2], k[4]), cex(k[3], k[5]);
        cex(k[3], k[4]);
    } else if (n > 1) {  // 5-comparator network on 4
        cex(k[0], k[1]), cex(k[2], k[3]), cex(k[0], k[2]), cex(k[1], k[3]), cex(k[1], k[2]);
    }
    const uint32_t tbase = (uint32_t)toff[g.c0 + (int)t - 1];
    const uint32_t qrel = (uint32_t)(qoff[g.qb0 + (int)qi] - qa);
#pragma unroll
    for (int i = 0; i < kCellSmall; i++)
        if ((uint32_t)i < n)
            cell_emit(off + i, k[i], i > 0, i > 0 ? k[i - 1] : 0u, (uint32_t)(i + 1) < n, i + 1 < kCellSmall ? k[i + 1] : 0u,
                      (uint32_t)(i + 2) < n, i + 2 < kCellSmall ? k[i + 2] : 0u, tbase, qrel, (qi << 16) | t, g, ssub, desc, cellid);
}

// one warp per queued cell (9..kCellWarp hits): rank sort out of shared memory (keys of a cell are distinct)
__global__ void __launch_bounds__(256) k_cell_warp(const uint32_t *__restrict__ cell_local, const uint32_t *__restrict__ ubase,
                                                   const uint32_t *__restrict__ wlist, const uint32_t *__restrict__ lcount,
                                                   uint32_t NB, uint32_t NBh, uint32_t nsplit, BlockGeom g,
                                                   const uint64_t *__restrict__ qoff, const uint64_t *__restrict__ toff, uint64_t qa,
                                                   const uint32_t *__restrict__ sub, uint32_t *__restrict__ ssub,
                                                   uint2 *__restrict__ desc, uint32_t *__restrict__ cellid,
                                                   const uint32_t *__restrict__ flags) {
    if (flags[0]) return;
    __shared__ uint32_t s_in[8][kCellWarp];
    __shared__ uint32_t s_out[8][kCellWarp + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nlist = lcount[0];
    uint32_t *si = s_in[warp], *so_ = s_out[warp];
    for (uint32_t w0 = (blockIdx.x * 8 + warp) * 32; w0 < nlist; w0 += gridDim.x * 8 * 32) {
        // 32 queued cells per warp at a time: list entries, extents and sequence bases are fetched by the lanes in parallel
        uint32_t my_c = 0, my_qi = 0, my_t = 0, my_off = 0, my_n = 0, my_tb = 0, my_qr = 0;
        if (w0 + lane < nlist) {
            my_c = wlist[w0 + lane];
            cell_extent(my_c, NB, NBh, nsplit, cell_local, ubase, my_qi, my_t, my_off, my_n);
            my_tb = (uint32_t)toff[g.c0 + (int)my_t - 1];
            my_qr = (uint32_t)(qoff[g.qb0 + (int)my_qi] - qa);
        }
        const int cells = (int)min(32u, nlist - w0);
        // cells of up to 32 hits (most of them): one key per lane, ranks by shuffles, the next cell's key is in
        // flight while the current one is ranked; larger cells go through shared memory
        uint32_t nx = 0;
        {
            const uint32_t off0 = __shfl_sync(0xffffffffu, my_off, 0), n0 = __shfl_sync(0xffffffffu, my_n, 0);
            if ((uint32_t)lane < n0 && n0 <= 32u) nx = sub[off0 + lane];
        }
        for (int t = 0; t < cells; t++) {
            const uint32_t off = __shfl_sync(0xffffffffu, my_off, t), n = __shfl_sync(0xffffffffu, my_n, t);
            const uint32_t tbase = __shfl_sync(0xffffffffu, my_tb, t), qrel = __shfl_sync(0xffffffffu, my_qr, t);
            const uint32_t cid = (__shfl_sync(0xffffffffu, my_qi, t) << 16) | __shfl_sync(0xffffffffu, my_t, t);
            const uint32_t x1 = nx;
            if (t + 1 < cells) {
                const uint32_t off1 = __shfl_sync(0xffffffffu, my_off, t + 1), n1 = __shfl_sync(0xffffffffu, my_n, t + 1);
                if ((uint32_t)lane < n1 && n1 <= 32u) nx = sub[off1 + lane];
            }
            if (n <= 32u) {
                uint32_t rank = 0;
                for (uint32_t j = 0; j < n; j++) {
                    const uint32_t y = __shfl_sync(0xffffffffu, x1, (int)j);
                    rank += (y < x1 || (y == x1 && j < (uint32_t)lane)) ? 1u : 0u;
                }
                if ((uint32_t)lane < n) so_[rank] = x1;
            } else {
                for (uint32_t i = lane; i < n; i += 32) si[i] = sub[off + i];
                __syncwarp();
                for (uint32_t i = lane; i < n; i += 32) {
                    const uint32_t x = si[i];
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < n; j++) {
                        const uint32_t y = si[j];
                        rank += (y < x || (y == x && j < i)) ? 1u : 0u;
                    }
                    so_[rank] = x;
                }
            }
            __syncwarp();
            for (uint32_t i = lane; i < n; i += 32)
                cell_emit(off + i, so_[i], i > 0, i > 0 ? so_[i - 1] : 0u, i + 1 < n, i + 1 < n ? so_[i + 1] : 0u, i + 2 < n,
                          i + 2 < n ? so_[i + 2] : 0u, tbase, qrel, cid, g, ssub, desc, cellid);
            __syncwarp();
        }
    }
}

// One warp per 32 consecutive cells of one query (cells of 1..kCellWarp hits; larger ones are queued for k_cell_block):
// replaces k_cell_small + k_cell_warp.  The hits of consecutive cells are contiguous in `sub`, so a warp loads the whole
// span coalesced into shared memory, every HIT ranks itself inside its cell (count of smaller keys: cells average a few
// hits), and every sorted position emits its key / descriptor / cell id -- one lane per hit instead of one thread
// (k_cell_small) or one warp (k_cell_warp) per cell, no divergence over the cell sizes.  Spans above kSpanCap hits are
// processed as several groups of whole cells.
enum { kSpanCap = 512 };
__global__ void __launch_bounds__(256) k_cell_span(const uint32_t *__restrict__ cell_local, const uint32_t *__restrict__ ubase,
                                                   uint32_t NB, uint32_t NBh, uint32_t nsplit, BlockGeom g,
                                                   const uint64_t *__restrict__ qoff, const uint64_t *__restrict__ toff, uint64_t qa,
                                                   const uint32_t *__restrict__ sub, uint32_t *__restrict__ ssub,
                                                   uint2 *__restrict__ desc, uint32_t *__restrict__ cellid,
                                                   uint32_t *__restrict__ blist, uint32_t *__restrict__ lcount) {
    __shared__ uint32_t s_in[8][kSpanCap];
    __shared__ uint32_t s_out[8][kSpanCap + 2];
    __shared__ uint8_t s_cell[8][kSpanCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = qi * NB + t;
    uint32_t off = 0, n = 0;
    if (t < NB) {
        const uint32_t unit = nsplit == 1 ? qi : qi * nsplit + t / NBh;
        const uint32_t ub = ubase[unit];
        off = ub + cell_local[c];
        const bool last = (t + 1 == NB) || (nsplit != 1 && (t + 1) % NBh == 0);
        n = (last ? ubase[unit + 1] : ub + cell_local[c + 1]) - off;
    }
    {   // cells above kCellWarp hits are queued: one atomic per warp
        const bool qb = n > (uint32_t)kCellWarp;
        const unsigned mb = __ballot_sync(0xffffffffu, qb);
        if (mb) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(lcount + 1, (uint32_t)__popc(mb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (qb) blist[base + __popc(mb & ((1u << lane) - 1u))] = c;
        }
    }
    const uint32_t nn = n <= (uint32_t)kCellWarp ? n : 0u;
    uint32_t pre = nn;  // inclusive prefix over the warp's cells, then exclusive
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += u;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, pre, 31);
    if (total == 0) return;
    pre -= nn;
    const uint32_t tbase = nn ? (uint32_t)toff[g.c0 + (int)t - 1] : 0u;
    const uint32_t qrel = (uint32_t)(qoff[g.qb0 + (int)qi] - qa);
    uint32_t *si = s_in[warp], *so_ = s_out[warp];
    uint8_t *sc = s_cell[warp];
    int a = 0;
    while (a < 32) {
        // group of whole cells [a, b) with at most kSpanCap hits (a cell has at most kCellWarp <= kSpanCap)
        const uint32_t pa = __shfl_sync(0xffffffffu, pre, a);
        const unsigned fits = __ballot_sync(0xffffffffu, lane >= a && pre + nn - pa <= (uint32_t)kSpanCap);
        const int b = a + __popc(fits);  // `fits` is a run of ones starting at lane a (prefixes are monotone)
        const uint32_t pb = b < 32 ? __shfl_sync(0xffffffffu, pre, b & 31) : total;
        const uint32_t gm = pb - pa;
        // ---- load: hit i of the group belongs to the last cell whose prefix is <= pa + i
        for (uint32_t i0 = 0; i0 < gm; i0 += 32) {
            const uint32_t i = i0 + lane, gi = pa + i;
            int lo = a, hi = b;  // invariant: pre[lo] <= gi < pre[hi] (pre[b] = pb)
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const int mid = (lo + hi) >> 1;
                const uint32_t pm = __shfl_sync(0xffffffffu, pre, mid);
                if (mid > lo && pm <= gi) lo = mid; else if (mid > lo) hi = mid;
            }
            // empty cells share their prefix with the next cell: the owner is the LAST cell with pre <= gi, found above
            // because the search keeps lo at the highest index whose prefix is <= gi
            const uint32_t coff = __shfl_sync(0xffffffffu, off, lo), cpre = __shfl_sync(0xffffffffu, pre, lo);
            if (i < gm) {
                si[i] = sub[coff + (gi - cpre)];
                sc[i] = (uint8_t)lo;
            }
        }
        __syncwarp();
        // ---- rank inside the cell
        for (uint32_t i0 = 0; i0 < gm; i0 += 32) {
            const uint32_t i = i0 + lane;
            const int cl = i < gm ? (int)sc[i] : a;
            const uint32_t cs = __shfl_sync(0xffffffffu, pre, cl) - pa, cn = __shfl_sync(0xffffffffu, nn, cl);
            if (i < gm) {
                // (the keys of a cell are distinct: one pattern, one alphabet -> one hit per (qst, sst))
                const uint32_t x = si[i];
                const uint32_t *sk = si + cs;
                uint32_t rank = 0, j = 0;
                for (; j + 4 <= cn; j += 4) rank += (sk[j] < x) + (sk[j + 1] < x) + (sk[j + 2] < x) + (sk[j + 3] < x);
                for (; j < cn; j++) rank += sk[j] < x;
                so_[cs + rank] = x;
            }
        }
        __syncwarp();
        // ---- emit
        for (uint32_t i0 = 0; i0 < gm; i0 += 32) {
            const uint32_t i = i0 + lane;
            const int cl = i < gm ? (int)sc[i] : a;
            const uint32_t cs = __shfl_sync(0xffffffffu, pre, cl) - pa, cn = __shfl_sync(0xffffffffu, nn, cl);
            const uint32_t coff = __shfl_sync(0xffffffffu, off, cl), ctb = __shfl_sync(0xffffffffu, tbase, cl);
            if (i < gm) {
                const uint32_t r = i - cs;
                const uint32_t ct = t - (uint32_t)lane + (uint32_t)cl;
                cell_emit(coff + r, so_[i], r > 0, r > 0 ? so_[i - 1] : 0u, r + 1 < cn, r + 1 < cn ? so_[i + 1] : 0u, r + 2 < cn,
                          r + 2 < cn ? so_[i + 2] : 0u, ctb, qrel, (qi << 16) | ct, g, ssub, desc, cellid);
            }
        }
        __syncwarp();
        a = b;
    }
}

// one CTA per queued cell (kCellWarp < hits <= kCellMax): shared-memory radix sort of the cell-local bits
template <int THREADS, int ITEMS>
__device__ __forceinline__ void cell_sort_tile(const uint32_t *__restrict__ sub, uint32_t off, uint32_t n, int Lb, void *smem) {
    typedef cub::BlockRadixSort<uint32_t, THREADS, ITEMS> Sort;
    typename Sort::TempStorage &tmp = *reinterpret_cast<typename Sort::TempStorage *>(smem);
    uint32_t k[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t idx = (uint32_t)i * THREADS + threadIdx.x;  // striped, coalesced
        k[i] = idx < n ? sub[off + idx] : 0xffffffffu;
    }
    Sort(tmp).SortBlockedToStriped(k, 0, Lb);
    __syncthreads();
    uint32_t *sorted = reinterpret_cast<uint32_t *>(smem);  // the sort's scratch becomes the sorted cell
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t idx = (uint32_t)i * THREADS + threadIdx.x;
        if (idx < n) sorted[idx] = k[i];
    }
    __syncthreads();
}

union CellBlockSmem {
    cub::BlockRadixSort<uint32_t, 512, 32>::TempStorage sort;
    uint32_t sorted[kCellMax + 8];
};

__global__ void __launch_bounds__(512) k_cell_block(const uint32_t *__restrict__ cell_local, const uint32_t *__restrict__ ubase,
                                                    const uint32_t *__restrict__ blist, const uint32_t *__restrict__ lcount,
                                                    uint32_t NB, uint32_t NBh, uint32_t nsplit, BlockGeom g,
                                                    const uint64_t *__restrict__ qoff, const uint64_t *__restrict__ toff, uint64_t qa,
                                                    const uint32_t *__restrict__ sub, uint32_t *__restrict__ ssub,
                                                    uint2 *__restrict__ desc, uint32_t *__restrict__ cellid,
                                                    const uint32_t *__restrict__ flags) {
    if (flags[0]) return;
    extern __shared__ __align__(16) unsigned char smem[];  // CellBlockSmem
    const uint32_t nlist = lcount[1];
    const int Lb = g.qst_bits + g.diag_bits;
    const uint32_t *sorted = reinterpret_cast<const uint32_t *>(smem);
    for (uint32_t w = blockIdx.x; w < nlist; w += gridDim.x) {
        const uint32_t c = blist[w];
        uint32_t qi, t, off, n;
        cell_extent(c, NB, NBh, nsplit, cell_local, ubase, qi, t, off, n);
        if (n <= 512u * 1)
            cell_sort_tile<512, 1>(sub, off, n, Lb, smem);
        else if (n <= 512u * 2)
            cell_sort_tile<512, 2>(sub, off, n, Lb, smem);
        else if (n <= 512u * 4)
            cell_sort_tile<512, 4>(sub, off, n, Lb, smem);
        else if (n <= 512u * 8)
            cell_sort_tile<512, 8>(sub, off, n, Lb, smem);
        else if (n <= 512u * 16)
            cell_sort_tile<512, 16>(sub, off, n, Lb, smem);
        else
            cell_sort_tile<512, 32>(sub, off, n, Lb, smem);
        const uint32_t tbase = (uint32_t)toff[g.c0 + (int)t - 1];
        const uint32_t qrel = (uint32_t)(qoff[g.qb0 + (int)qi] - qa);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            cell_emit(off + i, sorted[i], i > 0, i > 0 ? sorted[i - 1] : 0u, i + 1 < n, i + 1 < n ? sorted[i + 1] : 0u, i + 2 < n,
                      i + 2 < n ? sorted[i + 2] : 0u, tbase, qrel, (qi << 16) | t, g, ssub, desc, cellid);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Chained X-drop scoring, lane-persistent.
//
// The sorted hit array is cut into diagonal groups (query, target, diagonal).  A group costs anything
// from one to thousands of extension steps, so a thread-per-group mapping leaves ~3 of 32 lanes busy
// (ncu, round 1: 2.65 active threads per instruction).  Here every LANE runs a small state machine
// (NEED -> SEED -> RIGHT -> LEFT -> NEXT ...) and all 32 lanes execute the same "one extension step"
// body each iteration, whatever group / seed / direction they are in; a lane that finishes its group
// pulls the next group index from a global counter (warp-aggregated atomicAdd).
// Semantics: ungap / get_ungap_scores (fsearch.py:2454-2509): per seed (ascending unique qst) a right
// extension from (Q, S) and a left extension from (Q-1, S-1) continuing the right maximum, both
// X-drop 30, both confined to lo < q < hi where lo = max(0, diag) for the first seed (index 0 of either
// sequence is never scored) and the previous segment's max_qed afterwards; hi = min(ql, tl + diag).
// ---------------------------------------------------------------------------------------------
struct HeadFlag {
    const uint64_t *keys;
    int shift;
    __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t &p) const {
        const uint64_t k = keys[p];
        if (k == ~0ull) return 0u;
        return (p == 0 || (keys[p - 1] >> shift) != (k >> shift)) ? 1u : 0u;
    }
};

__global__ void __launch_bounds__(256) k_scatter_heads(const uint64_t *__restrict__ keys, uint32_t n, int shift,
                                                       const uint32_t *__restrict__ gidx, uint32_t *__restrict__ gheads) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint64_t k = keys[p];
    if (k == ~0ull) return;
    if (p == 0 || (keys[p - 1] >> shift) != (k >> shift)) gheads[gidx[p]] = p;
}

__global__ void __launch_bounds__(256) k_gather_keys(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ gheads,
                                                     uint32_t G, uint64_t *__restrict__ gkey) {
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi < G) gkey[gi] = keys[gheads[gi]];
}

__global__ void __launch_bounds__(256) k_group_ungap_generic(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                     uint32_t n, const uint32_t *__restrict__ gheads, uint32_t G, BlockGeom g,
                                                     const uint8_t *__restrict__ qcls, const uint64_t *__restrict__ qoff,
                                                     const uint8_t *__restrict__ tcls, const uint64_t *__restrict__ toff,
                                                     uint32_t *__restrict__ gscore, uint32_t *__restrict__ grank,
                                                     unsigned long long *__restrict__ counters) {
    __shared__ int8_t s_tbl[kClasses * kClasses];
    for (int k = threadIdx.x; k < kClasses * kClasses; k += blockDim.x) s_tbl[k] = c_score2[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t qmask = (1ull << g.qst_bits) - 1, dmask = (1ull << g.diag_bits) - 1;
    const int pair_shift = g.qst_bits + g.diag_bits;
    constexpr int U = 8;                      // extension steps per loop iteration
    const unsigned long long kBatch = 256;    // group indices taken per atomic
    // per-lane state
    bool has = false, fin = false, stopped = false;
    uint32_t gi = 0, e = 0, rank_min = 0;
    uint64_t grp = 0;
    const uint8_t *q = nullptr, *t = nullptr;  // t is pre-shifted by -diag: t[qq] faces q[qq]
    int lo = 0, hi = 0, total = 0, prev_q = -1;
    int Q = 0, qq = 0, score = 0, mx = 0, mx_qed = 0, dir = 1;
    unsigned long long steps = 0;
    uint32_t pool_next = 0, pool_end = 0;  // warp-uniform
    bool exhausted = false;                // warp-uniform
    for (;;) {
        // ---- (1) lanes without a group take the next one (a warp draws kBatch indices per atomic)
        const unsigned need = __ballot_sync(0xffffffffu, !has && !fin);
        if (need) {
            if (pool_next == pool_end && !exhausted) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counters + 4, kBatch);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= G)
                    exhausted = true;
                else {
                    pool_next = (uint32_t)base;
                    pool_end = (uint32_t)min((unsigned long long)G, base + kBatch);
                }
            }
            if (!has && !fin) {
                const uint32_t mine = pool_next + __popc(need & ((1u << lane) - 1));
                if (mine < pool_end) {
                    gi = mine;
                    e = gheads[gi];
                    const uint64_t key = keys[e];
                    grp = key >> g.qst_bits;
                    const int diag = (int)(grp & dmask) - g.diag_bias;
                    const uint64_t pair = key >> pair_shift;
                    const int hd1 = (int)(pair & ((1ull << g.hd_bits) - 1));
                    const int qi = (int)(pair >> g.hd_bits);
                    const uint64_t qb = qoff[g.qb0 + qi];
                    const int ql = (int)(qoff[g.qb0 + qi + 1] - qb);
                    const int tid = g.c0 + hd1 - 1;
                    const uint64_t tb = toff[tid];
                    const int tl = (int)(toff[tid + 1] - tb);
                    q = qcls + qb;
                    t = tcls + tb - diag;
                    lo = max(0, diag);            // first seed: 0 < q and 0 < s
                    hi = min(ql, tl + diag);      // q < ql and s < tl
                    total = 0;
                    rank_min = vals ? vals[e] : 0u;
                    prev_q = (int)(key & qmask);
                    Q = max(prev_q, lo);          // off = max(qlo - Q, slo - S, 0)
                    qq = Q, score = 0, mx = 0, mx_qed = Q, dir = 1;
                    stopped = false;
                    has = true;
                } else if (exhausted)
                    fin = true;  // otherwise the pool is refilled in the next iteration
            }
            pool_next = min(pool_end, pool_next + (uint32_t)__popc(need));
        }
        if (__all_sync(0xffffffffu, fin)) break;
        // ---- (2) U extension steps; right (dir = +1) and left (dir = -1) extensions share the body
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool act = has && !stopped;
            const bool in = act && qq > lo && qq < hi;
            int sc = 0;
            if (in) sc = s_tbl[(int)q[qq] * kClasses + t[qq]];
            if (in) {
                score += sc;
                steps++;
                if (score > mx) {
                    mx = score;
                    if (dir > 0) mx_qed = qq;
                } else if (score + 30 < mx)
                    stopped = true;
                qq += dir;
            } else if (act)
                stopped = true;
        }
        // ---- (3) transitions of the lanes whose extension ended
        if (has && stopped) {
            if (dir > 0) {  // right done -> left from (Q-1, S-1), continuing the right maximum
                dir = -1;
                qq = Q - 1;
                score = mx;
                stopped = false;
            } else {        // seed done
                total += mx;
                lo = mx_qed;  // next seed: qlo = max_qed, slo = max_sed (same diagonal)
                bool more = false;
                for (;;) {
                    e++;
                    if (e >= n) break;
                    const uint64_t k2 = keys[e];
                    if ((k2 >> g.qst_bits) != grp) break;
                    if (vals) rank_min = min(rank_min, vals[e]);
                    const int qst = (int)(k2 & qmask);
                    if (qst == prev_q) continue;  // same point again (other pattern / alphabet): lis() drops it
                    prev_q = qst;
                    more = true;
                    break;
                }
                if (more) {
                    Q = max(prev_q, lo);
                    qq = Q, score = 0, mx = 0, mx_qed = Q, dir = 1;
                    stopped = false;
                } else {
                    gscore[gi] = (uint32_t)total;
                    if (vals) grank[gi] = rank_min;
                    has = false;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if (lane == 0 && steps) atomicAdd(counters + 1, steps);
}

// ---------------------------------------------------------------------------------------------
// Single-seed diagonal groups (the bulk: ~95 % of the groups are one random seed hit).
//
// A single seed needs one right extension from (Q, S) and one left extension from (Q-1, S-1) that
// continues the right maximum (fsearch.py:2454-2494); the group score is the sum of the two gains,
// so the two phases are independent.  k_single_ungap runs them lane-serially:
//   * X-drop views of the sequences: residue classes with position 0 of every sequence replaced by
//     a terminator class (the reference never scores index 0 of either sequence and stops at
//     either end: `0 < q < ql and 0 < s < tl`), stored forward and reversed, so a left extension is
//     a forward walk over the reversed copies and no range arithmetic is left in the loop.
//   * 16 residues per lane per load (LDG.128, 16-byte aligned target chunk; leading bytes before the
//     start are rewritten to a "skip" class that scores 0 and is not counted); the matching 16 query
//     bytes are muxed out of two aligned chunks (the query side is shared by the lanes of a warp, so
//     its lines stay in L1).
//   * score table in shared memory, one private copy per LANE ([26 x 32 entries][32 lanes] int32):
//     random lookups are bank-conflict free.
//   * state per phase is two registers: v = S*8192 - n (S running score, n steps) and
//     d = (max - S)*8192 + (steps since the max).  Table entries are 1 - s*8192, so a step is
//     v -= e; d = max(d + e, 0); alive = d < 31*8192   (X-drop: S + 30 < max),
//     and the phase gain is (v + d + 8191) >> 13, the step count (-v) & 8191.  The terminator entry
//     is 2^29 (drops at once, leaves v + d and the count unchanged).
// Lanes are refilled in batches: the service block (result write, next group, phase switch, first
// loads) runs only when kRefill or more lanes are idle.  Requires sequences < 8192 residues.
// ---------------------------------------------------------------------------------------------
enum { kUngPad = 64, kUngStop = 24, kUngSkip = 25, kUngRows = 26, kUngTabBytes = kUngRows * 32 * 32 * 4 };
__global__ void __launch_bounds__(256) k_ung_fill(const uint8_t *__restrict__ cls, uint32_t n, uint8_t *__restrict__ dst,
                                                  uint32_t offF, uint32_t offR, uint32_t total, int shift) {
    const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= total) return;
    uint32_t v = kUngStop;
    if (pos >= offF && pos - offF < n)
        v = cls[pos - offF];
    else if (pos >= offR && pos - offR < n)
        v = cls[n - 1 - (pos - offR)];
    dst[pos] = (uint8_t)(v << shift);
}

__global__ void __launch_bounds__(256) k_ung_marks(const uint64_t *__restrict__ off, uint32_t nseq, uint64_t base,
                                                   uint32_t n, uint8_t *__restrict__ dst, uint32_t offF, uint32_t offR,
                                                   int shift) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseq) return;
    const uint64_t x = off[i] - base;
    if (x >= n) return;
    dst[offF + (uint32_t)x] = (uint8_t)(kUngStop << shift);
    dst[offR + (n - 1 - (uint32_t)x)] = (uint8_t)(kUngStop << shift);
}

// shifted copies of a view (copy s holds view[j + s * step] at j), each `stride` bytes long: with 16 copies one byte
// apart k_xdrop<FAST> reads its query window with one 16-byte aligned load whatever the alignment of the position; with
// 4 copies one word apart only the byte shift inside a word is left (no register multiplexing)
__global__ void __launch_bounds__(256) k_ung_shift(const uint8_t *__restrict__ src, uint32_t total, uint8_t *__restrict__ dst,
                                                   uint32_t stride, uint32_t step, uint8_t fill) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    if (j >= stride) return;
    dst[(size_t)s * stride + j] = j + s * step < total ? src[j + s * step] : fill;
}

static void ung_layout(uint32_t n, uint32_t off[2], uint32_t &total) {
    const uint32_t r = (n + 15u) & ~15u;
    off[0] = kUngPad;
    off[1] = kUngPad + r + kUngPad;
    total = off[1] + r + kUngPad;
}

static int ung_build(cudaStream_t st, so_stats &stats, const uint8_t *d_cls, const uint64_t *d_off, uint32_t nseq,
                     uint64_t base, uint32_t n, uint8_t *dst, const uint32_t off[2], uint32_t total, int shift) {
    k_ung_fill<<<(total + 255) / 256, 256, 0, st>>>(d_cls, n, dst, off[0], off[1], total, shift);
    if (nseq) k_ung_marks<<<(nseq + 255) / 256, 256, 0, st>>>(d_off, nseq, base, n, dst, off[0], off[1], shift);
    SO_CUDA(cudaGetLastError());
    stats.kernel_launches += 2;
    return SO_OK;
}

int build_ungap_targets(so_ctx *c) {
    if (c->d_tung) cudaFree(c->d_tung);
    c->d_tung = nullptr;
    const uint64_t R = c->t_off[(size_t)c->n_t];
    if (R == 0 || R >= 0x7fffff00ull) return SO_OK;  // chained kernels handle everything
    uint32_t total;
    ung_layout((uint32_t)R, c->tung_off, total);
    SO_CUDA(cudaMalloc((void **)&c->d_tung, total));
    return ung_build(c->stream, c->stats, c->d_tcls, c->d_toff, (uint32_t)c->n_t, 0, (uint32_t)R, c->d_tung, c->tung_off, total, 0);
}

// one thread per sorted hit: group heads, the X-drop descriptor of the group (global target residue
// index of the seed, query residue index inside the sub-block's query view, flags) and the list of
// chained (multi-seed) flag
__global__ void __launch_bounds__(256) k_group_desc(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                    uint32_t n, BlockGeom g, const uint32_t *__restrict__ gidx,
                                                    const uint64_t *__restrict__ qoff, const uint64_t *__restrict__ toff,
                                                    uint64_t qa, uint32_t *__restrict__ gheads, uint2 *__restrict__ desc,
                                                    uint64_t *__restrict__ gkey,
                                                    uint32_t *__restrict__ grank,
                                                    unsigned long long *__restrict__ counters) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const int shift = g.qst_bits;
    if (p < n) {
        const uint64_t k = keys[p];
        if (k != ~0ull && (p == 0 || (keys[p - 1] >> shift) != (k >> shift))) {
            const uint32_t gi = gidx[p];
            gheads[gi] = p;
            gkey[gi] = k;
            const bool multi = p + 1 < n && (keys[p + 1] >> shift) == (k >> shift);
            const int qst = (int)(k & ((1ull << g.qst_bits) - 1));
            const int diag = (int)((k >> g.qst_bits) & ((1ull << g.diag_bits) - 1)) - g.diag_bias;
            const uint64_t pair = k >> (g.qst_bits + g.diag_bits);
            const int hd1 = (int)(pair & ((1ull << g.hd_bits) - 1));
            const int qi = (int)(pair >> g.hd_bits);
            const int sst = qst - diag;
            const uint32_t xt = (uint32_t)toff[g.c0 + hd1 - 1] + (uint32_t)sst;
            const uint32_t uq = (uint32_t)(qoff[g.qb0 + qi] - qa) + (uint32_t)qst;
            uint32_t flags = (multi ? 0u : kDescSingle) | ((qst == 0 || sst == 0) ? kDescNoLeft : 0u);
            if (multi && !(p + 2 < n && (keys[p + 2] >> shift) == (k >> shift))) {
                const uint32_t delta = (uint32_t)(keys[p + 1] & ((1ull << g.qst_bits) - 1)) - (uint32_t)qst;
                if (delta >= 1 && delta <= 32) flags |= kDescPair | ((delta - 1) << 27);
            }
            desc[gi] = make_uint2(xt, uq | flags);
            if (vals) grank[gi] = vals[p];  // k_xdrop folds the other hits of a chained group in
        }
    }
}

__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// 16 query bytes starting at byte 4*ws + bs of the aligned chunk pair (a, b); w1 / w2 = bits of ws
__device__ __forceinline__ void ung_qwindow(const uint4 &a, const uint4 &b, bool w1, bool w2, uint32_t bs8, uint32_t W[4]) {
    const uint32_t X[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t Y[7], Z[5];
#pragma unroll
    for (int j = 0; j < 7; j++) Y[j] = w1 ? X[j + 1] : X[j];
#pragma unroll
    for (int j = 0; j < 5; j++) Z[j] = w2 ? Y[j + 2] : Y[j];
#pragma unroll
    for (int j = 0; j < 4; j++) W[j] = __funnelshift_r(Z[j], Z[j + 1], bs8);
}
// the same window when the chunks come from a word-shifted copy of the view (ws = 0: only the byte shift is left)
__device__ __forceinline__ void ung_qwindow0(const uint4 &a, const uint4 &b, uint32_t bs8, uint32_t W[4]) {
    W[0] = __funnelshift_r(a.x, a.y, bs8);
    W[1] = __funnelshift_r(a.y, a.z, bs8);
    W[2] = __funnelshift_r(a.z, a.w, bs8);
    W[3] = __funnelshift_r(a.w, b.x, bs8);
}
// shl.b32 clamps shift amounts above 31 (result 0), which C's << does not promise
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, int sh) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(sh));
    return r;
}
// bytes [0, lo) of the chunk -> class `cls`
__device__ __forceinline__ void ung_mask_below(uint4 &t, int lo, uint32_t cls) {
    uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t keep = shl_clamp(0xffffffffu, max(8 * lo - 32 * k, 0));  // bytes >= lo
        w[k] = (w[k] & keep) | ((cls * 0x01010101u) & ~keep);
    }
    t = make_uint4(w[0], w[1], w[2], w[3]);
}
// bytes [hi, 16) of the chunk -> class `cls`
__device__ __forceinline__ void ung_mask_from(uint4 &t, int hi, uint32_t cls) {
    uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t repl = shl_clamp(0xffffffffu, max(8 * hi - 32 * k, 0));  // bytes >= hi
        w[k] = (w[k] & ~repl) | ((cls * 0x01010101u) & repl);
    }
    t = make_uint4(w[0], w[1], w[2], w[3]);
}
// predicated 16-byte loads (no branch around them)
__device__ __forceinline__ void ldg128_if(uint4 &r, const uint4 *p, int pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w)
                 : "l"(p), "r"(pred));
}

// 16 chained X-drop steps on table entries e0..e15 (see the header comment):
//   dd = max(dd + e, 0)   unpredicated (one VIADDMNMX, alu pipe)
//   if alive: v -= e; d = dd   (d: value at the last live step; the copy is written as dd * one so that
//                               ptxas keeps it a predicated IMAD on the fma pipe instead of an alu SEL)
//   alive &= dd < 31 * 8192
// Per step: alu PRMT + VIADDMNMX + ISETP, fma IMAD (table address) + IMAD.IADD + IMAD, one LDS.
#define SO_XS(n)                                                                                         \
    "add.s32 t, dd, %" #n ";\n\tmax.s32 dd, t, 0;\n\t@p mad.lo.s32 %1, dd, %3, 0;\n\t@p sub.s32 %0, %0, %" #n \
    ";\n\tsetp.lt.and.s32 p, dd, 253952, p;\n\t"
__device__ __forceinline__ void ung_steps16(int &v, int &d, int &alive, int one, const int (&e)[16]) {
    asm("{\n\t.reg .pred p;\n\t.reg .s32 t, dd;\n\tsetp.ne.s32 p, %2, 0;\n\tmov.s32 dd, %1;\n\t"
        SO_XS(4) SO_XS(5) SO_XS(6) SO_XS(7) SO_XS(8) SO_XS(9) SO_XS(10) SO_XS(11)
        SO_XS(12) SO_XS(13) SO_XS(14) SO_XS(15) SO_XS(16) SO_XS(17) SO_XS(18) SO_XS(19)
        "selp.s32 %2, 1, 0, p;\n\t}"
        : "+r"(v), "+r"(d), "+r"(alive)
        : "r"(one), "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]),
          "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]));
}
#undef SO_XS

enum { kXdBuf = 48, kXdBufBytes = 16 * kXdBuf * (8 + 2) + 16 * 32 };  // per 16 warps
// the staging buffers are addressed with 32-bit shared-window addresses (generic 64-bit pointer arithmetic on them
// cost 2.4 % of the kernel's instructions)
__device__ __forceinline__ void sts_u64(uint32_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint64_t lds_u64(uint32_t a) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}  // per-warp record buffer of k_xdrop<FAST> (16 warps per CTA)
// warp-collective: the staged records go to their queries' regions, one global atomic per query present
__device__ __forceinline__ void xd_flush(uint32_t wrec, uint32_t wqi, int cnt, uint64_t *__restrict__ creg,
                                         size_t ccap, uint32_t *__restrict__ qcount, uint32_t *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int i = i0 + lane;
        const bool v = i < cnt;
        const unsigned act = __ballot_sync(0xffffffffu, v);
        if (v) {
            const uint64_t r = lds_u64(wrec + 8u * (uint32_t)i);
            const uint32_t q = lds_u16(wqi + 2u * (uint32_t)i);
            const unsigned peers = __match_any_sync(act, q);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&qcount[q], (uint32_t)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            const uint32_t pos = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if (pos < ccap)
                creg[(size_t)q * ccap + pos] = r;
            else
                atomicExch(flags, 1u);  // more passing diagonals than the region holds: the general path redoes the block
        }
    }
    __syncwarp();
}

// FAST = the sync-free cell path: descriptors are indexed by HIT position (one per sorted hit; hits that are not the
// head of their diagonal group carry kDescSkipX and are skipped), the number of hits comes from device memory
// (counters[0], written by k_unit_scan) and chains walk the sorted cell-local keys `ssub` up to the next head bit.
// kWarps = warps per CTA: 16 (two CTAs per SM, each with its own score table) or 32 (one CTA per SM: one table, which
// leaves ~100 KB of the SM to the L1 instead of 29 KB).  kQS = shifted copies of the query view (0: one view, the
// window is multiplexed out of two aligned chunks; 4: word-shifted copies; 16: byte-shifted copies).
template <int kRefill, bool FAST, int kWarps = 16, int kQS = 0>
__global__ void __launch_bounds__(kWarps * 32, kWarps == 16 ? 2 : 1) k_xdrop(const uint2 *__restrict__ desc, uint32_t G_,
                                                  const uint32_t *__restrict__ ssub,
                                                  const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                  uint32_t nhits, const uint32_t *__restrict__ gheads, int qst_bits,
                                                  const uint4 *__restrict__ T4, uint32_t toffF, uint32_t toffR, uint32_t R,
                                                  const uint4 *__restrict__ Q4, uint32_t qoffF, uint32_t qoffR, uint32_t Lq,
                                                  uint32_t *__restrict__ gscore, uint32_t *__restrict__ grank,
                                                  unsigned long long *__restrict__ counters, int one,
                                                  const uint32_t *__restrict__ cellid, int diag_bits,
                                                  uint64_t *__restrict__ creg, size_t ccap, uint32_t *__restrict__ qcount,
                                                  uint32_t *__restrict__ flags, uint32_t qstride4) {
    extern __shared__ int s_tab[];  // [(ct << 5 | cq)][lane]; FAST: + per-warp record buffers
    const uint32_t G = FAST ? (uint32_t)counters[0] : G_;
    // FAST: groups that pass (score >= 25, self.min: fsearch.py:2224, 2707) leave one record
    //   [ target + 1 | qst | diagonal + bias | score (20 bits) ]
    // in their query's region; records are staged per warp in shared memory and flushed kXdBuf / 3 at a time, so
    // a warp issues one global atomic per ~30 records instead of one per passing group (consecutive groups belong
    // to the same query: per-group atomics would serialise on one address)
    // (recomputed where they are used, in the service block only: holding them costs registers in the step loop)
#define XD_SBASE ((uint32_t)__cvta_generic_to_shared(s_tab) + (uint32_t)kUngTabBytes)
#define XD_WREC (XD_SBASE + (threadIdx.x >> 5) * (kXdBuf * 8u))
#define XD_WQI (XD_SBASE + (uint32_t)kWarps * (kXdBuf * 8u) + (threadIdx.x >> 5) * (kXdBuf * 2u))
#define XD_WIDX (XD_SBASE + (uint32_t)kWarps * (kXdBuf * 10u) + (threadIdx.x >> 5) * 32u)
    int wcount = 0;  // warp-uniform
    for (int k = threadIdx.x; k < kUngRows * 32 * 32; k += blockDim.x) s_tab[k] = g_xdrop_tab[k >> 5];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t lanebase = (uint32_t)__cvta_generic_to_shared(s_tab) + (uint32_t)lane * 4u;
    const uint32_t qmask = (1u << qst_bits) - 1u;
    constexpr int kNoLimit = 1 << 30;
    const unsigned long long kBatch = 512;  // group indices taken per atomic
    int alive = 0;
    bool has = false, fin = false, noleft = false, multi = false, chained = false, pairm = false;
    bool w1 = false, w2 = false;
    int phase = 0, acc = 0, v = 0, d = 0, lim = kNoLimit;
    int qcur = 0, lo = 0;                  // chains: qst of the seed being extended, max_qed of the last segment
    uint32_t gi = 0, xt = 0, uq = 0, tci = 0, qci = 0, bs8 = 0, e = 0;
    uint4 tc = make_uint4(0, 0, 0, 0), carry = make_uint4(0, 0, 0, 0);
    uint32_t W[4] = {0, 0, 0, 0};
    unsigned int steps = 0, nmulti = 0, ngroups = 0;
    uint32_t pool_next = 0, pool_end = 0;  // warp-uniform
    bool exhausted = false;                // warp-uniform
    if (FAST && G == 0) return;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !alive);
        if (__popc(idle) >= kRefill) {
            // ---- service: finished phases, next seeds / groups, first loads of the next phase
            bool setup = false, pass = false;
            if (!alive && has) {
                const int nst = (-v) & 8191;
                acc += (v + d + 8191) >> 13;
                steps += (unsigned)nst;
                bool seed_done = true;
                if (phase == 0) {
                    // later seeds: the left extension stops above the previous segment's max_qed
                    lim = chained ? qcur - 1 - lo : kNoLimit;
                    // max_qed: query index of the first step that reached the right maximum (fsearch.py:2470-2474)
                    lo = qcur + max(nst - (d & 8191), 1) - 1;
                    if (!noleft && lim > 0) {
                        phase = 1;
                        setup = true;
                        seed_done = false;
                    }
                }
                if (seed_done) {
                    bool more = false;
                    if (multi && pairm) {
                        // two-seed chain described by k_group_desc: e = qst distance of the second seed
                        if (!chained && (int)e > lo) {
                            xt += e, uq += e;
                            qcur = (int)e;
                            more = true;
                        }
                    } else if (multi) {
                        // next seed of the chain: seeds at or below max_qed extend nothing and leave it unchanged
                        for (;;) {
                            e++;
                            if (e >= (FAST ? G : nhits)) break;
                            uint32_t k2lo;
                            if (FAST) {
                                k2lo = ssub[e];
                                if (k2lo & kHeadBit) break;
                            } else {
                                const uint64_t k2 = keys[e];
                                if (((k2 ^ keys[e - 1]) >> qst_bits) != 0) break;
                                if (vals) grank[gi] = min(grank[gi], vals[e]);
                                k2lo = (uint32_t)k2;
                            }
                            const int qst = (int)(k2lo & qmask);
                            if (qst <= lo) continue;
                            const uint32_t delta = (uint32_t)(qst - qcur);
                            xt += delta, uq += delta;
                            qcur = qst;
                            more = true;
                            break;
                        }
                    }
                    if (more) {
                        phase = 0, noleft = false, chained = true, lim = kNoLimit;
                        setup = true;
                    } else {
                        if (FAST)
                            pass = acc >= 25;
                        else
                            gscore[gi] = (uint32_t)acc;
                        has = false;
                    }
                }
            }
            if (FAST) {
                const unsigned pm = __ballot_sync(0xffffffffu, pass);
                if (pm) {
                    if (pass) {
                        const uint32_t cid = cellid[gi], key = ssub[gi];
                        const uint32_t rk = ((key & qmask) << diag_bits) | ((key & ~kHeadBit) >> qst_bits);
                        const int slot = wcount + __popc(pm & ((1u << lane) - 1u));
                        sts_u64(XD_WREC + 8u * (uint32_t)slot,
                                ((uint64_t)(cid & 0xffffu) << (qst_bits + diag_bits + 20)) | ((uint64_t)rk << 20) | (uint64_t)(uint32_t)acc);
                        sts_u16(XD_WQI + 2u * (uint32_t)slot, cid >> 16);
                    }
                    wcount += __popc(pm);
                    __syncwarp();
                    if (wcount > kXdBuf - 32) {
                        xd_flush(XD_WREC, XD_WQI, wcount, creg, ccap, qcount, flags);
                        wcount = 0;
                    }
                }
            }
            if (FAST) {
                // The pool is a range of HIT positions; 32 consecutive descriptors are fetched at a time (one coalesced
                // load), the heads among them go to the lanes that need a group in order (head rank = need rank, the
                // holder's lane index travels through a 32-byte shared-memory table), positions behind the last head
                // that found a taker stay in the pool.  Skipped (non-head) entries cost nothing.
                unsigned need = __ballot_sync(0xffffffffu, !has && !fin);
                for (int tries = 0; need != 0u && tries < 3; tries++) {
                    if (pool_next == pool_end && !exhausted) {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(counters + 2, kBatch);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (base >= G)
                            exhausted = true;
                        else {
                            pool_next = (uint32_t)base;
                            pool_end = (uint32_t)min((unsigned long long)G, base + kBatch);
                            if (pool_next + (uint32_t)lane * 16u < pool_end)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(desc + pool_next + lane * 16));
                        }
                    }
                    const bool needer = !has && !fin;
                    if (pool_next == pool_end) {  // nothing left anywhere
                        if (needer && exhausted) fin = true;
                        break;
                    }
                    uint2 ds = make_uint2(kDescSkipX, 0u);
                    if (pool_next + (uint32_t)lane < pool_end) ds = desc[pool_next + lane];
                    const bool head = ds.x != kDescSkipX;
                    const unsigned hm = __ballot_sync(0xffffffffu, head);
                    const int nh = __popc(hm), nn = __popc(need);
                    if (head) sts_u8(XD_WIDX + (uint32_t)__popc(hm & ((1u << lane) - 1u)), (uint32_t)lane);
                    __syncwarp();
                    const int nr = __popc(need & ((1u << lane) - 1u));
                    const bool take = needer && nr < nh;
                    const int src = take ? (int)lds_u8(XD_WIDX + (uint32_t)nr) : lane;
                    const uint32_t adv = nh <= nn ? min(32u, pool_end - pool_next) : lds_u8(XD_WIDX + (uint32_t)(nn - 1)) + 1u;
                    const uint32_t dx = __shfl_sync(0xffffffffu, ds.x, src), dy = __shfl_sync(0xffffffffu, ds.y, src);
                    __syncwarp();
                    if (take) {
                        gi = pool_next + (uint32_t)src;
                        xt = dx;
                        uq = dy & 0xffffffu;
                        noleft = (dy & kDescNoLeft) != 0;
                        multi = (dy & kDescSingle) == 0;
                        qcur = 0;  // chains only use qst differences: the first seed of a described pair counts from 0
                        pairm = multi && (dy & kDescPair) != 0;
                        if (multi) nmulti++;
                        if (pairm)
                            e = (dy >> 27) + 1u;
                        else if (multi) {
                            e = gi;
                            qcur = (int)(ssub[e] & qmask);
                        }
                        phase = 0, acc = 0, chained = false, lim = kNoLimit;
                        has = true;
                        setup = true;
                        ngroups++;
                    }
                    pool_next += adv;
                    need = __ballot_sync(0xffffffffu, !has && !fin);
                }
            } else {
            unsigned need = __ballot_sync(0xffffffffu, !has && !fin);
                for (int tries = 0; need != 0u && tries < 1; tries++, need = __ballot_sync(0xffffffffu, !has && !fin)) {
                    if (pool_next == pool_end && !exhausted) {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(counters + 2, kBatch);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (base >= G)
                            exhausted = true;
                        else {
                            pool_next = (uint32_t)base;
                            pool_end = (uint32_t)min((unsigned long long)G, base + kBatch);
                            // the batch's descriptors (512 x 8 B = 32 lines) are pulled into L2 ahead of their use
                            if (pool_next + (uint32_t)lane * 16u < pool_end)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(desc + pool_next + lane * 16));
                        }
                    }
                    if (!has && !fin) {
                        const uint32_t mine = pool_next + __popc(need & ((1u << lane) - 1));
                        if (mine < pool_end) {
                            const uint2 ds = desc[mine];
                            {
                                gi = mine;
                                xt = ds.x;
                                uq = ds.y & 0xffffffu;
                                noleft = (ds.y & kDescNoLeft) != 0;
                                multi = (ds.y & kDescSingle) == 0;
                                qcur = 0;  // chains only use qst differences: the first seed of a described pair counts from 0
                                pairm = multi && (ds.y & kDescPair) != 0 && vals == nullptr;
                                if (multi) nmulti++;
                                if (pairm)
                                    e = (ds.y >> 27) + 1u;
                                else if (multi) {
                                    e = gheads[mine];
                                    qcur = (int)((uint32_t)keys[e] & qmask);
                                }
                                phase = 0, acc = 0, chained = false, lim = kNoLimit;
                                has = true;
                                setup = true;
                            }
                        } else if (exhausted)
                            fin = true;  // otherwise the pool is refilled at the next service
                    }
                    pool_next = min(pool_end, pool_next + (uint32_t)__popc(need));
                }
            }
            if (__all_sync(0xffffffffu, fin)) break;
            if (setup) {
                const uint32_t tpos = (phase ? toffR + (R - xt) : toffF + xt);
                const int o = (int)(tpos & 15u);
                const uint32_t qpos = (phase ? qoffR + (Lq - uq) : qoffF + uq) - (uint32_t)o;
                tci = tpos >> 4;
                tc = T4[tci];
                if (kQS == 16) {
                    // copy (qpos & 15) holds the wanted bytes 16-byte aligned: one aligned load per iteration
                    qci = (qpos & 15u) * qstride4 + (qpos >> 4);
                    const uint4 a = Q4[qci];
                    W[0] = a.x, W[1] = a.y, W[2] = a.z, W[3] = a.w;
                    tci += 1, qci += 1;
                } else if (kQS == 4) {
                    // copy ((qpos >> 2) & 3) holds the wanted words 16-byte aligned: only the byte shift is left
                    qci = ((qpos >> 2) & 3u) * qstride4 + (qpos >> 4);
                    bs8 = (qpos & 3u) * 8u;
                    const uint4 a = Q4[qci];
                    carry = Q4[qci + 1];
                    tci += 1, qci += 2;
                    ung_qwindow0(a, carry, bs8, W);
                } else {
                    qci = qpos >> 4;
                    w1 = (qpos & 4u) != 0, w2 = (qpos & 8u) != 0, bs8 = (qpos & 3u) * 8u;
                    const uint4 a = Q4[qci];
                    carry = Q4[qci + 1];
                    tci += 1, qci += 2;
                    ung_qwindow(a, carry, w1, w2, bs8, W);
                }
                ung_mask_below(tc, o, kUngSkip);  // bytes before the start score nothing
                if (lim < 16 - o) ung_mask_from(tc, o + lim, kUngStop);
                lim -= 16 - o;
                v = 0, d = 0;
                alive = 1;
            }
        }
        // ---- 16 extension steps on the current chunk: table lookups first (they consume the chunk
        // registers), then the loads of the next chunk, then the dependent chain
        {
            int ev[16];
            const uint32_t tw[4] = {tc.x, tc.y, tc.z, tc.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                ev[4 * j + 0] = lds_s32(lanebase + (prmt_b32(W[j], tw[j], 0xCC40u) << 4));
                ev[4 * j + 1] = lds_s32(lanebase + (prmt_b32(W[j], tw[j], 0xDD51u) << 4));
                ev[4 * j + 2] = lds_s32(lanebase + (prmt_b32(W[j], tw[j], 0xEE62u) << 4));
                ev[4 * j + 3] = lds_s32(lanebase + (prmt_b32(W[j], tw[j], 0xFF73u) << 4));
            }
            uint4 nq = kQS == 16 ? make_uint4(W[0], W[1], W[2], W[3]) : carry;
            ldg128_if(tc, T4 + tci, alive);
            ldg128_if(nq, Q4 + qci, alive);
            ung_steps16(v, d, alive, one, ev);
            if (__any_sync(0xffffffffu, alive && lim < 16)) {
                if (lim < 16) ung_mask_from(tc, lim, kUngStop);
            }
            lim -= 16;
            if (kQS == 16) {
                W[0] = nq.x, W[1] = nq.y, W[2] = nq.z, W[3] = nq.w;
            } else if (kQS == 4) {
                ung_qwindow0(carry, nq, bs8, W);
                carry = nq;
            } else {
                ung_qwindow(carry, nq, w1, w2, bs8, W);
                carry = nq;
            }
            tci += 1, qci += 1;
        }
    }
    unsigned long long st64 = steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) st64 += __shfl_xor_sync(0xffffffffu, st64, o);
    if (lane == 0 && st64) atomicAdd(counters + 1, st64);
    nmulti = __reduce_add_sync(0xffffffffu, nmulti);
    if (lane == 0 && nmulti) atomicAdd(counters + 3, (unsigned long long)nmulti);  // statistic
    if (FAST) {
        if (wcount) xd_flush(XD_WREC, XD_WQI, wcount, creg, ccap, qcount, flags);
        ngroups = __reduce_add_sync(0xffffffffu, ngroups);
        if (lane == 0 && ngroups) atomicAdd(counters + 4, (unsigned long long)ngroups);  // statistic: diagonal groups
    }
}
#undef XD_SBASE
#undef XD_WREC
#undef XD_WQI
#undef XD_WIDX

// Pair selection over the PASSING groups only (score >= 25, self.min: fsearch.py:2224, 2707): `plist` holds
// their group indices in ascending order (ordered stream compaction, cub::DeviceSelect), ~8 % of all groups.
// One thread per passing group; the first passing group of a (query, target) pair folds the pair: best
// diagonal (first appearance wins ties), candidate order = first passing rank.  gkey = head key of every
// group.  The first-appearance rank of a group is grank[] when the hit ordinals were carried through the
// sort; on the keys-only path (one pattern, one alphabet) the ordinal order is (qst ascending, then the
// bucket's descending (sequence, position) order), so the rank is rebuilt from the key itself:
// qst | (max - sequence) | (max - sst).  Output slots are claimed with one atomic per block.
struct PassPred {
    const uint32_t *gscore;
    __host__ __device__ __forceinline__ bool operator()(const uint32_t &gi) const { return gscore[gi] >= 25u; }
};

__global__ void __launch_bounds__(256) k_pair_select(const uint32_t *__restrict__ plist, const uint32_t *__restrict__ pcount,
                                                     const uint64_t *__restrict__ gkey, BlockGeom g,
                                                     const uint32_t *__restrict__ gscore,
                                                     const uint32_t *__restrict__ grank, int rank_bits,
                                                     uint64_t *__restrict__ ckeys, uint64_t *__restrict__ cvals,
                                                     unsigned long long *__restrict__ counters) {
    __shared__ uint32_t s_wcount[8];
    __shared__ unsigned long long s_base;
    const uint32_t n = *pcount;
    const int pair_shift = g.qst_bits + g.diag_bits;
    const uint64_t dmask = (1ull << g.diag_bits) - 1, hdmask = (1ull << g.hd_bits) - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {  // block-uniform trip count
        const uint32_t i = i0 + threadIdx.x;
        bool head = false;
        uint64_t pair = 0;
        if (i < n) {
            pair = gkey[plist[i]] >> pair_shift;
            head = i == 0 || (gkey[plist[i - 1]] >> pair_shift) != pair;
        }
        int best_score = 0, best_diag = 0;
        uint64_t best_rank = ~0ull, first_rank = ~0ull;
        if (head) {
            for (uint32_t j = i; j < n; j++) {
                const uint32_t k = plist[j];
                const uint64_t kk = gkey[k];
                if ((kk >> pair_shift) != pair) break;
                const int sc = (int)gscore[k];
                const int dg = (int)((kk >> g.qst_bits) & dmask) - g.diag_bias;
                uint64_t rk;
                if (grank)
                    rk = grank[k];
                else {
                    const uint64_t qst = kk & ((1ull << g.qst_bits) - 1);
                    rk = (qst << (g.hd_bits + g.diag_bits)) | ((hdmask - (pair & hdmask)) << g.diag_bits) |
                         (dmask - (uint64_t)((int)qst - dg));
                }
                first_rank = min(first_rank, rk);
                if (sc > best_score || (sc == best_score && rk < best_rank)) {
                    best_score = sc;
                    best_rank = rk;
                    best_diag = dg;
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_wcount[warp] = (uint32_t)__popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) {
                const uint32_t c = s_wcount[w];
                s_wcount[w] = tot;
                tot += c;
            }
            s_base = tot ? atomicAdd(counters, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (head) {
            const unsigned long long o = s_base + s_wcount[warp] + __popc(m & ((1u << lane) - 1));
            const int hd1 = (int)(pair & hdmask);
            const int qi = (int)(pair >> g.hd_bits);
            // between the candidates of one query (distinct targets) qst | (max - sequence) already orders the first
        // ranks, so the sst part is dropped from the sort key (one radix pass less)
        ckeys[o] = ((uint64_t)qi << rank_bits) | (grank ? first_rank : first_rank >> g.diag_bits);
            // value: target ordinal (24 bits) | score (20 bits) | diagonal + bias (20 bits)
            cvals[o] = ((uint64_t)(uint32_t)(g.c0 + hd1 - 1) << 40) | ((uint64_t)(uint32_t)best_score << 20) |
                       (uint64_t)(uint32_t)(best_diag + kCandDiagBias);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Pair fold + candidate order of the cell path (fsearch.py:2696-2719), one CTA per query at a time.
// k_xdrop<FAST> left one record per PASSING diagonal group (score >= 25) in the query's region:
//   [ target + 1 | qst of the group's first seed | diagonal + bias | score (20 bits) ]
// With one pattern and one alphabet the reference's dict order of a query's groups is (qst ascending, then target
// descending, then sst descending = diagonal ascending), so inside one (query, target) pair it is (qst, diagonal).
//   1. the records are ordered by (target, qst, diagonal): counting sort on the top bits (about 8 targets per bin),
//      then every bin (a few records) is put in order by one thread;
//   2. the records of one target fold into one candidate: best score, the FIRST diagonal in dict order wins ties,
//      and the candidate's place in the reference's list is the dict rank of the target's first passing group.
//      A candidate is one 64-bit word, rank on top:
//        [ qst of the first passing group | hdmask - (target + 1) | score (20 bits) | best diagonal + bias ]
//   3. the words are ordered (= the reference's candidate list order, fsearch.py:2715-2719): counting sort on the
//      top bits of the rank, position inside a bin by counting the smaller words (bins are skewed here: ranks start
//      with qst), and written, converted to the block format (target << 40 | score << 20 | diagonal +
//      kCandDiagBias), behind the query's candidates of the earlier chunks (select.cu).
// Both sorts run in passes over bin ranges that fit the shared-memory buffer (one pass for a typical query).
// ---------------------------------------------------------------------------------------------
enum { kCandBins = 8192, kCandSmem = 9216, kCandThreads = 512 };

struct CandSmem {
    uint64_t f[kCandSmem];
    uint32_t bin[kCandBins + 1];
    uint32_t part[kCandThreads / 32];
    int q;
    uint32_t b1, nc, flag;
};

// exclusive scan of s.bin[0..kCandBins) in place (kCandBins / kCandThreads consecutive bins per thread)
__device__ __forceinline__ void cand_scan_bins(CandSmem &s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int per = kCandBins / kCandThreads;
    uint32_t v[per], sum = 0;
#pragma unroll
    for (int k = 0; k < per; k++) v[k] = s.bin[tid * per + k], sum += v[k];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s.part[warp] = inc;
    __syncthreads();
    if (tid < 32) {
        const uint32_t t0 = tid < kCandThreads / 32 ? s.part[tid] : 0u;
        uint32_t w = t0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        if (tid < kCandThreads / 32) s.part[tid] = w - t0;
    }
    __syncthreads();
    uint32_t run = s.part[warp] + inc - sum;
#pragma unroll
    for (int k = 0; k < per; k++) s.bin[tid * per + k] = run, run += v[k];
    __syncthreads();
}

// histogram of src[0..n) on (word >> sh) and its exclusive scan: s.bin[b] = first position of bin b, s.bin[kCandBins] = n
__device__ __forceinline__ void cand_histogram(CandSmem &s, const uint64_t *__restrict__ src, uint32_t n, int sh) {
    for (int b = threadIdx.x; b <= kCandBins; b += kCandThreads) s.bin[b] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += kCandThreads) atomicAdd(&s.bin[(uint32_t)(src[i] >> sh)], 1u);
    __syncthreads();
    cand_scan_bins(s);
    if (threadIdx.x == 0) s.bin[kCandBins] = n;
    __syncthreads();
}

// next pass: bins [b0, b1) whose elements fit the buffer (b1 > b0; a single bin above the buffer gives b1 = b0)
__device__ __forceinline__ uint32_t cand_plan(CandSmem &s, uint32_t b0) {
    if (threadIdx.x == 0) {
        const uint32_t base = s.bin[b0];
        uint32_t lo = b0, hi = kCandBins;  // largest b1 in (b0, kCandBins] with bin[b1] - base <= kCandSmem
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (s.bin[mid] - base <= (uint32_t)kCandSmem)
                lo = mid;
            else
                hi = mid - 1;
        }
        s.b1 = lo;
    }
    __syncthreads();
    return s.b1;
}

__global__ void __launch_bounds__(kCandThreads) k_cand_sort(const uint64_t *__restrict__ creg, size_t ccap,
                                                           const uint32_t *__restrict__ qcount, int nq, BlockGeom g,
                                                           const uint64_t *__restrict__ qoff,
                                                           uint64_t *__restrict__ gbuf, uint64_t *__restrict__ bvals,
                                                           size_t bcap, uint32_t *__restrict__ bcount, int bq0,
                                                           uint32_t *__restrict__ next, uint32_t *__restrict__ flags,
                                                           unsigned long long *__restrict__ counters) {
    extern __shared__ __align__(16) unsigned char cs_raw[];
    CandSmem &s = *reinterpret_cast<CandSmem *>(cs_raw);
    const int tid = threadIdx.x;
    // (other CTAs of this kernel may raise the flag at any time: it is read by one thread and broadcast)
    if (tid == 0) s.flag = flags[0];
    __syncthreads();
    if (s.flag) return;
    const int Lb = g.qst_bits + g.diag_bits, esh = 20 + g.diag_bits, rank_bits = g.qst_bits + g.hd_bits;
    const int sh1 = 20 + Lb + max(0, g.hd_bits - 13);  // records: bin = top 13 bits of (target + 1)
    // words: bin = top 13 bits of the rank below the query's length (the rank starts with a query position)
    const uint32_t dmask = (1u << g.diag_bits) - 1u, hdmask = (1u << g.hd_bits) - 1u;
    uint64_t *C = gbuf + (size_t)blockIdx.x * ccap;    // this CTA's candidate words
    for (;;) {
        if (tid == 0) s.q = (int)atomicAdd(next, 1u);
        __syncthreads();
        const int q = s.q;
        __syncthreads();
        if (q >= nq) break;
        const uint32_t n = qcount[q];
        if (n == 0) continue;
        if ((size_t)n > ccap) continue;  // region overflow: k_xdrop raised the redo flag
        const uint64_t *src = creg + (size_t)q * ccap;
        int qlb = 1;
        {
            const uint32_t ql = (uint32_t)(qoff[g.qb0 + q + 1] - qoff[g.qb0 + q]);
            while (qlb < g.qst_bits && (1u << qlb) <= ql) qlb++;
        }
        const int sh2 = esh + max(0, qlb + g.hd_bits - 13);
        // ---- 1 + 2: records by target bin, fold per target -> candidate words in C
        cand_histogram(s, src, n, sh1);
        if (tid == 0) s.nc = 0;
        for (uint32_t b0 = 0; b0 < (uint32_t)kCandBins;) {
            const uint32_t b1 = cand_plan(s, b0);
            if (b1 == b0) {  // one bin above the buffer (thousands of passing diagonals on 8 targets): general path
                if (tid == 0) atomicExch(flags, 1u);
                break;
            }
            const uint32_t base = s.bin[b0];
            __syncthreads();
            for (uint32_t i = tid; i < n; i += kCandThreads) {
                const uint64_t e = src[i];
                const uint32_t b = (uint32_t)(e >> sh1);
                if (b >= b0 && b < b1) s.f[atomicAdd(&s.bin[b], 1u) - base] = e;
            }
            __syncthreads();
            // s.bin[b] is now the END of bin b.  One thread per record: the records of its target all lie in its bin
            // (a few records); the one that comes first in dict order (smallest qst | diagonal) folds the target
            const uint32_t m1 = s.bin[b1 - 1] - base;
            for (uint32_t i = tid; i < m1; i += kCandThreads) {
                const uint64_t e0 = s.f[i];
                const uint32_t b = (uint32_t)(e0 >> sh1);
                const uint32_t lo = (b == b0 ? base : s.bin[b - 1]) - base, hi = s.bin[b] - base;
                const uint32_t t = (uint32_t)(e0 >> (20 + Lb));
                const uint32_t first = (uint32_t)(e0 >> 20) & ((1u << Lb) - 1u);
                uint32_t best = (uint32_t)e0 & 0xfffffu, brk = first;
                bool head = true;
                for (uint32_t k = lo; k < hi; k++) {
                    const uint64_t e = s.f[k];
                    if (k == i || (uint32_t)(e >> (20 + Lb)) != t) continue;
                    const uint32_t rk = (uint32_t)(e >> 20) & ((1u << Lb) - 1u), sc = (uint32_t)e & 0xfffffu;
                    if (rk < first) {
                        head = false;
                        break;
                    }
                    if (sc > best || (sc == best && rk < brk)) best = sc, brk = rk;
                }
                if (head) {
                    const uint64_t rank = ((uint64_t)(first >> g.diag_bits) << g.hd_bits) | (uint64_t)(hdmask - t);
                    C[atomicAdd(&s.nc, 1u)] = (rank << esh) | ((uint64_t)best << g.diag_bits) | (uint64_t)(brk & dmask);
                }
            }
            __syncthreads();
            b0 = b1;
        }
        if (tid == 0) s.flag = flags[0];
        __syncthreads();
        const uint32_t nc = s.nc;
        const uint32_t stop = s.flag;
        __syncthreads();
        if (stop) continue;
        // ---- 3: words by rank -> the block's list of the query
        const uint32_t cbase = bcount[bq0 + q];
        if ((size_t)cbase + nc > bcap) {
            if (tid == 0) atomicExch(flags + 1, 2u);
            continue;
        }
        uint64_t *dst = bvals + (size_t)(bq0 + q) * bcap + cbase;
        cand_histogram(s, C, nc, sh2);
        for (uint32_t b0 = 0; b0 < (uint32_t)kCandBins;) {
            const uint32_t b1 = cand_plan(s, b0);
            if (b1 == b0) {
                if (tid == 0) atomicExch(flags, 1u);
                break;
            }
            const uint32_t base = s.bin[b0];
            __syncthreads();
            for (uint32_t i = tid; i < nc; i += kCandThreads) {
                const uint64_t e = C[i];
                const uint32_t b = (uint32_t)(e >> sh2);
                if (b >= b0 && b < b1) s.f[atomicAdd(&s.bin[b], 1u) - base] = e;
            }
            __syncthreads();
            const uint32_t m = s.bin[b1 - 1] - base;  // elements of the pass
            for (uint32_t i = tid; i < m; i += kCandThreads) {
                const uint64_t e = s.f[i];
                const uint32_t b = (uint32_t)(e >> sh2);
                const uint32_t lo = (b == b0 ? base : s.bin[b - 1]) - base, hi = s.bin[b] - base;
                uint32_t r = 0;
                for (uint32_t j = lo; j < hi; j++) r += s.f[j] < e ? 1u : 0u;
                const uint32_t t = hdmask - (uint32_t)((e >> esh) & hdmask);  // target + 1 inside the chunk
                const uint32_t score = (uint32_t)((e >> g.diag_bits) & 0xfffffu);
                const int diag = (int)((uint32_t)e & dmask) - g.diag_bias;
                dst[base + lo + r] = ((uint64_t)(uint32_t)(g.c0 + (int)t - 1) << 40) | ((uint64_t)score << 20) |
                                     (uint64_t)(uint32_t)(diag + kCandDiagBias);
            }
            __syncthreads();
            b0 = b1;
        }
        if (tid == 0) {
            bcount[bq0 + q] = cbase + nc;
            atomicAdd(counters + 5, (unsigned long long)nc);  // statistic: candidates
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_classify(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = c_code2[in[i]];
}

int classify_residues(so_ctx *c, const uint8_t *d_in, uint8_t *d_out, size_t n) {
    // the byte -> class table does not depend on the search parameters: one upload per device
    static bool uploaded[64] = {};
    if (c->device < 0 || c->device >= 64 || !uploaded[c->device]) {
        uint8_t code[256];
        make_code_table(code);
        SO_CUDA(cudaMemcpyToSymbol(c_code2, code, sizeof code));
        if (c->device >= 0 && c->device < 64) uploaded[c->device] = true;
    }
    if (n == 0) return SO_OK;
    k_classify<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_in, d_out, n);
    SO_CUDA(cudaGetLastError());
    c->stats.kernel_launches += 1;
    return SO_OK;
}

__global__ void k_query_bounds(const uint64_t *__restrict__ ckeys, uint32_t n, int nq, int rank_bits,
                               uint32_t *__restrict__ bounds) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > nq) return;
    // first index whose query field is >= q
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint32_t)(ckeys[mid] >> rank_bits) < (uint32_t)q)
            lo = mid + 1;
        else
            hi = mid;
    }
    bounds[q] = lo;
}

struct Widen {
    __host__ __device__ __forceinline__ uint64_t operator()(const uint32_t &v) const { return (uint64_t)v; }
};

static int bits_for(uint64_t maxval) {  // bits needed to hold values 0..maxval
    int b = 1;
    while (b < 63 && (1ull << b) <= maxval) b++;
    return b;
}

enum { SC_SLOTOFF = 12, SC_ST, SC_CNT, SC_OUT, SC_KA, SC_KB, SC_VA, SC_VB, SC_TMP, SC_CKA, SC_CKB, SC_CVA, SC_CVB, SC_MISC,
       SC_GIDX, SC_GHEAD, SC_GSCORE, SC_GRANK, SC_QUNG, SC_DESC, SC_MLIST, SC_GKEY, SC_PLIST, SC_CELLLOC, SC_UNIT, SC_SUB, SC_SSUB,
       SC_WLIST, SC_CREG, SC_QCOUNT, SC_GBUF, SC_CTL, SC_QSHIFT, SC_COUNT_ };
static_assert(SC_COUNT_ <= 64, "scratch slots");

static const uint64_t kHitCap = 300000000ull;  // seed hits per sub-block (memory: 24 B each)

static int g_xdrop_refill = 24;
static int g_xdrop_warps = 32;  // warps per k_xdrop<FAST> CTA: 16 (two CTAs per SM) or 32 (one, sharing one score table) (SO_XDROP_WARPS)
static int g_xdrop_qs = 16;     // shifted copies of the query view: 0, 4 or 16 (SO_XDROP_QSHIFT)
static int g_cell_span = 1;  // cells of up to kCellWarp hits: k_cell_span (1) or k_cell_small + k_cell_warp (0) (SO_CELL_SPAN)
static int g_xdrop_ctas = 2;  // resident k_xdrop<FAST> CTAs per SM (SO_XDROP_CTAS: tuning hook; 1 leaves half the SM to the other lane's kernels)
static uint32_t g_cell_max = kCellMax;  // cells above this many hits send the block to the general path (SO_CELL_MAX: test hook)
static uint32_t g_cell_split = 0;  // target ranges per query in the cell passes; 0 = as few as fit shared memory (SO_CELL_SPLIT)  // idle lanes that trigger a refill in k_xdrop (SO_XDROP_REFILL: tuning hook)

int upload_search_config(so_ctx *c) {
    const char *e = getenv("SO_XDROP_REFILL");
    g_xdrop_refill = e ? atoi(e) : 24;
    g_cell_span = 1;
    if (const char *cs2 = getenv("SO_CELL_SPAN")) g_cell_span = atoi(cs2) != 0;
    g_xdrop_ctas = 2;
    if (const char *xc = getenv("SO_XDROP_CTAS")) g_xdrop_ctas = std::max(1, std::min(2, atoi(xc)));
    // default: one 1024-thread CTA per SM (one score table: 122 KB of shared memory, which leaves ~124 KB of L1 instead
    // of 29 KB) and 16 byte-shifted copies of the query view (measured: 64 -> 55 ms per 4096 queries; the copies alone,
    // with two CTAs per SM, thrash the small L1: 85 ms).  Other refill thresholds use the two-CTA kernel.
    g_xdrop_warps = e && atoi(e) != 24 ? 16 : 32, g_xdrop_qs = e && atoi(e) != 24 ? 0 : 16;
    if (const char *xw = getenv("SO_XDROP_WARPS")) g_xdrop_warps = atoi(xw) == 32 ? 32 : 16;
    if (const char *xq = getenv("SO_XDROP_QSHIFT")) g_xdrop_qs = atoi(xq) == 16 ? 16 : atoi(xq) == 4 ? 4 : 0;
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<24, true, 16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<24, true, 16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<24, true, 32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + 2 * kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<24, true, 32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + 2 * kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<24, true, 32, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + 2 * kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<16, true, 32, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + 2 * kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<20, true, 32, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + 2 * kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<28, true, 32, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + 2 * kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<32, true, 32, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + 2 * kXdBufBytes));
    g_cell_split = 0;
    if (const char *cs = getenv("SO_CELL_SPLIT")) g_cell_split = (uint32_t)std::max(0, std::min(8, atoi(cs)));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<20, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<24, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<20, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_xdrop<24, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUngTabBytes + kXdBufBytes));
    SO_CUDA(cudaFuncSetAttribute(k_cell_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CellBlockSmem)));
    SO_CUDA(cudaFuncSetAttribute(k_cell_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    SO_CUDA(cudaFuncSetAttribute(k_cell_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    SO_CUDA(cudaFuncSetAttribute(k_cand_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CandSmem)));
    g_cell_max = kCellMax;
    if (const char *cm = getenv("SO_CELL_MAX")) g_cell_max = (uint32_t)std::max(1, std::min((int)kCellMax, atoi(cm)));  // test hook
    return upload_cfg(c->P);
}

void merge_lane_stats(so_ctx *c) {
    for (int l = 0; l < so_ctx::kMaxLanes; l++) {
        so_stats &a = c->stats_lane[l], &d = c->stats;
        d.seed_hits += a.seed_hits, d.groups += a.groups, d.candidates += a.candidates, d.ungap_steps += a.ungap_steps;
        d.kernel_launches += a.kernel_launches, d.lib_launches += a.lib_launches;
        d.ms_seed += a.ms_seed, d.ms_sort += a.ms_sort, d.ms_ungap += a.ms_ungap, d.ms_select += a.ms_select;
        d.ms_ungap_kernel += a.ms_ungap_kernel, d.multi_groups += a.multi_groups, d.redo_blocks += a.redo_blocks;
        d.h2d_bytes += a.h2d_bytes, d.d2h_bytes += a.d2h_bytes;
        c->prof.d2h_ms += c->d2h_ms_lane[l];
        memset(&a, 0, sizeof a);
        c->d2h_ms_lane[l] = 0;
    }
}

// The caller uploads the search configuration (upload_search_config) once before the first call.
int chunk_candidates(so_ctx *c, const ChunkIndex &ix, i64 q_begin, i64 q_end, PackedCands &out, int lane, BlockStore *bs,
                     int bs_q0) {
    so::DBuf<uint8_t> *scratch = c->lane_scratch(lane);
    cudaEvent_t *ev = c->lane_ev(lane);
    so_stats &stats = c->stats_lane[lane];
    const Params &P = c->P;
    const i64 nq_total = q_end - q_begin;
    out.offsets.assign((size_t)nq_total + 1, 0);
    out.n = 0;
    if (nq_total <= 0) return SO_OK;
    if (ix.n_seeds == 0) return SO_OK;
    int rc;
    const int AS = (int)(P.alphabets.size() * P.patterns.size());
    const i64 M = ix.c1 - ix.c0;
    i64 sub = c->sub_block > 0 ? c->sub_block : 256;
    if (const char *e = getenv("SO_SUB_BLOCK0")) sub = std::max<i64>(1, atoll(e));  // tuning hook: first sub-block
    i64 b0 = q_begin;
    cudaStream_t st = c->lane_stream(lane);
    while (b0 < q_end) {
        i64 b1 = std::min<i64>(q_end, b0 + sub);
        const int nq = (int)(b1 - b0);
        // slot offsets (queries shorter than the shortest seed span get no slots: the reference
        // indexes out of bounds there (fsearch.py:2648-2652); defined as "no hits")
        std::vector<uint32_t> slot_off((size_t)nq + 1, 0);
        uint32_t maxql = 1;
        uint64_t tot = 0;
        for (int k = 0; k < nq; k++) {
            uint64_t L = c->q_off[(size_t)(b0 + k + 1)] - c->q_off[(size_t)(b0 + k)];
            uint64_t n = (L >= (uint64_t)P.mink) ? L * (uint64_t)AS : 0;
            if (n) maxql = std::max<uint32_t>(maxql, (uint32_t)L);
            tot += n;
            if (tot > 0x7fffff00ull) break;
            slot_off[(size_t)k + 1] = (uint32_t)tot;
        }
        if (tot > 0x7fffff00ull) {
            if (nq == 1) {
                set_error("query %lld alone exceeds the seed slot limit", (long long)b0);
                return SO_ELIMIT;
            }
            sub = std::max<i64>(1, nq / 2);
            continue;
        }
        const uint32_t nslots = (uint32_t)tot;
        if (nslots == 0) {
            b0 = b1;
            continue;
        }
        BlockGeom g;
        g.nq = nq;
        g.qb0 = (int)b0;
        g.qst_bits = bits_for(maxql);
        g.diag_bias = (int)ix.max_tlen + 1;
        g.diag_bits = bits_for((uint64_t)maxql + ix.max_tlen + 2);
        g.hd_bits = bits_for((uint64_t)M + 1);
        const int q_bits = bits_for((uint64_t)nq);
        if (g.qst_bits + g.diag_bits + g.hd_bits + q_bits > 63) {
            if (nq == 1) {
                set_error("sequence too long for the seed key layout");
                return SO_ELIMIT;
            }
            sub = std::max<i64>(1, nq / 2);
            continue;
        }
        g.L = ix.n_seeds - 1;
        g.nc = P.nc;
        g.thr_mul = ix.threshold;
        g.mink = P.mink;
        g.c0 = (int)ix.c0;
        if ((rc = scratch[SC_SLOTOFF].reserve(((size_t)nq + 1) * 4)) != SO_OK) return rc;
        if ((rc = scratch[SC_ST].reserve((size_t)nslots * 4)) != SO_OK) return rc;
        if ((rc = scratch[SC_CNT].reserve(((size_t)nslots + 1) * 4)) != SO_OK) return rc;
        if ((rc = scratch[SC_OUT].reserve(((size_t)nslots + 1) * 8)) != SO_OK) return rc;
        if ((rc = scratch[SC_MISC].reserve(((size_t)nq + 2) * 4 + 256)) != SO_OK) return rc;
        uint32_t *d_slot_off = (uint32_t *)scratch[SC_SLOTOFF].p;
        uint32_t *d_st = (uint32_t *)scratch[SC_ST].p, *d_cnt = (uint32_t *)scratch[SC_CNT].p;
        uint64_t *d_out = (uint64_t *)scratch[SC_OUT].p;
        SO_CUDA(cudaMemcpyAsync(d_slot_off, slot_off.data(), ((size_t)nq + 1) * 4, cudaMemcpyHostToDevice, st));
        stats.h2d_bytes += ((i64)nq + 1) * 4;
        SO_CUDA(cudaEventRecord(ev[0], st));
        k_query_seeds<<<(nslots + 255) / 256, 256, 0, st>>>(c->d_qres, c->d_qoff, d_slot_off, g, ix.d_start, nslots, d_st,
                                                            d_cnt);
        k_filter<<<(nq * 32 + 255) / 256, 256, 0, st>>>(c->d_qoff, d_slot_off, c->d_perm, g, d_cnt);
        SO_CUDA(cudaMemsetAsync(d_cnt + nslots, 0, 4, st));
        size_t tmp = 0;
        cub::TransformInputIterator<uint64_t, Widen, const uint32_t *> cnt64(d_cnt, Widen());
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt64, d_out, (int)nslots + 1, st);
        if ((rc = scratch[SC_TMP].reserve(tmp)) != SO_OK) return rc;
        SO_CUDA(cub::DeviceScan::ExclusiveSum(scratch[SC_TMP].p, tmp, cnt64, d_out, (int)nslots + 1, st));
        stats.kernel_launches += 2;
        stats.lib_launches += 1;
        uint64_t H = 0;
        SO_CUDA(cudaMemcpyAsync(&H, d_out + nslots, 8, cudaMemcpyDeviceToHost, st));
        SO_CUDA(cudaStreamSynchronize(st));
        SO_CUDA(cudaGetLastError());
        if (H > kHitCap && nq > 1) {
            sub = std::max<i64>(1, nq / 2);
            continue;
        }
        if (H > 0x7fffff00ull) {  // (the library sorts below take 32-bit signed item counts)
            set_error("query %lld alone produces %llu seed hits; the limit is 2^31", (long long)b0, (unsigned long long)H);
            return SO_ELIMIT;
        }
        stats.seed_hits += (i64)H;
        if (H > 0) {
            if ((rc = scratch[SC_KA].reserve((size_t)H * 8)) != SO_OK) return rc;
            if ((rc = scratch[SC_KB].reserve((size_t)H * 8)) != SO_OK) return rc;
            // one pattern + one alphabet: the hit ordinal is not carried along (k_pair_select rebuilds the rank from the key)
            const bool keys_only = AS == 1 && !getenv("SO_FORCE_PAIRS");
            if (!keys_only) {
                if ((rc = scratch[SC_VA].reserve((size_t)H * 4)) != SO_OK) return rc;
                if ((rc = scratch[SC_VB].reserve((size_t)H * 4)) != SO_OK) return rc;
            }
            uint64_t *ka = (uint64_t *)scratch[SC_KA].p, *kb = (uint64_t *)scratch[SC_KB].p;
            uint32_t *va = keys_only ? nullptr : (uint32_t *)scratch[SC_VA].p;
            uint32_t *vb = keys_only ? nullptr : (uint32_t *)scratch[SC_VB].p;
            const int key_bits = g.qst_bits + g.diag_bits + g.hd_bits + q_bits;
            // candidate sort key: query | first-appearance rank (see k_pair_select)
            const int rank_bits = (AS == 1 && !getenv("SO_FORCE_PAIRS")) ? g.qst_bits + g.hd_bits : 32;
            cub::DoubleBuffer<uint64_t> dk(ka, kb);
            cub::DoubleBuffer<uint32_t> dv(va, vb);
            unsigned long long *d_counter = (unsigned long long *)(scratch[SC_MISC].p);
            SO_CUDA(cudaMemsetAsync(d_counter, 0, 128, st));
            // ---- grouping: device-wide radix sort (the cell partition lives in block_candidates_fast)
            {
                const int ewarps = 148 * 64;
                k_expand<<<ewarps * 32 / 256, 256, 0, st>>>(d_slot_off, g, nslots, c->d_qoff, d_st, d_cnt, d_out, ix.d_hdsst, ka,
                                                            va);
                SO_CUDA(cudaEventRecord(ev[1], st));
                // dropped hits carry ~0 and must sort last: include one extra bit above the fields
                const int end_bit = std::min(64, key_bits + 1);
                tmp = 0;
                if (keys_only) {
                    cub::DeviceRadixSort::SortKeys(nullptr, tmp, dk, (int)H, g.qst_bits, end_bit, st);
                    if ((rc = scratch[SC_TMP].reserve(tmp)) != SO_OK) return rc;
                    SO_CUDA(cub::DeviceRadixSort::SortKeys(scratch[SC_TMP].p, tmp, dk, (int)H, g.qst_bits, end_bit, st));
                } else {
                    cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)H, 0, end_bit, st);
                    if ((rc = scratch[SC_TMP].reserve(tmp)) != SO_OK) return rc;
                    SO_CUDA(cub::DeviceRadixSort::SortPairs(scratch[SC_TMP].p, tmp, dk, dv, (int)H, 0, end_bit, st));
                }
                stats.kernel_launches += 1;
                stats.lib_launches += 1;
            }
            SO_CUDA(cudaEventRecord(ev[2], st));
            // candidates: at most one per (query, target) pair
            const uint64_t ccap = std::min<uint64_t>(H, (uint64_t)nq * (uint64_t)(M + 1));
            if ((rc = scratch[SC_CKA].reserve((size_t)ccap * 8)) != SO_OK) return rc;
            if ((rc = scratch[SC_CKB].reserve((size_t)ccap * 8)) != SO_OK) return rc;
            if ((rc = scratch[SC_CVA].reserve((size_t)ccap * 8)) != SO_OK) return rc;
            if ((rc = scratch[SC_CVB].reserve((size_t)ccap * 8)) != SO_OK) return rc;
            uint64_t *cka = (uint64_t *)scratch[SC_CKA].p, *ckb = (uint64_t *)scratch[SC_CKB].p;
            uint64_t *cva = (uint64_t *)scratch[SC_CVA].p, *cvb = (uint64_t *)scratch[SC_CVB].p;
            uint32_t *d_bounds = (uint32_t *)(scratch[SC_MISC].p + 128);
            // diagonal groups: head flags -> exclusive scan -> compact head positions
            const int grp_shift = g.qst_bits;
            if ((rc = scratch[SC_GIDX].reserve(((size_t)H + 1) * 4)) != SO_OK) return rc;
            uint32_t *d_gidx = (uint32_t *)scratch[SC_GIDX].p;
            {
                HeadFlag hf{dk.Current(), grp_shift};
                cub::CountingInputIterator<uint32_t> cnt(0);
                cub::TransformInputIterator<uint32_t, HeadFlag, cub::CountingInputIterator<uint32_t>> flags(cnt, hf);
                tmp = 0;
                cub::DeviceScan::ExclusiveSum(nullptr, tmp, flags, d_gidx, (int)H, st);
                if ((rc = scratch[SC_TMP].reserve(tmp)) != SO_OK) return rc;
                SO_CUDA(cub::DeviceScan::ExclusiveSum(scratch[SC_TMP].p, tmp, flags, d_gidx, (int)H, st));
            }
            // number of groups = gidx[H-1] + flag(H-1); read it back to size the group arrays
            uint32_t last_idx = 0;
            uint64_t last_keys[2] = {0, 0};
            SO_CUDA(cudaMemcpyAsync(&last_idx, d_gidx + (H - 1), 4, cudaMemcpyDeviceToHost, st));
            SO_CUDA(cudaMemcpyAsync(last_keys, dk.Current() + (H >= 2 ? H - 2 : 0), H >= 2 ? 16 : 8, cudaMemcpyDeviceToHost, st));
            SO_CUDA(cudaStreamSynchronize(st));
            uint32_t G;
            {
                const uint64_t kl = H >= 2 ? last_keys[1] : last_keys[0];
                bool flag = kl != ~0ull && (H < 2 || (last_keys[0] >> grp_shift) != (kl >> grp_shift));
                G = last_idx + (flag ? 1u : 0u);
            }
            stats.groups += (i64)G;
            if ((rc = scratch[SC_GHEAD].reserve(((size_t)G + 1) * 4)) != SO_OK) return rc;
            if ((rc = scratch[SC_GSCORE].reserve(((size_t)G + 1) * 4)) != SO_OK) return rc;
            if ((rc = scratch[SC_GRANK].reserve(((size_t)G + 1) * 4)) != SO_OK) return rc;
            uint32_t *d_gheads = (uint32_t *)scratch[SC_GHEAD].p, *d_gscore = (uint32_t *)scratch[SC_GSCORE].p;
            uint32_t *d_grank = (uint32_t *)scratch[SC_GRANK].p;
            if (G > 0) {
                // packed fast variants need < 8192 steps per extension and 32-bit residue indices
                const bool fast_ungap = maxql < 8192 && ix.max_tlen < 8192 && c->q_off[(size_t)c->n_q] < 0xfff00000ull &&
                                        c->t_off[(size_t)c->n_t] < 0xfff00000ull && !getenv("SO_GENERIC_UNGAP");
                const uint64_t qa = c->q_off[(size_t)b0], Lq64 = c->q_off[(size_t)b1] - qa;
                const bool single_path = fast_ungap && c->d_tung && Lq64 + 2 * kUngPad < (1ull << 24) && !getenv("SO_NO_SINGLE");
                const uint32_t *d_vals = keys_only ? nullptr : dv.Current();
                const uint64_t *d_gkey = nullptr;
                if (single_path) {
                    // X-drop view of the sub-block's queries: (class << 3), forward + reversed
                    uint32_t qoff2[2], qtotal;
                    ung_layout((uint32_t)Lq64, qoff2, qtotal);
                    if ((rc = scratch[SC_QUNG].reserve(qtotal)) != SO_OK) return rc;
                    if ((rc = scratch[SC_DESC].reserve(((size_t)G + 1) * 8)) != SO_OK) return rc;
                    if ((rc = scratch[SC_GKEY].reserve(((size_t)G + 1) * 8)) != SO_OK) return rc;
                    uint8_t *d_qung = scratch[SC_QUNG].p;
                    uint2 *d_desc = (uint2 *)scratch[SC_DESC].p;
                    if ((rc = ung_build(st, stats, c->d_qcls + qa, c->d_qoff + b0, (uint32_t)nq, qa, (uint32_t)Lq64, d_qung, qoff2, qtotal,
                                        3)) != SO_OK)
                        return rc;
                    k_group_desc<<<(uint32_t)((H + 255) / 256), 256, 0, st>>>(dk.Current(), d_vals, (uint32_t)H, g, d_gidx, c->d_qoff,
                                                                             c->d_toff, qa, d_gheads, d_desc,
                                                                             (uint64_t *)scratch[SC_GKEY].p, d_grank,
                                                                             d_counter);
                    SO_CUDA(cudaEventRecord(ev[5], st));
                    const int refill = g_xdrop_refill;
                    auto kx = refill <= 8 ? k_xdrop<8, false> : refill <= 12 ? k_xdrop<12, false> : refill <= 16 ? k_xdrop<16, false>
                              : refill <= 20 ? k_xdrop<20, false> : k_xdrop<24, false>;
                    kx<<<148 * 2, 512, kUngTabBytes, st>>>(d_desc, G, nullptr, dk.Current(), d_vals, (uint32_t)H, d_gheads, g.qst_bits,
                                                          (const uint4 *)c->d_tung, c->tung_off[0], c->tung_off[1],
                                                          (uint32_t)c->t_off[(size_t)c->n_t], (const uint4 *)d_qung, qoff2[0],
                                                          qoff2[1], (uint32_t)Lq64, d_gscore, d_grank, d_counter, 1, nullptr, 0, nullptr, 0,
                                                          nullptr, nullptr, 0u);
                    d_gkey = (const uint64_t *)scratch[SC_GKEY].p;
                    stats.kernel_launches += 2;
                } else {
                    k_scatter_heads<<<(uint32_t)((H + 255) / 256), 256, 0, st>>>(dk.Current(), (uint32_t)H, grp_shift, d_gidx, d_gheads);
                    SO_CUDA(cudaEventRecord(ev[5], st));
                    stats.kernel_launches += 1;
                }
                if (!single_path) {
                    int per_sm = 4;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_group_ungap_generic, 256, 0);
                    const int ublocks = 148 * std::max(1, per_sm);
                    k_group_ungap_generic<<<ublocks, 256, 0, st>>>(dk.Current(), d_vals, (uint32_t)H, d_gheads, G, g, c->d_qcls,
                                                                   c->d_qoff, c->d_tcls, c->d_toff, d_gscore, d_grank, d_counter);
                    stats.kernel_launches += 1;
                }
                SO_CUDA(cudaEventRecord(ev[6], st));
                if (!d_gkey) {
                    if ((rc = scratch[SC_GKEY].reserve(((size_t)G + 1) * 8)) != SO_OK) return rc;
                    k_gather_keys<<<(G + 255) / 256, 256, 0, st>>>(dk.Current(), d_gheads, G, (uint64_t *)scratch[SC_GKEY].p);
                    d_gkey = (const uint64_t *)scratch[SC_GKEY].p;
                    stats.kernel_launches += 1;
                }
                // ordered compaction of the passing groups (library), then the pair fold over that list
                if ((rc = scratch[SC_PLIST].reserve(((size_t)G + 2) * 4)) != SO_OK) return rc;
                uint32_t *d_plist = (uint32_t *)scratch[SC_PLIST].p;
                uint32_t *d_pcount = (uint32_t *)(d_counter + 5);
                {
                    cub::CountingInputIterator<uint32_t> gi0(0);
                    PassPred pred{d_gscore};
                    tmp = 0;
                    cub::DeviceSelect::If(nullptr, tmp, gi0, d_plist, d_pcount, (int)G, pred, st);
                    if ((rc = scratch[SC_TMP].reserve(tmp)) != SO_OK) return rc;
                    SO_CUDA(cub::DeviceSelect::If(scratch[SC_TMP].p, tmp, gi0, d_plist, d_pcount, (int)G, pred, st));
                }
                k_pair_select<<<148 * 8, 256, 0, st>>>(d_plist, d_pcount, d_gkey, g, d_gscore, keys_only ? nullptr : d_grank,
                                                       rank_bits, cka, cva, d_counter);
                stats.kernel_launches += 1;
                stats.lib_launches += 1;
            }
            SO_CUDA(cudaEventRecord(ev[3], st));
            unsigned long long h_counter[4] = {0, 0, 0, 0};
            SO_CUDA(cudaMemcpyAsync(h_counter, d_counter, 32, cudaMemcpyDeviceToHost, st));
            SO_CUDA(cudaStreamSynchronize(st));
            const unsigned long long ncand = h_counter[0];
            stats.ungap_steps += (i64)h_counter[1];
            stats.multi_groups += (i64)h_counter[3];
            SO_CUDA(cudaGetLastError());
            stats.kernel_launches += 1;
            stats.lib_launches += 1;
            std::vector<uint32_t> bounds((size_t)nq + 1, 0);
            const size_t base_c = out.n;
            if (ncand > 0) {
                cub::DoubleBuffer<uint64_t> ck(cka, ckb), cv(cva, cvb);
                tmp = 0;
                cub::DeviceRadixSort::SortPairs(nullptr, tmp, ck, cv, (int)ncand, 0, rank_bits + q_bits, st);
                if ((rc = scratch[SC_TMP].reserve(tmp)) != SO_OK) return rc;
                SO_CUDA(cub::DeviceRadixSort::SortPairs(scratch[SC_TMP].p, tmp, ck, cv, (int)ncand, 0, rank_bits + q_bits, st));
                k_query_bounds<<<(nq + 1 + 127) / 128, 128, 0, st>>>(ck.Current(), (uint32_t)ncand, nq, rank_bits, d_bounds);
                if (bs) {
                    // H3 runs on the device: the candidates join the block's per-query lists (select.cu)
                    if ((rc = bs->append(cv.Current(), d_bounds, nq, bs_q0 + (int)(b0 - q_begin), st)) != SO_OK) return rc;
                    SO_CUDA(cudaEventRecord(ev[4], st));
                    SO_CUDA(cudaEventSynchronize(ev[4]));
                    out.n = base_c + (size_t)ncand;  // count only
                    stats.kernel_launches += 2;
                    stats.lib_launches += 1;
                } else {
                    SO_CUDA(cudaEventRecord(ev[4], st));
                    Timer td;
                    if ((rc = out.reserve(base_c + (size_t)ncand)) != SO_OK) return rc;
                    SO_CUDA(cudaMemcpyAsync(out.vals + base_c, cv.Current(), (size_t)ncand * 8, cudaMemcpyDeviceToHost, st));
                    SO_CUDA(cudaMemcpyAsync(bounds.data(), d_bounds, ((size_t)nq + 1) * 4, cudaMemcpyDeviceToHost, st));
                    SO_CUDA(cudaStreamSynchronize(st));
                    SO_CUDA(cudaGetLastError());
                    out.n = base_c + (size_t)ncand;
                    stats.kernel_launches += 1;
                    stats.lib_launches += 1;
                    stats.d2h_bytes += (i64)ncand * 8 + ((i64)nq + 1) * 4;
                    c->d2h_ms_lane[lane] += td.ms();
                }
                float ms = 0;
                cudaEventElapsedTime(&ms, ev[3], ev[4]);
                stats.ms_select += ms;
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[0], ev[1]);
            stats.ms_seed += ms;
            cudaEventElapsedTime(&ms, ev[1], ev[2]);
            stats.ms_sort += ms;
            cudaEventElapsedTime(&ms, ev[2], ev[3]);
            stats.ms_ungap += ms;
            if (G > 0) {
                cudaEventElapsedTime(&ms, ev[5], ev[6]);
                stats.ms_ungap_kernel += ms;
            }
            for (int k = 0; k < nq; k++)
                out.offsets[(size_t)(b0 - q_begin + k + 1)] = (uint64_t)base_c + bounds[(size_t)k + 1];
        }
        // offsets of queries without hits in this sub-block
        for (int k = 0; k < nq; k++) {
            size_t idx = (size_t)(b0 - q_begin + k + 1);
            if (out.offsets[idx] < out.offsets[idx - 1]) out.offsets[idx] = out.offsets[idx - 1];
        }
        // adapt the sub-block size to the observed hit density
        if (H > 0) {
            double per_q = (double)H / (double)nq;
            i64 want = (i64)((double)kHitCap * 0.5 / std::max(per_q, 1.0));
            sub = std::min<i64>(std::max<i64>(want, 16), 2047);
            if (c->sub_block > 0) sub = c->sub_block;
        }
        b0 = b1;
    }
    stats.candidates += (i64)out.n;
    return SO_OK;
}

// ---------------------------------------------------------------------------------------------
// Sync-free candidate production of one query block (the BASELINE configurations: one pattern, one alphabet,
// sequences below 8192 residues).  Per (sub-block, chunk): k_query_seeds, k_filter, k_cell_pass<false> (cell counts
// + in-CTA scan), k_unit_scan, k_cell_pass<true>, k_cell_small / k_cell_warp / k_cell_block (sorted cell keys +
// X-drop descriptors), k_xdrop<FAST> (passing groups leave a record in their query's region), k_cand_sort (pair fold, reference order,
// appended to the block's per-query lists).  Every launch is sized on the host from
// upper bounds (a query keeps at most threshold * len + max_bucket seed hits, fsearch.py:2667-2677), every
// data-dependent size stays in device memory, so the host never waits inside a block.
// ---------------------------------------------------------------------------------------------
struct FastPlan {
    int nsplit;
    uint32_t NB, NBh;
    BlockGeom g;
};

static bool fast_plan(const so_ctx *c, const ChunkIndex &ix, uint32_t maxql, int nq, i64 b0, FastPlan &pl) {
    const Params &P = c->P;
    const i64 M = ix.c1 - ix.c0;
    BlockGeom &g = pl.g;
    g.nq = nq;
    g.qb0 = (int)b0;
    g.qst_bits = bits_for(maxql);
    g.diag_bias = (int)ix.max_tlen + 1;
    g.diag_bits = bits_for((uint64_t)maxql + ix.max_tlen + 2);
    g.hd_bits = bits_for((uint64_t)M + 1);
    g.L = ix.n_seeds ? ix.n_seeds - 1 : 0;
    g.nc = P.nc;
    g.thr_mul = ix.threshold;
    g.mink = P.mink;
    g.c0 = (int)ix.c0;
    pl.NB = (uint32_t)(M + 2);
    pl.nsplit = g_cell_split ? (int)g_cell_split : (int)(((size_t)pl.NB * 4 + 220 * 1024 - 1) / (220 * 1024));
    pl.NBh = (pl.NB + (uint32_t)pl.nsplit - 1) / (uint32_t)pl.nsplit;
    if (pl.nsplit > 8 || (size_t)pl.NBh * 4 > 220 * 1024) return false;
    if (g.qst_bits + g.diag_bits > 31) return false;                           // cell-local key + head bit
    if (g.qst_bits + g.hd_bits + 20 + g.diag_bits > 64) return false;          // candidate word / passing-group record
    if (g.hd_bits > 16 || nq >= 65536) return false;                           // (query << 16 | target + 1), 16-bit query in the record buffers
    if ((uint64_t)nq * pl.NB >= 0x7fffff00ull || (uint64_t)nq * (uint64_t)pl.nsplit > 16384) return false;
    if (maxql >= 8192 || ix.max_tlen >= 8192) return false;                    // packed X-drop state
    return true;
}

int block_candidates_fast(so_ctx *c, i64 b0, i64 b1, int lane, BlockStore &bs, bool &eligible, int only_chunk) {
    eligible = false;
    const Params &P = c->P;
    const int AS = (int)(P.alphabets.size() * P.patterns.size());
    if (AS != 1 || getenv("SO_FORCE_PAIRS") || getenv("SO_CUB_SORT") || getenv("SO_GENERIC_UNGAP") || getenv("SO_NO_SINGLE") ||
        getenv("SO_NO_FAST"))
        return SO_OK;
    if (!c->d_tung || c->q_off[(size_t)c->n_q] >= 0xfff00000ull || c->t_off[(size_t)c->n_t] >= 0xfff00000ull) return SO_OK;
    so::DBuf<uint8_t> *scratch = c->lane_scratch(lane);
    so_stats &stats = c->stats_lane[lane];
    cudaStream_t st = c->lane_stream(lane);
    const size_t nch = c->chunks.size();
    // ---- sub-blocks: bounded seed hits (all chunks), query view below 2^24 bytes
    uint64_t thr_max = 0, maxb = 0;
    for (const auto &ix : c->chunks) {
        thr_max = std::max<uint64_t>(thr_max, (uint64_t)std::max<i64>(ix.threshold, 0));
        maxb = std::max<uint64_t>(maxb, ix.max_bucket);
    }
    struct Sub {
        i64 s0, s1;
        uint32_t maxql;
        uint64_t hit_cap;
    };
    std::vector<Sub> subs;
    i64 sub_max = c->sub_block > 0 ? c->sub_block : 2047;
    if (const char *e = getenv("SO_SUB_BLOCK0")) sub_max = std::max<i64>(1, std::min<i64>(2047, atoll(e)));  // test hook
    {
        i64 s0 = b0;
        uint64_t hits = 0, res = 0;
        uint32_t mq = 1;
        for (i64 q = b0; q <= b1; q++) {
            uint64_t L = 0, ub = 0;
            if (q < b1) {
                L = c->q_off[(size_t)q + 1] - c->q_off[(size_t)q];
                ub = L >= (uint64_t)P.mink ? thr_max * L + maxb : 0;
                if (ub > kHitCap) return SO_OK;  // a single query may exceed the hit buffers: general path
            }
            const bool full = q == b1 || hits + ub > kHitCap || res + L + 2 * kUngPad + 64 >= (1ull << 24) ||
                              q - s0 >= sub_max;
            if (full && q > s0) {
                subs.push_back(Sub{s0, q, mq, hits});
                s0 = q, hits = 0, res = 0, mq = 1;
            }
            if (q < b1) {
                hits += ub, res += L;
                if (L >= (uint64_t)P.mink) mq = std::max<uint32_t>(mq, (uint32_t)L);
            }
        }
    }
    // ---- every (sub-block, chunk) must fit the packed layouts, else nothing is launched
    std::vector<FastPlan> plans(subs.size() * nch);
    for (size_t si = 0; si < subs.size(); si++)
        for (size_t ch = 0; ch < nch; ch++)
            if (!fast_plan(c, c->chunks[ch], subs[si].maxql, (int)(subs[si].s1 - subs[si].s0), subs[si].s0, plans[si * nch + ch]))
                return SO_OK;
    eligible = true;
    int rc;
    // ---- control block: [0..7] u64 counters (0 hits of the launch, 1 X-drop steps, 2 group pool, 3 chained groups,
    //      4 groups, 5 candidates, 6 seed hits), [8..9] transient u32 (list lengths, work counters), [10] flags
    if ((rc = scratch[SC_CTL].reserve(256)) != SO_OK) return rc;
    unsigned long long *d_ctl = (unsigned long long *)scratch[SC_CTL].p;
    uint32_t *d_lcount = (uint32_t *)(d_ctl + 8);  // [0] warp list, [1] block list, [2] scatter-pass work counter, [3] cand-sort work counter
    uint32_t *d_flags = (uint32_t *)(d_ctl + 10);  // [0] redo through the general path, [1] hard error
    SO_CUDA(cudaMemsetAsync(d_ctl, 0, 128, st));
    std::vector<cudaEvent_t> &evp = c->ev_pool[lane];
    size_t ev_used = 0;
    auto stamp = [&]() {
        if (ev_used >= 640) return;
        if (ev_used >= evp.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            evp.push_back(e);
        }
        cudaEventRecord(evp[ev_used++], st);
    };
    const int refill = g_xdrop_refill;
    auto kx = refill <= 8 ? k_xdrop<8, true> : refill <= 12 ? k_xdrop<12, true> : refill <= 16 ? k_xdrop<16, true>
              : refill <= 20 ? k_xdrop<20, true> : k_xdrop<24, true>;
    const int xw = g_xdrop_warps, xqs = g_xdrop_qs;
    if (xw == 32 || xqs != 0)  // layout variants (refill 24 only)
        kx = xw == 32 ? (xqs == 16 ? k_xdrop<24, true, 32, 16> : xqs == 4 ? k_xdrop<24, true, 32, 4> : k_xdrop<24, true, 32, 0>)
                      : (xqs == 16 ? k_xdrop<24, true, 16, 16> : k_xdrop<24, true, 16, 4>);
    if (const char *r2 = getenv("SO_XDROP_REFILL2")) {  // tuning hook: refill threshold of the default layout
        const int r = atoi(r2);
        if (xw == 32 && xqs == 16)
            kx = r <= 16 ? k_xdrop<16, true, 32, 16> : r <= 20 ? k_xdrop<20, true, 32, 16> : r <= 24 ? k_xdrop<24, true, 32, 16>
                 : r <= 28 ? k_xdrop<28, true, 32, 16> : k_xdrop<32, true, 32, 16>;
    }
    for (size_t si = 0; si < subs.size(); si++) {
        const Sub &sb = subs[si];
        const int nq = (int)(sb.s1 - sb.s0);
        std::vector<uint32_t> slot_off((size_t)nq + 1, 0);
        uint64_t tot = 0;
        for (int k = 0; k < nq; k++) {
            const uint64_t L = c->q_off[(size_t)(sb.s0 + k + 1)] - c->q_off[(size_t)(sb.s0 + k)];
            tot += L >= (uint64_t)P.mink ? L : 0;
            slot_off[(size_t)k + 1] = (uint32_t)tot;
        }
        const uint32_t nslots = (uint32_t)tot;
        if (nslots == 0) continue;
        const uint64_t Hcap = std::max<uint64_t>(sb.hit_cap, 1);
        const uint64_t qa = c->q_off[(size_t)sb.s0], Lq64 = c->q_off[(size_t)sb.s1] - qa;
        uint32_t qoff2[2], qtotal;
        ung_layout((uint32_t)Lq64, qoff2, qtotal);
        if ((rc = scratch[SC_SLOTOFF].reserve(((size_t)nq + 1) * 4)) != SO_OK) return rc;
        if ((rc = scratch[SC_ST].reserve((size_t)nslots * 4)) != SO_OK) return rc;
        if ((rc = scratch[SC_CNT].reserve(((size_t)nslots + 1) * 4)) != SO_OK) return rc;
        if ((rc = scratch[SC_QUNG].reserve(qtotal)) != SO_OK) return rc;
        if ((rc = scratch[SC_SUB].reserve((size_t)Hcap * 4)) != SO_OK) return rc;
        if ((rc = scratch[SC_SSUB].reserve((size_t)Hcap * 4)) != SO_OK) return rc;
        if ((rc = scratch[SC_DESC].reserve((size_t)Hcap * 8)) != SO_OK) return rc;
        if ((rc = scratch[SC_GSCORE].reserve((size_t)Hcap * 4)) != SO_OK) return rc;
        if ((rc = scratch[SC_WLIST].reserve(((size_t)Hcap / kCellSmall + 16) * 4 * 2)) != SO_OK) return rc;
        uint32_t *d_slot_off = (uint32_t *)scratch[SC_SLOTOFF].p;
        uint32_t *d_st = (uint32_t *)scratch[SC_ST].p, *d_cnt = (uint32_t *)scratch[SC_CNT].p;
        uint8_t *d_qung = scratch[SC_QUNG].p;
        uint32_t *d_sub = (uint32_t *)scratch[SC_SUB].p, *d_ssub = (uint32_t *)scratch[SC_SSUB].p;
        uint2 *d_desc = (uint2 *)scratch[SC_DESC].p;
        uint32_t *d_cellid = (uint32_t *)scratch[SC_GSCORE].p;  // (query << 16 | target + 1) of every group head
        uint32_t *d_wlist = (uint32_t *)scratch[SC_WLIST].p, *d_blist = d_wlist + ((size_t)Hcap / kCellSmall + 16);
        SO_CUDA(cudaMemcpyAsync(d_slot_off, slot_off.data(), ((size_t)nq + 1) * 4, cudaMemcpyHostToDevice, st));
        stats.h2d_bytes += ((i64)nq + 1) * 4;
        // X-drop view of the sub-block's queries: (class << 3), forward + reversed (shared by all chunks)
        if ((rc = ung_build(st, stats, c->d_qcls + qa, c->d_qoff + sb.s0, (uint32_t)nq, qa, (uint32_t)Lq64, d_qung, qoff2, qtotal,
                            3)) != SO_OK)
            return rc;
        const uint8_t *d_qview = d_qung;
        uint32_t qstride = 0;
        if (xqs) {
            qstride = (qtotal + 64u + 15u) & ~15u;
            if ((rc = scratch[SC_QSHIFT].reserve((size_t)qstride * (size_t)xqs)) != SO_OK) return rc;
            k_ung_shift<<<dim3((qstride + 255) / 256, (unsigned)xqs), 256, 0, st>>>(d_qung, qtotal, scratch[SC_QSHIFT].p, qstride,
                                                                                  xqs == 16 ? 1u : 4u, (uint8_t)(kUngStop << 3));
            stats.kernel_launches += 1;
            d_qview = scratch[SC_QSHIFT].p;
        }
        for (size_t ch = 0; ch < nch; ch++) {
            const ChunkIndex &ix = c->chunks[ch];
            if (ix.n_seeds == 0 || (only_chunk >= 0 && (size_t)only_chunk != ch)) continue;
            const FastPlan &pl = plans[si * nch + ch];
            const BlockGeom &g = pl.g;
            const i64 M = ix.c1 - ix.c0;
            const uint32_t NB = pl.NB, NBh = pl.NBh, nsplit = (uint32_t)pl.nsplit;
            const uint32_t ncells = (uint32_t)nq * NB, U = (uint32_t)nq * nsplit;
            const size_t ccap = 2 * ((size_t)M + 1);  // passing diagonal groups per query (overflow raises the redo flag)
            if ((rc = scratch[SC_CELLLOC].reserve(((size_t)ncells + 1) * 4)) != SO_OK) return rc;
            if ((rc = scratch[SC_UNIT].reserve(((size_t)U * 2 + 8) * 4)) != SO_OK) return rc;
            if ((rc = scratch[SC_CREG].reserve((size_t)nq * ccap * 8)) != SO_OK) return rc;
            if ((rc = scratch[SC_QCOUNT].reserve(((size_t)nq + 8) * 4)) != SO_OK) return rc;
            const int sgrid = std::min(nq, 148 * 2);
            if ((rc = scratch[SC_GBUF].reserve((size_t)sgrid * ccap * 8)) != SO_OK) return rc;
            uint32_t *d_cloc = (uint32_t *)scratch[SC_CELLLOC].p;
            uint32_t *d_utot = (uint32_t *)scratch[SC_UNIT].p, *d_ubase = d_utot + U + 4;
            uint64_t *d_creg = (uint64_t *)scratch[SC_CREG].p;
            uint32_t *d_qcount = (uint32_t *)scratch[SC_QCOUNT].p;  // [nq] candidates per query, [nq] count-pass work counter
            SO_CUDA(cudaMemsetAsync(d_qcount, 0, ((size_t)nq + 8) * 4, st));
            stamp();
            k_query_seeds<<<(nslots + 255) / 256, 256, 0, st>>>(c->d_qres, c->d_qoff, d_slot_off, g, ix.d_start, nslots, d_st, d_cnt);
            k_filter<<<(nq * 32 + 255) / 256, 256, 0, st>>>(c->d_qoff, d_slot_off, c->d_perm, g, d_cnt);
            stamp();
            const int pblocks = std::min<int>(nq * (int)nsplit, 148 * (int)nsplit);
            k_cell_pass<false><<<pblocks, 1024, NBh * 4, st>>>(d_slot_off, g, c->d_qoff, d_st, d_cnt, ix.d_hdsst, NB, nsplit, d_cloc,
                                                             d_utot, nullptr, nullptr, g_cell_max, d_flags, d_qcount + nq);
            k_unit_scan<<<1, 1024, 0, st>>>(d_utot, U, d_ubase, d_ctl, d_flags);
            k_cell_pass<true><<<pblocks, 1024, NBh * 4, st>>>(d_slot_off, g, c->d_qoff, d_st, d_cnt, ix.d_hdsst, NB, nsplit, d_cloc,
                                                            nullptr, d_ubase, d_sub, g_cell_max, d_flags, d_lcount + 2);
            if (g_cell_span) {
                k_cell_span<<<dim3((NB + 255) / 256, (unsigned)nq), 256, 0, st>>>(d_cloc, d_ubase, NB, NBh, nsplit, g, c->d_qoff, c->d_toff, qa,
                                                                                 d_sub, d_ssub, d_desc, d_cellid, d_blist, d_lcount);
            } else {
                k_cell_small<<<dim3((NB + 255) / 256, (unsigned)nq), 256, 0, st>>>(d_cloc, d_ubase, ncells, NB, NBh, nsplit, g, c->d_qoff,
                                                                                  c->d_toff, qa, d_sub, d_ssub, d_desc, d_cellid, d_wlist,
                                                                                  d_blist, d_lcount, d_flags);
                k_cell_warp<<<148 * 4, 256, 0, st>>>(d_cloc, d_ubase, d_wlist, d_lcount, NB, NBh, nsplit, g, c->d_qoff, c->d_toff, qa, d_sub,
                                                    d_ssub, d_desc, d_cellid, d_flags);
            }
            k_cell_block<<<148 * 2, 512, sizeof(CellBlockSmem), st>>>(d_cloc, d_ubase, d_blist, d_lcount, NB, NBh, nsplit, g, c->d_qoff,
                                                                     c->d_toff, qa, d_sub, d_ssub, d_desc, d_cellid, d_flags);
            stamp();
            kx<<<xw == 32 ? 148 : 148 * g_xdrop_ctas, xw * 32, kUngTabBytes + (xw / 16) * kXdBufBytes, st>>>(
                d_desc, 0u, d_ssub, nullptr, nullptr, 0u, nullptr, g.qst_bits, (const uint4 *)c->d_tung, c->tung_off[0], c->tung_off[1],
                (uint32_t)c->t_off[(size_t)c->n_t], (const uint4 *)d_qview, qoff2[0], qoff2[1], (uint32_t)Lq64, nullptr, nullptr, d_ctl,
                1, d_cellid, g.diag_bits, d_creg, ccap, d_qcount, d_flags, qstride / 16);
            stamp();
            k_cand_sort<<<sgrid, kCandThreads, sizeof(CandSmem), st>>>(d_creg, ccap, d_qcount, nq, g, c->d_qoff, (uint64_t *)scratch[SC_GBUF].p,
                                                                      bs.vals.p, bs.capq, bs.count.p, (int)(sb.s0 - b0), d_lcount + 3,
                                                                      d_flags, d_ctl);
            stamp();
            SO_CUDA(cudaGetLastError());
            stats.kernel_launches += 10;
        }
    }
    c->ev_used[lane] = ev_used;
    return SO_OK;
}

// the block's control block (flags, counters) goes to the host behind the block's kernels, BEFORE the lane waits: on a
// shared stream a copy enqueued after the wait would queue behind the next block of another lane
int enqueue_fast_ctl(so_ctx *c, int lane) {
    SO_CUDA(cudaMemcpyAsync(c->h_ctl[lane], c->lane_scratch(lane)[SC_CTL].p, sizeof c->h_ctl[lane], cudaMemcpyDeviceToHost,
                            c->lane_stream(lane)));
    return SO_OK;
}

// after the lane's wait: flags, counters and stage times of block_candidates_fast
int finish_fast_block(so_ctx *c, int lane, bool &redo) {
    so_stats &stats = c->stats_lane[lane];
    redo = false;
    const unsigned long long *h = c->h_ctl[lane];
    const uint32_t f0 = (uint32_t)h[10], f1 = (uint32_t)(h[10] >> 32);
    if (f1) {
        set_error("candidate production failed on the device (flag %u)", f1);
        return SO_ELIMIT;
    }
    const std::vector<cudaEvent_t> &evp = c->ev_pool[lane];
    const size_t n = c->ev_used[lane] / 5;
    for (size_t k = 0; k < n; k++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, evp[k * 5], evp[k * 5 + 1]);
        stats.ms_seed += ms;
        cudaEventElapsedTime(&ms, evp[k * 5 + 1], evp[k * 5 + 2]);
        stats.ms_sort += ms;
        cudaEventElapsedTime(&ms, evp[k * 5 + 2], evp[k * 5 + 3]);
        stats.ms_ungap += ms, stats.ms_ungap_kernel += ms;
        cudaEventElapsedTime(&ms, evp[k * 5 + 3], evp[k * 5 + 4]);
        stats.ms_select += ms;
    }
    c->ev_used[lane] = 0;
    if (f0) {
        redo = true;  // a (query, target) cell above the shared-memory sort: the general path redoes the block
        return SO_OK;
    }
    stats.ungap_steps += (i64)h[1], stats.multi_groups += (i64)h[3], stats.groups += (i64)h[4];
    stats.candidates += (i64)h[5], stats.seed_hits += (i64)h[6];
    stats.d2h_bytes += (i64)sizeof c->h_ctl[lane];
    return SO_OK;
}

}  // namespace so

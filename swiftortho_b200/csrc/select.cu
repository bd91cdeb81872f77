// H3 on the device: per-query candidate merge, the reference's quicksort by -score and the [:vmax] cut
// (lib/fsearch.py:3039-3059 with qsort / quicksort / partition / insort, lib/fsearch.py:260-327).
//
// The reference sorts the chunk-major concatenation of a query's candidates with its own UNSTABLE but
// deterministic quicksort and aligns the first vmax of them, so the order among equal scores is part of the
// result.  The sort is emulated exactly, in parallel:
//   * pivot index l + int(0.3745401188473625 * gap) (Rand.init_genrand(42) runs on every call, so random() is a
//     constant), l + 3 for a range of 7, stable insertion sort (= any stable sort) below 7;
//   * the Hoare partition (fsearch.py:281-297) is data parallel: with a_1 < a_2 < ... the positions in (l, r]
//     whose key is >= pivot (where the i scan stops) and b_1 > b_2 > ... the positions in [l, r] whose key is
//     <= pivot (where the j scan stops), both taken on the array BEFORE the partition, the sequential loop
//     swaps exactly the pairs (a_k, b_k), k = 1..K, K = #{k : a_k <= b_k} (the condition is monotone in k), and
//     ends with j = max(b_{K+1}, a_K') where K' = K if a_K < b_K else K - 1 (checked against the sequential loop
//     on 2*10^5 random ranges with heavy ties, tests/test_oracle_golden.py::test_parallel_hoare_formula);
//   * only ranges that intersect [0, need) are refined (the others cannot influence that prefix), like the
//     host's quicksort_pruned.
// One CTA per query: ranges above kSmall (512) elements are partitioned by the whole CTA in global memory (stopper
// lists by ballot + prefix over the warps), smaller ranges are handed to single warps that finish the whole
// sub-tree in shared memory.  Only the selected <= vmax candidates per query go back to the host.
#include "context.h"

namespace so {

enum { kSelThreads = 256, kSelWarps = 8, kSmall = 512, kBigStack = 64, kSmallList = 256, kWarpStack = 48 };
constexpr double kPivotFracD = 0.3745401188473625;  // random() after init_genrand(42)

struct SelSmem {
    uint64_t data[kSelWarps][kSmall];
    uint16_t la[kSelWarps][kSmall];
    uint16_t lb[kSelWarps][kSmall];
    int2 wstack[kSelWarps][kWarpStack];
    int2 bstack[kBigStack];
    int2 slist[kSmallList];
    int wc[2][kSelWarps][2];
    int nb, ns, snext, j, q;
};

__device__ __forceinline__ uint32_t sel_key(uint64_t e) { return (uint32_t)(e >> 32); }

// Hoare partition of x[l..r] (r - l + 1 > kSmall) by the whole CTA; returns the pivot's final index
__device__ int sel_partition_cta(uint64_t *__restrict__ x, int l, int r, uint32_t *__restrict__ LA, uint32_t *__restrict__ LB,
                                 SelSmem &S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int gap = r - l + 1;
    if (tid == 0) {
        const int m = gap == 7 ? l + 3 : l + (int)(kPivotFracD * (double)gap);
        const uint64_t t = x[l];
        x[l] = x[m];
        x[m] = t;
    }
    __syncthreads();
    const uint32_t pv = sel_key(x[l]);
    constexpr int U = 4;
    int nA = 0, nB = 0, buf = 0;
    for (int base = l; base <= r; base += kSelThreads * U) {
        const int wb = base + warp * 32 * U;
        uint32_t k[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int p = wb + u * 32 + lane;
            k[u] = p <= r ? sel_key(x[p]) : 0u;
        }
        unsigned ma[U], mb[U];
        int ca = 0, cb = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int p = wb + u * 32 + lane;
            const bool valid = p <= r;
            ma[u] = __ballot_sync(0xffffffffu, valid && p > l && k[u] >= pv);
            mb[u] = __ballot_sync(0xffffffffu, valid && k[u] <= pv);
            ca += __popc(ma[u]);
            cb += __popc(mb[u]);
        }
        if (lane == 0) S.wc[buf][warp][0] = ca, S.wc[buf][warp][1] = cb;
        __syncthreads();
        int oa = nA, ob = nB;
#pragma unroll
        for (int w = 0; w < kSelWarps; w++) {
            const int a = S.wc[buf][w][0], b = S.wc[buf][w][1];
            if (w < warp) oa += a, ob += b;
            nA += a, nB += b;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t p = (uint32_t)(wb + u * 32 + lane);
            if ((ma[u] >> lane) & 1u) LA[oa + __popc(ma[u] & lt)] = p;
            if ((mb[u] >> lane) & 1u) LB[ob + __popc(mb[u] & lt)] = p;
            oa += __popc(ma[u]);
            ob += __popc(mb[u]);
        }
        buf ^= 1;
    }
    __syncthreads();
    const int np = min(nA, nB);
    int K = 0;
    for (int k0 = 0; k0 < np; k0 += kSelThreads) {
        const int k = k0 + tid;
        bool ok = false;
        if (k < np) {
            const uint32_t a = LA[k], b = LB[nB - 1 - k];
            ok = a <= b;
            if (a < b) {
                const uint64_t t = x[a];
                x[a] = x[b];
                x[b] = t;
            }
        }
        const int c = __syncthreads_count(ok);
        K += c;
        if (c < kSelThreads) break;
    }
    if (tid == 0) {
        int j;
        if (K == 0)
            j = (int)LB[nB - 1];
        else {
            const int bn = (int)LB[nB - 1 - K];  // b_{K+1}: exists, position l is the lowest j-stopper and never pairs
            const int Kp = LA[K - 1] < LB[nB - K] ? K : K - 1;
            const int aK = Kp >= 1 ? (int)LA[Kp - 1] : -1;
            j = max(bn, aK);
        }
        const uint64_t t = x[l];
        x[l] = x[j];
        x[j] = t;
        S.j = j;
    }
    __syncthreads();
    return S.j;
}

// the whole (pruned) sub-tree of x[l..r], r - l + 1 <= kSmall, by one warp in shared memory
__device__ void sel_small_warp(uint64_t *__restrict__ x, int l, int r, int need, SelSmem &S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    uint64_t *d = S.data[warp];
    uint16_t *la = S.la[warp], *lb = S.lb[warp];
    int2 *st = S.wstack[warp];
    const int gap0 = r - l + 1;
    const int needl = need - l;
    for (int i = lane; i < gap0; i += 32) d[i] = x[l + i];
    __syncwarp();
    int sp = 0;
    if (lane == 0) st[0] = make_int2(0, gap0 - 1);
    sp = 1;
    __syncwarp();
    while (sp > 0) {
        const int2 rg = st[sp - 1];
        sp--;
        __syncwarp();
        const int a = rg.x, b = rg.y;
        if (b <= a || a >= needl) continue;
        const int g = b - a + 1;
        if (g < 7) {  // insort: stable
            uint64_t v = 0;
            int rank = 0;
            if (lane < g) {
                v = d[a + lane];
                const uint32_t kv = sel_key(v);
                for (int j = 0; j < g; j++) {
                    const uint32_t kj = sel_key(d[a + j]);
                    rank += (kj < kv || (kj == kv && j < lane)) ? 1 : 0;
                }
            }
            __syncwarp();
            if (lane < g) d[a + rank] = v;
            __syncwarp();
            continue;
        }
        if (lane == 0) {
            const int m = g == 7 ? a + 3 : a + (int)(kPivotFracD * (double)g);
            const uint64_t t = d[a];
            d[a] = d[m];
            d[m] = t;
        }
        __syncwarp();
        const uint32_t pv = sel_key(d[a]);
        int nA = 0, nB = 0;
        for (int t0 = a; t0 <= b; t0 += 32) {
            const int p = t0 + lane;
            const bool valid = p <= b;
            const uint32_t k = valid ? sel_key(d[p]) : 0u;
            const unsigned ma = __ballot_sync(0xffffffffu, valid && p > a && k >= pv);
            const unsigned mb = __ballot_sync(0xffffffffu, valid && k <= pv);
            if ((ma >> lane) & 1u) la[nA + __popc(ma & lt)] = (uint16_t)p;
            if ((mb >> lane) & 1u) lb[nB + __popc(mb & lt)] = (uint16_t)p;
            nA += __popc(ma);
            nB += __popc(mb);
        }
        __syncwarp();
        const int np = min(nA, nB);
        int K = 0;
        for (int k0 = 0; k0 < np; k0 += 32) {
            const int k = k0 + lane;
            bool ok = false;
            if (k < np) {
                const int pa = la[k], pb = lb[nB - 1 - k];
                ok = pa <= pb;
                if (pa < pb) {
                    const uint64_t t = d[pa];
                    d[pa] = d[pb];
                    d[pb] = t;
                }
            }
            const int c = __popc(__ballot_sync(0xffffffffu, ok));
            K += c;
            if (c < 32) break;
        }
        __syncwarp();
        int j;
        if (K == 0)
            j = lb[nB - 1];
        else {
            const int bn = lb[nB - 1 - K];
            const int Kp = la[K - 1] < lb[nB - K] ? K : K - 1;
            const int aK = Kp >= 1 ? (int)la[Kp - 1] : -1;
            j = max(bn, aK);
        }
        __syncwarp();
        // the larger side is pushed first, so the stack stays below log2(kSmall) + 2 entries
        const int2 L = make_int2(a, j - 1), R = make_int2(j + 1, b);
        const bool lbig = (L.y - L.x) >= (R.y - R.x);
        const int2 first = lbig ? L : R, second = lbig ? R : L;
        const bool p1 = first.y > first.x && first.x < needl && sp < kWarpStack;
        const bool p2 = second.y > second.x && second.x < needl && sp + (p1 ? 1 : 0) < kWarpStack;
        if (lane == 0) {
            const uint64_t t = d[a];
            d[a] = d[j];
            d[j] = t;
            if (p1) st[sp] = first;
            if (p2) st[sp + (p1 ? 1 : 0)] = second;
        }
        sp += (p1 ? 1 : 0) + (p2 ? 1 : 0);
        __syncwarp();
    }
    for (int i = lane; i < gap0; i += 32) x[l + i] = d[i];
    __syncwarp();
}

// positions [0, need) of the reference quicksort of x[0..n) (packed key << 32 | payload), by one CTA
__device__ void sel_qsort_prefix_cta(uint64_t *__restrict__ x, int n, int need, uint32_t *__restrict__ LA,
                                     uint32_t *__restrict__ LB, SelSmem &S, uint32_t *__restrict__ err) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (n <= 1 || need <= 0) return;
    if (tid == 0) {
        S.nb = S.ns = S.snext = 0;
        if (n > kSmall)
            S.bstack[S.nb++] = make_int2(0, n - 1);
        else
            S.slist[S.ns++] = make_int2(0, n - 1);
    }
    for (;;) {
        // ---- ranges above kSmall: whole CTA
        for (;;) {
            __syncthreads();
            const int nb = S.nb, ns = S.ns;
            if (nb == 0 || ns + 2 > kSmallList) break;
            const int2 rg = S.bstack[nb - 1];
            __syncthreads();
            const int j = sel_partition_cta(x, rg.x, rg.y, LA, LB, S);
            if (tid == 0) {
                int nbb = nb - 1, nss = ns;
                const int2 ch[2] = {make_int2(j + 1, rg.y), make_int2(rg.x, j - 1)};
                for (int c = 0; c < 2; c++) {
                    if (ch[c].y <= ch[c].x || ch[c].x >= need) continue;
                    if (ch[c].y - ch[c].x + 1 > kSmall) {
                        if (nbb < kBigStack)
                            S.bstack[nbb++] = ch[c];
                        else
                            atomicExch(err, 1u);
                    } else
                        S.slist[nss++] = ch[c];
                }
                S.nb = nbb, S.ns = nss;
            }
        }
        // ---- ranges of at most kSmall elements: one warp each
        if (S.ns == 0 && S.nb == 0) break;
        for (;;) {
            int i = 0;
            if (lane == 0) i = atomicAdd(&S.snext, 1);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= S.ns) break;
            const int2 rg = S.slist[i];
            sel_small_warp(x, rg.x, rg.y, need, S);
        }
        __syncthreads();
        if (tid == 0) S.ns = 0, S.snext = 0;
    }
    __syncthreads();
}

// candidates of one chunk / sub-block (sorted by query, then reference order; bounds[nq + 1]) appended to the
// per-query lists of the query block; one warp per query
__global__ void __launch_bounds__(256) k_block_append(const uint64_t *__restrict__ cv, const uint32_t *__restrict__ bounds, int nq,
                                                      int q0, uint64_t *__restrict__ vals, size_t capq,
                                                      uint32_t *__restrict__ count, uint32_t *__restrict__ err) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nq) return;
    const uint32_t b0 = bounds[w], n = bounds[w + 1] - b0;
    const uint32_t base = count[q0 + w];
    if ((size_t)base + n > capq) {
        if (lane == 0) atomicExch(err, 2u);
        return;
    }
    uint64_t *dst = vals + (size_t)(q0 + w) * capq + base;
    for (uint32_t i = lane; i < n; i += 32) dst[i] = cv[b0 + i];
    if (lane == 0) count[q0 + w] = base + n;
}

// PASS 2 head (fsearch.py:3039-3059): one CTA per query at a time
__global__ void __launch_bounds__(kSelThreads) k_h3_select(const uint64_t *__restrict__ vals, size_t capq,
                                                           const uint32_t *__restrict__ count, int nq, int vmax,
                                                           uint64_t *__restrict__ keys_all, uint32_t *__restrict__ la_all,
                                                           uint32_t *__restrict__ lb_all, size_t scap,
                                                           uint64_t *__restrict__ sel, uint32_t *__restrict__ sel_n,
                                                           uint32_t *__restrict__ next, uint32_t *__restrict__ err) {
    extern __shared__ __align__(16) unsigned char sel_smem_raw[];
    SelSmem &S = *reinterpret_cast<SelSmem *>(sel_smem_raw);
    uint64_t *x = keys_all + (size_t)blockIdx.x * scap;
    uint32_t *LA = la_all + (size_t)blockIdx.x * scap, *LB = lb_all + (size_t)blockIdx.x * scap;
    for (;;) {
        if (threadIdx.x == 0) S.q = (int)atomicAdd(next, 1u);
        __syncthreads();
        const int q = S.q;
        __syncthreads();
        if (q >= nq) break;
        const int n = (int)min((size_t)count[q], scap);
        const uint64_t *v = vals + (size_t)q * capq;
        const int limit = min(vmax, n);
        for (int i = threadIdx.x; i < n; i += kSelThreads) {
            const uint32_t score = (uint32_t)((v[i] >> 20) & 0xfffffu);
            x[i] = ((uint64_t)(0xffffffffu - score) << 32) | (uint32_t)i;
        }
        __syncthreads();
        sel_qsort_prefix_cta(x, n, limit, LA, LB, S, err);
        __syncthreads();
        for (int i = threadIdx.x; i < limit; i += kSelThreads) sel[(size_t)q * vmax + i] = v[(uint32_t)x[i]];
        if (threadIdx.x == 0) sel_n[q] = (uint32_t)n;
        __syncthreads();
    }
}

// test hook: prefix of the reference quicksort of packed elements, one array per CTA
__global__ void __launch_bounds__(kSelThreads) k_qsort_prefix_test(uint64_t *__restrict__ x, int n, int need,
                                                                   uint32_t *__restrict__ LA, uint32_t *__restrict__ LB,
                                                                   uint32_t *__restrict__ err) {
    extern __shared__ __align__(16) unsigned char sel_smem_raw[];
    SelSmem &S = *reinterpret_cast<SelSmem *>(sel_smem_raw);
    sel_qsort_prefix_cta(x, n, need, LA, LB, S, err);
}

static int sel_attr_done[64] = {};
static int sel_attrs(int device) {
    if (device >= 0 && device < 64 && sel_attr_done[device]) return SO_OK;
    SO_CUDA(cudaFuncSetAttribute(k_h3_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SelSmem)));
    SO_CUDA(cudaFuncSetAttribute(k_qsort_prefix_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SelSmem)));
    if (device >= 0 && device < 64) sel_attr_done[device] = 1;
    return SO_OK;
}

int BlockStore::prepare(so_ctx *c, i64 nq_, size_t capq_, int vmax_, cudaStream_t st) {
    int rc;
    nq = nq_, capq = std::max<size_t>(capq_, 1), vmax = std::max(vmax_, 1);
    if ((rc = sel_attrs(c->device)) != SO_OK) return rc;
    if ((rc = vals.reserve((size_t)nq * capq)) != SO_OK) return rc;
    if ((rc = count.reserve((size_t)nq + 8)) != SO_OK) return rc;
    grid = 148 * 4;  // 54 KB of shared memory per CTA: four resident CTAs per SM
    if ((i64)grid > nq) grid = (int)std::max<i64>(nq, 1);
    if ((rc = keys.reserve((size_t)grid * capq)) != SO_OK) return rc;
    if ((rc = la.reserve((size_t)grid * capq)) != SO_OK) return rc;
    if ((rc = lb.reserve((size_t)grid * capq)) != SO_OK) return rc;
    if ((rc = sel.reserve((size_t)nq * (size_t)vmax)) != SO_OK) return rc;
    if ((rc = sel_n.reserve((size_t)nq + 8)) != SO_OK) return rc;
    // count[nq] = work counter of k_h3_select, count[nq + 1] = error flag
    SO_CUDA(cudaMemsetAsync(count.p, 0, ((size_t)nq + 8) * 4, st));
    return SO_OK;
}

int BlockStore::append(const uint64_t *d_cv, const uint32_t *d_bounds, int n, int q0, cudaStream_t st) {
    if (n <= 0) return SO_OK;
    k_block_append<<<(n * 32 + 255) / 256, 256, 0, st>>>(d_cv, d_bounds, n, q0, vals.p, capq, count.p, count.p + nq + 1);
    SO_CUDA(cudaGetLastError());
    return SO_OK;
}

int BlockStore::select(cudaStream_t st) {
    k_h3_select<<<grid, kSelThreads, sizeof(SelSmem), st>>>(vals.p, capq, count.p, (int)nq, vmax, keys.p, la.p, lb.p, capq, sel.p,
                                                           sel_n.p, count.p + nq, count.p + nq + 1);
    SO_CUDA(cudaGetLastError());
    return SO_OK;
}

void BlockStore::release() {
    vals.release(), count.release(), keys.release(), la.release(), lb.release(), sel.release(), sel_n.release();
}

// so_qsort_prefix_device (test hook): the first `need` positions of the reference quicksort of keys[n]
int qsort_prefix_device(so_ctx *c, const uint32_t *keys, i64 n, i64 need, uint32_t *perm_out) {
    int rc;
    if ((rc = sel_attrs(c->device)) != SO_OK) return rc;
    if (n <= 0 || need <= 0) return SO_OK;
    need = std::min(need, n);
    std::vector<uint64_t> h((size_t)n);
    for (i64 i = 0; i < n; i++) h[(size_t)i] = ((uint64_t)keys[i] << 32) | (uint32_t)i;
    DBuf<uint64_t> x;
    DBuf<uint32_t> la, lb, err;
    if ((rc = x.reserve((size_t)n)) != SO_OK) return rc;
    if ((rc = la.reserve((size_t)n)) != SO_OK || (rc = lb.reserve((size_t)n)) != SO_OK || (rc = err.reserve(4)) != SO_OK) {
        x.release(), la.release(), lb.release(), err.release();
        return rc;
    }
    cudaMemcpyAsync(x.p, h.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(err.p, 0, 16, c->stream);
    k_qsort_prefix_test<<<1, kSelThreads, sizeof(SelSmem), c->stream>>>(x.p, (int)n, (int)need, la.p, lb.p, err.p);
    uint32_t herr = 0;
    cudaMemcpyAsync(h.data(), x.p, (size_t)need * 8, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    x.release(), la.release(), lb.release(), err.release();
    if (e != cudaSuccess) {
        set_error("CUDA error in k_qsort_prefix_test: %s", cudaGetErrorString(e));
        return SO_ENODEV;
    }
    if (herr) {
        set_error("device quicksort: range stack overflow");
        return SO_ELIMIT;
    }
    for (i64 i = 0; i < need; i++) perm_out[i] = (uint32_t)h[(size_t)i];
    return SO_OK;
}

}  // namespace so

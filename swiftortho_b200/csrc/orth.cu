// Orthology inference, device half (reference: bin/find_orth.py).
//
// so_orth_classify: one CTA per query group of the filtered hit table (bin/find_orth.py:158-234 `blastparse`
// yields one group per run of equal query ids; :298-348 `get_qIPO` classifies its hits):
//   * inside a group a target id keeps its best score, the earlier row on ties (`output[key][-1] < Score`, :229);
//   * sco_max[taxon] = best score per target taxon (Counter: starts at 0), out_max = best score against another taxon;
//   * same taxon: in-paralog candidate (IP) iff score >= out_max and query != target; other taxon: ortholog
//     candidate (OT) iff score >= sco_max[taxon of the target], else co-ortholog candidate (CO).
// Ids are ranks in byte order of the id strings, taxa are small integers, scores are the doubles the reference
// computes (raw bit score, score / first score of the query, score / alignment length): comparisons only, so the
// device result is bit exact.
// so_sort_pairs_u64: device radix sort (CUB, library) of 64-bit keys with 32-bit payloads, used for the reciprocal
// joins of the candidate lists (the reference shells out to GNU sort, bin/find_orth.py:476-478, 499-501, 552-554).
#include <cub/cub.cuh>

#include "context.h"

namespace so {

__global__ void __launch_bounds__(128) k_orth_classify(const uint64_t *__restrict__ goff, int64_t ngroups,
                                                       const uint32_t *__restrict__ qrank, const uint32_t *__restrict__ srank,
                                                       const uint32_t *__restrict__ qtax, const uint32_t *__restrict__ stax,
                                                       const double *__restrict__ score, uint32_t ntaxa,
                                                       uint8_t *__restrict__ cls) {
    extern __shared__ unsigned long long s_max[];  // [ntaxa] best score per target taxon (bit pattern of a double >= 0), [ntaxa] out_max
    for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const uint64_t a = goff[g], b = goff[g + 1];
        for (uint32_t t = threadIdx.x; t <= ntaxa; t += blockDim.x) s_max[t] = 0ull;
        __syncthreads();
        // representative rows (best score of their target, earliest on ties) + maxima
        for (uint64_t i = a + threadIdx.x; i < b; i += blockDim.x) {
            const uint32_t sid = srank[i];
            const double sc = score[i];
            bool rep = true;
            for (uint64_t j = a; j < b; j++) {
                if (j == i || srank[j] != sid) continue;
                const double sj = score[j];
                if (sj > sc || (sj == sc && j < i)) {
                    rep = false;
                    break;
                }
            }
            cls[i] = rep ? 1 : 0;
            if (sc > 0.) {  // (every row takes part in the maxima, like the reference's first loop over the deduplicated
                            //  list: a non-representative row never exceeds its representative)
                const unsigned long long bits = (unsigned long long)__double_as_longlong(sc);
                atomicMax(&s_max[stax[i]], bits);
                if (qtax[i] != stax[i]) atomicMax(&s_max[ntaxa], bits);
            }
        }
        __syncthreads();
        const double out_max = __longlong_as_double((long long)s_max[ntaxa]);
        for (uint64_t i = a + threadIdx.x; i < b; i += blockDim.x) {
            if (!cls[i]) continue;
            const double sc = score[i];
            uint8_t c = 0;
            if (qtax[i] == stax[i]) {
                if (sc >= out_max && qrank[i] != srank[i]) c = 1;  // IP
            } else {
                c = sc >= __longlong_as_double((long long)s_max[stax[i]]) ? 2 : 3;  // OT : CO
            }
            cls[i] = c;
        }
        __syncthreads();
    }
}

}  // namespace so

using namespace so;

static int orth_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: swiftortho_b200 has no CPU fallback");
        return SO_ENODEV;
    }
    if (device < 0 || device >= ndev) {
        set_error("device %d out of range (%d visible)", device, ndev);
        return SO_EINVAL;
    }
    SO_CUDA(cudaSetDevice(device));
    return SO_OK;
}

extern "C" int so_orth_classify(int device, const uint64_t *group_offsets, int64_t n_groups, const uint32_t *qrank,
                                const uint32_t *srank, const uint32_t *qtax, const uint32_t *stax, const double *score,
                                uint32_t n_taxa, uint8_t *cls) {
    if (n_groups < 0 || (n_groups > 0 && (!group_offsets || !qrank || !srank || !qtax || !stax || !score || !cls))) {
        set_error("so_orth_classify: bad argument");
        return SO_EINVAL;
    }
    int rc = orth_device(device);
    if (rc != SO_OK) return rc;
    if (n_groups == 0) return SO_OK;
    if ((size_t)(n_taxa + 1) * 8 > 200 * 1024) {
        set_error("so_orth_classify: more than 25599 taxa");
        return SO_ELIMIT;
    }
    const size_t n = (size_t)group_offsets[n_groups];
    uint64_t *d_goff = nullptr;
    uint32_t *d_u32 = nullptr;
    double *d_sc = nullptr;
    uint8_t *d_cls = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_goff, ((size_t)n_groups + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_u32, std::max<size_t>(n, 1) * 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_sc, std::max<size_t>(n, 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_cls, std::max<size_t>(n, 1));
    if (e == cudaSuccess) e = cudaMemcpy(d_goff, group_offsets, ((size_t)n_groups + 1) * 8, cudaMemcpyHostToDevice);
    const uint32_t *src[4] = {qrank, srank, qtax, stax};
    for (int k = 0; k < 4 && e == cudaSuccess && n; k++) e = cudaMemcpy(d_u32 + (size_t)k * n, src[k], n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n) e = cudaMemcpy(d_sc, score, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const size_t smem = ((size_t)n_taxa + 1) * 8;
        e = cudaFuncSetAttribute(k_orth_classify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024));
        if (e == cudaSuccess) {
            const int grid = (int)std::min<int64_t>(n_groups, 148 * 16);
            k_orth_classify<<<grid, 128, smem>>>(d_goff, n_groups, d_u32, d_u32 + n, d_u32 + 2 * n, d_u32 + 3 * n, d_sc, n_taxa, d_cls);
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess && n) e = cudaMemcpy(cls, d_cls, n, cudaMemcpyDeviceToHost);
    cudaFree(d_goff), cudaFree(d_u32), cudaFree(d_sc), cudaFree(d_cls);
    if (e != cudaSuccess) {
        set_error("CUDA error in so_orth_classify: %s", cudaGetErrorString(e));
        return SO_ENODEV;
    }
    return SO_OK;
}

extern "C" int so_sort_pairs_u64(int device, uint64_t *keys, uint32_t *vals, int64_t n) {
    if (n < 0 || (n > 0 && (!keys || !vals)) || n > 0x7fffff00ll) {
        set_error("so_sort_pairs_u64: bad argument");
        return SO_EINVAL;
    }
    int rc = orth_device(device);
    if (rc != SO_OK) return rc;
    if (n == 0) return SO_OK;
    uint64_t *dk = nullptr, *dk2 = nullptr;
    uint32_t *dv = nullptr, *dv2 = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    cudaError_t e = cudaMalloc((void **)&dk, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&dk2, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&dv, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&dv2, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(dk, keys, (size_t)n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dv, vals, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dk2, dv, dv2, (int)n);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, std::max<size_t>(tb, 16));
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dk2, dv, dv2, (int)n);
    if (e == cudaSuccess) e = cudaMemcpy(keys, dk2, (size_t)n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(vals, dv2, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(dk), cudaFree(dk2), cudaFree(dv), cudaFree(dv2), cudaFree(tmp);
    if (e != cudaSuccess) {
        set_error("CUDA error in so_sort_pairs_u64: %s", cudaGetErrorString(e));
        return SO_ENODEV;
    }
    return SO_OK;
}

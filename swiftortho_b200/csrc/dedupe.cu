// Redundancy pre-filter: 64-bit content hash of every sequence on the device.
// scripts/nr_flt.py:8-27 keys a dict with the full sequence string to collapse exact duplicates before the
// all-vs-all search (scripts/run_all_fast.py:110-119); here the sequences are hashed by one warp each (every lane
// runs FNV-1a-64 over the residues i = lane (mod 32), the 32 lane states and the length are folded in lane order),
// the host groups by hash and confirms equality byte for byte, so the grouping is exact whatever the hash does.
#include "context.h"

namespace so {

__global__ void __launch_bounds__(256) k_seq_hash(const uint8_t *__restrict__ res, const uint64_t *__restrict__ off, int64_t n,
                                                  uint64_t *__restrict__ out) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint64_t a = off[w], b = off[w + 1];
    uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)lane;
    for (uint64_t i = a + (uint64_t)lane; i < b; i += 32) h = (h ^ (uint64_t)res[i]) * 0x100000001b3ull;
    uint64_t acc = (b - a) * 0x9e3779b97f4a7c15ull;
    for (int l = 0; l < 32; l++) {
        const uint64_t hl = __shfl_sync(0xffffffffu, h, l);
        acc = (acc ^ hl) * 0x100000001b3ull;
        acc ^= acc >> 29;
    }
    if (lane == 0) out[w] = acc;
}

}  // namespace so

extern "C" int so_seq_hash(int device, const uint8_t *residues, const uint64_t *offsets, int64_t n, uint64_t *hashes) {
    using namespace so;
    if (n < 0 || (n > 0 && (!offsets || !hashes))) {
        set_error("so_seq_hash: bad argument");
        return SO_EINVAL;
    }
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
    if (n == 0) return SO_OK;
    SO_CUDA(cudaSetDevice(device));
    const uint64_t base = offsets[0], bytes = offsets[n] - base;
    uint8_t *d_res = nullptr;
    uint64_t *d_off = nullptr, *d_out = nullptr;
    std::vector<uint64_t> rel((size_t)n + 1);
    for (int64_t i = 0; i <= n; i++) rel[(size_t)i] = offsets[i] - base;
    cudaError_t e = cudaMalloc((void **)&d_res, (size_t)bytes + 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_off, ((size_t)n + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, (size_t)n * 8);
    if (e == cudaSuccess && bytes) e = cudaMemcpy(d_res, residues + base, (size_t)bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_off, rel.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const int64_t blocks = (n * 32 + 255) / 256;
        k_seq_hash<<<(unsigned)blocks, 256>>>(d_res, d_off, n, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(hashes, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_res), cudaFree(d_off), cudaFree(d_out);
    if (e != cudaSuccess) {
        set_error("CUDA error in so_seq_hash: %s", cudaGetErrorString(e));
        return SO_ENODEV;
    }
    return SO_OK;
}

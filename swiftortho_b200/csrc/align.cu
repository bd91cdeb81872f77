// K7 / K8 / K9: the reference's banded single-matrix local alignment (kswat_st,
// lib/fsearch.py:1357-1476) as two sm_100a kernels.
//
//   k_banded_dp   one THREAD per alignment (inter-sequence parallelism, tasks sorted by length so
//                 the 32 lanes of a warp walk similar row counts).  The band of the reference is
//                 j - i in [-16, +15] (fsearch.py:1393: start=max(1,i-16), end=min(i+16,l0)) =
//                 32 cells per row, kept entirely in registers: S[d] = score of lane d of the
//                 previous row, Dc[d] = that cell's contribution as a vertical predecessor
//                 (score + ge if its trace is '|' else score + go).  Cells are evaluated in the
//                 reference's row-major order, so "first strict maximum" (fsearch.py:1401) needs
//                 no reduction.  max(0, I, M, D) is one DPX instruction (__vimax3_s32_relu).
//                 BLOSUM62 lives in shared memory as a 32x32 int8 table over 5-bit residue classes.
//                 The 2-bit trace (3 '\', 2 '-', 1 '|', 0 '*'; priority M > I > D as in
//                 fsearch.py:1404-1411) is written as one 64-bit word per row, laid out
//                 [warp][row][lane] so a warp's store is one coalesced 256-byte line.
//   k_traceback   one thread per alignment walks the trace from (i_max, j_max) with the reference's
//                 `while i > 0 or j > 0` loop (fsearch.py:1418-1443), including the walk along
//                 row 0 / column 0, and accumulates the statistics of fsearch.py:1454-1469
//                 (identity on raw bytes, mismatches incl. gap columns, gap count = ceil(run/2)).
//
// Guard cells of the reference (fsearch.py:1379-1389) are constants here: left of the band and
// column 0 contribute I = go; the cell right of the band in the previous row contributes D = go
// (its score is always 0 and its trace never '|').
#include <algorithm>
#include <numeric>

#include "context.h"

namespace so {

__constant__ int8_t c_score[kClasses * kClasses];
__constant__ uint8_t c_code[256];

int upload_tables() {
    int8_t tbl[kClasses * kClasses];
    uint8_t code[256];
    make_score_table(tbl);
    make_code_table(code);
    SO_CUDA(cudaMemcpyToSymbol(c_score, tbl, sizeof tbl));
    SO_CUDA(cudaMemcpyToSymbol(c_code, code, sizeof code));
    return SO_OK;
}

struct AlnTask {
    const uint8_t *s0;  // columns (the side with the shorter remainder), already offset to its start
    const uint8_t *s1;  // rows
    int len0, len1;
};

struct DpOut {
    int score, imax, jmax, rows;
};

struct TbOut {
    int i0, j0, al, nid, mis, gap, bad, pad;
};

// One band row.  Cell values are PACKED: P = score * 4 + code with code 3 '\' (came from M), 2 '-' (from I),
// 1 '|' (from D), 0 '*' (none).  Taking max(0, I', M', D') over packed candidates yields the score AND the
// reference's trace priority M > I > D (fsearch.py:1404-1411) in one DPX max3.relu: equal scores differ in
// the low bits, a zero-score M gives 3 ('\'), all-negative gives 0 ('*').  No equality test against the max
// result is needed (ptxas 12.9 miscompiles `B == operand` after max.s32.relu on sm_100a, tools/test_vimax.cu).
//   S3[d]  = P | 3 of lane d in the previous row      -> M' = S3[d] + 4*sub
//   Dc[d]  = vertical candidate of that cell: P - 4 if its code is '|' (extend, -1) else (P|3) - 46 (open, -11)
//   Ip     = horizontal candidate of the cell to the left: P - 4 if its code is '-' else (P|3) - 45
// Row maximum with "first column wins" = max over (P|3)*32 + (31 - d).
// Instruction diet of the cell (ncu: the kernel is issue / alu-pipe bound):
//   * substitution score: the table is [class1][256] int8, 256-byte aligned, so ONE PRMT splices the class-0
//     byte of the window into the row's base address (no extract + add), then LDS.S8;
//   * the two "extend or open" candidates are P - adj[t] with adj looked up by PRMT from a packed constant
//     (t = trace code in the low bits of P), instead of compare + select + subtract;
//   * the 2-bit trace codes are shifted into the row word with one funnel shift (SHF) per cell.
__device__ __forceinline__ uint32_t prmt_u32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ int lds_s8(uint32_t addr) {
    int v;
    asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <bool EDGE>
__device__ __forceinline__ void dp_row(int i, int len0, bool live, uint32_t rowbase, uint32_t ksel, int (&S3)[32],
                                       int (&Dc)[33], const uint32_t (&W)[8], uint32_t &tlo, uint32_t &thi,
                                       int &rowkey) {
    // horizontal / vertical candidate = P - adj[t]:  extend (-1) when the trace code allows it, else open (-11);
    // every candidate carries its own trace code in the low bits (I: 2, D: 1)
    const uint32_t KI = 42u | (43u << 8) | (4u << 16) | (45u << 24);   // t = 0, 1, 2 ('-': P - 4), 3
    const uint32_t KD = 43u | (4u << 8) | (45u << 16) | (46u << 24);   // t = 0, 1 ('|': P - 4), 2, 3
    int Ip = -42;  // left of lane 0: guard cell / column 0 -> score 0, not '-' -> (0 - 11) * 4 + 2
    uint32_t acc = 0;
    int keyprev = 0;
    rowkey = 0;
#pragma unroll
    for (int d = 0; d < 32; d++) {
        const uint32_t addr = prmt_u32(W[d >> 2], rowbase, 0x7650u | (uint32_t)(d & 3));
        const int Mp = S3[d] + lds_s8(addr);
        int P = __vimax3_s32_relu(Ip, Mp, Dc[d + 1]);
        if (EDGE) {
            const bool valid = live && (unsigned)(i + d - 17) < (unsigned)len0;  // 1 <= j < l0
            if (!valid) P = 0;  // score 0, code '*': contributes I = D = open from 0, M from 0
        }
        const uint32_t tsel = ((uint32_t)P & 3u) | ksel;
        Ip = P - (int)prmt_u32(KI, 0u, tsel);
        Dc[d] = P - (int)prmt_u32(KD, 0u, tsel);
        const int P3 = P | 3;
        S3[d] = P3;
        const int key = P3 * 32 + (31 - d);
        if (d & 1)
            rowkey = __vimax3_s32(rowkey, keyprev, key);
        else
            keyprev = key;
        acc = __funnelshift_r(acc, (uint32_t)P, 2);  // the code of cell d ends up at bits 2d (mod 32)
        if (d == 15) tlo = acc, acc = 0;
    }
    thi = acc;
}

__global__ void __launch_bounds__(128) k_banded_dp(const AlnTask *__restrict__ tasks, int n,
                                                   const uint64_t *__restrict__ warp_base,
                                                   uint64_t *__restrict__ trace, DpOut *__restrict__ out, uint32_t ksel) {
    // ksel = 0x4440 (PRMT selector: byte t of the constant, zeros above); a kernel argument so that it stays in a
    // register and (P & 3) | ksel is ONE LOP3 instead of two with immediates
    __shared__ __align__(256) int8_t s_tbl4[kClasses * 256];  // 4 * BLOSUM62 (fits int8: -16 .. 44), row stride 256
    __shared__ uint8_t s_code[256];
    for (int k = threadIdx.x; k < kClasses * kClasses; k += blockDim.x)
        s_tbl4[(k / kClasses) * 256 + (k % kClasses)] = (int8_t)(4 * c_score[k]);
    const uint32_t tblbase = (uint32_t)__cvta_generic_to_shared(s_tbl4);

    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_code[k] = c_code[k];
    __syncthreads();

    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool active = t < n;
    AlnTask tk;
    if (active)
        tk = tasks[t];
    else {
        tk.s0 = tk.s1 = nullptr;
        tk.len0 = tk.len1 = 0;
    }
    const int len0 = tk.len0;
    const int nrows = active ? min(tk.len1, len0 + 16) : 0;  // rows that hold at least one band cell
    int wrows = nrows;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wrows = max(wrows, __shfl_xor_sync(0xffffffffu, wrows, o));
    uint64_t *tr = trace + warp_base[t >> 5] + lane;

    int S3[32], Dc[33];
#pragma unroll
    for (int d = 0; d < 32; d++) S3[d] = 3, Dc[d] = -43;  // row 0: score 0, trace '-' (never '|')
    Dc[32] = -43;  // the cell right of the band in the previous row: score 0, never '|'
    // window of s0 classes: byte d holds class(s0[i + d - 17]) for the current row i
    uint32_t W[8];
#pragma unroll
    for (int k = 0; k < 8; k++) W[k] = 0;
#pragma unroll
    for (int d = 16; d < 32; d++) {
        const int idx = d - 16;  // row 1: s0[d - 16]
        const uint32_t cls = (idx < len0) ? s_code[tk.s0[idx]] : 0;
        W[d >> 2] |= cls << ((d & 3) * 8);
    }
    int best3 = 3, besti = 0, bestd = 0;  // best3 = (best score) * 4 + 3
    for (int i = 1; i <= wrows; i++) {
        const bool live = i <= nrows;
        const uint32_t rowbase = tblbase + (live ? (uint32_t)s_code[tk.s1[i - 1]] << 8 : 0u);
        uint32_t tlo, thi;
        int rowkey;
        const bool interior = live && i >= 17 && i + 15 <= len0;
        if (__all_sync(0xffffffffu, interior))
            dp_row<false>(i, len0, live, rowbase, ksel, S3, Dc, W, tlo, thi, rowkey);
        else
            dp_row<true>(i, len0, live, rowbase, ksel, S3, Dc, W, tlo, thi, rowkey);
        if (live) tr[(size_t)(i - 1) * 32] = ((uint64_t)thi << 32) | tlo;
        // first strict maximum in row-major order (fsearch.py:1401)
        const int rs = rowkey >> 5;
        if (rs > best3) {
            best3 = rs;
            besti = i;
            bestd = 31 - (rowkey & 31);
        }
        // slide the window: drop byte 0, append class(s0[i + 15]) for row i + 1
        const int nidx = i + 15;
        const uint32_t ncls = (nidx < len0) ? s_code[tk.s0[nidx]] : 0;
#pragma unroll
        for (int k = 0; k < 7; k++) W[k] = __funnelshift_r(W[k], W[k + 1], 8);
        W[7] = (W[7] >> 8) | (ncls << 24);
    }
    if (active) {
        DpOut o;
        o.score = best3 >> 2;
        o.imax = o.score > 0 ? besti : 0;
        o.jmax = o.score > 0 ? (besti + bestd - 16) : 0;
        o.rows = nrows;
        out[t] = o;
    }
}

// WAVE = the trace layout of k_banded_dp_wave: per task, words [row block of 16][lane h], 4 bits per row
template <bool WAVE>
__global__ void __launch_bounds__(128) k_traceback(const AlnTask *__restrict__ tasks, int n,
                                                   const uint64_t *__restrict__ warp_base,
                                                   const uint64_t *__restrict__ trace,
                                                   const DpOut *__restrict__ dp, TbOut *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const AlnTask tk = tasks[t];
    const uint64_t *tr = WAVE ? trace + warp_base[t] : trace + warp_base[t >> 5] + (threadIdx.x & 31);
    int i = dp[t].imax, j = dp[t].jmax;
    int al = 0, nid = 0, mis = 0, gap = 0, bad = 0;
    int run_type = 0, run_len = 0;  // 2: trace '-', 1: trace '|'
    int cached_row = -1;
    uint64_t word = 0;
    while (i > 0 || j > 0) {
        int code;
        if (i == 0)
            code = 2;  // row 0 holds '-' (fsearch.py:1379-1382)
        else if (j == 0)
            code = 1;  // column 0 holds '|' (fsearch.py:1383-1386)
        else {
            const int d = j - i + 16;
            if (d < 0 || d > 31) {
                bad = 1;
                break;
            }
            if (WAVE) {
                const int key = ((i - 1) >> 4) * 16 + (d >> 1);  // (row block, lane)
                if (key != cached_row) {
                    word = tr[key];
                    cached_row = key;
                }
                code = (int)((word >> (4 * ((i - 1) & 15) + 2 * (d & 1))) & 3);
            } else {
                if (i != cached_row) {
                    word = tr[(size_t)(i - 1) * 32];
                    cached_row = i;
                }
                code = (int)((word >> (2 * d)) & 3);
            }
        }
        if (code == 0) break;
        al++;
        if (code == 3) {
            if (tk.s0[j - 1] == tk.s1[i - 1])
                nid++;
            else
                mis++;
            i--;
            j--;
            gap += (run_len + 1) >> 1;
            run_len = 0;
            run_type = 0;
        } else {
            mis++;
            if (run_type != code) {
                gap += (run_len + 1) >> 1;
                run_len = 0;
                run_type = code;
            }
            run_len++;
            if (code == 2)
                j--;
            else
                i--;
        }
    }
    gap += (run_len + 1) >> 1;
    TbOut o;
    o.i0 = i, o.j0 = j, o.al = al, o.nid = nid, o.mis = mis, o.gap = gap, o.bad = bad, o.pad = 0;
    out[t] = o;
}

// -----------------------------------------------------------------------------------------------
// k_banded_dp_wave: the same recurrence with 16 LANES per alignment, for launches too small to fill the GPU with one
// thread per alignment (an alignment round of the search pipeline carries ~2*10^4 tasks = 4 warps per SM for
// k_banded_dp, whose cell chain is latency bound at that occupancy).
//
// Lane h of a half-warp owns band columns d = 2h and 2h + 1.  Cell (i, d) needs (i, d-1) [I], (i-1, d) [M] and
// (i-1, d+1) [D]; with row i = T - h at macro-step T every lane computes its two cells per macro-step:
//   sub-step 0: (i, 2h)   I <- lane h-1's cell (i, 2h-1), computed at macro-step T-1 (shuffle up, left guard for h = 0)
//                         M, D <- own cells of row i-1 (macro-step T-1)
//   sub-step 1: (i, 2h+1) I <- own cell of sub-step 0;  M <- own;  D <- lane h+1's cell (i-1, 2h+2), computed in
//                         sub-step 0 of THIS macro-step (shuffle down, right-of-band constant for h = 15)
// so all 16 lanes are busy from macro-step 16 on (pipeline fill / drain: 15 macro-steps per alignment).  Cells are the
// packed values of k_banded_dp (score * 4 + trace code, DPX max3.relu, PRMT-looked-up gap adjustments); cells outside
// the matrix evaluate to 0, which reproduces the guard cells.  "First strict maximum in row-major order"
// (fsearch.py:1401): every lane keeps its own first maximum (its cells come in row-major order), the 16 lanes are
// reduced on (score desc, row asc, column asc).  Trace: 4 bits per lane and row, 16 rows per 64-bit word, words laid out
// [row block][lane] (one 128-byte line per half-warp and 16 rows); k_traceback<true> reads that layout.
// -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_banded_dp_wave(const AlnTask *__restrict__ tasks, int n,
                                                        const uint64_t *__restrict__ task_base,
                                                        uint64_t *__restrict__ trace, DpOut *__restrict__ out, uint32_t ksel) {
    __shared__ __align__(256) int8_t s_tbl4[kClasses * 256];
    __shared__ uint8_t s_code[256];
    for (int k = threadIdx.x; k < kClasses * kClasses; k += blockDim.x)
        s_tbl4[(k / kClasses) * 256 + (k % kClasses)] = (int8_t)(4 * c_score[k]);
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_code[k] = c_code[k];
    __syncthreads();
    const uint32_t tblbase = (uint32_t)__cvta_generic_to_shared(s_tbl4);
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;  // one task per 16 lanes
    const int h = threadIdx.x & 15;
    const bool active = t < n;
    AlnTask tk;
    if (active)
        tk = tasks[t];
    else {
        tk.s0 = tk.s1 = nullptr;
        tk.len0 = tk.len1 = 0;
    }
    const int len0 = tk.len0;
    const int nrows = active ? min(tk.len1, len0 + 16) : 0;
    int wsteps = nrows ? nrows + 15 : 0;  // macro-steps of this alignment; the warp runs the longer of its two
    wsteps = max(wsteps, __shfl_xor_sync(0xffffffffu, wsteps, 16));
    uint64_t *tr = active ? trace + task_base[t] + h : trace;
    const uint32_t KI = 42u | (43u << 8) | (4u << 16) | (45u << 24);
    const uint32_t KD = 43u | (4u << 8) | (45u << 16) | (46u << 24);
    // state of the lane's two columns after the previous row (row 0: score 0, not extendable)
    int S3A = 3, S3B = 3, DcB = -43;  // DcB: vertical candidate of (i-1, 2h+1), used by cell A of row i
    int IpB = -42;                    // horizontal candidate of the lane's cell B of row i, wanted by lane h+1 next macro-step
    int best3 = 3, besti = 0, bestd = 0;
    uint64_t acc = 0;
    // column residues: cell A faces s0[i + 2h - 17], cell B faces s0[i + 2h - 16]; with i = T - h: s0[T + h - 17 (+1)]
    int cA = 0;
    {
        const int j0 = 1 + h - 17;  // macro-step T = 1
        cA = (active && j0 >= 0 && j0 < len0) ? s_code[tk.s0[j0]] : 0;
    }
    for (int T = 1; T <= wsteps; T++) {
        const int i = T - h;
        const bool vrow = i >= 1 && i <= nrows;
        const int jb = T + h - 16;  // index into s0 of cell B's residue (= j - 1)
        const int cB = (active && jb >= 0 && jb < len0) ? (int)s_code[tk.s0[jb]] : 0;
        const uint32_t rowbase = tblbase + (vrow ? (uint32_t)s_code[tk.s1[i - 1]] << 8 : 0u);
        // ---- sub-step 0: cell (i, 2h);  j = i + 2h - 16
        int IpA = __shfl_up_sync(0xffffffffu, IpB, 1, 16);
        if (h == 0) IpA = -42;
        const int jA = i + 2 * h - 16;
        const int MpA = S3A + lds_s8(rowbase + (uint32_t)cA);
        int PA = __vimax3_s32_relu(IpA, MpA, DcB);
        if (!(vrow && jA >= 1 && jA <= len0)) PA = 0;
        const uint32_t tselA = ((uint32_t)PA & 3u) | ksel;
        const int IpFromA = PA - (int)prmt_u32(KI, 0u, tselA);
        const int DcFromA = PA - (int)prmt_u32(KD, 0u, tselA);
        // ---- sub-step 1: cell (i, 2h + 1); its vertical neighbour (i-1, 2h+2) is lane h+1's cell A of this macro-step
        int DcN = __shfl_down_sync(0xffffffffu, DcFromA, 1, 16);
        if (h == 15) DcN = -43;
        const int MpB = S3B + lds_s8(rowbase + (uint32_t)cB);
        int PB = __vimax3_s32_relu(IpFromA, MpB, DcN);
        if (!(vrow && jA + 1 >= 1 && jA + 1 <= len0)) PB = 0;
        const uint32_t tselB = ((uint32_t)PB & 3u) | ksel;
        IpB = PB - (int)prmt_u32(KI, 0u, tselB);
        DcB = PB - (int)prmt_u32(KD, 0u, tselB);
        S3A = PA | 3, S3B = PB | 3;
        if (vrow) {
            // first strict maximum, row-major inside the lane: (i, 2h) before (i, 2h+1), rows ascending
            if (S3A > best3) best3 = S3A, besti = i, bestd = 2 * h;
            if (S3B > best3) best3 = S3B, besti = i, bestd = 2 * h + 1;
            const uint32_t nib = ((uint32_t)PA & 3u) | (((uint32_t)PB & 3u) << 2);
            const int r = (i - 1) & 15;
            acc |= (uint64_t)nib << (4 * r);
            if (r == 15 || i == nrows) {
                tr[(size_t)((i - 1) >> 4) * 16] = acc;
                acc = 0;
            }
        }
        cA = cB;  // next macro-step: cell A faces this macro-step's cell B residue
    }
    // reduce the 16 lanes: score desc, row asc, column asc
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const int b3 = __shfl_xor_sync(0xffffffffu, best3, o, 16), bi = __shfl_xor_sync(0xffffffffu, besti, o, 16);
        const int bd = __shfl_xor_sync(0xffffffffu, bestd, o, 16);
        if (b3 > best3 || (b3 == best3 && (bi < besti || (bi == besti && bd < bestd)))) best3 = b3, besti = bi, bestd = bd;
    }
    if (active && h == 0) {
        DpOut o;
        o.score = best3 >> 2;
        o.imax = o.score > 0 ? besti : 0;
        o.jmax = o.score > 0 ? (besti + bestd - 16) : 0;
        o.rows = nrows;
        out[t] = o;
    }
}

// cells the reference fills for (len0, len1): rows i = 1..l1-1, columns [max(1,i-16), min(i+16,l0))
static inline i64 band_cells(i64 len0, i64 len1) {
    const i64 l0 = len0 + 1, rows = std::min<i64>(len1, len0 + 16);
    i64 cl = 0;
    // closed form per row: hi(i) - lo(i), hi = min(i+16, l0), lo = max(1, i-16)
    // rows 1..16: lo = 1; rows >= 17: lo = i-16.   rows <= l0-16: hi = i+16; beyond: hi = l0.
    for (i64 i = 1; i <= std::min<i64>(rows, 16); i++) cl += std::min<i64>(i + 16, l0) - 1;
    if (rows > 16) {
        const i64 a = 17, b = rows;                    // lo = i - 16
        const i64 m = std::min<i64>(b, l0 - 16);       // rows a..m: hi = i + 16 -> 32 cells
        if (m >= a) cl += 32 * (m - a + 1);
        const i64 a2 = std::max<i64>(a, m + 1);        // rows a2..b: hi = l0 -> l0 - i + 16 cells
        if (b >= a2) cl += (l0 + 16) * (b - a2 + 1) - (a2 + b) * (b - a2 + 1) / 2;
    }
    return cl;
}

void merge_align_stats(so_ctx *c) {
    so_stats &a = c->stats_aln, &d = c->stats;
    d.alignments += a.alignments, d.dp_cells += a.dp_cells, d.kernel_launches += a.kernel_launches;
    d.ms_align += a.ms_align, d.ms_dp += a.ms_dp, d.ms_traceback += a.ms_traceback, d.ms_host += a.ms_host;
    d.h2d_bytes += a.h2d_bytes, d.d2h_bytes += a.d2h_bytes;
    memset(&a, 0, sizeof a);
}

// -----------------------------------------------------------------------------------------------
// Host driver: resolves pairs to device tasks, sorts by length, launches, converts coordinates.
// -----------------------------------------------------------------------------------------------
int align_pairs(so_ctx *c, const so_pair *pairs, i64 n, so_aln *out) {
    if (n <= 0) return SO_OK;
    if (!c->d_tres || !c->d_qres) {
        set_error("so_align_batch: targets and queries must be loaded first");
        return SO_EINVAL;
    }
    Timer tm;
    struct Prep {
        int len0, len1;
        int qst, sst;  // clamped starts inside the slices
        bool swap;
    };
    std::vector<Prep> prep((size_t)n);
    std::vector<AlnTask> tasks((size_t)n);
    std::vector<int> order((size_t)n);
    for (i64 k = 0; k < n; k++) {
        const so_pair &p = pairs[k];
        if (p.query < 0 || p.query >= c->n_q || p.target < 0 || p.target >= c->n_t) {
            set_error("so_align_batch: pair %lld addresses a sequence out of range", (long long)k);
            return SO_EINVAL;
        }
        i64 ql_full = (i64)(c->q_off[p.query + 1] - c->q_off[p.query]);
        i64 tl_full = (i64)(c->t_off[p.target + 1] - c->t_off[p.target]);
        i64 qo = std::min<i64>(std::max<i64>(p.q_off, 0), ql_full), to = std::min<i64>(std::max<i64>(p.t_off, 0), tl_full);
        i64 qlen = std::min<i64>(std::max<i64>(p.q_len, 0), ql_full - qo), tlen = std::min<i64>(std::max<i64>(p.t_len, 0), tl_full - to);
        // kswat_st argument clamping (fsearch.py:1359-1362)
        i64 qst = std::min<i64>(std::max<i64>(p.qst, 0), qlen), sst = std::min<i64>(std::max<i64>(p.sst, 0), tlen);
        i64 qrem = qlen - qst, trem = tlen - sst;
        if (qrem > 4096 || trem > 4096) {
            set_error("so_align_batch: slice remainder above 4096 (tile the pair like kswat_st_long)");
            return SO_ELIMIT;
        }
        Prep &r = prep[(size_t)k];
        r.qst = (int)qst, r.sst = (int)sst;
        const uint8_t *qp = c->d_qres + c->q_off[p.query] + qo + qst;
        const uint8_t *tp = c->d_tres + c->t_off[p.target] + to + sst;
        // s0 (columns) is the query iff its remainder is strictly shorter (fsearch.py:1364-1369)
        if (qrem < trem) {
            r.swap = false, r.len0 = (int)qrem, r.len1 = (int)trem;
            tasks[(size_t)k] = AlnTask{qp, tp, r.len0, r.len1};
        } else {
            r.swap = true, r.len0 = (int)trem, r.len1 = (int)qrem;
            tasks[(size_t)k] = AlnTask{tp, qp, r.len0, r.len1};
        }
        order[(size_t)k] = (int)k;
    }
    // order: len0 descending, then len1 descending, then pair index (tasks of similar length share a warp); two stable
    // counting passes (lengths are at most 4096) instead of a comparison sort: a round carries up to ~10^5 tasks
    if (n > 64) {
        std::vector<int> tmp((size_t)n), cnt(4098);
        for (int pass = 0; pass < 2; pass++) {
            std::fill(cnt.begin(), cnt.end(), 0);
            const std::vector<int> &src = pass == 0 ? order : tmp;
            std::vector<int> &dst = pass == 0 ? tmp : order;
            auto key = [&](int k) { return 4096 - (pass == 0 ? prep[(size_t)k].len1 : prep[(size_t)k].len0); };
            for (i64 k = 0; k < n; k++) cnt[(size_t)key(src[(size_t)k]) + 1]++;
            for (int b = 0; b < 4097; b++) cnt[(size_t)b + 1] += cnt[(size_t)b];
            for (i64 k = 0; k < n; k++) dst[(size_t)cnt[(size_t)key(src[(size_t)k])]++] = src[(size_t)k];
        }
    } else
        std::sort(order.begin(), order.end(), [&](int a, int b) {
            if (prep[(size_t)a].len0 != prep[(size_t)b].len0) return prep[(size_t)a].len0 > prep[(size_t)b].len0;
            if (prep[(size_t)a].len1 != prep[(size_t)b].len1) return prep[(size_t)a].len1 > prep[(size_t)b].len1;
            return a < b;
        });
    std::vector<AlnTask> sorted((size_t)n);
    for (i64 k = 0; k < n; k++) sorted[(size_t)k] = tasks[(size_t)order[(size_t)k]];
    // small launches (alignment rounds of the search) use the 16-lanes-per-alignment kernel; SO_DP_WAVE=0/1 forces
    i64 wave_below = 60000;
    if (const char *e = getenv("SO_DP_WAVE")) wave_below = atoi(e) ? (i64)1 << 40 : 0;
    const bool wave = n < wave_below;
    const i64 nwarps = (n + 31) / 32;
    const i64 nwarps_pad = wave ? n : ((n + 127) / 128) * 4;  // wave: one base per TASK
    std::vector<uint64_t> wbase((size_t)nwarps_pad + 1, 0);
    i64 cells = 0;
    if (wave) {
        for (i64 k = 0; k < n; k++) {
            const AlnTask &t = sorted[(size_t)k];
            const int rows = std::min(t.len1, t.len0 + 16);
            wbase[(size_t)k + 1] = wbase[(size_t)k] + (uint64_t)((rows + 15) / 16) * 16;
        }
    } else
        for (i64 w = 0; w < nwarps_pad; w++) {
            int rows = 0;
            for (i64 k = w * 32; k < std::min<i64>(n, w * 32 + 32); k++) {
                const AlnTask &t = sorted[(size_t)k];
                rows = std::max(rows, std::min(t.len1, t.len0 + 16));
            }
            wbase[(size_t)w + 1] = wbase[(size_t)w] + (uint64_t)rows * 32;
        }
    (void)nwarps;
    size_t need_trace = (size_t)wbase[(size_t)nwarps_pad] + 64;
    int rc;
    if ((rc = c->trace.reserve(need_trace)) != SO_OK) return rc;
    size_t b_tasks = (size_t)n * sizeof(AlnTask), b_wb = wbase.size() * sizeof(uint64_t);
    if ((rc = c->scratch[0].reserve(b_tasks)) != SO_OK) return rc;
    if ((rc = c->scratch[1].reserve(b_wb)) != SO_OK) return rc;
    if ((rc = c->scratch[2].reserve((size_t)n * sizeof(DpOut))) != SO_OK) return rc;
    if ((rc = c->scratch[3].reserve((size_t)n * sizeof(TbOut))) != SO_OK) return rc;
    AlnTask *d_tasks = (AlnTask *)c->scratch[0].p;
    uint64_t *d_wb = (uint64_t *)c->scratch[1].p;
    DpOut *d_dp = (DpOut *)c->scratch[2].p;
    TbOut *d_tb = (TbOut *)c->scratch[3].p;
    SO_CUDA(cudaMemcpyAsync(d_tasks, sorted.data(), b_tasks, cudaMemcpyHostToDevice, c->stream_aln));
    SO_CUDA(cudaMemcpyAsync(d_wb, wbase.data(), b_wb, cudaMemcpyHostToDevice, c->stream_aln));
    const int grid = (int)((n + 127) / 128);
    SO_CUDA(cudaEventRecord(c->ev_aln[0], c->stream_aln));
    if (wave)
        k_banded_dp_wave<<<(int)((n * 16 + 127) / 128), 128, 0, c->stream_aln>>>(d_tasks, (int)n, d_wb, c->trace.p, d_dp, 0x4440u);
    else
        k_banded_dp<<<grid, 128, 0, c->stream_aln>>>(d_tasks, (int)n, d_wb, c->trace.p, d_dp, 0x4440u);
    SO_CUDA(cudaEventRecord(c->ev_aln[1], c->stream_aln));
    if (wave)
        k_traceback<true><<<grid, 128, 0, c->stream_aln>>>(d_tasks, (int)n, d_wb, c->trace.p, d_dp, d_tb);
    else
        k_traceback<false><<<grid, 128, 0, c->stream_aln>>>(d_tasks, (int)n, d_wb, c->trace.p, d_dp, d_tb);
    SO_CUDA(cudaEventRecord(c->ev_aln[2], c->stream_aln));
    SO_CUDA(cudaGetLastError());
    std::vector<DpOut> h_dp((size_t)n);
    std::vector<TbOut> h_tb((size_t)n);
    SO_CUDA(cudaMemcpyAsync(h_dp.data(), d_dp, (size_t)n * sizeof(DpOut), cudaMemcpyDeviceToHost, c->stream_aln));
    SO_CUDA(cudaMemcpyAsync(h_tb.data(), d_tb, (size_t)n * sizeof(TbOut), cudaMemcpyDeviceToHost, c->stream_aln));
    SO_CUDA(cudaStreamSynchronize(c->stream_aln));
    if (getenv("SO_DEBUG_TRACE") && n == 1) {
        int rows = std::min(sorted[0].len1, sorted[0].len0 + 16);
        std::vector<uint64_t> tr((size_t)rows * 32);
        cudaMemcpy(tr.data(), c->trace.p, tr.size() * 8, cudaMemcpyDeviceToHost);
        fprintf(stderr, "DBG len0 %d len1 %d score %d imax %d jmax %d rows %d | tb i0 %d j0 %d al %d\n", sorted[0].len0,
                sorted[0].len1, h_dp[0].score, h_dp[0].imax, h_dp[0].jmax, h_dp[0].rows, h_tb[0].i0, h_tb[0].j0, h_tb[0].al);
        for (int i = 1; i <= rows && i <= 12; i++) fprintf(stderr, "DBG row %d %016llx\n", i, (unsigned long long)tr[(size_t)(i - 1) * 32]);
    }
    float ms_dp = 0, ms_tb = 0;
    cudaEventElapsedTime(&ms_dp, c->ev_aln[0], c->ev_aln[1]);
    cudaEventElapsedTime(&ms_tb, c->ev_aln[1], c->ev_aln[2]);
    c->stats_aln.ms_align += ms_dp + ms_tb;
    c->stats_aln.ms_dp += ms_dp;
    c->stats_aln.ms_traceback += ms_tb;
    c->stats_aln.kernel_launches += 2;
    c->stats_aln.alignments += n;
    c->stats_aln.h2d_bytes += (i64)(b_tasks + b_wb);
    c->stats_aln.d2h_bytes += (i64)((size_t)n * (sizeof(DpOut) + sizeof(TbOut)));
    for (i64 s = 0; s < n; s++) {
        const int k = order[(size_t)s];
        const Prep &r = prep[(size_t)k];
        const DpOut &d = h_dp[(size_t)s];
        const TbOut &b = h_tb[(size_t)s];
        if (b.bad) {
            set_error("traceback left the band (internal error) for pair %d", k);
            return SO_EINVAL;
        }
        so_aln &o = out[k];
        o.raw_score = d.score;
        o.aln_len = b.al;
        o.n_ident = b.nid;
        o.mismatch = b.mis;
        o.gaps = b.gap;
        // (i, j) = rows / columns.  fsearch.py:1473-1476
        if (r.swap) {  // rows are the query
            o.qst = b.i0 + r.qst, o.qed = d.imax + r.qst, o.sst = b.j0 + r.sst, o.sed = d.jmax + r.sst;
        } else {
            o.qst = b.j0 + r.qst, o.qed = d.jmax + r.qst, o.sst = b.i0 + r.sst, o.sed = d.imax + r.sst;
        }
        const i64 cl = band_cells(r.len0, r.len1);
        o.cells = (int32_t)cl;
        cells += cl;
    }
    c->stats_aln.dp_cells += cells;
    c->stats_aln.ms_host += tm.ms() - (ms_dp + ms_tb);
    return SO_OK;
}

}  // namespace so

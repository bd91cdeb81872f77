/*
 * swiftortho_b200 — C ABI of the B200-native all-vs-all homology search.
 *
 * The reference (Rinoahu/SwiftOrtho) has no FFI on this path: bin/find_hit.py shells out to the
 * RPython-translated binary lib/fsearch-c (bin/find_hit.py:119-129) and the two exchange files.
 * This library is what a maintainer binds instead of that process boundary (ctypes stub in
 * INTEGRATION.md).  Every entry point names the reference function(s) it replaces
 * (file:line relative to the reference root).  Plain C types only; all buffers are caller-owned
 * host memory unless stated; every function returns 0 on success or a negative SO_E* code and
 * leaves a message retrievable with so_last_error().  There is no CPU fallback: entry points
 * that compute on the device fail with SO_ENODEV when no CUDA device is present.
 */
#ifndef SWIFTORTHO_B200_H
#define SWIFTORTHO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_ABI_VERSION 1

enum {
    SO_OK = 0,
    SO_EINVAL = -1,  /* bad argument */
    SO_ENODEV = -2,  /* no CUDA device / CUDA error */
    SO_ENOMEM = -3,  /* host or device allocation failed */
    SO_EIO = -4,     /* file could not be read / written */
    SO_ELIMIT = -5   /* input exceeds a documented limit (sequence >= 65536 residues, ...) */
};

int so_abi_version(void);
const char *so_last_error(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int so_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * H0  FASTA container — replaces lib/fsearch.py:1543-1553 `index` and 2180-2202 `Fasta`.
 * Record starts are offset 0 and every '>' preceded by '\n'; header = first line without '>',
 * sequence = remaining lines joined (no strip, no upper-casing, '\r' survives).
 * ------------------------------------------------------------------------------------------- */
typedef struct so_fasta so_fasta;
int so_fasta_open(const char *path, so_fasta **out);
void so_fasta_close(so_fasta *f);
int64_t so_fasta_count(const so_fasta *f);
/* total residues, and pointers to the packed residue buffer / offsets[count+1] owned by `f` */
int64_t so_fasta_residues(const so_fasta *f, const uint8_t **residues, const uint64_t **offsets);
/* header of record i (not NUL terminated) */
int so_fasta_header(const so_fasta *f, int64_t i, const char **hd, int64_t *len);

/* ---------------------------------------------------------------------------------------------
 * H1  low-complexity mask of a query — replaces lib/fsearch.py:2872-2946 `seg` (+ `entropy`
 * 2854-2868, `Counter` 157-177).  out[n] receives the upper-cased sequence with masked windows
 * replaced by lower-case 'x'.
 * ------------------------------------------------------------------------------------------- */
int so_seg(const uint8_t *seq, int64_t n, uint8_t *out);

/* ---------------------------------------------------------------------------------------------
 * Q   the reference's deterministic unstable quicksort — lib/fsearch.py:260-327 (`qsort`,
 * `quicksort`, `partition`, `insort`).  perm[n] receives the original indices in final order
 * (ascending key).
 * ------------------------------------------------------------------------------------------- */
int so_qsort_perm(const int64_t *keys, int64_t n, int32_t *perm);

/* The same sort on the device (test hook of the H3 selection kernel, csrc/select.cu): perm[min(need, n)]
 * receives the original indices at positions [0, need) of the reference quicksort of keys[n] (ascending,
 * 32-bit keys).  Declared after so_ctx below. */

/* ---------------------------------------------------------------------------------------------
 * F   score2bit / e-value text — lib/fsearch.py:1066-1071 `score2bit`, 1086 `bit2e`,
 * 43-61 `f2s`.  so_f2s writes a NUL-terminated string into out[cap].
 * ------------------------------------------------------------------------------------------- */
int64_t so_score2bit(int64_t raw_score);
double so_bit2e(int64_t n_targets, int64_t qlen, int64_t tlen, int64_t bit);
int so_f2s(double e, char *out, int cap);

/* ---------------------------------------------------------------------------------------------
 * Search context: one per process / GPU.  Parameters carry the same meaning as the fsearch-c
 * flags (lib/fsearch.py:3187-3188, 3215-3216).
 * ------------------------------------------------------------------------------------------- */
typedef struct so_params {
    const char *seeds;     /* -s  comma separated spaced-seed patterns, e.g. "111111"          */
    const char *alphabets; /* -r  reduced alphabet(s), '/'-separated, e.g. "AST,CFILMVY,..."    */
    uint32_t n_buckets;    /* -M  hash table size NC (bucket = fnv32 % NC)                      */
    int32_t step;          /* -j  distance between target seed starts                           */
    double expect;         /* -e                                                                */
    int64_t max_hits;      /* -v                                                                */
    double max_miss;       /* -m                                                                */
    int64_t threshold;     /* -t  (<1: use mu+2sd of the chunk)                                 */
    int32_t filter_query;  /* -F  1 = seg-mask queries                                          */
    int64_t chunk;         /* -c  target sequences per index chunk                              */
    int64_t ref_start;     /* -L  (-1 = 0)                                                      */
    int64_t ref_end;       /* -U  (-1 = all)                                                    */
} so_params;

typedef struct so_ctx so_ctx;
int so_ctx_create(int device, const so_params *p, so_ctx **out);
void so_ctx_destroy(so_ctx *c);

/* Load the packed target / query sets (H0 layout: residues + offsets[n+1]) into HBM.
 * Queries are seg-masked on the host first when filter_query is set (fsearch.py:2995-2998).
 * `n_db` is D = len(DB), the e-value database size (fsearch.py:2979). */
int so_set_targets(so_ctx *c, const uint8_t *residues, const uint64_t *offsets, int64_t n);
int so_set_queries(so_ctx *c, const uint8_t *residues, const uint64_t *offsets, int64_t n);
/* The same in two halves, for callers that stream query blocks (the reference recomputes seg and the position order per
 * query inside its loop, fsearch.py:2995-2998, 2647-2668): so_queries_prepare does the host work (seg masks, S3 position
 * order) without touching the device or the context's query state -- it may run on another thread while so_search works
 * on the previous block -- and so_set_queries_prepared makes the prepared set the context's query set (H2D). */
typedef struct so_qprep so_qprep;
int so_queries_prepare(const so_ctx *c, const uint8_t *residues, const uint64_t *offsets, int64_t n, so_qprep **out);
int so_set_queries_prepared(so_ctx *c, so_qprep *p);
void so_qprep_free(so_qprep *p);

/* K1-K3  build the index of every target chunk — replaces Fasta.makedb / build_msav
 * (lib/fsearch.py:2283-2295, 2208-2280: generate_nr_tbl 406-422, spseeds_fnv 519-556,
 * get_mu_sd 746-761).  All chunk indexes stay resident in HBM. */
typedef struct so_index_info {
    int64_t chunk_start, chunk_end; /* target ordinals [start, end)                    */
    int64_t n_seeds;                /* len(locus)                                      */
    int64_t n_buckets_used;         /* non-empty buckets                               */
    int64_t threshold;              /* int(mu + 2 sd) or -t                            */
    double build_ms;                /* device time of the build                        */
} so_index_info;
int so_index_build(so_ctx *c);
int64_t so_index_chunks(const so_ctx *c);
int so_index_info_get(const so_ctx *c, int64_t chunk, so_index_info *info);
/* test hook: copy start[NC+1] / locus[n_seeds] of a chunk to the host (either may be NULL) */
int so_index_export(so_ctx *c, int64_t chunk, uint32_t *start, uint32_t *locus);

/* S3+K4-K6  candidates of queries [q_begin, q_end) against one chunk — replaces
 * Fasta.find_msav_m(sort=False) (lib/fsearch.py:2645-2724: get_bin_mem 2530-2541, get_locs_m
 * 2638-2642, bisect 134-153, ungap 2454-2494, get_ungap_scores 2497-2509, guess_start
 * 2544-2553).  cand_offsets[q_end-q_begin+1] and the candidate array (reference order) are
 * malloc'ed by the library; free with so_free. */
typedef struct so_cand {
    uint32_t target; /* global target ordinal hd              */
    uint32_t score;  /* summed X-drop segment score (>= 25)   */
    uint32_t qi, qj; /* head of the diagonal (guess_start)    */
} so_cand;
int so_candidates(so_ctx *c, int64_t chunk, int64_t q_begin, int64_t q_end, uint64_t **cand_offsets,
                  so_cand **cands);
void so_free(void *p);

/* K7-K9  banded gapped alignment + traceback of explicit pairs — replaces kswat_st
 * (lib/fsearch.py:1357-1476) and, for sequences >= 4096, the tiles of kswat_st_long
 * (lib/fsearch.py:1480-1498; one so_pair per tile).  Sequences are addressed inside the loaded
 * query / target sets; (q_off,q_len) / (t_off,t_len) select the slice S0 / S1 handed to kswat_st
 * and (qst, sst) its start arguments. */
typedef struct so_pair {
    int64_t query, target;
    int32_t q_off, q_len, t_off, t_len;
    int32_t qst, sst;
} so_pair;
typedef struct so_aln {
    int32_t raw_score; /* maxscore                                            */
    int32_t aln_len;   /* AL                                                  */
    int32_t n_ident;   /* identical columns (idy = n_ident * (100. / AL))     */
    int32_t mismatch, gaps;
    int32_t qst, qed, sst, sed; /* as returned by kswat_st (0-based start, 1-based end), slice-relative */
    int32_t cells;              /* DP cells filled                            */
} so_aln;
int so_align_batch(so_ctx *c, const so_pair *pairs, int64_t n, so_aln *out);

/* H2-H4 + F  the whole search of queries [q_begin, q_end) against every chunk — replaces the body
 * of blastp (lib/fsearch.py:2968-3121) including candidate merge / sort / stop rule
 * (3039-3106), e-value filter (3071-3072) and per-query final order (3108-3110).  Rows come
 * back in output order as fixed-width records (malloc'ed; so_free). */
typedef struct so_hit {
    int64_t query, target;
    int32_t qlen, tlen;
    int32_t aln_len, mismatch, gaps;
    int32_t qst, qed, sst, sed; /* 1-based inclusive, as printed */
    int32_t raw_score;
    int32_t n_ident;
    int32_t pad;
    int64_t bit;
    double identity;
    double evalue;
} so_hit;
int so_search(so_ctx *c, int64_t q_begin, int64_t q_end, so_hit **rows, int64_t *n_rows);
/* Q on the device — lib/fsearch.py:260-327 as run by the candidate selection of so_search (3051, 3059-3062) */
int so_qsort_prefix_device(so_ctx *c, const uint32_t *keys, int64_t n, int64_t need, uint32_t *perm);

/* F  format rows as the reference's 16-column text (lib/fsearch.py:3233-3243).  Headers come from
 * the two FASTA containers.  Appends to `path` when append != 0. */
int so_write_rows(const so_hit *rows, int64_t n, const so_fasta *queries, const so_fasta *targets, const char *path,
                  int append);

/* R   redundancy pre-filter — the device half of scripts/nr_flt.py:8-27 (exact-duplicate collapse
 * before the search, scripts/run_all_fast.py:110-119): hashes[n] receives a 64-bit content hash of every sequence
 * of a packed set (H0 layout).  The host groups by hash and confirms equality byte for byte (swiftortho_b200/nr.py). */
int so_seq_hash(int device, const uint8_t *residues, const uint64_t *offsets, int64_t n, uint64_t *hashes);

/* O   orthology inference, device half — bin/find_orth.py:158-234 (`blastparse`: best score per
 * target inside a query group) and :298-348 (`get_qIPO`: per-taxon maxima, in-paralog / ortholog / co-ortholog call).
 * Rows are the filtered hit table: group_offsets[n_groups + 1] delimits the runs of equal query id; ids are ranks in
 * byte order of the id strings, taxa small integers, score the (normalised) double.  cls[row] = 0 not a candidate
 * (or not the best row of its target), 1 IP, 2 OT, 3 CO. */
int so_orth_classify(int device, const uint64_t *group_offsets, int64_t n_groups, const uint32_t *qrank,
                     const uint32_t *srank, const uint32_t *qtax, const uint32_t *stax, const double *score,
                     uint32_t n_taxa, uint8_t *cls);
/* device radix sort of 64-bit keys with 32-bit payloads, in place (the reference's `sort` calls,
 * bin/find_orth.py:476-478, 499-501, 552-554, on (id rank << 32 | id rank) keys) */
int so_sort_pairs_u64(int device, uint64_t *keys, uint32_t *vals, int64_t n);

/* C   clustering, device half — bin/find_cluster.py.  Host arrays in, host arrays out.
 * so_cc_labels: connected components of an undirected graph with n vertices and m edges (the reference's
 *   nx.connected_components calls, bin/find_cluster.py:1508-1520, 1550-1556, 1690-1692, 1708-1718);
 *   labels[v] = smallest vertex id of v's component.
 * so_apc: affinity propagation as `apclust_blk` / `apclust` run it (bin/find_cluster.py:310-516) on the float32 table
 *   fc2mat writes (:767-860): entries (row[e], col[e], sim[e]) in FILE ORDER, ks items, `sweeps` sweeps (the reference
 *   always runs itr = 100), labels[i] = exemplar of item i (column of the first maximum of R + A in row i).
 * so_mcl: Markov clustering as `mcl` runs it (bin/find_cluster.py:636-690) on a CSR float32 matrix with sorted,
 *   duplicate-free rows: at most max_iter (100) iterations of column normalisation, X @ X, x ** inflation, convergence
 *   test every check_every-th (5th) iteration, pruning below 1e-5.  Returns the final matrix (library-owned, so_free);
 *   entries above 1e-5 are the edges whose connected components are the clusters.  iterations = iterations run. */
int so_cc_labels(int device, int64_t n, int64_t m, const uint32_t *eu, const uint32_t *ev, uint32_t *labels);
int so_apc(int device, int64_t n_rows, int64_t ks, const uint32_t *row, const uint32_t *col, const float *sim, double damp,
           int sweeps, int32_t *labels);
int so_mcl(int device, int64_t n, const int64_t *indptr, const uint32_t *indices, const float *data, double inflation,
           int max_iter, int check_every, int64_t **out_indptr, uint32_t **out_indices, float **out_data, int *iterations);

/* counters of the last so_search / so_align_batch call (for bench.py) */
typedef struct so_stats {
    int64_t queries, seed_hits, groups, candidates, alignments, dp_cells, rows;
    int64_t ungap_steps;     /* residue pairs scored by the X-drop extensions (K6)   */
    int64_t kernel_launches; /* launches of this library's own kernels */
    int64_t lib_launches;    /* CUB (library) launches                 */
    double ms_seed, ms_sort, ms_ungap, ms_select, ms_align, ms_dp, ms_traceback, ms_host, ms_total;
    int64_t h2d_bytes, d2h_bytes;
    double ms_ungap_kernel;  /* X-drop kernels alone (k_single_ungap + k_group_ungap), inside ms_ungap */
    int64_t multi_groups;    /* diagonal groups holding more than one seed (chained path)            */
    int64_t redo_blocks;     /* query blocks the sync-free path handed back to the general path      */
    int64_t alignments_used; /* alignments the sequential stop rule consumed (alignments - this = wasted) */
} so_stats;
int so_stats_get(const so_ctx *c, so_stats *s);
/* tuning hooks (no reference counterpart): queries per seeding sub-block (0 = adaptive), and the number
 * of candidate-production lanes (streams) so_search overlaps (1 .. 4, default 1; 0 = one lane and alignment rounds
 * serialised with candidate production: measurement mode for per-kernel CUDA-event times) */
int so_set_sub_block(so_ctx *c, int64_t n);
int so_set_lanes(so_ctx *c, int n);
int so_stats_reset(so_ctx *c);

#ifdef __cplusplus
}
#endif
#endif

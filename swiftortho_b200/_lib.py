"""ctypes binding of libswiftortho_b200.so (include/swiftortho_b200.h).

The library is the product: if it is missing or a device call fails this module raises — there
is no Python / CPU fallback for any stage of the search.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libswiftortho_b200.so')


class SoError(RuntimeError):
    pass


class so_params(C.Structure):
    _fields_ = [('seeds', C.c_char_p), ('alphabets', C.c_char_p), ('n_buckets', C.c_uint32), ('step', C.c_int32),
                ('expect', C.c_double), ('max_hits', C.c_int64), ('max_miss', C.c_double), ('threshold', C.c_int64),
                ('filter_query', C.c_int32), ('chunk', C.c_int64), ('ref_start', C.c_int64), ('ref_end', C.c_int64)]


class so_index_info(C.Structure):
    _fields_ = [('chunk_start', C.c_int64), ('chunk_end', C.c_int64), ('n_seeds', C.c_int64),
                ('n_buckets_used', C.c_int64), ('threshold', C.c_int64), ('build_ms', C.c_double)]


class so_cand(C.Structure):
    _fields_ = [('target', C.c_uint32), ('score', C.c_uint32), ('qi', C.c_uint32), ('qj', C.c_uint32)]


class so_pair(C.Structure):
    _fields_ = [('query', C.c_int64), ('target', C.c_int64), ('q_off', C.c_int32), ('q_len', C.c_int32),
                ('t_off', C.c_int32), ('t_len', C.c_int32), ('qst', C.c_int32), ('sst', C.c_int32)]


class so_aln(C.Structure):
    _fields_ = [('raw_score', C.c_int32), ('aln_len', C.c_int32), ('n_ident', C.c_int32), ('mismatch', C.c_int32),
                ('gaps', C.c_int32), ('qst', C.c_int32), ('qed', C.c_int32), ('sst', C.c_int32), ('sed', C.c_int32),
                ('cells', C.c_int32)]


class so_hit(C.Structure):
    _fields_ = [('query', C.c_int64), ('target', C.c_int64), ('qlen', C.c_int32), ('tlen', C.c_int32),
                ('aln_len', C.c_int32), ('mismatch', C.c_int32), ('gaps', C.c_int32), ('qst', C.c_int32),
                ('qed', C.c_int32), ('sst', C.c_int32), ('sed', C.c_int32), ('raw_score', C.c_int32),
                ('n_ident', C.c_int32), ('pad', C.c_int32), ('bit', C.c_int64), ('identity', C.c_double),
                ('evalue', C.c_double)]


class so_stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ('queries', 'seed_hits', 'groups', 'candidates', 'alignments', 'dp_cells',
                                         'rows', 'ungap_steps', 'kernel_launches', 'lib_launches')] + \
               [(n, C.c_double) for n in ('ms_seed', 'ms_sort', 'ms_ungap', 'ms_select', 'ms_align', 'ms_dp',
                                          'ms_traceback', 'ms_host', 'ms_total')] + \
               [('h2d_bytes', C.c_int64), ('d2h_bytes', C.c_int64), ('ms_ungap_kernel', C.c_double),
                ('multi_groups', C.c_int64), ('redo_blocks', C.c_int64), ('alignments_used', C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/swiftortho_b200.h declares: (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    'so_abi_version': (C.c_int, []),
    'so_last_error': (C.c_char_p, []),
    'so_device_count': (C.c_int, []),
    'so_fasta_open': (C.c_int, [C.c_char_p, _P(C.c_void_p)]),
    'so_fasta_close': (None, [C.c_void_p]),
    'so_fasta_count': (C.c_int64, [C.c_void_p]),
    'so_fasta_residues': (C.c_int64, [C.c_void_p, _P(C.c_void_p), _P(C.c_void_p)]),
    'so_fasta_header': (C.c_int, [C.c_void_p, C.c_int64, _P(C.c_char_p), _P(C.c_int64)]),
    'so_seg': (C.c_int, [C.c_char_p, C.c_int64, C.c_char_p]),
    'so_qsort_perm': (C.c_int, [_P(C.c_int64), C.c_int64, _P(C.c_int32)]),
    'so_qsort_prefix_device': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    'so_score2bit': (C.c_int64, [C.c_int64]),
    'so_bit2e': (C.c_double, [C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    'so_f2s': (C.c_int, [C.c_double, C.c_char_p, C.c_int]),
    'so_ctx_create': (C.c_int, [C.c_int, _P(so_params), _P(C.c_void_p)]),
    'so_ctx_destroy': (None, [C.c_void_p]),
    'so_set_targets': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    'so_set_queries': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    'so_queries_prepare': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _P(C.c_void_p)]),
    'so_set_queries_prepared': (C.c_int, [C.c_void_p, C.c_void_p]),
    'so_qprep_free': (None, [C.c_void_p]),
    'so_index_build': (C.c_int, [C.c_void_p]),
    'so_index_chunks': (C.c_int64, [C.c_void_p]),
    'so_index_info_get': (C.c_int, [C.c_void_p, C.c_int64, _P(so_index_info)]),
    'so_index_export': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'so_candidates': (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, _P(_P(C.c_uint64)), _P(_P(so_cand))]),
    'so_free': (None, [C.c_void_p]),
    'so_align_batch': (C.c_int, [C.c_void_p, _P(so_pair), C.c_int64, _P(so_aln)]),
    'so_search': (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _P(_P(so_hit)), _P(C.c_int64)]),
    'so_write_rows': (C.c_int, [_P(so_hit), C.c_int64, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]),
    'so_seq_hash': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'so_orth_classify': (C.c_int, [C.c_int, C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_uint32, C.c_void_p]),
    'so_sort_pairs_u64': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    'so_cc_labels': (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'so_apc': (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p]),
    'so_mcl': (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int,
                         _P(C.c_void_p), _P(C.c_void_p), _P(C.c_void_p), _P(C.c_int)]),
    'so_stats_get': (C.c_int, [C.c_void_p, _P(so_stats)]),
    'so_stats_reset': (C.c_int, [C.c_void_p]),
    'so_set_sub_block': (C.c_int, [C.c_void_p, C.c_int64]),
    'so_set_lanes': (C.c_int, [C.c_void_p, C.c_int]),
}

_lib = None


def load():
    """Load the native library (never builds it, never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SoError('%s is missing: run `python -m swiftortho_b200.build` (nvcc, sm_100a). '
                      'There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.so_abi_version() != 1:
        raise SoError('ABI mismatch')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SoError('swiftortho_b200 error %d: %s' % (rc, load().so_last_error().decode('utf-8', 'replace')))

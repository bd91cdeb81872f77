"""Python host of the B200 search path — mirrors the reference's operator surface.

`Fasta` and `blastp(...)` keep the names, argument meaning and defaults of
lib/fsearch.py:2180 (`Fasta`) and lib/fsearch.py:2968 (`blastp`), but every stage runs in
libswiftortho_b200.so (CUDA, sm_100a) through ctypes.  No torch, no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, so_aln, so_cand, so_hit, so_index_info, so_pair, so_params, so_stats

AA9 = 'AST,CFILMVY,DN,EQ,G,H,KR,P,W'                       # bin/find_hit.py:219
AA20 = 'A,S,T,C,F,I,L,M,V,Y,D,N,E,Q,G,H,K,R,P,W'           # bin/find_hit.py:220


class Fasta:
    """FASTA container with the reference's record semantics (lib/fsearch.py:2180-2202)."""

    def __init__(self, path):
        self.lib = _lib.load()
        self.path = path
        h = C.c_void_p()
        check(self.lib.so_fasta_open(str(path).encode(), C.byref(h)))
        self.h = h
        self.N = int(self.lib.so_fasta_count(h))
        res, off = C.c_void_p(), C.c_void_p()
        self.n_residues = int(self.lib.so_fasta_residues(h, C.byref(res), C.byref(off)))
        self._res, self._off = res, off
        self.offsets = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint64)), shape=(self.N + 1,))

    def __len__(self):
        return self.N

    def header(self, i):
        hd, n = C.c_char_p(), C.c_int64()
        check(self.lib.so_fasta_header(self.h, i, C.byref(hd), C.byref(n)))
        return C.string_at(hd, n.value).decode('latin-1')

    def sequence(self, i):
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return C.string_at(self._res.value + a, b - a).decode('latin-1')

    def __getitem__(self, i):
        if i < 0:
            i += self.N
        if not 0 <= i < self.N:
            return ['', '']
        return [self.header(i), self.sequence(i)]

    def close(self):
        if self.h:
            self.lib.so_fasta_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Searcher:
    """One search context = one process = one GPU."""

    def __init__(self, device=0, ssd='111111', nr=AA9, ht=120000000, step=1, expect=1e-3, v=500, max_miss=1e-3,
                 thr=-1, flt='T', chk=50000, rst=-1, red=-1):
        self.lib = _lib.load()
        if nr.strip() == 'aa9':
            nr = AA9
        elif nr.strip() == 'aa20':
            nr = AA20
        self._keep = (ssd.encode(), nr.encode())
        p = so_params(self._keep[0], self._keep[1], int(ht), int(step), float(expect), int(v), float(max_miss),
                      int(thr), 1 if str(flt).upper() == 'T' else 0, int(chk), int(rst), int(red))
        h = C.c_void_p()
        check(self.lib.so_ctx_create(int(device), C.byref(p), C.byref(h)))
        self.h = h
        self.queries = self.targets = None

    def set_targets(self, fasta):
        self.targets = fasta
        check(self.lib.so_set_targets(self.h, fasta._res, fasta._off, fasta.N))

    def set_queries(self, fasta):
        self.queries = fasta
        check(self.lib.so_set_queries(self.h, fasta._res, fasta._off, fasta.N))

    def build_index(self):
        check(self.lib.so_index_build(self.h))
        out = []
        for k in range(int(self.lib.so_index_chunks(self.h))):
            info = so_index_info()
            check(self.lib.so_index_info_get(self.h, k, C.byref(info)))
            out.append({n: getattr(info, n) for n, _ in info._fields_})
        return out

    def index_export(self, chunk, n_buckets, n_seeds):
        start = np.zeros(n_buckets + 1, dtype=np.uint32)
        locus = np.zeros(max(n_seeds, 1), dtype=np.uint32)
        check(self.lib.so_index_export(self.h, chunk, start.ctypes.data, locus.ctypes.data))
        return start, locus[:n_seeds]

    def candidates(self, chunk, q_begin, q_end):
        """[(target, score, qi, qj), ...] per query, reference order (find_msav_m, sort=False)."""
        off, cd = C.POINTER(C.c_uint64)(), C.POINTER(so_cand)()
        check(self.lib.so_candidates(self.h, chunk, q_begin, q_end, C.byref(off), C.byref(cd)))
        n = q_end - q_begin
        o = [int(off[i]) for i in range(n + 1)]
        arr = np.ctypeslib.as_array(C.cast(cd, C.POINTER(C.c_uint32)), shape=(max(o[-1], 1), 4)).copy()
        self.lib.so_free(off)
        self.lib.so_free(cd)
        return [arr[o[i]:o[i + 1]] for i in range(n)]

    def align(self, pairs):
        """pairs: list of (query, target, q_off, q_len, t_off, t_len, qst, sst) -> list of dict."""
        n = len(pairs)
        P = (so_pair * max(n, 1))()
        for i, t in enumerate(pairs):
            P[i] = so_pair(*[int(x) for x in t])
        A = (so_aln * max(n, 1))()
        check(self.lib.so_align_batch(self.h, P, n, A))
        return [{k: getattr(A[i], k) for k, _ in so_aln._fields_} for i in range(n)]

    def search(self, q_begin, q_end):
        """Rows (numpy structured array of so_hit) of queries [q_begin, q_end) in output order."""
        rows, n = C.POINTER(so_hit)(), C.c_int64()
        check(self.lib.so_search(self.h, q_begin, q_end, C.byref(rows), C.byref(n)))
        return _Rows(self.lib, rows, n.value)

    def write(self, rows, path, append=False):
        check(self.lib.so_write_rows(rows.ptr, rows.n, self.queries.h, self.targets.h, str(path).encode(),
                                     1 if append else 0))

    def search_to_file(self, q_begin, q_end, path, block=8192, query_base=0, fasta=None):
        """Search queries [q_begin, q_end) block by block and append the rows to `path`; a writer thread formats and
        writes block i while block i + 1 is searched (the library calls release the GIL).  `query_base` is added to
        the query ordinals (the loaded query set may be a slice of `fasta`)."""
        import queue
        import threading
        fasta = fasta or self.queries
        wq, err, nrows = queue.Queue(maxsize=2), [], [0]

        def writer():
            while True:
                rows = wq.get()
                if rows is None:
                    return
                try:
                    if not err:
                        check(self.lib.so_write_rows(rows.ptr, rows.n, fasta.h, self.targets.h, str(path).encode(), 1))
                        nrows[0] += rows.n
                except BaseException as e:  # noqa: BLE001
                    err.append(e)

        th = threading.Thread(target=writer)
        th.start()
        try:
            for b in range(q_begin, q_end, block):
                rows = self.search(b, min(q_end, b + block))
                if query_base:
                    rows.view()['query'] += query_base
                wq.put(rows)
                if err:
                    break
        finally:
            wq.put(None)
            th.join()
        if err:
            raise err[0]
        return nrows[0]

    def search_stream(self, fasta, ranges, path, append=True):
        """Stream the query blocks `ranges` = [(a, b), ...] of `fasta` (a host FASTA container) through the search and
        append their rows to `path`.  Three threads: the host preparation of block i + 1 (seg masks, S3 position order:
        so_queries_prepare) and the text formatting of block i - 1 (so_write_rows) overlap the device work of block i
        (so_set_queries_prepared: H2D, so_search: kernels + D2H of the rows); the library calls release the GIL.
        Returns the number of rows written."""
        import queue
        import threading
        lib = self.lib
        pq, wq, err, nrows = queue.Queue(maxsize=2), queue.Queue(maxsize=2), [], [0]
        if not append:
            open(path, 'wb').close()

        def preparer():
            try:
                for a, b in ranges:
                    if err:
                        break
                    off = np.ascontiguousarray(fasta.offsets[a:b + 1])
                    h = C.c_void_p()
                    check(lib.so_queries_prepare(self.h, fasta._res, off.ctypes.data, b - a, C.byref(h)))
                    pq.put((a, b, h))
            except BaseException as e:  # noqa: BLE001
                err.append(e)
            finally:
                pq.put(None)

        def writer():
            while True:
                rows = wq.get()
                if rows is None:
                    return
                try:
                    if not err:
                        check(lib.so_write_rows(rows.ptr, rows.n, fasta.h, self.targets.h, str(path).encode(), 1))
                        nrows[0] += rows.n
                except BaseException as e:  # noqa: BLE001
                    err.append(e)

        tp, tw = threading.Thread(target=preparer), threading.Thread(target=writer)
        tp.start()
        tw.start()
        try:
            while True:
                item = pq.get()
                if item is None:
                    break
                a, b, h = item
                try:
                    if not err:
                        check(lib.so_set_queries_prepared(self.h, h))
                        rows = self.search(0, b - a)
                        rows.view()['query'] += a
                        wq.put(rows)
                except BaseException as e:  # noqa: BLE001
                    err.append(e)
                finally:
                    lib.so_qprep_free(h)
        finally:
            wq.put(None)
            tp.join()
            tw.join()
        self.queries = None  # the context now holds the last block only
        if err:
            raise err[0]
        return nrows[0]

    def stats(self, reset=False):
        s = so_stats()
        check(self.lib.so_stats_get(self.h, C.byref(s)))
        if reset:
            check(self.lib.so_stats_reset(self.h))
        return s.as_dict()

    def set_sub_block(self, n):
        check(self.lib.so_set_sub_block(self.h, int(n)))

    def set_lanes(self, n):
        check(self.lib.so_set_lanes(self.h, int(n)))

    def close(self):
        if self.h:
            self.lib.so_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Rows:
    def __init__(self, lib, ptr, n):
        self.lib, self.ptr, self.n = lib, ptr, n

    def view(self):
        """numpy structured view of the library-owned row records (no copy; valid while this object lives)."""
        dt = np.dtype([(k, np.dtype(t)) for k, t in so_hit._fields_])
        if self.n == 0:
            return np.zeros(0, dtype=dt)
        return np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)),
                                     shape=(self.n * C.sizeof(so_hit),)).view(dt)

    def as_array(self):
        return self.view().copy()

    def __del__(self):
        try:
            if self.ptr:
                self.lib.so_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def blastp(qry, ref, out, expect=1e-5, v=500, max_miss=1e-3, st=-1, ed=-1, rst=-1, red=-1, thr=-1, flt='T',
           ssd='111111', nr=AA9, step=4, ht=-1, chk=50000, wrt='w', device=0, block=16384):
    """Drop-in for the `fsearch-c -p blastp` run (lib/fsearch.py:2968 + 3231-3256): searches queries
    [st, ed) of `qry` against `ref` and writes the 16-column rows to `out`.  Returns the stats dict."""
    if ht < 2:
        raise _lib.SoError('-M (hash table size) must be given; the reference derives a degenerate size otherwise')
    Q = Fasta(qry)
    T = Q if str(ref) == str(qry) else Fasta(ref)
    N, D = len(Q), len(T)
    st = min(max(0, st), N)                      # lib/fsearch.py:2980-2981
    ed = min(D if ed < 0 else ed, N)
    S = Searcher(device=device, ssd=ssd, nr=nr, ht=ht, step=step, expect=expect, v=v, max_miss=max_miss, thr=thr,
                 flt=flt, chk=chk, rst=rst, red=red)
    S.set_targets(T)
    S.build_index()
    first = 'a' not in wrt
    if first:
        open(out, 'wb').close()
    # query blocks are prepared (seg, S3 order), searched and written as a three-stage pipeline; only [st, ed) is prepared
    S.search_stream(Q, [(a, min(ed, a + block)) for a in range(st, ed, block)], out)
    stats = S.stats()
    S.close()
    return stats

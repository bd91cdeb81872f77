import ctypes
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def build_oracle():
    """Compile oracle/fsearch_oracle.cpp (the checker) if it is not built yet."""
    so = os.path.join(ROOT, 'oracle', '_build', 'liboracle.so')
    exe = os.path.join(ROOT, 'oracle', '_build', 'fsearch_oracle')
    src = os.path.join(ROOT, 'oracle', 'fsearch_oracle.cpp')
    if (not os.path.exists(so) or not os.path.exists(exe)
            or os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')], stdout=subprocess.DEVNULL)
    return so, exe


class Oracle:
    """ctypes view of the CPU oracle (test infrastructure)."""

    def __init__(self):
        so, self.exe = build_oracle()
        L = ctypes.CDLL(so)
        self.L = L
        L.orc_kswat_st.restype = ctypes.c_double
        L.orc_kswat_st.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
        L.orc_ungap.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int] + [ctypes.c_int] * 4 + [
            ctypes.POINTER(ctypes.c_longlong)]
        L.orc_seg.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p]
        L.orc_score2bit.restype = ctypes.c_longlong
        L.orc_score2bit.argtypes = [ctypes.c_longlong]
        L.orc_f2s.argtypes = [ctypes.c_double, ctypes.c_char_p, ctypes.c_int]
        L.orc_qsort_perm.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.orc_spseeds.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p,
                                  ctypes.c_uint, ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_int)]
        L.orc_blastp.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_double,
                                 ctypes.c_longlong, ctypes.c_double] + [ctypes.c_longlong] * 5 + [
            ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_longlong] * 3 + [
            ctypes.c_char_p, ctypes.POINTER(ctypes.c_longlong)]

    def kswat_st(self, s0, s1, qst=0, sst=0):
        out = (ctypes.c_longlong * 9)()
        a, b = s0.encode('latin-1'), s1.encode('latin-1')
        idy = self.L.orc_kswat_st(a, len(a), b, len(b), qst, sst, out)
        return idy, list(out)

    def ungap(self, q, s, Q, S, qlo=-1, slo=-1):
        out = (ctypes.c_longlong * 5)()
        a, b = q.encode('latin-1'), s.encode('latin-1')
        self.L.orc_ungap(a, len(a), b, len(b), Q, S, qlo, slo, out)
        return list(out)

    def seg(self, s):
        a = s.encode('latin-1')
        out = ctypes.create_string_buffer(len(a) + 1)
        self.L.orc_seg(a, len(a), out)
        return out.raw[:len(a)].decode('latin-1')

    def f2s(self, e):
        out = ctypes.create_string_buffer(64)
        self.L.orc_f2s(e, out, 64)
        return out.value.decode()

    def qsort_perm(self, keys):
        n = len(keys)
        k = (ctypes.c_longlong * max(n, 1))(*keys)
        p = (ctypes.c_int * max(n, 1))()
        self.L.orc_qsort_perm(k, n, p)
        return list(p)[:n]

    def spseeds(self, seq, step, nr, ssd, mod):
        a = seq.encode('latin-1')
        cap = max(1, len(a) * (ssd.count(',') + 1) * (nr.count('/') + 1))
        b = (ctypes.c_uint * cap)()
        p = (ctypes.c_int * cap)()
        n = self.L.orc_spseeds(a, len(a), step, nr.encode(), ssd.encode(), mod, b, p)
        return [[b[i], p[i]] for i in range(n)]

    def candidates(self, qry, ref, c0, c1, q0, q1, flags, cap=4000000):
        import numpy as np
        f = {'-F': 'T', '-s': '111111', '-r': 'AST,CFILMVY,DN,EQ,G,H,KR,P,W', '-j': '1', '-M': '1000003', '-t': '-1'}
        f.update(flags)
        fn = self.L.orc_candidates
        fn.restype = ctypes.c_longlong
        fn.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_longlong] * 4 + [ctypes.c_char_p] * 3 + [
            ctypes.c_longlong] * 3 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
        off = np.zeros(q1 - q0 + 1, dtype=np.uint64)
        out = np.zeros((cap, 4), dtype=np.uint32)
        thr = ctypes.c_longlong()
        n = fn(qry.encode(), ref.encode(), c0, c1, q0, q1, f['-F'].encode(), f['-s'].encode(), f['-r'].encode(),
               int(f['-j']), int(f['-M']), int(f['-t']), off.ctypes.data, out.ctypes.data, cap, ctypes.byref(thr))
        assert 0 <= n <= cap
        return [out[int(off[i]):int(off[i + 1])] for i in range(q1 - q0)], thr.value

    def index(self, ref, c0, c1, flags, cap=50000000):
        import numpy as np
        f = {'-s': '111111', '-r': 'AST,CFILMVY,DN,EQ,G,H,KR,P,W', '-j': '1', '-M': '1000003'}
        f.update(flags)
        fn = self.L.orc_index
        fn.restype = ctypes.c_longlong
        fn.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_char_p,
                       ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong]
        nc = int(f['-M'])
        start = np.zeros(nc, dtype=np.uint32)
        locus = np.zeros(cap, dtype=np.uint32)
        n = fn(ref.encode(), c0, c1, f['-s'].encode(), f['-r'].encode(), int(f['-j']), nc, start.ctypes.data,
               locus.ctypes.data, cap)
        assert 0 <= n <= cap
        return start, locus[:n]

    def blastp(self, qry, ref, out, flags):
        """flags: dict of fsearch-c flag letters (same as the golden cases)."""
        stats = (ctypes.c_longlong * 7)()
        f = {'-e': '1e-3', '-v': '500', '-m': '1e-3', '-l': '-1', '-u': '-1', '-L': '-1', '-U': '-1', '-t': '-1',
             '-F': 'T', '-s': '111111', '-r': 'AST,CFILMVY,DN,EQ,G,H,KR,P,W', '-j': '4', '-M': '-1', '-c': '50000',
             '-O': 'wb'}
        f.update(flags)
        rc = self.L.orc_blastp(qry.encode(), ref.encode(), out.encode(), float(f['-e']), int(f['-v']),
                               float(f['-m']), int(f['-l']), int(f['-u']), int(f['-L']), int(f['-U']),
                               int(f['-t']), f['-F'].encode(), f['-s'].encode(), f['-r'].encode(), int(f['-j']),
                               int(f['-M']), int(f['-c']), f['-O'].encode(), stats)
        assert rc == 0
        return dict(zip(['queries', 'seed_hits', 'groups', 'candidates', 'alignments', 'dp_cells', 'rows'],
                        list(stats)))


@pytest.fixture(scope='session')
def oracle():
    return Oracle()


@pytest.fixture(scope='session')
def kat():
    with open(os.path.join(GOLDEN, 'kat.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def golden_cases():
    with open(os.path.join(GOLDEN, 'cases.json')) as f:
        return json.load(f)


class ClusterOracle:
    """ctypes view of oracle/cluster_oracle.cpp (test infrastructure) with the backend interface of
    swiftortho_b200.find_cluster.DeviceBackend, so the host logic can be checked on CPU against the reference goldens
    and the CUDA kernels against the same restatement."""

    def __init__(self):
        import numpy as np
        self.np = np
        so = os.path.join(ROOT, 'oracle', '_build', 'libcluster_oracle.so')
        src = os.path.join(ROOT, 'oracle', 'cluster_oracle.cpp')
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(so)
        self.L = L
        V, P = ctypes.c_void_p, ctypes.POINTER
        L.orc_cc_labels.argtypes = [ctypes.c_int64, ctypes.c_int64, V, V, V]
        L.orc_apc.argtypes = [ctypes.c_int64, ctypes.c_int64, V, V, V, ctypes.c_double, ctypes.c_int, V]
        L.orc_mcl.argtypes = [ctypes.c_int64, V, V, V, ctypes.c_double, ctypes.c_int, ctypes.c_int, P(V), P(V), P(V),
                              P(ctypes.c_int)]
        L.orc_cluster_free.argtypes = [V]

    def cc_labels(self, n, eu, ev):
        np = self.np
        eu, ev = np.ascontiguousarray(eu, dtype=np.uint32), np.ascontiguousarray(ev, dtype=np.uint32)
        lab = np.empty(n, dtype=np.uint32)
        self.L.orc_cc_labels(n, len(eu), eu.ctypes.data, ev.ctypes.data, lab.ctypes.data)
        return lab

    def apc(self, ks, row, col, sim, damp, sweeps=100):
        np = self.np
        row, col = np.ascontiguousarray(row, dtype=np.uint32), np.ascontiguousarray(col, dtype=np.uint32)
        sim = np.ascontiguousarray(sim, dtype=np.float32)
        lab = np.empty(ks, dtype=np.int32)
        self.L.orc_apc(len(row), ks, row.ctypes.data, col.ctypes.data, sim.ctypes.data, float(damp), int(sweeps), lab.ctypes.data)
        return lab

    def mcl(self, n, indptr, indices, data, inflation, max_iter=100, check_every=5):
        np = self.np
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        data = np.ascontiguousarray(data, dtype=np.float32)
        op, oi, od, it = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int(0)
        self.L.orc_mcl(n, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, float(inflation), int(max_iter),
                       int(check_every), ctypes.byref(op), ctypes.byref(oi), ctypes.byref(od), ctypes.byref(it))
        ptr = np.ctypeslib.as_array(ctypes.cast(op, ctypes.POINTER(ctypes.c_int64)), shape=(n + 1,)).copy()
        nnz = int(ptr[n]) if n else 0
        col = np.ctypeslib.as_array(ctypes.cast(oi, ctypes.POINTER(ctypes.c_uint32)), shape=(max(nnz, 1),))[:nnz].copy()
        val = np.ctypeslib.as_array(ctypes.cast(od, ctypes.POINTER(ctypes.c_float)), shape=(max(nnz, 1),))[:nnz].copy()
        for p in (op, oi, od):
            self.L.orc_cluster_free(p)
        return ptr, col, val, it.value


@pytest.fixture(scope='session')
def cluster_oracle():
    return ClusterOracle()

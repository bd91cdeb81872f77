"""Orthology inference row: swiftortho_b200.find_orth against the outputs of the reference's own
bin/find_orth.py (tests/golden/make_orth_golden.py: synth60*.orth, synth600*.orth; three score normalisations, two
filter settings).  The CPU test drives the host half (parsing, ranks, reciprocal joins, averages, formatting) with a
plain-Python statement of the two device entry points as the checker; the GPU test runs the real kernels."""
import ctypes
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = [('synth60.sc', 'synth60.orth', dict(coverage=.5, identity=0., norm='no')),
         ('synth60.sc', 'synth60_bsr.orth', dict(coverage=.5, identity=0., norm='bsr')),
         ('synth600.sc', 'synth600.orth', dict(coverage=.5, identity=0., norm='no')),
         ('synth600.sc', 'synth600_bal.orth', dict(coverage=.3, identity=25., norm='bal')),
         ('synth600.sc', 'synth600_bsr.orth', dict(coverage=.6, identity=0., norm='bsr'))]


class OracleOrthLib:
    """bin/find_orth.py:158-234 + 298-348 restated on the integer layout of so_orth_classify, and a stable argsort for
    so_sort_pairs_u64 (test infrastructure: the checker of the device kernels)."""

    def so_orth_classify(self, dev, goff, ng, qr, sr, qt, st, sc, ntaxa, cls):
        goff = np.ctypeslib.as_array(ctypes.cast(goff, ctypes.POINTER(ctypes.c_uint64)), shape=(ng + 1,))
        n = int(goff[-1])

        def arr(p, t):
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(t)), shape=(n,))
        qr, sr, qt, st = (arr(x, ctypes.c_uint32) for x in (qr, sr, qt, st))
        sc, cls = arr(sc, ctypes.c_double), arr(cls, ctypes.c_uint8)
        cls[:] = 0
        for g in range(ng):
            a, b = int(goff[g]), int(goff[g + 1])
            best = {}
            for i in range(a, b):                        # output[key][-1] < Score: strict, the earlier row wins ties
                k = int(sr[i])
                if k not in best or sc[best[k]] < sc[i]:
                    best[k] = i
            smax, omax = {}, 0.0
            for i in best.values():
                smax[int(st[i])] = max(smax.get(int(st[i]), 0.0), sc[i])
                if qt[i] != st[i]:
                    omax = max(omax, sc[i])
            for i in best.values():
                if qt[i] == st[i]:
                    cls[i] = 1 if (sc[i] >= omax and qr[i] != sr[i]) else 0
                else:
                    cls[i] = 2 if sc[i] >= smax[int(st[i])] else 3
        return 0

    def so_sort_pairs_u64(self, dev, keys, vals, n):
        k = np.ctypeslib.as_array(ctypes.cast(keys, ctypes.POINTER(ctypes.c_uint64)), shape=(n,))
        v = np.ctypeslib.as_array(ctypes.cast(vals, ctypes.POINTER(ctypes.c_uint32)), shape=(n,))
        o = np.argsort(k, kind='stable')
        k[:], v[:] = k[o], v[o]
        return 0


@pytest.mark.parametrize('sc,gold,kw', CASES)
def test_find_orth_host_half_reference_golden(monkeypatch, sc, gold, kw):
    from swiftortho_b200 import _lib, find_orth
    monkeypatch.setattr(_lib, 'load', lambda: OracleOrthLib())
    monkeypatch.setattr(_lib, 'check', lambda rc: None)
    out = io.StringIO()
    find_orth.find_orth(os.path.join(GOLDEN, sc), sep='|', out=out, **kw)
    assert out.getvalue() == open(os.path.join(GOLDEN, gold)).read()


@pytest.mark.gpu
@pytest.mark.parametrize('sc,gold,kw', CASES)
def test_find_orth_device_reference_golden(sc, gold, kw):
    """The whole path with the CUDA kernels (so_orth_classify, so_sort_pairs_u64): byte-identical to the reference."""
    from swiftortho_b200 import build, find_orth
    build.build()
    out = io.StringIO()
    find_orth.find_orth(os.path.join(GOLDEN, sc), sep='|', out=out, **kw)
    assert out.getvalue() == open(os.path.join(GOLDEN, gold)).read()


@pytest.mark.gpu
def test_orth_classify_against_oracle_random():
    """so_orth_classify on random groups (duplicate targets, score ties, many taxa) against the checker."""
    from swiftortho_b200 import _lib, build
    build.build()
    lib = _lib.load()
    rnd = np.random.default_rng(3)
    sizes = rnd.integers(1, 700, size=300)
    goff = np.concatenate(([0], np.cumsum(sizes))).astype(np.uint64)
    n = int(goff[-1])
    ntaxa = 37
    qr = np.repeat(rnd.integers(0, 5000, size=len(sizes)), sizes).astype(np.uint32)
    sr = rnd.integers(0, 400, size=n).astype(np.uint32)
    tax_of = rnd.integers(0, ntaxa, size=5000).astype(np.uint32)
    qt, st = tax_of[qr], tax_of[sr % 5000]
    sc = rnd.integers(20, 60, size=n).astype(np.float64) / rnd.choice([1., 3., 7.], size=n)
    got, want = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8)
    args = lambda c: (0, goff.ctypes.data, len(sizes), qr.ctypes.data, sr.ctypes.data, qt.ctypes.data, st.ctypes.data,
                      sc.ctypes.data, ntaxa, c.ctypes.data)
    _lib.check(lib.so_orth_classify(*args(got)))
    OracleOrthLib().so_orth_classify(*args(want))
    assert (got == want).all()
    keys = rnd.integers(0, 1 << 40, size=100000).astype(np.uint64)
    vals = np.arange(len(keys), dtype=np.uint32)
    k2 = keys.copy()
    _lib.check(lib.so_sort_pairs_u64(0, k2.ctypes.data, vals.ctypes.data, len(keys)))
    assert (k2 == np.sort(keys)).all() and (keys[vals] == k2).all()

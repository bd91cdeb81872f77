"""CPU-side checks of the product library: it loads, exports every symbol the header declares, and
its host-only entry points (H0 FASTA, H1 seg, Q quicksort order, F text) reproduce the golden
vectors generated from the reference source.  No device compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


@pytest.fixture(scope='module')
def lib():
    from swiftortho_b200 import build, _lib
    build.build()
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, 'include', 'swiftortho_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = set(re.findall(r'\b(so_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) >= 25
    from swiftortho_b200 import _lib
    assert names == set(_lib.SYMBOLS), names ^ set(_lib.SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.so_abi_version() == 1


def test_seg_matches_reference(lib, kat):
    for r in kat['seg']:
        a = r['in'].encode('latin-1')
        out = C.create_string_buffer(len(a) + 1)
        assert lib.so_seg(a, len(a), out) == 0
        assert out.raw[:len(a)].decode('latin-1') == r['out'], r['in']


def test_qsort_matches_reference(lib, kat):
    for r in kat['qsort']:
        n = len(r['keys'])
        k = (C.c_int64 * max(n, 1))(*r['keys'])
        p = (C.c_int32 * max(n, 1))()
        assert lib.so_qsort_perm(k, n, p) == 0
        assert list(p)[:n] == r['perm']


def test_qsort_prefix_equals_full_sort_prefix(lib, oracle):
    # the pruned quicksort used for candidates must give the same first `need` entries as the full one
    rng = np.random.default_rng(5)
    for n in [10, 100, 1000, 20000]:
        keys = rng.integers(0, 40, size=n).tolist()
        k = (C.c_int64 * n)(*keys)
        p = (C.c_int32 * n)()
        lib.so_qsort_perm(k, n, p)
        assert list(p) == oracle.qsort_perm(keys)


def test_bits_and_evalue_text(lib, kat):
    for s, b in kat['score2bit']:
        assert lib.so_score2bit(s) == b
    buf = C.create_string_buffer(64)
    for e, s in kat['f2s']:
        assert lib.so_f2s(float(e), buf, 64) == 0
        assert buf.value.decode() == s, e
    assert lib.so_bit2e(15028, 450, 450, 897) == 15028 * 450 * 450 * 2.0 ** -897


def test_fasta_container(lib):
    from swiftortho_b200.search import Fasta
    f = Fasta(os.path.join(GOLDEN, 'g4.fsa'))
    assert len(f) == 4
    assert f.header(0) == 'A|a1 first protein' and f.header(3) == 'D|d1'
    assert f.sequence(0).startswith('MENIHDLWERALAEMEKK') and len(f.sequence(0)) == 90
    assert len(f.sequence(2)) == 66 and f.n_residues == 90 + 90 + 66 + 60
    f2 = Fasta(os.path.join(GOLDEN, 'odd24.fsa'))
    assert len(f2) == 24


def test_no_cpu_fallback(lib):
    """Without a GPU the compute entry points must fail loudly (SO_ENODEV), never fall back."""
    if lib.so_device_count() > 0:
        pytest.skip('a GPU is present')
    from swiftortho_b200 import _lib
    from swiftortho_b200.search import Searcher
    with pytest.raises(_lib.SoError, match='no CUDA device'):
        Searcher(device=0, ht=1000003)


def test_find_hit_cli_parsing():
    from swiftortho_b200 import find_hit
    a = find_hit.parse_args(['find_hit.py', '-p', 'blastp', '-i', 'q.fsa', '-dq.fsa', '-e', '1e-5', '-s111111', 'junk'])
    assert a['-p'] == 'blastp' and a['-d'] == 'q.fsa' and a['-s'] == '111111' and a['-e'] == '1e-5'
    assert a['-M'] == '120000000' and a['-c'] == '50000' and a['-j'] == '1' and a['-v'] == '500'


def test_query_slices_balanced():
    from swiftortho_b200 import find_hit

    class F:
        offsets = np.concatenate([[0], np.cumsum(np.arange(1, 101))]).astype(np.uint64)
    sl = find_hit.slices_by_residues(F, 0, 100, 4)
    assert sl[0][0] == 0 and sl[-1][1] == 100
    assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
    w = [int(F.offsets[e]) - int(F.offsets[s]) for s, e in sl]
    assert max(w) - min(w) < 250

"""Pins the CPU oracle (oracle/fsearch_oracle.cpp) to the reference.

Every expected value under tests/golden/ was produced by executing the reference's own
lib/fsearch.py (tests/golden/make_golden.py, oracle/ref_shim/run_reference.py).  These tests run
on CPU (no GPU marker).
"""
import math
import os

import pytest

from conftest import GOLDEN


def test_b62_table(oracle, kat):
    # fsearch.py:330-346
    for i, j, v in kat['b62_samples']:
        assert oracle.L.orc_b62(i, j) == v
    chk = sum((i * 257 + j) * (oracle.L.orc_b62(i, j) + 5) for i in range(256) for j in range(256)) % (1 << 61)
    assert chk == kat['b62_checksum']


def test_kswat_st_known_answers(oracle, kat):
    # fsearch.py:1357-1476; tuple = (idy, AL, mis, gap, qst, qed, sst, sed, bit)
    assert kat['kswat_state_leaks'] == 0
    n = 0
    for r in kat['kswat_st']:
        idy, out = oracle.kswat_st(r['s0'], r['s1'], r['qst'], r['sst'])
        assert out[:8] == r['out'], (r['s0'], r['s1'], out, r['out'])
        if r['idy'] is None:
            assert math.isnan(idy)
        else:
            assert idy == r['idy']
        n += 1
    assert n >= 200


def test_kswat_survey_vectors(oracle):
    # SURVEY.md section 9.2
    core = 'MENIHDLWERALAEMEKKVSKPSYETWLKSTKANDIQ'
    idy, o = oracle.kswat_st('GGGG' + core, core + 'P' * 10)
    assert o[:8] == [41, 4, 2, 0, 41, 0, 37, 80] and abs(idy - 90.2439) < 1e-3
    idy, o = oracle.kswat_st(core, 'GGGG' + core + 'PPPP')
    assert o[:8] == [41, 4, 2, 0, 37, 0, 41, 80]
    idy, o = oracle.kswat_st(core[:20] + 'WWWWWW' + core[20:], core + 'P' * 12)
    assert o[:8] == [43, 6, 3, 0, 43, 0, 37, 73]


def test_ungap(oracle, kat):
    # fsearch.py:2454-2494
    for r in kat['ungap']:
        assert oracle.ungap(r['q'], r['s'], r['Q'], r['S'], r['qlo'], r['slo']) == r['out'], r


def test_seg(oracle, kat):
    # fsearch.py:2872-2928
    for r in kat['seg']:
        assert oracle.seg(r['in']) == r['out'], r['in']


def test_score2bit_f2s(oracle, kat):
    # fsearch.py:1066-1071, 43-61
    for s, b in kat['score2bit']:
        assert oracle.L.orc_score2bit(s) == b
    assert oracle.L.orc_score2bit(2317) == 897  # README.md:52
    for e, s in kat['f2s']:
        assert oracle.f2s(float(e)) == s, e


def test_qsort_permutation(oracle, kat):
    # fsearch.py:260-327
    for r in kat['qsort']:
        assert oracle.qsort_perm(r['keys']) == r['perm']


def test_spseeds(oracle, kat):
    # fsearch.py:519-556, 406-422
    for r in kat['spseeds']:
        assert oracle.spseeds(r['seq'], r['step'], r['nr'], r['ssd'], r['mod']) == r['out'], r


def _case_names():
    import json
    p = os.path.join(GOLDEN, 'cases.json')
    if not os.path.exists(p):
        return []
    return [c['name'] for c in json.load(open(p))]


@pytest.mark.parametrize('name', _case_names())
def test_end_to_end_golden(oracle, golden_cases, name, tmp_path):
    """Whole fsearch-c runs: the oracle's output file must equal the reference's byte for byte."""
    case = [c for c in golden_cases if c['name'] == name][0]
    fsa = os.path.join(GOLDEN, name + '.fsa')
    qry = os.path.join(GOLDEN, name + '.qry.fsa') if case['separate_query'] else fsa
    out = str(tmp_path / (name + '.sc'))
    flags = {k: v for k, v in case['flags'].items() if k != '-p'}
    oracle.blastp(qry, fsa, out, flags)
    got = open(out, 'rb').read()
    exp = open(os.path.join(GOLDEN, name + '.sc'), 'rb').read()
    assert got.count(b'\n') == case['rows']
    assert got == exp


def test_readme_known_answer(oracle, tmp_path):
    # README.md:52: A|a1 A|a1 100.00 450 0 0 1 450 1 450 2.88e-261 897 450 450
    exp = open(os.path.join(GOLDEN, 'qry450.sc'), 'rb').read().split(b'\t')
    assert exp[2:10] == [b'100.00', b'450', b'0', b'0', b'1', b'450', b'1', b'450']
    assert exp[11:14] == [b'897', b'450', b'450']


def test_parallel_hoare_formula():
    """The data-parallel statement of the reference's Hoare partition (lib/fsearch.py:281-297) used by the
    device-side candidate sort (csrc/select.cu): with a_k the k-th position from the left whose key is >= pivot
    and b_k the k-th from the right whose key is <= pivot (both on the array before the partition), the loop
    swaps exactly (a_k, b_k) for k <= K = #{k: a_k <= b_k} and ends at j = max(b_{K+1}, a_K') with K' = K if
    a_K < b_K else K - 1.  Checked against the sequential loop on random ranges with heavy ties."""
    import random

    def seq_partition(x, l, r):
        pv = x[l][0]
        i, j = l, r + 1
        while True:
            i += 1
            while i <= r and x[i][0] < pv:
                i += 1
            j -= 1
            while x[j][0] > pv:
                j -= 1
            if i > j:
                break
            x[i], x[j] = x[j], x[i]
        x[l], x[j] = x[j], x[l]
        return j

    def par_partition(x, l, r):
        pv = x[l][0]
        A = [p for p in range(l + 1, r + 1) if x[p][0] >= pv]
        B = [p for p in range(l, r + 1) if x[p][0] <= pv]
        nA, nB = len(A), len(B)
        ok = [A[k] <= B[nB - 1 - k] for k in range(min(nA, nB))]
        K = sum(ok)
        assert ok == [True] * K + [False] * (len(ok) - K)      # monotone
        for k in range(K):
            a, b = A[k], B[nB - 1 - k]
            x[a], x[b] = x[b], x[a]
        if K == 0:
            j = B[nB - 1]
        else:
            assert K < nB                                      # position l is a j-stopper that never pairs
            Kp = K if A[K - 1] < B[nB - K] else K - 1
            j = max(B[nB - 1 - K], A[Kp - 1] if Kp >= 1 else -1)
        x[l], x[j] = x[j], x[l]
        return j

    rnd = random.Random(7)
    for _ in range(20000):
        n = rnd.randint(2, 48)
        kr = rnd.choice([0, 1, 2, 3, 5, 10, 100])
        x = [(rnd.randint(0, kr), i) for i in range(n)]
        l = rnd.randint(0, n - 2)
        r = rnd.randint(l + 1, n - 1)
        x1, x2 = list(x), list(x)
        assert seq_partition(x1, l, r) == par_partition(x2, l, r) and x1 == x2

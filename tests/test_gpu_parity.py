"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and against the golden
vectors generated from the reference source.  Integer / index work must be bit exact; the e-value and
identity doubles are compared through the printed text (byte-identical files) and to 1e-12 relative."""
import math
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

AA9 = 'AST,CFILMVY,DN,EQ,G,H,KR,P,W'


@pytest.fixture(scope='module')
def so():
    from swiftortho_b200 import build, search
    build.build()
    return search


def _write_fasta(path, seqs, prefix='s'):
    with open(path, 'wb') as f:
        for i, s in enumerate(seqs):
            f.write(b'>%s%d\n%s\n' % (prefix.encode(), i, s.encode('latin-1')))


def _align_setup(so, tmp_path, s0s, s1s):
    q, t = str(tmp_path / 'q.fsa'), str(tmp_path / 't.fsa')
    _write_fasta(q, s0s, 'q')
    _write_fasta(t, s1s, 't')
    Q, T = so.Fasta(q), so.Fasta(t)
    S = so.Searcher(device=0, ht=1000003, flt='F')
    S.set_targets(T)
    S.set_queries(Q)
    return S, Q, T


def _check_aln(lib, got, idy, out, what):
    # out = [AL, mis, gap, qst, qed, sst, sed, bit(, raw)]
    assert [got['aln_len'], got['mismatch'], got['gaps'], got['qst'], got['qed'], got['sst'], got['sed']] == out[:7], what
    assert lib.so_score2bit(got['raw_score']) == out[7], what
    if got['aln_len'] == 0:
        assert idy is None or (isinstance(idy, float) and math.isnan(idy))
    else:
        assert got['n_ident'] * (100. / got['aln_len']) == idy, what


def test_align_reference_known_answers(so, kat, tmp_path):
    """kswat_st vectors produced by the reference source (tests/golden/kat.json)."""
    rows = kat['kswat_st']
    S, Q, T = _align_setup(so, tmp_path, [r['s0'] for r in rows], [r['s1'] for r in rows])
    pairs = [(i, i, 0, len(r['s0']), 0, len(r['s1']), r['qst'], r['sst']) for i, r in enumerate(rows)]
    got = S.align(pairs)
    for g, r in zip(got, rows):
        _check_aln(S.lib, g, r['idy'], r['out'], (r['s0'], r['s1'], r['qst'], r['sst']))
    st = S.stats()
    assert st['kernel_launches'] >= 2 and st['alignments'] == len(rows)


def test_align_random_against_oracle(so, oracle, tmp_path):
    rnd = random.Random(3)
    A = 'ARNDCQEGHILKMFPSTWYV'

    def rp(n):
        return ''.join(rnd.choice(A) for _ in range(n))

    def mut(s, r):
        o = []
        for c in s:
            x = rnd.random()
            if x < r * 0.2:
                continue
            if x < r * 0.4:
                o.append(rnd.choice(A))
            o.append(rnd.choice(A) if x < r else c)
        return ''.join(o) or 'A'
    s0s, s1s, starts = [], [], []
    for k in range(1500):
        n = rnd.choice([1, 2, 5, 16, 17, 18, 31, 32, 33, 40, 100, 350, 351, 700, 1500]) if k % 3 == 0 else rnd.randrange(1, 900)
        a = rp(n)
        b = mut(a, rnd.uniform(0.02, 0.7)) if rnd.random() < 0.7 else rp(rnd.randrange(1, 900))
        if rnd.random() < 0.15:
            a = a[:n // 2] + 'x' * 12 + a[n // 2:]
        if rnd.random() < 0.1:
            b = b.lower()
        d = rnd.randrange(-40, 41) if rnd.random() < 0.6 else 0
        s0s.append(a)
        s1s.append(b)
        starts.append((0, d) if d > 0 else (-d, 0))
    # long rows: the 4096-residue tile limit
    a = rp(4095)
    s0s += [a, a, rp(4095)]
    s1s += [mut(a, 0.1)[:4095], a, rp(300)]
    starts += [(0, 0), (0, 0), (3, 0)]
    S, Q, T = _align_setup(so, tmp_path, s0s, s1s)
    pairs = [(i, i, 0, len(s0s[i]), 0, len(s1s[i]), starts[i][0], starts[i][1]) for i in range(len(s0s))]
    got = S.align(pairs)
    cells = 0
    for i, g in enumerate(got):
        idy, out = oracle.kswat_st(s0s[i], s1s[i], starts[i][0], starts[i][1])
        _check_aln(S.lib, g, idy, out, i)
        assert g['raw_score'] == out[8]
        cells += g['cells']
    assert cells == S.stats()['dp_cells']


def test_align_slices_and_clamping(so, oracle, tmp_path):
    """(q_off, q_len, t_off, t_len) slices = kswat_st on sub-strings (the kswat_st_long tiles)."""
    rnd = random.Random(9)
    A = 'ARNDCQEGHILKMFPSTWYV'
    a = ''.join(rnd.choice(A) for _ in range(600))
    b = a[:200] + ''.join(rnd.choice(A) for _ in range(30)) + a[200:]
    S, Q, T = _align_setup(so, tmp_path, [a], [b])
    cases = [(0, 600, 0, 630, 0, 0), (100, 300, 90, 400, 0, 0), (100, 300, 90, 400, 5, 2), (0, 600, 0, 630, 700, 0),
             (0, 600, 0, 630, 0, 630), (590, 10, 600, 30, 0, 0), (0, 600, 0, 630, -5, -7)]
    got = S.align([(0, 0) + c for c in cases])
    for c, g in zip(cases, got):
        qo, ql, to, tl, qs, ss = c
        idy, out = oracle.kswat_st(a[qo:qo + ql], b[to:to + tl], qs, ss)
        _check_aln(S.lib, g, idy, out, c)


CASES = ['g4', 'g4_spaced', 'qry450', 'synth60', 'synth60_chunk25', 'synth60_v3', 'synth60_noflt_j2', 'synth60_window',
         'synth60_thr', 'synth40_multi', 'synth40_alph2', 'synth40_aa20', 'odd24', 'long8']


def _run_case(so, case, name, out, **kw):
    fsa = os.path.join(GOLDEN, name + '.fsa')
    qry = os.path.join(GOLDEN, name + '.qry.fsa') if case['separate_query'] else fsa
    f = case['flags']
    return so.blastp(qry, fsa, out, expect=float(f['-e']), v=int(f['-v']), max_miss=float(f['-m']), st=int(f['-l']),
                     ed=int(f['-u']), rst=int(f['-L']), red=int(f['-U']), thr=int(f['-t']), flt=f['-F'], ssd=f['-s'],
                     nr=f['-r'], step=int(f['-j']), ht=int(f['-M']), chk=int(f['-c']), **kw)


@pytest.mark.parametrize('name', CASES)
def test_end_to_end_reference_golden(so, golden_cases, name, tmp_path):
    """Whole search: output file byte-identical to the one the reference source produced."""
    case = [c for c in golden_cases if c['name'] == name][0]
    out = str(tmp_path / 'out.sc')
    st = _run_case(so, case, name, out)
    got = open(out, 'rb').read()
    exp = open(os.path.join(GOLDEN, name + '.sc'), 'rb').read()
    assert got.count(b'\n') == case['rows']
    assert got == exp
    assert st['kernel_launches'] > 0


def _synth(tmp_path, n, taxa, seed, max_len=None, lengths='gamma'):
    from swiftortho_b200 import synth
    h, s = synth.generate(n, taxa, lengths, seed=seed, max_len=max_len)
    p = str(tmp_path / ('synth_%d_%d.fsa' % (n, seed)))
    with open(p, 'wb') as f:
        f.write(synth.to_fasta_bytes(h, s))
    return p


def test_index_against_oracle(so, oracle, tmp_path):
    """K1-K3: bucket starts, locus order (reverse insertion) and threshold, two chunks, two patterns."""
    p = _synth(tmp_path, 300, 5, 11, max_len=400)
    for ssd, nr, step in [('111111', AA9, 1), ('11111,1101011', AA9 + '/A,S,T,C,F,I,L,M,V,Y,D,N,E,Q,G,H,K,R,P,W', 2)]:
        T = so.Fasta(p)
        S = so.Searcher(device=0, ssd=ssd, nr=nr, ht=1000003, step=step, chk=170)
        S.set_targets(T)
        infos = S.build_index()
        assert [(i['chunk_start'], i['chunk_end']) for i in infos] == [(0, 170), (170, 300)]
        for k, info in enumerate(infos):
            st_o, loc_o = oracle.index(p, info['chunk_start'], info['chunk_end'], {'-s': ssd, '-r': nr, '-j': str(step)})
            st_g, loc_g = S.index_export(k, 1000003, info['n_seeds'])
            assert info['n_seeds'] == len(loc_o)
            assert np.array_equal(loc_g, loc_o)
            assert np.array_equal(st_g[:1000003], st_o)
            assert st_g[1000003] == len(loc_o)
            _, thr = oracle.candidates(p, p, info['chunk_start'], info['chunk_end'], 0, 0, {'-s': ssd, '-r': nr, '-j': str(step)})
            assert info['threshold'] == thr
        S.close()


@pytest.mark.parametrize('ssd,nr,flt', [('111111', AA9, 'T'), ('111111,1110100111', AA9, 'T'), ('1111', AA9, 'F')])
def test_candidates_against_oracle(so, oracle, tmp_path, ssd, nr, flt):
    """S3 + K4-K6: candidate lists (target, score, qi, qj) in the reference's order, per chunk."""
    p = _synth(tmp_path, 240, 6, 21, max_len=350)
    T = so.Fasta(p)
    S = so.Searcher(device=0, ssd=ssd, nr=nr, ht=1000003, step=1, chk=100, flt=flt)
    S.set_targets(T)
    S.set_queries(T)
    infos = S.build_index()
    tot = 0
    for k, info in enumerate(infos):
        exp, _ = oracle.candidates(p, p, info['chunk_start'], info['chunk_end'], 0, 240, {'-s': ssd, '-r': nr, '-F': flt})
        got = S.candidates(k, 0, 240)
        S.set_sub_block(7)
        got7 = S.candidates(k, 0, 240)
        S.set_sub_block(0)
        for q in range(240):
            assert np.array_equal(got[q], exp[q]), (k, q)
            assert np.array_equal(got7[q], exp[q]), (k, q)
            tot += len(exp[q])
    assert tot > 1000
    S.close()


def test_search_baseline_flags_against_oracle(so, oracle, tmp_path):
    """The BASELINE flag set (-M 120000000 -c 50000 -e 1e-5 -s 111111 aa9) on a scaled-down config 2,
    and independence from the query slicing (find_hit.py -l/-u)."""
    p = _synth(tmp_path, 1200, 12, 20261019)
    ref = str(tmp_path / 'oracle.sc')
    stats_o = oracle.blastp(p, p, ref, {'-e': '1e-5', '-j': '1', '-M': '120000000', '-c': '50000', '-s': '111111'})
    out = str(tmp_path / 'gpu.sc')
    st = so.blastp(p, p, out, expect=1e-5, step=1, ht=120000000, chk=50000, ssd='111111')
    exp = open(ref, 'rb').read()
    assert open(out, 'rb').read() == exp
    assert st['seed_hits'] <= stats_o['seed_hits']       # hits on "sequence -1" are dropped before grouping
    assert st['candidates'] == stats_o['candidates']
    # sliced run (two windows) == full run
    out2 = str(tmp_path / 'gpu2.sc')
    so.blastp(p, p, out2, expect=1e-5, step=1, ht=120000000, chk=50000, ssd='111111', st=0, ed=500)
    so.blastp(p, p, out2, expect=1e-5, step=1, ht=120000000, chk=50000, ssd='111111', st=500, ed=1200, wrt='a')
    assert open(out2, 'rb').read() == exp
    # properties of the output: grouped by ascending query ordinal, bit score non-increasing inside a query
    last_q, last_bit = -1, None
    for line in exp.split(b'\n')[:-1]:
        c = line.split(b'\t')
        q, bit = int(c[14]), int(c[11])
        assert q >= last_q
        if q == last_q:
            assert bit <= last_bit
        last_q, last_bit = q, bit


@pytest.mark.parametrize('env', [{'SO_CUB_SORT': '1'}, {'SO_GENERIC_UNGAP': '1'}, {'SO_NO_SINGLE': '1'},
                                 {'SO_FORCE_PAIRS': '1'}, {'SO_XDROP_REFILL': '8'},
                                 {'SO_QUERY_BLOCK': '64', 'SO_SUB_BLOCK0': '37'}])
def test_alternative_code_paths_against_oracle(so, oracle, tmp_path, env):
    """Every fallback / alternative device path gives the oracle's file too: device-wide radix sort instead of the
    cell partition, the generic chained X-drop kernel instead of k_xdrop, hit ordinals carried through a pairs
    sort, another refill threshold, odd query-block / sub-block sizes.  (The library reads these variables per
    call.)"""
    p = _synth(tmp_path, 700, 8, 20261023)
    ref = str(tmp_path / 'oracle.sc')
    oracle.blastp(p, p, ref, {'-e': '1e-5', '-j': '1', '-M': '120000000', '-c': '400', '-s': '111111'})
    out = str(tmp_path / 'gpu.sc')
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        so.blastp(p, p, out, expect=1e-5, step=1, ht=120000000, chk=400, ssd='111111')
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert open(out, 'rb').read() == open(ref, 'rb').read()


def test_search_one_lane_equals_two_lanes(so, tmp_path):
    """so_search with one candidate-production lane returns the rows of the default two-lane pipeline."""
    p = _synth(tmp_path, 900, 9, 20261024)
    F = so.Fasta(p)
    S = so.Searcher(device=0, ssd='111111', expect=1e-5, chk=300)
    S.set_targets(F)
    S.set_queries(F)
    S.build_index()
    a = S.search(0, F.N).as_array().copy()
    S.set_lanes(1)
    b = S.search(0, F.N).as_array().copy()
    assert len(a) > 900 and a.tobytes() == b.tobytes()
    S.close()


def test_full_size_config2_window_against_oracle(so, oracle, tmp_path):
    """BASELINE config 2 at FULL size (100 000 synthetic proteins, 2 index chunks, NC = 120 M, the cell-partition
    path with ~137 k seed hits per query and chunk): a window of queries against the complete target set must
    give the oracle's file byte for byte (queries are independent, find_hit.py -l/-u), plus the size-independent
    properties of the whole block the window lies in."""
    from swiftortho_b200 import synth
    p = str(tmp_path / 'c2.fsa')
    synth.write_config(p, 2, n=100000, taxa=20)
    flags = {'-e': '1e-5', '-j': '1', '-M': '120000000', '-c': '50000', '-s': '111111'}
    ref = str(tmp_path / 'oracle.sc')
    oracle.blastp(p, p, ref, dict(flags, **{'-l': '1000', '-u': '1024'}))
    out = str(tmp_path / 'gpu.sc')
    so.blastp(p, p, out, expect=1e-5, step=1, ht=120000000, chk=50000, ssd='111111', st=1000, ed=1024)
    exp = open(ref, 'rb').read()
    assert exp.count(b'\n') > 24
    assert open(out, 'rb').read() == exp
    # the same rows come out of a 1024-query block (other sub-block / lane / alignment-round boundaries)
    out2 = str(tmp_path / 'gpu2.sc')
    so.blastp(p, p, out2, expect=1e-5, step=1, ht=120000000, chk=50000, ssd='111111', st=512, ed=1536)
    rows = [ln for ln in open(out2, 'rb').read().split(b'\n')[:-1]]
    win = b''.join(ln + b'\n' for ln in rows if 1000 <= int(ln.split(b'\t')[14]) < 1024)
    assert win == exp
    last_q, last_bit, firsts = -1, None, 0
    for ln in rows:
        c = ln.split(b'\t')
        q, bit = int(c[14]), int(c[11])
        assert q >= last_q
        if q == last_q:
            assert bit <= last_bit
        else:
            firsts += c[0] == c[1]          # the best hit of an unmasked query is the query itself
        last_q, last_bit = q, bit
    assert firsts > 0.95 * 1024


def test_large_chunk_split_cells_against_oracle(so, oracle, tmp_path):
    """-c 60000: the first chunk holds more targets than one CTA's shared-memory counters cover, so the cell passes
    split the targets of a query into two ranges (two CTAs per query)."""
    from swiftortho_b200 import synth
    p = str(tmp_path / 'c2.fsa')
    synth.write_config(p, 2, n=100000, taxa=20)
    ref = str(tmp_path / 'oracle.sc')
    oracle.blastp(p, p, ref, {'-e': '1e-5', '-j': '1', '-M': '120000000', '-c': '60000', '-s': '111111', '-l': '70000',
                              '-u': '70012'})
    out = str(tmp_path / 'gpu.sc')
    so.blastp(p, p, out, expect=1e-5, step=1, ht=120000000, chk=60000, ssd='111111', st=70000, ed=70012)
    exp = open(ref, 'rb').read()
    assert exp.count(b'\n') >= 12
    assert open(out, 'rb').read() == exp


@pytest.mark.parametrize('cfg,n,taxa,ssd,ev,lo', [(5, 100000, 20, '111111', '1e-5', 2000),
                                                  (4, 100000, 20, '1110100111', '1e-3', 3000)])
def test_full_size_configs_4_5_window_against_oracle(so, oracle, tmp_path, cfg, n, taxa, ssd, ev, lo):
    """BASELINE config 5 (long-tailed lengths 50-5000: >= 4096 tiles, 13-bit qst, 100 000 targets) at full size and
    config 4's flag set (spaced seed 1110100111, -e 1e-3: denser candidates) on 100 000 of its 250 000 proteins:
    a 16-query window against the complete target set gives the oracle's file byte for byte."""
    from swiftortho_b200 import synth
    p = str(tmp_path / 'c.fsa')
    synth.write_config(p, cfg, n=n, taxa=taxa)
    flags = {'-e': ev, '-j': '1', '-M': '120000000', '-c': '50000', '-s': ssd, '-l': str(lo), '-u': str(lo + 16)}
    ref = str(tmp_path / 'oracle.sc')
    oracle.blastp(p, p, ref, flags)
    out = str(tmp_path / 'gpu.sc')
    so.blastp(p, p, out, expect=float(ev), step=1, ht=120000000, chk=50000, ssd=ssd, st=lo, ed=lo + 16)
    exp = open(ref, 'rb').read()
    assert exp.count(b'\n') >= 16
    assert open(out, 'rb').read() == exp


def test_search_long_tailed_config5_against_oracle(so, oracle, tmp_path):
    """Config 5 shape (log-normal lengths up to 5000: the >= 4096 tile path) scaled down."""
    p = _synth(tmp_path, 300, 6, 20261022, lengths='lognormal')
    ref = str(tmp_path / 'oracle.sc')
    oracle.blastp(p, p, ref, {'-e': '1e-5', '-j': '1', '-M': '12000017', '-c': '120', '-s': '111111'})
    out = str(tmp_path / 'gpu.sc')
    so.blastp(p, p, out, expect=1e-5, step=1, ht=12000017, chk=120, ssd='111111')
    assert open(out, 'rb').read() == open(ref, 'rb').read()


def test_search_config4_seed_against_oracle(so, oracle, tmp_path):
    """Config 4 flags (spaced seed 1110100111, -e 1e-3) scaled down."""
    p = _synth(tmp_path, 600, 10, 20261021)
    ref = str(tmp_path / 'oracle.sc')
    oracle.blastp(p, p, ref, {'-e': '1e-3', '-j': '1', '-M': '120000000', '-c': '50000', '-s': '1110100111'})
    out = str(tmp_path / 'gpu.sc')
    so.blastp(p, p, out, expect=1e-3, step=1, ht=120000000, chk=50000, ssd='1110100111')
    assert open(out, 'rb').read() == open(ref, 'rb').read()


def test_find_hit_cli_drop_in(so, golden_cases, tmp_path):
    """The find_hit.py-compatible command line (bin/find_hit.py flags) writes the reference's file."""
    import subprocess
    import sys
    from conftest import ROOT
    fsa = os.path.join(GOLDEN, 'synth60.fsa')
    out = str(tmp_path / 'cli.sc')
    r = subprocess.run([sys.executable, '-m', 'swiftortho_b200.find_hit', '-p', 'blastp', '-i', fsa, '-d', fsa, '-o', out,
                        '-e', '1e-5', '-s', '111111', '-M', '1000003', '-a', '1'], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert 'chk size 50000' in r.stdout
    assert open(out, 'rb').read() == open(os.path.join(GOLDEN, 'synth60.sc'), 'rb').read()
    assert not os.path.exists(out + '_sc_tmpdir')

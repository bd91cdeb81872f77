#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by executing the REFERENCE SOURCE itself.

Runs only in the build container (needs /root/reference); the outputs are committed so that the
tests, smoke() and bench.py never touch /root/reference at run time.

    python tests/golden/make_golden.py            # regenerates everything (several minutes)

Two kinds of vectors:
  * end-to-end: small FASTA fixtures + the exact `fsearch-c` flag set -> expected output file
    (`cases.json` lists them; `<case>.fsa` / `<case>.sc`)
  * function level: kswat_st, ungap, seg, f2s, score2bit, qsort permutation, spseeds_fnv,
    generate_nr_tbl (`kat.json`)
"""
import json
import os
import random
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_shim'))
sys.setrecursionlimit(100000)

import run_reference  # noqa: E402
from swiftortho_b200 import synth  # noqa: E402

AA9 = 'AST,CFILMVY,DN,EQ,G,H,KR,P,W'
AA20 = 'A,S,T,C,F,I,L,M,V,Y,D,N,E,Q,G,H,K,R,P,W'

G4 = b""">A|a1 first protein
MENIHDLWERALAEMEKKVSKPSYETWLKSTKANDIQNDVITITAPNEFARDWLEEHYAG
LTSDTIEHLTGARLTPRFVIPQNELEDDFL
>B|b1
MENLHDLWERSLAEMEKKVSKPSFETWLKSTKANDIQNDVITAPNEFARDWLEDHYAGLT
SDTIEHLTGGGARLTPRFVIPQNEIEDDFL
>C|c1 low complexity
MKKAAAAAAAAAAAAAAAAAAAAAAAAAAAAHDLWERALAEMEKKVSKPSYETWLKSTKANDIQND
>D|d1
MSTNPKPQRKTKRNTNRRPQDVKFPGGGQIVGGVYLLPRRGPRLGVRATRKTSERSQPRG
"""


def flags(**kw):
    d = {'-p': 'blastp', '-e': '1e-5', '-v': '500', '-l': '-1', '-u': '-1', '-L': '-1', '-U': '-1',
         '-m': '1e-3', '-t': '-1', '-j': '1', '-F': 'T', '-O': 'wb', '-M': '1000003', '-c': '50000',
         '-s': '111111', '-r': AA9}
    d.update(kw)
    return d


def run_case(name, fasta_bytes, fl, cases, query_bytes=None):
    fsa = os.path.join(HERE, name + '.fsa')
    with open(fsa, 'wb') as f:
        f.write(fasta_bytes)
    qfsa = fsa
    if query_bytes is not None:
        qfsa = os.path.join(HERE, name + '.qry.fsa')
        with open(qfsa, 'wb') as f:
            f.write(query_bytes)
    tmp = tempfile.mkdtemp()
    out = os.path.join(HERE, name + '.sc')
    argv = []
    for k, v in fl.items():
        argv += [k, v]
    argv += ['-i', qfsa, '-d', fsa, '-o', out, '-T', tmp, '-D', '']
    print('running', name, flush=True)
    run_reference.run_entry_point(argv)
    shutil.rmtree(tmp, ignore_errors=True)
    cases.append({'name': name, 'flags': fl, 'separate_query': query_bytes is not None,
                  'rows': sum(1 for _ in open(out, 'rb'))})


def synth_fasta(n, taxa, seed, max_len=260, lengths='gamma'):
    h, s = synth.generate(n, taxa, lengths, seed=seed, max_len=max_len)
    return synth.to_fasta_bytes(h, s), h, s


def end_to_end():
    cases = []
    run_case('g4', G4, flags(), cases)
    run_case('g4_spaced', G4, flags(**{'-s': '1110100111', '-e': '1e-3'}), cases)
    # README.md:52 known answer: the 450-aa example protein against itself
    q450 = open('/root/reference/example/qry.fsa', 'rb').read()
    run_case('qry450', q450, flags(), cases)
    s60, _, _ = synth_fasta(60, 6, 101)
    run_case('synth60', s60, flags(), cases)
    run_case('synth60_chunk25', s60, flags(**{'-c': '25'}), cases)
    run_case('synth60_v3', s60, flags(**{'-v': '3', '-e': '10'}), cases)
    run_case('synth60_noflt_j2', s60, flags(**{'-F': 'F', '-j': '2'}), cases)
    run_case('synth60_window', s60, flags(**{'-l': '7', '-u': '19', '-L': '10', '-U': '50'}), cases)
    run_case('synth60_thr', s60, flags(**{'-t': '2', '-M': '5003'}), cases)
    s40, _, _ = synth_fasta(40, 4, 202, max_len=200)
    run_case('synth40_multi', s40, flags(**{'-s': '111111,1110100111,110011', '-e': '1e-3'}), cases)
    run_case('synth40_alph2', s40, flags(**{'-r': AA9 + '/' + AA20, '-s': '11111,1101011', '-c': '17'}), cases)
    run_case('synth40_aa20', s40, flags(**{'-r': AA20, '-s': '1111'}), cases)
    # odd residues / lower case / X runs, separate query file
    rnd = random.Random(7)
    _, hh, ss = synth_fasta(24, 3, 303, max_len=180)
    recs = []
    for k, (h, s) in enumerate(zip(hh, ss)):
        b = bytearray(s.tobytes())
        for _ in range(6):
            b[rnd.randrange(len(b))] = rnd.choice(b'XBZUJO*xabcz')
        if k % 5 == 0:
            b = bytearray(bytes(b).lower())
        recs.append((h + (' desc %d' % k if k % 3 == 0 else ''), bytes(b)))
    odd = b''.join(b'>' + h.encode() + b'\n' + s + b'\n' for h, s in recs)
    qodd = b''.join(b'>' + h.encode() + b'\n' + s + b'\n' for h, s in recs[3:15])
    run_case('odd24', odd, flags(**{'-e': '1e-3'}), cases, query_bytes=qodd)
    # long path (>= 4096 residues): one long family + short ones
    h, s = synth.generate(6, 3, 'gamma', seed=404, max_len=150)
    rng = __import__('numpy').random.Generator(__import__('numpy').random.PCG64(5))
    anc = synth.AA[rng.choice(20, size=4500, p=synth.RR)]
    longs = [synth._mutate(rng, anc, 0.15) for _ in range(2)]
    h += ['L0|long_a', 'L1|long_b']
    s += longs
    run_case('long8', synth.to_fasta_bytes(h, s), flags(), cases)
    with open(os.path.join(HERE, 'cases.json'), 'w') as f:
        json.dump(cases, f, indent=1)


def function_level():
    g = run_reference.load_reference()
    rnd = random.Random(11)
    kat = {}
    core = 'MENIHDLWERALAEMEKKVSKPSYETWLKSTKANDIQ'

    def fresh():
        return [[0] * 1000 for _ in range(1000)], [["*"] * 1000 for _ in range(1000)]

    def rand_prot(n, alphabet='ARNDCQEGHILKMFPSTWYV'):
        return ''.join(rnd.choice(alphabet) for _ in range(n))

    def mutate(s, rate):
        out = []
        for c in s:
            r = rnd.random()
            if r < rate * 0.15:
                continue
            if r < rate * 0.3:
                out.append(rnd.choice('ARNDCQEGHILKMFPSTWYV'))
            if r < rate:
                out.append(rnd.choice('ARNDCQEGHILKMFPSTWYV'))
            else:
                out.append(c)
        return ''.join(out)

    # kswat_st
    rows = []
    pairs = [('GGGG' + core, core + 'P' * 10, 0, 0), (core, 'GGGG' + core + 'PPPP', 0, 0),
             (core[:20] + 'WWWWWW' + core[20:], core + 'P' * 12, 0, 0), (core, core, 0, 0),
             ('A', 'A', 0, 0), ('AW', 'W', 0, 0), (core, core, 5, 0), (core, core, 0, 7),
             (core, core, 40, 3), (core, core, 37, 37)]
    for _ in range(220):
        a = rand_prot(rnd.randrange(1, 400))
        kind = rnd.random()
        if kind < 0.6:
            b = mutate(a, rnd.uniform(0.05, 0.6))
            if not b:
                b = 'A'
        else:
            b = rand_prot(rnd.randrange(1, 400))
        if rnd.random() < 0.3:
            a = a[:rnd.randrange(len(a))] + 'x' * 12 + a[rnd.randrange(len(a)):]
        if rnd.random() < 0.2:
            b = b.lower()
        if rnd.random() < 0.2:
            b = b + rnd.choice('XBZUJO*') + b[:5]
        d = rnd.randrange(-30, 31) if rnd.random() < 0.7 else 0
        qi, qj = (0, d) if d > 0 else (-d, 0)
        pairs.append((a, b, qi, qj))
    for a, b, qi, qj in pairs:
        sm, tm = fresh()
        al0, al1 = [], []
        r = g['kswat_st'](a, b, qst=qi, sst=qj, score=sm, trace=tm, al0=al0, al1=al1)
        idy = None if r[0] != r[0] else r[0]
        rows.append({'s0': a, 's1': b, 'qst': qi, 'sst': qj, 'idy': idy, 'out': [int(x) for x in r[1:]]})
    kat['kswat_st'] = rows
    # shared dirty matrices == fresh matrices (no state leak): run all pairs through ONE matrix pair
    sm, tm = fresh()
    leak = 0
    for a, b, qi, qj in pairs:
        r = g['kswat_st'](a, b, qst=qi, sst=qj, score=sm, trace=tm, al0=[], al1=[])
        sm2, tm2 = fresh()
        r2 = g['kswat_st'](a, b, qst=qi, sst=qj, score=sm2, trace=tm2, al0=[], al1=[])
        if repr(r) != repr(r2):
            leak += 1
    kat['kswat_state_leaks'] = leak

    # ungap (method of Fasta; does not touch self)
    ung = []
    U = g['Fasta'].ungap
    for _ in range(200):
        a = rand_prot(rnd.randrange(8, 200))
        b = mutate(a, rnd.uniform(0.0, 0.5)) or 'AAAA'
        if rnd.random() < 0.3:
            b = rand_prot(rnd.randrange(8, 200))
        Q, S = rnd.randrange(0, len(a)), rnd.randrange(0, len(b) + 1)
        qlo = rnd.choice([-1, -1, rnd.randrange(0, len(a))])
        slo = -1 if qlo == -1 else max(0, S - (Q - qlo))
        r = U(None, a, b, Q, S, qlo=qlo, slo=slo)
        ung.append({'q': a, 's': b, 'Q': Q, 'S': S, 'qlo': qlo, 'slo': slo, 'out': [int(x) for x in r[:5]]})
    for args in [(core, core, 0, 0), (core, core, 5, 5), (core, core, 1, 1)]:
        r = U(None, *args)
        ung.append({'q': args[0], 's': args[1], 'Q': args[2], 'S': args[3], 'qlo': -1, 'slo': -1,
                    'out': [int(x) for x in r[:5]]})
    kat['ungap'] = ung

    # seg
    segs = []
    cases = ['MKK' + 'A' * 40 + 'MENIHDLWERALAEMEKKVSKPSYETWLKS', core, 'A' * 11, 'ACDEFGHIKLM', 'AAAAAAAAAAAA',
             'ACDEFGHIKLMN', 'acdefghiklmnpqrstvwy' * 3, 'SSSSSSSSSSSSGGGGGGGGGGGGG' + core]
    for _ in range(150):
        n = rnd.randrange(1, 300)
        kind = rnd.random()
        if kind < 0.4:
            s = rand_prot(n)
        elif kind < 0.7:
            s = rand_prot(n, 'AAAASG')
        else:
            s = rand_prot(n // 2 + 1) + rnd.choice('AQSP') * rnd.randrange(5, 40) + rand_prot(n // 2 + 1, 'KRED')
        if rnd.random() < 0.2:
            s = s.lower()
        cases.append(s)
    for s in cases:
        segs.append({'in': s, 'out': g['seg'](s)[0]})
    kat['seg'] = segs

    kat['score2bit'] = [[s, int(g['score2bit'](s))] for s in [0, 1, 24, 25, 196, 2317, 2398, 30000] +
                        [rnd.randrange(0, 5000) for _ in range(50)]]
    es = [1.14e-264, 0.00234, 0.5, 3.0, 0.0, 1e-3, 9.99e-4, 1e-5, 1e-300, 9.9999996e-7, 2.88e-261, 1e-10, 3.1e-310, 7.7e-316]
    es += [10 ** rnd.uniform(-300, 2) for _ in range(200)]
    kat['f2s'] = [[repr(e), g['f2s'](e)] for e in es]

    qs = []
    for _ in range(120):
        n = rnd.choice([0, 1, 2, 5, 6, 7, 8, 9, 20, 50, 333, 1000, 2500])
        span = rnd.choice([1, 3, 10, 1000000])
        keys = [rnd.randrange(-span, span + 1) for _ in range(n)]
        x = [[k, i] for i, k in enumerate(keys)]
        g['qsort'](x, key=lambda e: e[0])
        qs.append({'keys': keys, 'perm': [e[1] for e in x]})
    kat['qsort'] = qs

    sp = []
    for ssd, nr in [('111111', AA9), ('1110100111', AA9), ('111111,1110100111,110011', AA9),
                    ('11111,1101011', AA9 + '/' + AA20), ('1111', AA20), ('11,101', 'AST,CFILMVY')]:
        codes = [g['generate_nr_tbl'](e) for e in nr.split('/')]
        for step in (1, 2):
            for _ in range(4):
                s = rand_prot(rnd.randrange(3, 120), 'ARNDCQEGHILKMFPSTWYVxX')
                out = list(g['spseeds_fnv'](s, step=step, codes=codes, ssps=ssd, mod=1000003))
                sp.append({'seq': s, 'step': step, 'ssd': ssd, 'nr': nr, 'mod': 1000003,
                           'out': [[int(a), int(b)] for a, b in out]})
                out = list(g['spseeds_fnv'](s, step=step, codes=codes, ssps=ssd, mod=13))
                sp.append({'seq': s, 'step': step, 'ssd': ssd, 'nr': nr, 'mod': 13,
                           'out': [[int(a), int(b)] for a, b in out]})
    kat['spseeds'] = sp
    kat['nr_tbl'] = {nr: [int(v) for v in g['generate_nr_tbl'](nr)][:256] for nr in [AA9, AA20, 'ast,CFIL']}
    b = g['b62']
    kat['b62_checksum'] = sum((i * 257 + j) * (b[i][j] + 5) for i in range(256) for j in range(256)) % (1 << 61)
    kat['b62_samples'] = [[i, j, b[i][j]] for i, j in
                          [(ord(a), ord(c)) for a in 'AWxX*UbZ-' for c in 'AWxXyC*J\r']]
    with open(os.path.join(HERE, 'kat.json'), 'w') as f:
        json.dump(kat, f)


if __name__ == '__main__':
    what = sys.argv[1:] or ['kat', 'e2e']
    if 'kat' in what:
        function_level()
    if 'e2e' in what:
        end_to_end()

import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftortho_b200 import search as so
kat = json.load(open('tests/golden/kat.json'))['kswat_st']
d = tempfile.mkdtemp()
def wf(p, seqs, pre):
    with open(p, 'wb') as f:
        for i, s in enumerate(seqs):
            f.write(b'>%s%d\n%s\n' % (pre, i, s.encode('latin-1')))
wf(d + '/q.fsa', [r['s0'] for r in kat], b'q'); wf(d + '/t.fsa', [r['s1'] for r in kat], b't')
Q, T = so.Fasta(d + '/q.fsa'), so.Fasta(d + '/t.fsa')
print('records', len(Q), len(T), len(kat))
for i, r in enumerate(kat):
    assert Q.sequence(i) == r['s0'] and T.sequence(i) == r['s1'], i
S = so.Searcher(device=0, ht=1000003, flt='F')
S.set_targets(T); S.set_queries(Q)
pairs = [(i, i, 0, len(r['s0']), 0, len(r['s1']), r['qst'], r['sst']) for i, r in enumerate(kat)]
nbad = 0
for i, p in enumerate(pairs):
    try:
        g = S.align([p])[0]
        o = [g['aln_len'], g['mismatch'], g['gaps'], g['qst'], g['qed'], g['sst'], g['sed']]
        if o != kat[i]['out'][:7]:
            nbad += 1; print('single mismatch', i, o, kat[i]['out'], len(kat[i]['s0']), len(kat[i]['s1']))
    except Exception as e:
        nbad += 1; print('single fail', i, e)
print('single-pair batches bad:', nbad)
for n in (2, 3, 32, 33, 64, 128, 129, len(pairs)):
    try:
        got = S.align(pairs[:n])
        bad = [i for i in range(n) if [got[i][k] for k in ('aln_len','mismatch','gaps','qst','qed','sst','sed')] != kat[i]['out'][:7]]
        print('batch', n, 'mismatches', bad[:10])
    except Exception as e:
        print('batch', n, 'fail', e)

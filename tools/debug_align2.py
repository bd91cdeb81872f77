import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['SO_DEBUG_TRACE'] = '1'
from swiftortho_b200 import search as so
kat = json.load(open('tests/golden/kat.json'))['kswat_st']
d = tempfile.mkdtemp()
def wf(p, seqs, pre):
    with open(p, 'wb') as f:
        for i, s in enumerate(seqs):
            f.write(b'>%s%d\n%s\n' % (pre, i, s.encode('latin-1')))
wf(d + '/q.fsa', [r['s0'] for r in kat], b'q'); wf(d + '/t.fsa', [r['s1'] for r in kat], b't')
Q, T = so.Fasta(d + '/q.fsa'), so.Fasta(d + '/t.fsa')
S = so.Searcher(device=0, ht=1000003, flt='F')
S.set_targets(T); S.set_queries(Q)
for i in (0, 196):
    r = kat[i]
    print('pair', i, flush=True)
    try:
        print(S.align([(i, i, 0, len(r['s0']), 0, len(r['s1']), r['qst'], r['sst'])]), r['out'], flush=True)
    except Exception as e:
        print(e)

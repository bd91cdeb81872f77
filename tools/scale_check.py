"""Larger-than-test sanity run: N proteins (several index chunks), a query window, property checks."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from swiftortho_b200 import search as so
n = int(sys.argv[1]); taxa = int(sys.argv[2]); nq = int(sys.argv[3])
p = bench.dataset(n, taxa)
F = so.Fasta(p)
S = so.Searcher(device=0, **bench.FLAGS)
t = time.time(); S.set_targets(F); S.set_queries(F); info = S.build_index(); print('setup %.1f s, %d chunks' % (time.time() - t, len(info)), flush=True)
t = time.time(); rows = S.search(0, nq); dt = time.time() - t
print('cold: %.2f s' % dt, flush=True)
S.stats(reset=True)
t = time.time(); rows = S.search(nq, 2 * nq); dt = time.time() - t   # warm: buffers allocated
a = rows.as_array()
st = S.stats()
print('search %d queries vs %d targets: %.2f s (%.0f q/s) rows %d' % (nq, n, dt, nq / dt, len(a)))
print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items()}))
# properties: rows grouped by ascending query, bit non-increasing within a query, every unmasked query hits itself first
import numpy as np
q = a['query']; assert np.all(np.diff(q) >= 0)
assert q.min() >= nq
same = np.diff(q) == 0
assert np.all(np.diff(a['bit'])[same] <= 0)
first = np.concatenate([[True], ~same])
selfhit = (a['query'][first] == a['target'][first]).mean()
print('queries with rows: %d, first row is the self hit: %.3f' % (first.sum(), selfhit))
assert selfhit > 0.95

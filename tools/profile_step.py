"""One search step on config 2 for ncu (small query block so replays stay short)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from swiftortho_b200 import search as so
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = bench.dataset(100000, 20)
F = so.Fasta(p)
S = so.Searcher(device=0, **bench.FLAGS)
S.set_targets(F); S.set_queries(F); S.build_index()
for r in range(reps):
    S.stats(reset=True)
    t = time.time(); rows = S.search(r * nq, (r + 1) * nq); dt = time.time() - t
    print('search %d queries %.3f s rows %d' % (nq, dt, rows.n), json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in S.stats().items()}), flush=True)

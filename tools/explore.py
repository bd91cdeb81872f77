import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftortho_b200 import search as so, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
p = '/tmp/c2_%d.fsa' % n
if not os.path.exists(p):
    t = time.time(); synth.write_config(p, 2, n=n, taxa=max(2, n // 5000)); print('gen', time.time() - t, flush=True)
t = time.time(); F = so.Fasta(p); print('fasta', time.time() - t, len(F), F.n_residues, flush=True)
S = so.Searcher(device=0, ssd='111111', ht=120000000, step=1, expect=1e-5, chk=50000)
t = time.time(); S.set_targets(F); print('set_targets', time.time() - t, flush=True)
t = time.time(); S.set_queries(F); print('set_queries', time.time() - t, flush=True)
t = time.time(); info = S.build_index(); print('index', time.time() - t, info, flush=True)
for rep in range(2):
    S.stats(reset=True)
    t = time.time(); rows = S.search(rep * nq, (rep + 1) * nq); dt = time.time() - t
    st = S.stats()
    print('search %d queries: %.3f s -> %.1f q/s; rows %d' % (nq, dt, nq / dt, rows.n), flush=True)
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}), flush=True)

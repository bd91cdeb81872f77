"""Gapped-extension-only benchmark: one big so_align_batch on config-2 shaped pairs -> GCUPS."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from swiftortho_b200 import search as so
import ctypes as C
M = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = bench.dataset(100000, 20)
F = so.Fasta(p)
S = so.Searcher(device=0, **bench.FLAGS)
S.set_targets(F); S.set_queries(F)
n = len(F)
P = (so.so_pair * M)()
off = F.offsets
for i in range(M):
    q = i % n; t = (i * 7919 + 13) % n
    P[i] = so.so_pair(q, t, 0, int(off[q + 1] - off[q]), 0, int(off[t + 1] - off[t]), 0, 0)
A = (so.so_aln * M)()
for r in range(reps):
    S.stats(reset=True)
    t0 = time.time(); so.check(S.lib.so_align_batch(S.h, P, M, A)); dt = time.time() - t0
    st = S.stats()
    print(json.dumps({'pairs': M, 'wall_s': round(dt, 3), 'ms_dp': round(st['ms_dp'], 2), 'ms_traceback': round(st['ms_traceback'], 2),
                      'cells': st['dp_cells'], 'gcups_dp': round(st['dp_cells'] / st['ms_dp'] / 1e6, 1),
                      'gcups_dp_tb': round(st['dp_cells'] / (st['ms_dp'] + st['ms_traceback']) / 1e6, 1)}), flush=True)

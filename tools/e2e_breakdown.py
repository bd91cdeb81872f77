"""Where the end-to-end arm of bench.py spends its time (per step of 4096 queries)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from swiftortho_b200 import search as so
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = bench.dataset(100000, 20)
F = so.Fasta(p)
S = so.Searcher(device=0, **bench.FLAGS)
S.set_targets(F); S.build_index(); S.set_queries(F)
lib = S.lib
res_ptr = F._res.value
T = dict(setq=0., search=0., rows=0., write=0.)
for s in range(5):
    a, b = (s + 3) * B, (s + 4) * B
    t0 = time.perf_counter()
    off = np.ascontiguousarray(F.offsets[a:b + 1])
    so.check(lib.so_set_queries(S.h, C.c_void_p(res_ptr), off.ctypes.data, b - a))
    t1 = time.perf_counter()
    rows = S.search(0, b - a)
    t2 = time.perf_counter()
    arr = rows.as_array(); arr['query'] += a
    hits = (so.so_hit * max(1, len(arr))).from_buffer_copy(arr.tobytes())
    t3 = time.perf_counter()
    so.check(lib.so_write_rows(hits, len(arr), F.h, F.h, b'/dev/shm/e2e_test.sc', 0))
    t4 = time.perf_counter()
    if s >= 2:
        T['setq'] += t1 - t0; T['search'] += t2 - t1; T['rows'] += t3 - t2; T['write'] += t4 - t3
print({k: round(1e3 * v / 3, 1) for k, v in T.items()})

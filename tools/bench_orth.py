"""Timing of the find_orth row on the BASELINE config-2 hit table: the table is produced by the search
(100 000 proteins all-vs-all), then `swiftortho_b200.find_orth` (device classification + device sorts, host joins) is
timed end to end; the share of the two device entry points is reported beside it.

    python tools/bench_orth.py [n_proteins] > profiles/bench_orth_r02.json
"""
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from swiftortho_b200 import _lib, find_orth, search  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
fsa = bench.dataset(n, 20, config=2)
shm = '/dev/shm' if os.path.isdir('/dev/shm') else '/tmp'
sc = os.path.join(shm, 'bench_orth.sc')
t0 = time.perf_counter()
search.blastp(fsa, fsa, sc, expect=1e-5, step=1, ht=120000000, chk=50000, ssd='111111', block=16384)
t_search = time.perf_counter() - t0
rows = sum(1 for _ in open(sc, 'rb'))
lib = _lib.load()
spent = {'so_orth_classify': 0.0, 'so_sort_pairs_u64': 0.0}


class Timed:
    def __getattr__(self, k):
        f = getattr(lib, k)
        if k not in spent:
            return f

        def g(*a):
            t = time.perf_counter()
            r = f(*a)
            spent[k] += time.perf_counter() - t
            return r
        return g


find_orth._lib.load = lambda: Timed()
res = {}
for norm in ('no', 'bsr'):
    for k in spent:
        spent[k] = 0.0
    out = io.StringIO()
    t0 = time.perf_counter()
    find_orth.find_orth(sc, .5, 0., norm, '|', out)
    dt = time.perf_counter() - t0
    lines = out.getvalue().count('\n')
    res[norm] = {'seconds': dt, 'rows_per_s': rows / dt, 'orth_lines': lines, 'device_entry_points_s': dict(spent)}
os.remove(sc)
print(json.dumps({'what': 'find_orth on the config-2 hit table (%d proteins, %d rows)' % (n, rows), 'search_seconds': t_search,
                  'find_orth': res, 'note': 'the reference find_orth.py cannot run on the GPU box (no /root/reference there); '
                  'in the build container it takes 0.3 s on the 5 472-row synth600 fixture (17 k rows/s, one core)'}))

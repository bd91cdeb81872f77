"""Small workload for compute-sanitizer (memcheck / racecheck): the device quicksort on a few sizes, one golden search
through the sync-free cell path and through the general path, find_orth and the sequence hash.

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import ctypes as C
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swiftortho_b200 import find_orth, nr, search  # noqa: E402

G = os.path.join(ROOT, 'tests', 'golden')
S = search.Searcher(device=0, ht=1000003)
rnd = np.random.default_rng(1)
for n, kr in ((50, 3), (1500, 7), (5000, 40)):
    keys = rnd.integers(0, kr + 1, size=n, dtype=np.uint32)
    got = np.zeros(n, dtype=np.uint32)
    search.check(S.lib.so_qsort_prefix_device(S.h, keys.ctypes.data, n, min(n, 600), got.ctypes.data))
    k64 = (C.c_int64 * n)(*[int(v) for v in keys])
    perm = (C.c_int32 * n)()
    search.check(S.lib.so_qsort_perm(k64, n, perm))
    assert list(got[:min(n, 600)]) == list(perm)[:min(n, 600)]
S.close()
d = tempfile.mkdtemp()
fsa = os.path.join(G, 'synth60.fsa')
want = open(os.path.join(G, 'synth60.sc'), 'rb').read()
for env in ({}, {'SO_NO_FAST': '1'}, {'SO_CELL_MAX': '3'}):
    os.environ.update(env)
    out = os.path.join(d, 'o.sc')
    search.blastp(fsa, fsa, out, expect=1e-5, step=1, ht=1000003, chk=50000, ssd='111111')
    assert open(out, 'rb').read() == want, env
    for k in env:
        del os.environ[k]
o = io.StringIO()
find_orth.find_orth(os.path.join(G, 'synth60.sc'), .5, 0., 'no', '|', o)
assert o.getvalue() == open(os.path.join(G, 'synth60.orth')).read()
o = io.StringIO()
nr.nr_flt(os.path.join(G, 'synth60_dup.fsa'), o)
assert o.getvalue() == open(os.path.join(G, 'synth60_dup_nr.fsa')).read()
print('sanitize_small ok')

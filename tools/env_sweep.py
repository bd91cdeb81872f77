"""Tuning-hook sweep on config 2: one process, the index built once, every setting timed over the same query
blocks (two warm-up steps, six timed steps of 4096 queries, two production lanes = the bench.py `value`
pipeline), then one measurement-mode step for the per-stage CUDA-event times and the in-pipeline DP rate.

    python tools/env_sweep.py [out.json] [B] [--stages]

The hooks are read by so_search on every call, so the environment can change between calls."""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from swiftortho_b200 import search as so

SETTINGS = [  # (name, production lanes, environment); edit for the experiment at hand
    ('default', 1, {}),
    ('xdrop_16_warps_1_view', 1, {'SO_XDROP_WARPS': '16', 'SO_XDROP_QSHIFT': '0'}),     # layout of the first half of round 2
    ('cell_small_warp', 1, {'SO_CELL_SPAN': '0'}),                                       # k_cell_small + k_cell_warp
    ('both_old', 1, {'SO_XDROP_WARPS': '16', 'SO_XDROP_QSHIFT': '0', 'SO_CELL_SPAN': '0'}),
    ('query_block_512', 1, {'SO_QUERY_BLOCK': '512'}),
    ('lanes_2', 2, {}),
    ('lanes_2_own_streams', 2, {'SO_SHARED_STREAM': '0'}),
    ('default_again', 1, {}),
]
HOOKS = sorted({k for _, _, e in SETTINGS for k in e})


def main():
    pos = [a for a in sys.argv[1:] if not a.startswith('--')]
    out = pos[0] if pos else 'gpurun_out/env_sweep.json'
    B = int(pos[1]) if len(pos) > 1 else 4096
    p = bench.dataset(100000, 20)
    F = so.Fasta(p)
    S = so.Searcher(device=0, **bench.FLAGS)
    S.set_targets(F)
    S.build_index()
    S.set_queries(F)
    res = []
    for name, lanes, env in SETTINGS:
        for k in HOOKS:
            os.environ.pop(k, None)
        os.environ.update(env)
        rec = {'name': name, 'lanes': lanes, 'env': env, 'queries_per_step': B}
        try:
            S.set_lanes(lanes)
            for s in range(2):
                S.search(s * B, (s + 1) * B)
            S.stats(reset=True)
            t0 = time.perf_counter()
            nrows = 0
            md5 = hashlib.md5()
            for s in range(6):
                r = S.search((2 + s) * B, (3 + s) * B)
                nrows += r.n
                md5.update(r.as_array().tobytes())
            dt = time.perf_counter() - t0
            st = S.stats(reset=True)
            rec.update(ms_per_step=1e3 * dt / 6, proteins_per_s=6 * B / dt, rows=nrows, rows_md5=md5.hexdigest(),
                       alignments=st['alignments'], alignments_used=st.get('alignments_used'))
            if '--no-stages' not in sys.argv:
                S.set_lanes(0)
                S.search(2 * B, 3 * B)
                S.stats(reset=True)
                S.search(2 * B, 3 * B)
                k = S.stats(reset=True)
                rec['stage_ms'] = {x: round(k[x], 2) for x in ('ms_seed', 'ms_sort', 'ms_ungap', 'ms_select', 'ms_dp', 'ms_traceback')}
                rec['dp_gcups_in_pipeline'] = k['dp_cells'] / (k['ms_dp'] * 1e-3) / 1e9 if k['ms_dp'] > 0 else 0.0
        except Exception as e:  # a setting the library rejects is recorded, the sweep goes on
            rec['error'] = repr(e)
        print(json.dumps(rec), flush=True)
        res.append(rec)
    os.makedirs(os.path.dirname(out) or '.', exist_ok=True)
    with open(out, 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()

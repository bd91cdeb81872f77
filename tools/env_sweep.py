"""Tuning-hook sweep on config 2: one process, the index built once, every setting timed over the same query
blocks (two warm-up steps, three timed steps of 4096 queries, two production lanes = the bench.py `value`
pipeline), then one measurement-mode step for the per-stage CUDA-event times and the in-pipeline DP rate.

    python tools/env_sweep.py [out.json] [B]

The hooks are read by so_search on every call, so the environment can change between calls."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from swiftortho_b200 import search as so

SETTINGS = [
    ('baseline', {}),
    ('align_batch_2048', {'SO_ALIGN_BATCH': '2048'}),
    ('align_batch_4096', {'SO_ALIGN_BATCH': '4096'}),
    ('query_block_592', {'SO_QUERY_BLOCK': '592'}),
    ('query_block_1024', {'SO_QUERY_BLOCK': '1024'}),
    ('xdrop_ctas_1', {'SO_XDROP_CTAS': '1'}),
    ('xdrop_ctas_1_split_2', {'SO_XDROP_CTAS': '1', 'SO_CELL_SPLIT': '2'}),
    ('baseline_again', {}),
]
HOOKS = sorted({k for _, e in SETTINGS for k in e})


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/env_sweep.json'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    p = bench.dataset(100000, 20)
    F = so.Fasta(p)
    S = so.Searcher(device=0, **bench.FLAGS)
    S.set_targets(F)
    S.build_index()
    S.set_queries(F)
    res = []
    for name, env in SETTINGS:
        for k in HOOKS:
            os.environ.pop(k, None)
        os.environ.update(env)
        rec = {'name': name, 'env': env}
        try:
            S.set_lanes(2)
            rows0 = None
            for s in range(2):
                S.search(s * B, (s + 1) * B)
            S.stats(reset=True)
            t0 = time.perf_counter()
            nrows = 0
            for s in range(3):
                r = S.search((2 + s) * B, (3 + s) * B)
                nrows += r.n
            dt = time.perf_counter() - t0
            st = S.stats(reset=True)
            rec.update(ms_per_step=1e3 * dt / 3, proteins_per_s=3 * B / dt, rows=nrows,
                       alignments=st['alignments'], alignments_used=st.get('alignments_used'))
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

"""Clustering row measurement: a synthetic ortholog table (F families x T taxa, every family a dense
block with a few weak links between families) through swiftortho_b200.find_cluster on the GPU; stage wall times (host
parsing / rounds / batches vs the device calls) and, with --reference, the reference script on the same file on the
host cores of the machine that runs it (/root/reference exists in the build container only).

    python tools/bench_cluster.py [--families 5000] [--taxa 20] [--alg mcl|apc] [--reference] [--out x.json]"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_table(path, F, T, seed=7):
    rng = np.random.default_rng(seed)
    with open(path, 'w') as f:
        for fam in range(F):
            names = ['T%03d|g%07d_0' % (t, fam) for t in range(T)]
            for a in range(T):
                for b in range(a + 1, T):
                    if rng.random() < 0.9:
                        f.write('OT\t%s\t%s\t%.6f\n' % (names[a], names[b], rng.uniform(0.5, 1.5)))
            if fam and rng.random() < 0.05:
                o = int(rng.integers(0, fam))
                f.write('CO\tT%03d|g%07d_0\tT%03d|g%07d_0\t%.6f\n' % (1, o, 0, fam, rng.uniform(0.05, 0.3)))


class Timed:
    def __init__(self, b):
        self.b, self.t, self.calls = b, {}, {}

    def __getattr__(self, k):
        fn = getattr(self.b, k)

        def w(*a, **kw):
            t0 = time.perf_counter()
            r = fn(*a, **kw)
            self.t[k] = self.t.get(k, 0.) + time.perf_counter() - t0
            self.calls[k] = self.calls.get(k, 0) + 1
            return r
        return w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--families', type=int, default=5000)
    ap.add_argument('--taxa', type=int, default=20)
    ap.add_argument('--alg', default='both')
    ap.add_argument('--reference', action='store_true')
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    d = '/dev/shm' if os.path.isdir('/dev/shm') else '/tmp'
    p = os.path.join(d, 'bench_cluster_%d_%d.orth' % (a.families, a.taxa))
    make_table(p, a.families, a.taxa)
    nlines = sum(1 for _ in open(p))
    rec = {'families': a.families, 'taxa': a.taxa, 'lines': nlines, 'runs': []}
    algs = ['mcl', 'apc'] if a.alg == 'both' else [a.alg]
    if a.reference:
        for alg in algs:
            t0 = time.perf_counter()
            r = subprocess.run([sys.executable, '/root/reference/bin/find_cluster.py', '-i', p, '-a', alg], stdout=subprocess.PIPE,
                               stderr=subprocess.DEVNULL, text=True)
            dt = time.perf_counter() - t0
            rec['runs'].append({'impl': 'reference', 'alg': alg, 'wall_s': dt, 'clusters': r.stdout.count('\n'),
                                'cores': os.cpu_count()})
            print(json.dumps(rec['runs'][-1]), flush=True)
    else:
        from swiftortho_b200 import find_cluster as fc
        B = Timed(fc.DeviceBackend(0))
        B.cc_labels(4, [0], [1])  # context creation outside the timed region
        for alg in algs:
            for rep in range(2):
                B.t, B.calls = {}, {}
                t0 = time.perf_counter()
                lines = fc.cluster(p, alg, backend=B)
                dt = time.perf_counter() - t0
            rec['runs'].append({'impl': 'swiftortho_b200', 'alg': alg, 'wall_s': dt, 'clusters': len(lines),
                                'device_calls_s': dict(B.t), 'device_calls': dict(B.calls), 'lines_per_s': nlines / dt})
            print(json.dumps(rec['runs'][-1]), flush=True)
    if a.out:
        with open(a.out, 'w') as f:
            json.dump(rec, f, indent=1)


if __name__ == '__main__':
    main()

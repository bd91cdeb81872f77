"""Clustering row (reference bin/find_cluster.py).

CPU (`-m "not gpu"`): oracle/cluster_oracle.cpp against the known answers of the reference's own functions
(`mcl`, `apclust_blk`, exec'ed from the reference source by tests/golden/make_cluster_golden.py) and the host logic
of swiftortho_b200.find_cluster -- driven by that oracle through its backend seam -- against the reference script's
output on five inputs x five parameter sets.
GPU (`-m gpu`): the CUDA kernels through the C ABI against the oracle (bit-exact: labels, float32 matrices) and the
CLI against the reference goldens.

Tolerance: the reference's outputs are partitions (sets of names); they must be equal.  Float32 values of the MCL
matrix are compared GPU vs oracle bit for bit (same canonical summation order); against scipy's product order they
differ by float32 rounding only, which changes none of the golden partitions."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from swiftortho_b200 import find_cluster as fc


@pytest.fixture(scope='module')
def golden():
    with open(os.path.join(GOLDEN, 'cluster_golden.json')) as f:
        return json.load(f)


def _canon(lines):
    return sorted(sorted(l.split('\t')) for l in lines if l)


def _partition_of_matrix(backend, n, ptr, col, val):
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr))
    m = val > np.float32(1e-5)
    eu, ev = rows[m], col[m].astype(np.int64)
    cc = backend.cc_labels(n, eu, ev)
    nodes = np.unique(np.concatenate([eu, ev]))
    groups = {}
    for v in nodes:
        groups.setdefault(int(cc[v]), []).append(int(v))
    return sorted(sorted(g) for g in groups.values())


# ----------------------------------------------------------------------------------------------- CPU
def test_oracle_mcl_against_reference_function(cluster_oracle, golden):
    for c in golden['functions']['mcl']:
        ptr, col, val, it = cluster_oracle.mcl(c['n'], c['indptr'], c['indices'], c['data'], c['inflation'])
        assert 0 < it <= 100
        assert _partition_of_matrix(cluster_oracle, c['n'], ptr, col, val) == c['partition']


def test_oracle_apc_against_reference_function(cluster_oracle, golden):
    for c in golden['functions']['apc']:
        lab = cluster_oracle.apc(c['ks'], c['row'], c['col'], c['sim'], c['damp'])
        assert lab.tolist() == c['labels']       # exemplars, exactly


def test_oracle_cc_labels():
    from conftest import ClusterOracle
    rng = np.random.default_rng(5)
    B = ClusterOracle()
    n = 3000
    eu, ev = rng.integers(0, n, 2500), rng.integers(0, n, 2500)
    lab = B.cc_labels(n, eu, ev)
    par = list(range(n))

    def find(x):
        while par[x] != x:
            par[x] = par[par[x]]
            x = par[x]
        return x
    for a, b in zip(eu, ev):
        ra, rb = find(int(a)), find(int(b))
        if ra != rb:
            par[max(ra, rb)] = min(ra, rb)
    comp = {}
    for v in range(n):
        comp.setdefault(find(v), []).append(v)
    for members in comp.values():
        assert set(lab[members].tolist()) == {min(members)}


def test_host_logic_against_reference_goldens(cluster_oracle, golden):
    """find_cluster's host side (parsing, the two clustering rounds of `cnc` with their falsy-zero quirks, sort and
    batches, fc2mat's table, output) on top of the oracle's numeric kernels = the reference's partition; the clusters
    also come out in the reference's order."""
    assert len(golden["cases"]) == 25
    for c in golden['cases']:
        a = dict(zip(c['args'][::2], c['args'][1::2]))
        lines = fc.cluster(os.path.join(GOLDEN, c['input']), a['-a'], float(a.get('-d', 0.5)), float(a.get('-I', 1.5)),
                           backend=cluster_oracle)
        assert _canon(lines) == c['partition'], (c['input'], c['args'])
        assert [sorted(l.split('\t')) for l in lines] == [sorted(l.split('\t')) for l in c['raw']]


def test_find_cluster_argument_parsing():
    a = fc.parse_args(['x', '-i', 'foo', '-amcl', '-I', '2', 'junk', '-d0.7'])
    assert a['-i'] == 'foo' and a['-a'] == 'mcl' and a['-I'] == '2' and a['-d'] == '0.7' and a['-b'] == '25000000'


def test_batches_cut_at_cluster_boundaries(cluster_oracle, tmp_path):
    """More than `chk` lines flush a batch only when the next cluster starts (bin/find_cluster.py:1623-1640); the
    partition does not depend on the batch size."""
    p = os.path.join(GOLDEN, 'synth600.orth')
    whole = fc.mcl_clusters(p, 1.5, cluster_oracle)
    small = fc.mcl_clusters(p, 1.5, cluster_oracle, chk=50)
    assert _canon(whole) == _canon(small) and len(whole) > 10


def test_product_never_imports_the_oracle():
    src = open(os.path.join(ROOT, 'swiftortho_b200', 'find_cluster.py')).read()
    assert 'oracle' not in src.replace('cluster_oracle.cpp', '')


# ----------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope='module')
def device_backend():
    return fc.DeviceBackend(0)


def _random_blocks(rng, nblocks, max_size, noise, wide=0):
    sizes = list(rng.integers(1, max_size, size=nblocks)) + ([wide] if wide else [])
    blk = np.repeat(np.arange(len(sizes)), sizes)
    m = len(blk)
    n = m + 1
    a, b = np.triu_indices(m, 1)
    same = blk[a] == blk[b]
    keep = rng.random(len(a)) < np.where(same, 0.6, noise)
    a, b, same = a[keep], b[keep], same[keep]
    z = np.where(same, rng.uniform(0.3, 2.0, len(a)), rng.uniform(0.01, 0.3, len(a))).astype(np.float32)
    diag = np.zeros(n, dtype=np.float32)
    np.maximum.at(diag, a, z)
    np.maximum.at(diag, b, z)
    d = np.flatnonzero(diag)
    r = np.concatenate([a, b, d])
    c = np.concatenate([b, a, d])
    v = np.concatenate([z, z, diag[d]])
    o = np.lexsort((c, r))
    r, c, v = r[o], c[o], v[o]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, r + 1, 1)
    return n, np.cumsum(ptr), c.astype(np.uint32), v


@pytest.mark.gpu
def test_gpu_cc_labels_against_oracle(device_backend, cluster_oracle):
    rng = np.random.default_rng(11)
    for n, m in ((1, 0), (10, 3), (5000, 4000), (1000000, 900000), (200000, 2000000)):
        eu, ev = rng.integers(0, n, m), rng.integers(0, n, m)
        assert np.array_equal(device_backend.cc_labels(n, eu, ev), cluster_oracle.cc_labels(n, eu, ev))
    # a path: the longest pointer chains
    n = 300000
    perm = rng.permutation(n)
    assert np.array_equal(device_backend.cc_labels(n, perm[:-1], perm[1:]), np.zeros(n, dtype=np.uint32))


@pytest.mark.gpu
def test_gpu_apc_against_oracle_and_reference(device_backend, cluster_oracle, golden):
    for c in golden['functions']['apc']:
        assert device_backend.apc(c['ks'], c['row'], c['col'], c['sim'], c['damp']).tolist() == c['labels']
    rng = np.random.default_rng(12)
    ks = 20000
    a, b = rng.integers(0, ks, 150000), rng.integers(0, ks, 150000)
    keep = a != b
    a, b = a[keep], b[keep]
    z = rng.uniform(0.1, 2.0, len(a)).astype(np.float32)
    row = np.empty(2 * len(a) + ks, dtype=np.uint32)
    col = np.empty_like(row)
    sim = np.empty(len(row), dtype=np.float32)
    row[0:2 * len(a):2], col[0:2 * len(a):2], row[1:2 * len(a):2], col[1:2 * len(a):2] = a, b, b, a
    sim[0:2 * len(a):2] = sim[1:2 * len(a):2] = z
    row[2 * len(a):] = col[2 * len(a):] = np.arange(ks)
    sim[2 * len(a):] = -240.
    for damp, sweeps in ((0.5, 100), (0.9, 7)):
        assert np.array_equal(device_backend.apc(ks, row, col, sim, damp, sweeps), cluster_oracle.apc(ks, row, col, sim, damp, sweeps))


@pytest.mark.gpu
def test_gpu_mcl_against_oracle_and_reference(device_backend, cluster_oracle, golden):
    """The final float32 matrix equals the oracle's bit for bit (structure and values), the partitions equal the
    reference's; a 1500-wide block exercises the CTA-per-row product, another inflation the pow() path."""
    def same_matrix(g, o, infl):
        # I = 1.5 and 2 are computed as x * sqrt(x) and x * x in float64 on both sides (correctly rounded operations):
        # bit-exact.  Other inflations go through pow(), whose last float64 bit may differ between CUDA and glibc;
        # after rounding to float32 that shows up about once in 1e8 values, so those cases are compared to 1e-6.
        assert g[3] == o[3]
        assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
        if infl in (1.5, 2.0):
            assert np.array_equal(g[2], o[2])
        else:
            assert np.allclose(g[2], o[2], rtol=1e-6, atol=0)

    for c in golden['functions']['mcl']:
        g = device_backend.mcl(c['n'], c['indptr'], c['indices'], c['data'], c['inflation'])
        o = cluster_oracle.mcl(c['n'], c['indptr'], c['indices'], c['data'], c['inflation'])
        same_matrix(g, o, c['inflation'])
        assert _partition_of_matrix(device_backend, c['n'], *g[:3]) == c['partition']
    rng = np.random.default_rng(13)
    for nblocks, max_size, noise, wide, infl in ((300, 40, 0.0002, 0, 1.5), (20, 30, 0.001, 1500, 1.5), (60, 50, 0.002, 0, 1.7)):
        n, ptr, col, val = _random_blocks(rng, nblocks, max_size, noise, wide)
        g = device_backend.mcl(n, ptr, col, val, infl)
        o = cluster_oracle.mcl(n, ptr, col, val, infl)
        assert g[3] > 1
        same_matrix(g, o, infl)


@pytest.mark.gpu
def test_gpu_mcl_rejects_unsorted_rows(device_backend):
    from swiftortho_b200._lib import SoError
    with pytest.raises(SoError):
        device_backend.mcl(3, [0, 2, 2, 2], [1, 0], [1., 1.], 1.5)


@pytest.mark.gpu
def test_find_cluster_against_reference_goldens(golden, device_backend):
    """find_cluster on the CUDA kernels = the reference script's partition and cluster order on every golden case
    (in process: the call `main` makes), and through the command line on four of them."""
    for c in golden['cases']:
        a = dict(zip(c['args'][::2], c['args'][1::2]))
        lines = fc.cluster(os.path.join(GOLDEN, c['input']), a['-a'], float(a.get('-d', 0.5)), float(a.get('-I', 1.5)),
                           backend=device_backend)
        assert _canon(lines) == c['partition'], (c['input'], c['args'])
        assert [sorted(l.split('\t')) for l in lines] == [sorted(l.split('\t')) for l in c['raw']]
    cli = [c for c in golden['cases'] if c['input'] in ('synth600.orth', 'cluster_quirks.xyz')
           and c['args'] in (['-a', 'mcl', '-I', '1.5'], ['-a', 'apc'])]
    assert len(cli) == 4
    for c in cli:
        r = subprocess.run([sys.executable, '-m', 'swiftortho_b200.find_cluster', '-i', os.path.join(GOLDEN, c['input'])] + c['args'],
                           cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        lines = [l for l in r.stdout.split('\n') if l]
        assert _canon(lines) == c['partition'], (c['input'], c['args'])
        assert [sorted(l.split('\t')) for l in lines] == [sorted(l.split('\t')) for l in c['raw']]


@pytest.mark.gpu
def test_find_cluster_device_equals_oracle_on_a_larger_table(device_backend, cluster_oracle, tmp_path):
    """400 families x 20 taxa (~70 000 lines) with weak links between families: the CUDA-backed and the oracle-backed
    host give the same lines, in the same order, for both algorithms."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import bench_cluster
    p = str(tmp_path / 'fam.orth')
    bench_cluster.make_table(p, 400, 20, seed=3)
    for alg in ('mcl', 'apc'):
        a = fc.cluster(p, alg, backend=device_backend)
        b = fc.cluster(p, alg, backend=cluster_oracle)
        assert a == b and len(a) >= 390

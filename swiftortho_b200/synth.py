"""Deterministic synthetic proteomes for the BASELINE.json configs.

Families model: F ancestral proteins (i.i.d. Robinson-Robinson residues), one mutated copy per
taxon (substitutions U(0.05,0.55), 1 % insertion and 1 % deletion events of geometric length,
mean 3), 2 % of proteins carry a 15-40 aa low-complexity insert (exercises the query `seg`
mask, reference lib/fsearch.py:2872-2946), 5 % of families get an in-paralog.  Records are
written taxon-major like concatenated proteome files, headers `>Ttaxon|gfamily_copy`
(README.md:40-48 of the reference: `taxon|gene`), 60-column lines, upper case, no X.
"""
import io

import numpy as np

AA = np.frombuffer(b'ARNDCQEGHILKMFPSTWYV', dtype=np.uint8)
# Robinson & Robinson background frequencies, same order as AA
RR = np.array([7.8, 5.1, 4.5, 5.4, 1.9, 4.3, 6.3, 7.4, 2.2, 5.1, 9.0, 5.7, 2.2, 3.9, 5.2, 7.1,
               5.8, 1.3, 3.2, 6.4], dtype=np.float64)
RR = RR / RR.sum()

CONFIGS = {
    # id: (N, taxa, length model, seed pattern, e-value)
    1: dict(n=15000, taxa=5, lengths='gamma', seeds='111111', evalue='1e-5'),
    2: dict(n=100000, taxa=20, lengths='gamma', seeds='111111', evalue='1e-5'),
    3: dict(n=1000000, taxa=200, lengths='gamma', seeds='111111', evalue='1e-5'),
    4: dict(n=250000, taxa=50, lengths='gamma', seeds='1110100111', evalue='1e-3'),
    5: dict(n=100000, taxa=20, lengths='lognormal', seeds='111111', evalue='1e-5'),
}


def _mutate(rng, anc, sub_rate, indel_rate=0.01):
    L = anc.shape[0]
    seq = anc.copy()
    m = rng.random(L) < sub_rate
    k = int(m.sum())
    if k:
        seq[m] = AA[rng.choice(20, size=k, p=RR)]
    ndel = rng.binomial(L, indel_rate)
    if ndel:
        keep = np.ones(L, dtype=bool)
        pos = rng.integers(0, L, size=ndel)
        ln = rng.geometric(1.0 / 3.0, size=ndel)
        for p, l in zip(pos, ln):
            keep[p:p + l] = False
        if keep.sum() >= 30:
            seq = seq[keep]
    nins = rng.binomial(seq.shape[0], indel_rate)
    if nins:
        pos = np.sort(rng.integers(0, seq.shape[0] + 1, size=nins))
        ln = rng.geometric(1.0 / 3.0, size=nins)
        parts = []
        last = 0
        for p, l in zip(pos, ln):
            parts.append(seq[last:p])
            parts.append(AA[rng.choice(20, size=int(l), p=RR)])
            last = p
        parts.append(seq[last:])
        seq = np.concatenate(parts)
    return seq


def _low_complexity(rng, seq):
    n = int(rng.integers(15, 41))
    if rng.random() < 0.5:
        ins = np.full(n, AA[rng.integers(0, 20)], dtype=np.uint8)
    else:
        two = AA[rng.integers(0, 20, size=2)]
        ins = np.tile(two, n // 2 + 1)[:n]
    p = int(rng.integers(0, seq.shape[0] + 1))
    return np.concatenate([seq[:p], ins, seq[p:]])


def generate(n, taxa, lengths='gamma', seed=20261017, max_len=None):
    """Return (headers: list[str], seqs: list[np.uint8 array]) of exactly `n` proteins."""
    rng = np.random.Generator(np.random.PCG64(seed))
    fam = max(1, int(n / (taxa + 0.05)))
    while fam * taxa > n:
        fam -= 1
    npar = n - fam * taxa
    if lengths == 'gamma':
        ln = np.clip(rng.gamma(2.0, 175.0, size=fam), 50, 2000).astype(np.int64)
    else:
        ln = np.clip(rng.lognormal(5.6, 0.9, size=fam), 50, 5000).astype(np.int64)
    if max_len is not None:
        ln = np.minimum(ln, max_len)
    par_fams = rng.choice(fam, size=npar, replace=npar > fam)
    par_taxon = rng.integers(0, taxa, size=npar)
    par_of = {}
    for f, t in zip(par_fams, par_taxon):
        par_of.setdefault(int(t), []).append(int(f))
    ancestors = [AA[rng.choice(20, size=int(l), p=RR)] for l in ln]
    headers, seqs = [], []
    for t in range(taxa):
        rates = rng.uniform(0.05, 0.55, size=fam)
        lc = rng.random(fam) < 0.02
        for f in range(fam):
            s = _mutate(rng, ancestors[f], rates[f])
            if lc[f]:
                s = _low_complexity(rng, s)
            headers.append('T%03d|g%07d_0' % (t, f))
            seqs.append(s)
        for c, f in enumerate(par_of.get(t, [])):
            s = _mutate(rng, ancestors[f], rng.uniform(0.02, 0.2))
            headers.append('T%03d|g%07d_%d' % (t, f, c + 1))
            seqs.append(s)
    assert len(seqs) == n, (len(seqs), n)
    return headers, seqs


def _taxon_copies(job):
    """One taxon of generate_parallel: its own PCG64 stream (seed, taxon), same per-copy model as generate()."""
    seed, t, ancestors, pars = job
    rng = np.random.Generator(np.random.PCG64([seed, t]))
    fam = len(ancestors)
    rates = rng.uniform(0.05, 0.55, size=fam)
    lc = rng.random(fam) < 0.02
    headers, seqs = [], []
    for f in range(fam):
        s = _mutate(rng, ancestors[f], rates[f])
        if lc[f]:
            s = _low_complexity(rng, s)
        headers.append('T%03d|g%07d_0' % (t, f))
        seqs.append(s)
    for c, f in enumerate(pars):
        s = _mutate(rng, ancestors[f], rng.uniform(0.02, 0.2))
        headers.append('T%03d|g%07d_%d' % (t, f, c + 1))
        seqs.append(s)
    return to_fasta_bytes(headers, seqs), len(seqs)


def write_parallel(path, n, taxa, lengths='gamma', seed=20261017, workers=None):
    """The same families model with one independent random stream per taxon (PCG64([seed, taxon])), so the taxa
    are generated by a process pool: used for the 1 M-protein config 3 (the single-stream generate() needs two
    minutes there).  Deterministic in (n, taxa, lengths, seed), independent of the worker count."""
    import multiprocessing as mp
    import os
    rng = np.random.Generator(np.random.PCG64([seed, 1 << 20]))
    fam = max(1, int(n / (taxa + 0.05)))
    while fam * taxa > n:
        fam -= 1
    npar = n - fam * taxa
    if lengths == 'gamma':
        ln = np.clip(rng.gamma(2.0, 175.0, size=fam), 50, 2000).astype(np.int64)
    else:
        ln = np.clip(rng.lognormal(5.6, 0.9, size=fam), 50, 5000).astype(np.int64)
    par_fams = rng.choice(fam, size=npar, replace=npar > fam)
    par_taxon = rng.integers(0, taxa, size=npar)
    par_of = {}
    for f, t in zip(par_fams, par_taxon):
        par_of.setdefault(int(t), []).append(int(f))
    ancestors = [AA[rng.choice(20, size=int(l), p=RR)] for l in ln]
    jobs = [(seed, t, ancestors, par_of.get(t, [])) for t in range(taxa)]
    workers = workers or min(32, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1))
    total = 0
    with open(path, 'wb') as f:
        if workers <= 1 or taxa <= 1:
            for j in jobs:
                b, k = _taxon_copies(j)
                f.write(b)
                total += k
        else:
            with mp.get_context('fork').Pool(workers) as pool:
                for b, k in pool.imap(_taxon_copies, jobs, chunksize=1):
                    f.write(b)
                    total += k
    assert total == n, (total, n)
    return n


def to_fasta_bytes(headers, seqs, width=60):
    out = io.BytesIO()
    for h, s in zip(headers, seqs):
        out.write(b'>' + h.encode() + b'\n')
        b = s.tobytes()
        for i in range(0, len(b), width):
            out.write(b[i:i + width])
            out.write(b'\n')
    return out.getvalue()


def write_config(path, config_id, n=None, taxa=None, max_len=None):
    """Write the FASTA of BASELINE config `config_id` (optionally down-scaled) to `path`."""
    c = CONFIGS[config_id]
    n = c['n'] if n is None else n
    taxa = c['taxa'] if taxa is None else taxa
    if config_id == 3 and max_len is None:
        return write_parallel(path, n, taxa, c['lengths'], seed=20261017 + config_id)
    h, s = generate(n, taxa, c['lengths'], seed=20261017 + config_id, max_len=max_len)
    with open(path, 'wb') as f:
        f.write(to_fasta_bytes(h, s))
    return n

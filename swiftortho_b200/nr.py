#!/usr/bin/env python3
"""Redundancy pre-filter and hit re-expansion (reference: scripts/nr_flt.py, scripts/nr2full.py,
used by scripts/run_all_fast.py:110-119 around the all-vs-all search).

    python -m swiftortho_b200.nr flt  proteome.fsa > proteome.fsa_nr.fsa      # scripts/nr_flt.py
    python -m swiftortho_b200.nr full proteome_nr.fsa.sc > proteome.sc        # scripts/nr2full.py

`flt` collapses exact duplicate sequences: one record per distinct sequence, in first-appearance order, whose header
is the ids of the duplicates joined by ';;;' (scripts/nr_flt.py:13-27).  The sequences are hashed on the GPU
(so_seq_hash); the host groups by hash and confirms every group byte for byte, so the result does not depend on the
hash.  `full` rewrites a hit table of such records into one row per (query id, target id) pair
(scripts/nr2full.py:14-44).
"""
import ctypes as C
import sys

import numpy as np

from . import _lib


def parse_fasta(path):
    """Bio.SeqIO.parse(path, 'fasta') as used by scripts/nr_flt.py:13-19: (id, sequence) with id = the title up to the
    first whitespace and sequence = the record's lines right-stripped and joined, blanks and '\\r' removed."""
    ids, seqs = [], []
    cur, title = None, None
    with open(path, 'r') as f:
        for line in f:
            if line.startswith('>'):
                if title is not None:
                    ids.append(title)
                    seqs.append(''.join(cur).replace(' ', '').replace('\r', ''))
                t = line[1:].rstrip()
                title = t.split(None, 1)[0] if t.split() else ''
                cur = []
            elif title is not None:
                cur.append(line.rstrip())
    if title is not None:
        ids.append(title)
        seqs.append(''.join(cur).replace(' ', '').replace('\r', ''))
    return ids, seqs


def duplicate_groups(seqs, device=0):
    """[[index, ...], ...]: the exact-duplicate groups of `seqs` in first-appearance order (device hash + exact check)."""
    lib = _lib.load()
    n = len(seqs)
    if n == 0:
        return []
    enc = [s.encode('latin-1') for s in seqs]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in enc], dtype=np.uint64)
    res = np.frombuffer(b''.join(enc) or b'\0', dtype=np.uint8)
    hashes = np.zeros(n, dtype=np.uint64)
    _lib.check(lib.so_seq_hash(int(device), res.ctypes.data, off.ctypes.data, n, hashes.ctypes.data))
    groups, by_hash = [], {}
    for i in range(n):
        cands = by_hash.setdefault(int(hashes[i]), [])
        for g in cands:                       # equal hash: confirm on the bytes
            if enc[groups[g][0]] == enc[i]:
                groups[g].append(i)
                break
        else:
            cands.append(len(groups))
            groups.append([i])
    return groups


def nr_flt(path, out=None, device=0):
    out = out or sys.stdout
    ids, seqs = parse_fasta(path)
    for g in duplicate_groups(seqs, device):
        out.write('>' + ';;;'.join(ids[i] for i in g) + '\n')
        out.write(seqs[g[0]] + '\n')


def nr2full(path, out=None):
    """scripts/nr2full.py:14-44: rows of one query record are expanded to every (query id, target id) pair of the
    ';;;'-joined headers, columns 3..14 kept, the last two columns replaced by the two full ids, and printed grouped
    by query id in first-appearance order."""
    out = out or sys.stdout

    def flush(hits):
        outs = {}
        for j in hits:
            qds, rds = j[:2]
            for qd in qds.split(';;;'):
                for rd in rds.split(';;;'):
                    q, r = qd.split(' ')[0], rd.split(' ')[0]
                    outs.setdefault(q, []).append('\t'.join([q, r] + j[2:-2] + [qd, rd]))
        for vals in outs.values():
            for v in vals:
                out.write(v + '\n')

    hits = []
    with open(path, 'r') as f:
        for line in f:
            j = line[:-1].split('\t')
            if hits and hits[0][0] != j[0]:
                flush(hits)
                hits = [j]
            else:
                hits.append(j)
    if hits:
        flush(hits)


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) < 3 or argv[1] not in ('flt', 'full'):
        print(__doc__)
        raise SystemExit(2)
    if argv[1] == 'flt':
        nr_flt(argv[2])
    else:
        nr2full(argv[2])


if __name__ == '__main__':
    main()

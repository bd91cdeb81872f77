#!/usr/bin/env python3
"""find_orth.py-compatible orthology inference (reference: bin/find_orth.py).

    python -m swiftortho_b200.find_orth -i foo.sc [-c .5] [-y 0] [-n no|bsr|bal] [-s '|'] > foo.sc.orth

Same flags and the same stdout lines (`IP|OT|CO  id  id  normalised score`) as the reference.  The reference works
on text: it classifies the hits of every query (bin/find_orth.py:298-348), writes three candidate files, sorts them
with GNU sort (LC_ALL=C) and joins them through mmap binary searches (:384-611).  Here every id becomes its rank in
byte order of the id strings (what the C-locale line sort compares first), so the whole pipeline runs on integers:

  * the per-query classification (best score per target, per-taxon maxima, IP / OT / CO call) is one CUDA kernel
    (so_orth_classify, csrc/orth.cu), one CTA per query group;
  * the reciprocal joins of the candidate lists are a device radix sort of (rank, rank) keys (so_sort_pairs_u64)
    followed by a vectorised scan for keys that occur exactly twice;
  * averages are accumulated strictly in the reference's file order (np.add.at: sequential, unbuffered), scores stay the
    doubles the reference computes, and the lines are formatted with Python's str(float) like the reference.

Quirks kept on purpose (bin/find_orth.py): the LAST pair of a sorted candidate file is scored max(a, b) instead of
(a + b) / 2 (:374-376); co-orthologs are only looked up in the orientation (in-paralog of the query id, in-paralog of
the target id) as written (:588-589); `get_sam_tax` never recognises a repeat of the first pair of a taxon run
(:676).  There is no CPU fallback: the classification and the sorts fail without the CUDA library.
"""
import sys

import numpy as np

from . import _lib


def manual_print():
    print('Usage:')
    print('    python %s -i foo.sc [-c .5] [-y 50] [-n no]' % sys.argv[0])
    print('Parameters:')
    print('  -i: tab-delimited file which contain 14 columns')
    print('  -c: min coverage of sequence [0~1]')
    print('  -y: identity [0~100]')
    print('  -n: normalization score [no|bsr|bal]. bsr: bit sore ratio; bal:  bit score over anchored length. Default: no')
    print('  -a: cpu number for sorting. Default: 1')
    print('  -t: keep tmpdir[y|n]. Default: n')
    print('  -T: tmpdir for sort command. Default: ./tmp/')
    print('  -s: separator between taxa and sequence id. Default is |.')


def parse_args(argv):
    # bin/find_orth.py:42-72
    args = {'-i': '', '-c': .5, '-y': 0, '-n': 'no', '-t': 'n', '-a': '4', '-T': './tmp/', '-s': '|'}
    n = len(argv)
    for i in range(1, n):
        k = argv[i]
        if k in args:
            if i + 1 >= n:
                break
            args[k] = argv[i + 1]
        elif k[:2] in args and len(k) > 2:
            args[k[:2]] = k[2:]
    return args


def read_table(path):
    """Columns of a .sc / m8 file the reference reads (bin/find_orth.py:165-197): ids (as sorted categories), identity,
    alignment length, query coordinates, score, query length (or, for 12-column m8 input, max(qst, qed) of the id's
    first row).  Rows with a non-numeric field among columns 3..12 (13, 14) are skipped like the reference's
    try / except (:178-187)."""
    import pandas as pd
    with open(path, 'rb') as f:
        first = f.readline()
    ncol = first.count(b'\t') + 1 if first else 0
    if ncol < 12:
        if not first:
            return None
        raise ValueError('%s: at least 12 tab-separated columns expected' % path)
    numeric = list(range(2, 12)) + ([12, 13] if ncol > 13 else [])
    kw = dict(sep='\t', header=None, keep_default_na=False, na_filter=False, quoting=3, engine='c', usecols=[0, 1] + numeric)
    try:                                         # fast path: every numeric field parses
        df = pd.read_csv(path, dtype={**{0: 'category', 1: 'category'}, **{c: np.float64 for c in numeric}}, **kw)
        ok = np.ones(len(df), dtype=bool)
        num = {c: df[c].to_numpy() for c in numeric}
    except (ValueError, TypeError):              # some row holds text in a numeric column: parse leniently, drop those rows
        df = pd.read_csv(path, dtype=str, **kw)
        ok = np.ones(len(df), dtype=bool)
        num = {}
        for c in numeric:
            v = pd.to_numeric(df[c], errors='coerce')
            ok &= ~v.isna().to_numpy() | df[c].str.lower().isin(['nan', '+nan', '-nan']).to_numpy()
            num[c] = v.to_numpy(dtype=np.float64)
        df[0], df[1] = df[0].astype('category'), df[1].astype('category')
    keep = np.nonzero(ok)[0]
    qcat, scat = df[0].cat, df[1].cat
    t = dict(qnames=np.asarray(qcat.categories, dtype=object), snames=np.asarray(scat.categories, dtype=object),
             qcode=qcat.codes.to_numpy()[keep].astype(np.int64), scode=scat.codes.to_numpy()[keep].astype(np.int64),
             idy=num[2][keep], aln=num[3][keep], qst=num[6][keep], qed=num[7][keep], score=num[11][keep])
    if ncol > 13:
        t['qln'] = num[12][keep]
    else:                                        # len_dict (:188-193): the first row of a query id decides
        _, first_row = np.unique(t['qcode'], return_index=True)
        ln = np.zeros(len(t['qnames']), dtype=np.float64)
        ln[t['qcode'][first_row]] = np.maximum(t['qst'], t['qed'])[first_row]
        t['qln'] = ln[t['qcode']]
    return t


def _fmt(x):
    return str(float(x))


def find_orth(path, coverage=.5, identity=0., norm='no', sep='|', out=None, device=0):
    out = out or sys.stdout
    lib = _lib.load()
    t = read_table(path)
    if t is None:
        return
    # ---- ids -> ranks in byte order of the id strings (what the reference's LC_ALL=C line sorts compare first)
    key = lambda x: x.encode('latin-1', 'replace')                      # noqa: E731
    names = np.array(sorted(set(t['qnames']) | set(t['snames']), key=key), dtype=object)
    for x in names:
        assert sep in x                                                  # bin/find_orth.py:173
    rank_of = {x: i for i, x in enumerate(names)}
    qmap = np.fromiter((rank_of[x] for x in t['qnames']), dtype=np.uint32, count=len(t['qnames']))
    smap = np.fromiter((rank_of[x] for x in t['snames']), dtype=np.uint32, count=len(t['snames']))
    # ---- filter (bin/find_orth.py:195-198)
    qcv = (1. + np.abs(t['qed'] - t['qst'])) / t['qln']
    keep = ~((qcv < coverage) | (t['idy'] < identity))
    qr, sr = np.ascontiguousarray(qmap[t['qcode'][keep]]), np.ascontiguousarray(smap[t['scode'][keep]])
    score, aln = t['score'][keep], t['aln'][keep]
    n = len(qr)
    if n == 0:
        return
    tax_names = [x.split(sep)[0] for x in names]
    tax_code = {}
    tax = np.fromiter((tax_code.setdefault(x, len(tax_code)) for x in tax_names), dtype=np.uint32, count=len(names))
    # ---- query groups = runs of equal query id among the kept rows (:200-206)
    starts = np.concatenate(([0], np.nonzero(qr[1:] != qr[:-1])[0] + 1, [n])).astype(np.uint64)
    # ---- normalised score (:208-226)
    if norm == 'bsr':
        _, first = np.unique(qr, return_index=True)                      # first kept row of every query id (mbsc_dict)
        mb = np.empty(len(names), dtype=np.float64)
        mb[qr[first]] = score[first]
        S = score / mb[qr]
    elif norm == 'bal':
        S = score / aln
    else:
        S = score.copy()
    # ---- classification on the device (:298-348)
    import ctypes as C
    cls = np.zeros(n, dtype=np.uint8)
    qt, stx = np.ascontiguousarray(tax[qr]), np.ascontiguousarray(tax[sr])
    S = np.ascontiguousarray(S)
    _lib.check(lib.so_orth_classify(int(device), starts.ctypes.data, len(starts) - 1, qr.ctypes.data, sr.ctypes.data,
                                    qt.ctypes.data, stx.ctypes.data, S.ctypes.data, len(tax_code), cls.ctypes.data))
    lo, hi = np.minimum(qr, sr), np.maximum(qr, sr)

    def reciprocal(a, b, s):
        """sorted candidate lines -> pairs found from both sides (get_IPO, :352-381): keys that occur exactly twice;
        score = mean of the two, the file's last pair: max of the two."""
        m = len(a)
        if m == 0:
            return a[:0], b[:0], s[:0]
        keys = (a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)
        idx = np.arange(m, dtype=np.uint32)
        keys = np.ascontiguousarray(keys)
        _lib.check(lib.so_sort_pairs_u64(int(device), keys.ctypes.data, idx.ctypes.data, m))
        s = s[idx]
        brk = np.concatenate(([0], np.nonzero(keys[1:] != keys[:-1])[0] + 1, [m]))
        cnt = np.diff(brk)
        two = np.nonzero(cnt == 2)[0]
        p = brk[two]
        sc = (s[p] + s[p + 1]) / 2.
        if len(two) and two[-1] == len(cnt) - 1:                         # the last group of the file (:374-376)
            sc[-1] = max(s[p[-1]], s[p[-1] + 1])
        k = keys[p]
        return (k >> np.uint64(32)).astype(np.uint32), (k & np.uint64(0xffffffff)).astype(np.uint32), sc

    # ---- OTs (:466-492)
    ot = cls == 2
    ot_a, ot_b, ot_s = reciprocal(lo[ot], hi[ot], S[ot])
    inots = np.zeros(len(names), dtype=bool)
    inots[ot_a] = True
    inots[ot_b] = True
    # ---- IPs (:498-541): every candidate is written in both orientations
    ip = cls == 1
    ia = np.concatenate((lo[ip], hi[ip]))
    ib = np.concatenate((hi[ip], lo[ip]))
    ip_a, ip_b, ip_s = reciprocal(ia, ib, np.concatenate((S[ip], S[ip])))
    fw = ip_a < ip_b
    ntx = len(tax_code)
    tot, cnt = np.zeros(ntx), np.zeros(ntx)
    tot_o, cnt_o = np.zeros(ntx), np.zeros(ntx)
    np.add.at(tot, tax[ip_a[fw]], ip_s[fw])                              # sequential, in file order
    np.add.at(cnt, tax[ip_a[fw]], 1.)
    sel = fw & (inots[ip_a] | inots[ip_b])
    np.add.at(tot_o, tax[ip_a[sel]], ip_s[sel])
    np.add.at(cnt_o, tax[ip_a[sel]], 1.)
    with np.errstate(divide='ignore', invalid='ignore'):
        ip_avg = np.where(cnt_o > 0, tot_o / np.where(cnt_o > 0, cnt_o, 1.), tot / np.where(cnt > 0, cnt, 1.))
    # ---- COs (:547-611)
    co = cls == 3
    co_key = (lo[co].astype(np.uint64) << np.uint64(32)) | hi[co].astype(np.uint64)
    co_best = {}
    for k, v in zip(co_key.tolist(), S[co].tolist()):                    # max over the lines of a key (:597)
        if k not in co_best or co_best[k] < v:
            co_best[k] = v
    partners = {}
    for a, b in zip(ip_a.tolist(), ip_b.tolist()):                       # IPs file order: sorted by (a, b)
        partners.setdefault(a, []).append(b)
    co_a, co_b, co_s = [], [], []
    if len(ip_a) and co_best:
        for q, s_ in zip(ot_a.tolist(), ot_b.tolist()):
            pq, ps = partners.get(q), partners.get(s_)
            if not pq and not ps:
                continue
            qips, sips = (pq or []) + [q], (ps or []) + [s_]
            visit = set()
            for x in qips:
                for y in sips:
                    # (`qip != qid or sip != sid` compares bytes with str in the reference: always true under
                    # Python 3, so the pair itself is looked up as well, :588)
                    if (x, y) in visit:
                        continue
                    visit.add((x, y))
                    v = co_best.get((x << 32) | y)
                    if v is not None:
                        co_a.append(x), co_b.append(y), co_s.append(v)
    # ---- output (:617-760)
    w = out.write
    for a, b, s in zip(ip_a[fw].tolist(), ip_b[fw].tolist(), ip_s[fw].tolist()):
        avg = ip_avg[tax[a]]
        if avg == 0:
            continue                                                     # ZeroDivisionError -> `continue` (:632-635)
        w('IP\t%s\t%s\t%s\n' % (names[a], names[b], _fmt(s / avg)))

    def same_taxon_runs(A, B, Sc):
        """get_sam_tax (:659-681) + n_co_ot (:703-722)"""
        flag, run, visit = None, [], set()
        for a, b, s in zip(A, B, Sc):
            tx = tax[a]
            if tx != flag:
                if run:
                    yield run
                flag, run = tx, [(a, b, s)]
                visit = {('id', a), ('id', b)}                           # set((qid, sid)): the two ids, not the pair
            elif (a, b) not in visit:
                run.append((a, b, s))
                visit.add((a, b))
        if run:
            yield run

    def emit(label, A, B, Sc):
        for run in same_taxon_runs(A, B, Sc):
            tot, cnt = {}, {}
            for a, b, s in run:
                k = tax[b]
                tot[k] = tot.get(k, 0.) + s if k in tot else s
                cnt[k] = cnt.get(k, 0.) + 1.
            for a, b, s in run:
                k = tax[b]
                w('%s\t%s\t%s\t%s\n' % (label, names[a], names[b], _fmt(s / (tot[k] / cnt[k]))))

    emit('OT', ot_a.tolist(), ot_b.tolist(), ot_s.tolist())
    emit('CO', co_a, co_b, co_s)


def main(argv=None):
    argv = sys.argv if argv is None else argv
    args = parse_args(argv)
    if args['-i'] == '':
        manual_print()
        raise SystemExit()
    try:
        qry, coverage, identity, norm, sep = args['-i'], float(args['-c']), float(args['-y']), args['-n'], args['-s']
    except Exception:
        manual_print()
        raise SystemExit()
    find_orth(qry, coverage, identity, norm, sep)


if __name__ == '__main__':
    main()

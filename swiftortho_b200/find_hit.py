#!/usr/bin/env python3
"""find_hit.py-compatible command line (reference: bin/find_hit.py:194-358).

    python -m swiftortho_b200.find_hit -p blastp -i qry.fsa -d db.fsa -o out.sc -e 1e-5 -s 111111 [...]

Same flag letters, defaults and output file as the reference CLI (bin/find_hit.py:227-228, 155-188).
The reference fans query slices out to `-a` CPU processes that each exec `fsearch-c`
(bin/find_hit.py:95-151); here `-a` is the number of GPUs: one worker process per GPU, each
searching a contiguous query slice against the replicated target index, part files concatenated in
ascending query order exactly like bin/find_hit.py:135-146.
"""
import multiprocessing as mp
import os
import shutil
import sys

from . import _lib
from .search import AA9, AA20, Fasta, blastp


def manual_print():
    print('Usage:')
    print('  search:')
    print('    python %s -p blastp -i qry.fsa -d db.fsa' % sys.argv[0])
    print('Parameters:')
    print('  -p: program')
    print('  -i: query sequences in fasta format')
    print('  -l: start index of query sequences')
    print('  -u: end index of query sequences')
    print('  -L: start index of reference')
    print('  -U: end index of reference')
    print('  -d: ref database')
    print('  -o: output file')
    print('  -O: write mode of output file. w: overwrite, a: append')
    print('  -s: spaced seed in comma separated format: 1111,1110,1001')
    print('  -r: reduced amino acid alphabet: aa9 (default), aa20, or comma separated groups')
    print('  -v: number of hits to show')
    print('  -e: expect value')
    print('  -m: max ratio of pseudo hits that will trigger stop')
    print('  -j: distance between start sites of two neighbor seeds')
    print('  -t: filter high frequency kmers whose counts > t')
    print('  -F: filter query sequence')
    print('  -M: bucket size of hash table')
    print('  -c: chunck size of reference (sequences per index chunk)')
    print('  -a: number of GPUs to use (the reference: number of processors)')
    print('  -T: tmpdir to store tmp file')


def parse_args(argv):
    # bin/find_hit.py:227-242 ("-k v" and "-kv" forms, unknown tokens skipped)
    args = {'-p': '', '-v': '500', '-s': '11111111', '-i': '', '-d': '', '-e': '1e-3', '-l': '-1', '-u': '-1',
            '-m': '1e-3', '-t': '-1', '-r': AA9, '-j': '1', '-F': 'T', '-o': '', '-D': '', '-O': 'wb', '-L': '-1',
            '-U': '-1', '-M': '120000000', '-c': '50000', '-a': '1', '-T': ''}
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


def _worker(job):
    (dev, qry, ref, part, exp, bv, st, ed, rstart, rend, miss, thr, step, flt, ht, chk, ssd, nr) = job
    blastp(qry, ref, part, expect=exp, v=bv, max_miss=miss, st=st, ed=ed, rst=rstart, red=rend, thr=thr, flt=flt,
           ssd=ssd, nr=nr, step=step, ht=ht, chk=chk, wrt='w', device=dev)
    return part


def slices_by_residues(fasta, start, end, parts):
    """Contiguous query ranges with balanced residue counts."""
    off = fasta.offsets
    total = int(off[end]) - int(off[start])
    cuts = [start]
    for p in range(1, parts):
        want = int(off[start]) + total * p // parts
        lo = max(cuts[-1], start)
        import numpy as np
        k = int(np.searchsorted(off[lo:end + 1], want)) + lo
        cuts.append(min(max(k, cuts[-1]), end))
    cuts.append(end)
    return [(cuts[i], cuts[i + 1]) for i in range(parts) if cuts[i + 1] > cuts[i]]


def main(argv=None):
    argv = sys.argv if argv is None else argv
    args = parse_args(argv)
    if args['-p'] != 'blastp' or args['-i'] == '' or args['-d'] == '':
        manual_print()
        raise SystemExit()
    try:
        qry, ref, exp, bv = args['-i'], args['-d'], float(args['-e']), int(args['-v'])
        start, end, rstart, rend = int(args['-l']), int(args['-u']), int(args['-L']), int(args['-U'])
        miss, thr, step, flt = float(args['-m']), int(args['-t']), int(args['-j']), args['-F'].upper()
        outfile, wrt, ht, chk = args['-o'], args['-O'], int(args['-M']), int(args['-c'])
        ssd, nr, ngpu, tmpdir = args['-s'], args['-r'], int(args['-a']), args['-T']
        tmpdir = tmpdir or outfile + '_sc_tmpdir'                     # bin/find_hit.py:263
        if nr.strip() == 'aa9':
            nr = AA9
        elif nr.strip() == 'aa20':
            nr = AA20
        chk = int(chk / (nr.count('/') + 1))                          # bin/find_hit.py:273-274
        print('chk size', chk)
    except Exception:
        manual_print()
        raise SystemExit()
    if not outfile:
        print('an output file (-o) is required')
        raise SystemExit(2)
    lib = _lib.load()
    ndev = lib.so_device_count()
    if ndev <= 0:
        raise _lib.SoError('no CUDA device visible: swiftortho_b200 has no CPU fallback')
    Q = Fasta(qry)
    N = len(Q)
    Start = 0 if start < 0 else start
    End = N if end < 0 else min(end, N)
    # under torchrun every rank runs this CLI: shard by RANK / WORLD_SIZE
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', str(rank)))
    os.makedirs(tmpdir, exist_ok=True)
    tmp_name = outfile.split(os.sep)[-1]
    if world > 1:
        run_sharded(Q, Start, End, rank, world, outfile, tmpdir, wrt,
                    lambda s, e, part: _worker((local % ndev, qry, ref, part, exp, bv, s, e, rstart, rend, miss, thr,
                                                step, flt, ht, chk, ssd, nr)))
        return 0
    ngpu = max(1, min(ngpu, ndev))
    sl = slices_by_residues(Q, Start, End, ngpu) if End > Start else []
    jobs = []
    for k, (s, e) in enumerate(sl):
        part = '%s/%s.%012d' % (tmpdir, tmp_name, s)
        jobs.append((k % ndev, qry, ref, part, exp, bv, s, e, rstart, rend, miss, thr, step, flt, ht, chk, ssd, nr))
    if len(jobs) <= 1:
        parts = [_worker(j) for j in jobs]
    else:
        with mp.get_context('spawn').Pool(len(jobs)) as pool:
            parts = pool.map(_worker, jobs)
    _concat(outfile, parts, wrt)
    shutil.rmtree(tmpdir, ignore_errors=True)                          # bin/find_hit.py:354-355
    return 0


def run_sharded(Q, Start, End, rank, world, outfile, tmpdir, wrt, worker):
    """One rank of a multi-process run: search this rank's query slice into a part file; rank 0 waits for
    every part and concatenates them in ascending query order (bin/find_hit.py:135-146).  No collective
    is needed on this path: the exchange is the part files, exactly like the reference."""
    import time
    os.makedirs(tmpdir, exist_ok=True)
    tmp_name = outfile.split(os.sep)[-1]
    sl = slices_by_residues(Q, Start, End, world) if End > Start else []
    mine = sl[rank] if rank < len(sl) else None
    if mine:
        part = '%s/%s.%012d' % (tmpdir, tmp_name, mine[0])
        worker(mine[0], mine[1], part)
        open(part + '.done', 'w').close()
    if rank == 0:
        for s in sl:
            while not os.path.exists('%s/%s.%012d.done' % (tmpdir, tmp_name, s[0])):
                time.sleep(0.05)
        _concat(outfile, ['%s/%s.%012d' % (tmpdir, tmp_name, s[0]) for s in sl], wrt)
        shutil.rmtree(tmpdir, ignore_errors=True)
    return sl


def _concat(outfile, parts, wrt):
    mode = 'ab' if 'a' in wrt else 'wb'
    with open(outfile, mode) as out:
        for p in parts:
            if not os.path.isfile(p):
                continue                                               # bin/find_hit.py:136-138
            with open(p, 'rb') as f:
                shutil.copyfileobj(f, out, 1 << 24)
            os.remove(p)


if __name__ == '__main__':
    main()

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
    os.makedirs(tmpdir, exist_ok=True)
    params = (exp, bv, rstart, rend, miss, thr, step, flt, ht, chk, ssd, nr)
    max_chr = int(os.environ.get('SO_MAX_CHR', '4200000000'))          # bin/find_hit.py:286 (env: test hook)
    if os.path.getsize(ref) < max_chr:
        search_fanout(qry, ref, outfile, start, end, ngpu, ndev, tmpdir, params)
    else:
        search_split_reference(qry, ref, outfile, start, end, ngpu, ndev, tmpdir, params, max_chr)
    if _rank_world()[0] == 0:
        shutil.rmtree(tmpdir, ignore_errors=True)                      # bin/find_hit.py:354-355
    return 0


def _rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def search_fanout(qry, ref, outfile, start, end, ngpu, ndev, tmpdir, params):
    """bin/find_hit.py:95-151 (`blastp`): query slices -> one worker per GPU -> part files -> OUT.
    The reference always removes OUT first (`rm -f`, :126) and `cat`s the parts in ascending start order; `-O`
    only reaches the part files."""
    (exp, bv, rstart, rend, miss, thr, step, flt, ht, chk, ssd, nr) = params
    os.makedirs(tmpdir, exist_ok=True)
    Q = Fasta(qry)
    N = len(Q)
    Start = 0 if start < 0 else start
    End = N if end < 0 else min(end, N)
    rank, world = _rank_world()
    local = int(os.environ.get('LOCAL_RANK', str(rank)))
    tmp_name = outfile.split(os.sep)[-1]
    if world > 1:
        # under torchrun every rank runs this CLI: shard by RANK / WORLD_SIZE
        run_sharded(Q, Start, End, rank, world, outfile, tmpdir, 'wb',
                    lambda s, e, part: _worker((local % ndev, qry, ref, part, exp, bv, s, e, rstart, rend, miss, thr,
                                                step, flt, ht, chk, ssd, nr)))
        return
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
    _concat(outfile, parts)


def split_reference(ref, ref_dir, max_chr):
    """bin/find_hit.py:303-340: the reference FASTA is cut into parts; a part is closed when the characters
    (header + sequence) written to it exceed max_chr, records are rewritten as `>header\nsequence\n`
    (fasta_parse, :23-36: header = line without '>' and newline, sequence lines stripped and joined).
    Yields the part paths in order (the same file name is reused by the reference; here every part keeps its own
    file so the parts can be listed)."""
    os.makedirs(ref_dir, exist_ok=True)

    def records(f):
        head, seq = '', []
        for line in f:
            if line.startswith('>'):
                if seq:
                    yield head, ''.join(seq)
                head, seq = line[1:-1], []
            else:
                seq.append(line.strip())
        if seq:
            yield head, ''.join(seq)

    parts, flag_chr, idx = [], 0, 0
    path = '%s/ref.%d.fsa' % (ref_dir, idx)
    out = open(path, 'w')
    with open(ref, 'r') as f:
        for hd, sq in records(f):
            l_chr = len(hd) + len(sq)
            if flag_chr > max_chr:
                out.close()
                parts.append(path)
                idx += 1
                path = '%s/ref.%d.fsa' % (ref_dir, idx)
                out = open(path, 'w')
                flag_chr = l_chr
            flag_chr += l_chr
            out.write('>%s\n%s\n' % (hd, sq))
    out.close()
    if os.path.getsize(path) > 0:
        parts.append(path)
    return parts


def merge_part_tables(scs, bv, outfile):
    """bin/find_hit.py:343-345, the reference's own pipeline (GNU sort + awk are part of the reference's
    behaviour here): merge the per-part tables on column 15 (query ordinal) then -bit, keep the first `-v` rows
    of every query id."""
    import subprocess
    cmd = "sort -m -k15,15n -k12,12nr %s | awk '{if(c[$1]<%s) print $0;c[$1]+=1}' > %s" % (
        ' '.join(_shq(p) for p in scs), bv, _shq(outfile))
    subprocess.check_call(cmd, shell=True)


def _shq(p):
    return "'" + str(p).replace("'", "'\\''") + "'"


def search_split_reference(qry, ref, outfile, start, end, ngpu, ndev, tmpdir, params, max_chr):
    """bin/find_hit.py:296-351: references of max_chr bytes or more are searched part by part (every part is its
    own database: D, chunking and target ordinals restart in each part) and the part tables are merged."""
    ref_dir = '%s_parts' % ref
    rank, world = _rank_world()
    if rank == 0:
        parts = split_reference(ref, ref_dir, max_chr)
        with open(os.path.join(ref_dir, 'parts.list.tmp'), 'w') as f:
            f.write('\n'.join(parts))
        os.replace(os.path.join(ref_dir, 'parts.list.tmp'), os.path.join(ref_dir, 'parts.list.%s' % _run_id()))
    else:
        parts = _wait_for(os.path.join(ref_dir, 'parts.list.%s' % _run_id()), lambda p: open(p).read().split('\n'))
    scs = []
    for idx, part in enumerate(parts):
        sc = '%s/%d.sc' % (ref_dir, idx)
        search_fanout(qry, part, sc, start, end, ngpu, ndev, os.path.join(tmpdir, 'part%d' % idx), params)
        scs.append(sc)
    if rank == 0:
        # `*.sc` in the reference's command expands in lexicographic order
        merge_part_tables(sorted(scs), params[1], outfile)
        if not os.environ.get('SO_KEEP_PARTS'):                        # debugging aid
            shutil.rmtree(ref_dir, ignore_errors=True)


def _run_id():
    """One token per launch, the same in every rank of a node: SO_RUN_ID, the torchrun run id, else the launcher's
    pid (all ranks of one node are children of one launcher process)."""
    rid = os.environ.get('SO_RUN_ID') or os.environ.get('TORCHELASTIC_RUN_ID')
    if not rid or rid == 'none':
        rid = 'ppid%d' % os.getppid()
    return rid


def _wait_for(path, read, other_err=None, timeout=None):
    import time
    timeout = float(os.environ.get('SO_SHARD_TIMEOUT', '86400')) if timeout is None else timeout
    t0 = time.time()
    while True:
        if other_err and os.path.exists(other_err):
            raise RuntimeError('another rank failed: %s' % open(other_err).read().strip())
        if os.path.exists(path):
            return read(path)
        if time.time() - t0 > timeout:
            raise TimeoutError('timed out after %.0f s waiting for %s' % (timeout, path))
        time.sleep(0.05)


def run_sharded(Q, Start, End, rank, world, outfile, tmpdir, wrt, worker):
    """One rank of a multi-process run: search this rank's query slice into a part file; rank 0 waits for
    every part and concatenates them in ascending query order (bin/find_hit.py:135-146).  No collective
    is needed on this path: the exchange is the part files, exactly like the reference.
    Every marker carries the launch's run id, so files left behind by a crashed run are never mistaken for this
    run's; a part is written under a private name and renamed when complete; a rank that fails leaves an error
    marker that makes rank 0 raise instead of waiting; the wait has a timeout (SO_SHARD_TIMEOUT seconds)."""
    os.makedirs(tmpdir, exist_ok=True)
    rid = _run_id()
    tmp_name = outfile.split(os.sep)[-1]
    sl = slices_by_residues(Q, Start, End, world) if End > Start else []
    mine = sl[rank] if rank < len(sl) else None
    name = lambda s: '%s/%s.%012d' % (tmpdir, tmp_name, s)           # noqa: E731
    err = '%s/%s.%s.err' % (tmpdir, tmp_name, rid)
    if mine:
        part = name(mine[0])
        try:
            worker(mine[0], mine[1], part + '.%s.tmp' % rid)
            os.replace(part + '.%s.tmp' % rid, part + '.%s' % rid)
            open(part + '.%s.done' % rid, 'w').close()
        except BaseException as e:                                     # noqa: BLE001
            with open(err, 'w') as f:
                f.write('rank %d: %r' % (rank, e))
            raise
    if rank == 0:
        for s in sl:
            _wait_for(name(s[0]) + '.%s.done' % rid, lambda p: None, other_err=err)
        _concat(outfile, [name(s[0]) + '.%s' % rid for s in sl])
        for s in sl:
            try:
                os.remove(name(s[0]) + '.%s.done' % rid)
            except OSError:
                pass
    return sl


def _concat(outfile, parts, wrt='wb'):
    # the reference removes OUT first whatever -O says (bin/find_hit.py:126) and appends the parts in order
    with open(outfile, 'wb') as out:
        for p in parts:
            if not os.path.isfile(p):
                continue                                               # bin/find_hit.py:136-138
            with open(p, 'rb') as f:
                shutil.copyfileobj(f, out, 1 << 24)
            os.remove(p)


if __name__ == '__main__':
    main()

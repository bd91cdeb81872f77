#!/usr/bin/env python3
"""Golden vector of the large-reference path (bin/find_hit.py:296-351), generated from the reference itself:

  * the reference FASTA is cut into parts by the reference's own split loop: the lines bin/find_hit.py:304-346 are
    exec'ed here with max_chr lowered (their `blastp(...)` call is replaced by a hook that records the part);
  * every part is searched by the reference core (oracle/ref_shim/run_reference.py, lib/fsearch.py executed under
    CPython) with the flags find_hit.py passes (bin/find_hit.py:119-120);
  * the part tables are merged by the reference's own command (bin/find_hit.py:350-351).

    python tests/golden/make_split_golden.py        # needs /root/reference (this container only)

Output: tests/golden/synth60_split.sc (reference FASTA = tests/golden/synth60.fsa, SO_MAX_CHR = 4000, -v 3).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('SWIFTORTHO_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_shim'))
sys.setrecursionlimit(100000)
import run_reference  # noqa: E402

MAX_CHR = 4000
BV = '3'
AA9 = 'AST,CFILMVY,DN,EQ,G,H,KR,P,W'


def reference_lines(path, first, last):
    with open(path) as f:
        lines = f.readlines()
    return lines[first - 1:last]


def main():
    src = os.path.join(REF, 'bin', 'find_hit.py')
    # fasta_parse (bin/find_hit.py:23-36) and the split loop (:304-346), verbatim from the reference at run time
    ns = {}
    exec(''.join(reference_lines(src, 23, 36)), ns)
    work = tempfile.mkdtemp()
    REFFSA = os.path.join(work, 'ref.fsa')
    shutil.copy(os.path.join(HERE, 'synth60.fsa'), REFFSA)
    NQ = sum(1 for line in open(REFFSA) if line.startswith('>'))
    parts = []

    def blastp(start, end):                      # the hook: keep a copy of the part the reference would search now
        k = len(parts)
        p = os.path.join(work, 'part%d.fsa' % k)
        if not ns['_o'].closed:
            ns['_o'].flush()
        shutil.copy(ns['ref'], p)
        parts.append((p, ns['outfile']))

    ns.update(dict(os=os, ref=REFFSA, outfile=os.path.join(work, 'OUT.sc'), max_chr=MAX_CHR, blastp=blastp, start=-1, end=-1))
    body = reference_lines(src, 304, 346)
    indent = len(body[0]) - len(body[0].lstrip())
    code = ''.join(l[indent:] if l.strip() else l for l in body)
    # the reference closes the part file before the last blastp call only implicitly (`_o.close()` at :336)
    exec(code, ns)
    print('parts:', [(os.path.getsize(p), o) for p, o in parts])
    assert len(parts) >= 2
    ref_dir = '%s_parts' % REFFSA
    for p, sc in parts:
        tmp = tempfile.mkdtemp()
        # find_hit.py always passes the query range explicitly (-l 0 -u N for one worker, bin/find_hit.py:107-120); with -u -1
        # the core would stop at query len(part) (lib/fsearch.py:2980-2981)
        argv = ['-p', 'blastp', '-i', REFFSA, '-d', p, '-e', '1e-5', '-v', BV, '-l', '0', '-u', str(NQ), '-L', '-1', '-U', '-1',
                '-m', '1e-3', '-t', '-1', '-j', '1', '-F', 'T', '-D', '', '-O', 'wb', '-M', '1000003', '-c', '50000',
                '-s', '111111', '-r', AA9, '-o', sc, '-T', tmp]
        print('reference search against', os.path.basename(p), flush=True)
        run_reference.run_entry_point(argv)
        shutil.rmtree(tmp, ignore_errors=True)
    out = os.path.join(HERE, 'synth60_split.sc')
    # bin/find_hit.py:350-351, verbatim command (without the trailing rm)
    cmd = "sort -m -k15,15n -k12,12nr %s/*.sc | awk '{if(c[$1]<%s) print $0;c[$1]+=1}'  > %s" % (ref_dir, BV, out)
    subprocess.check_call(cmd, shell=True, env=dict(os.environ, LC_ALL='C'))
    print('rows:', sum(1 for _ in open(out, 'rb')))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
    main()

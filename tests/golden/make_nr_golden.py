#!/usr/bin/env python3
"""Golden vectors of the redundancy pre-filter row, generated from the reference scripts:

  * synth60_dup.fsa        synth60.fsa with 9 exact duplicates (other ids, one with a description, other line width)
  * synth60_dup_nr.fsa     output of the reference's scripts/nr_flt.py.  Biopython is not installed here, so the script
                           runs against a 20-line stand-in for Bio.SeqIO.parse(..., 'fasta') (id = title up to the
                           first blank, seq = lines right-stripped and joined): the SeqIO details are therefore
                           restated, everything else is the reference's code
  * synth60_dup_nr.sc      reference search of the nr set (oracle/ref_shim, flags of scripts/run_all_fast.py:117 scaled:
                           -e 1e-5 -s 111111 -m 5e-2 -M 1000003)
  * synth60_dup_full.sc    output of the reference's scripts/nr2full.py on that table (pure Python, run as is)

    python tests/golden/make_nr_golden.py        # needs /root/reference (this container only)
"""
import os
import random
import subprocess
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('SWIFTORTHO_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_shim'))
sys.setrecursionlimit(100000)
AA9 = 'AST,CFILMVY,DN,EQ,G,H,KR,P,W'

STUB = '''
class _Rec(object):
    def __init__(self, title, seq):
        self.description = title
        self.id = title.split(None, 1)[0] if title.split() else ''
        self.seq = seq
def parse(handle, fmt):
    f = open(handle) if isinstance(handle, str) else handle
    title, lines = None, []
    for line in f:
        if line.startswith('>'):
            if title is not None:
                yield _Rec(title, ''.join(lines).replace(' ', '').replace('\\r', ''))
            title, lines = line[1:].rstrip(), []
        elif title is not None:
            lines.append(line.rstrip())
    if title is not None:
        yield _Rec(title, ''.join(lines).replace(' ', '').replace('\\r', ''))
'''


def main():
    rnd = random.Random(11)
    recs = []
    cur = None
    for line in open(os.path.join(HERE, 'synth60.fsa')):
        if line.startswith('>'):
            cur = [line[1:].strip(), []]
            recs.append(cur)
        else:
            cur[1].append(line.strip())
    recs = [(h, ''.join(s)) for h, s in recs]
    out = list(recs)
    for k, src in enumerate(rnd.sample(range(len(recs)), 9)):
        h, s = recs[src]
        tax = 'T%03d' % rnd.randrange(6)
        nh = '%s|dup%02d' % (tax, k) + (' copy of %s' % h if k == 3 else '')
        out.insert(rnd.randrange(len(out) + 1), (nh, s))
    dup = os.path.join(HERE, 'synth60_dup.fsa')
    with open(dup, 'w') as f:
        for k, (h, s) in enumerate(out):
            w = 60 if k % 2 == 0 else 47
            f.write('>%s\n' % h)
            for i in range(0, len(s), w):
                f.write(s[i:i + w] + '\n')
    # scripts/nr_flt.py with the Bio.SeqIO stand-in
    d = tempfile.mkdtemp()
    os.makedirs(os.path.join(d, 'Bio'))
    open(os.path.join(d, 'Bio', '__init__.py'), 'w').close()
    open(os.path.join(d, 'Bio', 'SeqIO.py'), 'w').write(STUB)
    nr = os.path.join(HERE, 'synth60_dup_nr.fsa')
    r = subprocess.run([sys.executable, os.path.join(REF, 'scripts', 'nr_flt.py'), dup], stdout=subprocess.PIPE, check=True,
                       env=dict(os.environ, PYTHONPATH=d))
    open(nr, 'wb').write(r.stdout)
    print('nr records:', r.stdout.count(b'>'), 'of', len(out))
    # reference search of the nr set
    import run_reference
    sc = os.path.join(HERE, 'synth60_dup_nr.sc')
    tmp = tempfile.mkdtemp()
    run_reference.run_entry_point(['-p', 'blastp', '-i', nr, '-d', nr, '-e', '1e-5', '-v', '500', '-l', '-1', '-u', '-1', '-L', '-1',
                                   '-U', '-1', '-m', '5e-2', '-t', '-1', '-j', '1', '-F', 'T', '-D', '', '-O', 'wb', '-M', '1000003',
                                   '-c', '50000', '-s', '111111', '-r', AA9, '-o', sc, '-T', tmp])
    # scripts/nr2full.py as is
    r = subprocess.run([sys.executable, os.path.join(REF, 'scripts', 'nr2full.py'), sc], stdout=subprocess.PIPE, check=True)
    open(os.path.join(HERE, 'synth60_dup_full.sc'), 'wb').write(r.stdout)
    print('rows nr:', sum(1 for _ in open(sc)), 'rows full:', r.stdout.count(b'\n'))


if __name__ == '__main__':
    main()

"""Groundwork for SURVEY.md section 8(f) rank 1 (the bin/find_orth.py consumer, not built yet): runs the
reference's own find_orth.py (plain Python 3) on a golden hit table and stores its output as a fixture, so a later
round has a pinned answer for the orthology-inference row.

    python tests/golden/make_orth_golden.py            # needs /root/reference (this container only)
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('SWIFTORTHO_REFERENCE', '/root/reference')


def run(sc_name, out_name, extra=()):
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(HERE, sc_name), d)
        r = subprocess.run([sys.executable, os.path.join(REF, 'bin', 'find_orth.py'), '-i', sc_name,
                            *(('-c', '0.5', '-y', '0') if '-c' not in extra else ()), *extra], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, LC_ALL='C'))
        assert r.returncode == 0, r.stderr.decode()
        with open(os.path.join(HERE, out_name), 'wb') as f:
            f.write(r.stdout)
    return r.stdout.count(b'\n')


def big_table():
    """synth600.sc: a larger hit table (600 proteins, 8 taxa, many in-paralogs) from the CPU oracle port
    (oracle/fsearch_oracle.cpp, itself pinned to the reference core): input of the larger find_orth goldens."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), 'tests'))
    import numpy as np
    from conftest import Oracle
    from swiftortho_b200 import synth
    h, s = synth.generate(600, 8, seed=20261031, max_len=180)
    # more in-paralogs than the generator's 5 %: every 6th protein gets a close copy in its own taxon
    rng = np.random.Generator(np.random.PCG64(5))
    hh, ss = [], []
    for k, (a, b) in enumerate(zip(h, s)):
        hh.append(a), ss.append(b)
        if k % 6 == 0:
            c = b.copy()
            m = rng.random(len(c)) < 0.04
            c[m] = synth.AA[rng.integers(0, 20, size=int(m.sum()))]
            hh.append(a + 'p'), ss.append(c)
    fsa = os.path.join(HERE, 'synth600.fsa')
    with open(fsa, 'wb') as f:
        f.write(synth.to_fasta_bytes(hh, ss))
    Oracle().blastp(fsa, fsa, os.path.join(HERE, 'synth600.sc'), {'-e': '1e-5', '-j': '1', '-M': '1000003', '-s': '111111'})
    os.remove(fsa)


if __name__ == '__main__':
    print('synth60.orth rows:', run('synth60.sc', 'synth60.orth'))
    print('synth60_bsr.orth rows:', run('synth60.sc', 'synth60_bsr.orth', ('-n', 'bsr')))
    big_table()
    print('synth600.orth rows:', run('synth600.sc', 'synth600.orth'))
    print('synth600_bal.orth rows:', run('synth600.sc', 'synth600_bal.orth', ('-n', 'bal', '-c', '0.3', '-y', '25')))
    print('synth600_bsr.orth rows:', run('synth600.sc', 'synth600_bsr.orth', ('-n', 'bsr', '-c', '0.6')))

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
        r = subprocess.run([sys.executable, os.path.join(REF, 'bin', 'find_orth.py'), '-i', sc_name, '-c', '0.5', '-y', '0',
                            *extra], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, LC_ALL='C'))
        assert r.returncode == 0, r.stderr.decode()
        with open(os.path.join(HERE, out_name), 'wb') as f:
            f.write(r.stdout)
    return r.stdout.count(b'\n')


if __name__ == '__main__':
    print('synth60.orth rows:', run('synth60.sc', 'synth60.orth'))
    print('synth60_bsr.orth rows:', run('synth60.sc', 'synth60_bsr.orth', ('-n', 'bsr')))

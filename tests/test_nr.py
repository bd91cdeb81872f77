"""Redundancy pre-filter row: scripts/nr_flt.py / scripts/nr2full.py of the reference.
Goldens: tests/golden/make_nr_golden.py (reference scripts run in the build container)."""
import io
import os

import pytest

from conftest import GOLDEN


def test_nr2full_reference_golden():
    """scripts/nr2full.py:14-44 on the reference's own table of the nr set: byte-identical expansion."""
    from swiftortho_b200 import nr
    out = io.StringIO()
    nr.nr2full(os.path.join(GOLDEN, 'synth60_dup_nr.sc'), out)
    assert out.getvalue() == open(os.path.join(GOLDEN, 'synth60_dup_full.sc')).read()


def test_nr_parse_matches_golden_grouping():
    """The host half of nr_flt (record parsing + exact grouping, here with a plain dict standing in for the device
    hash) reproduces the reference script's nr FASTA."""
    from swiftortho_b200 import nr
    ids, seqs = nr.parse_fasta(os.path.join(GOLDEN, 'synth60_dup.fsa'))
    groups = {}
    for i, s in enumerate(seqs):
        groups.setdefault(s, []).append(i)
    text = ''.join('>' + ';;;'.join(ids[i] for i in g) + '\n' + s + '\n' for s, g in groups.items())
    assert text == open(os.path.join(GOLDEN, 'synth60_dup_nr.fsa')).read()
    assert len(groups) == 60 and len(ids) == 69


@pytest.mark.gpu
def test_nr_flt_device_hash_and_search(tmp_path):
    """nr_flt through so_seq_hash (one warp per sequence) gives the reference's nr FASTA; the search of that set and
    the re-expansion give the reference's full table (scripts/run_all_fast.py:110-119 end to end)."""
    from swiftortho_b200 import build, nr, search
    build.build()
    out = io.StringIO()
    nr.nr_flt(os.path.join(GOLDEN, 'synth60_dup.fsa'), out)
    assert out.getvalue() == open(os.path.join(GOLDEN, 'synth60_dup_nr.fsa')).read()
    nrf = str(tmp_path / 'nr.fsa')
    open(nrf, 'w').write(out.getvalue())
    sc = str(tmp_path / 'nr.sc')
    search.blastp(nrf, nrf, sc, expect=1e-5, max_miss=5e-2, step=1, ht=1000003, chk=50000, ssd='111111')
    assert open(sc, 'rb').read() == open(os.path.join(GOLDEN, 'synth60_dup_nr.sc'), 'rb').read()
    full = io.StringIO()
    nr.nr2full(sc, full)
    assert full.getvalue() == open(os.path.join(GOLDEN, 'synth60_dup_full.sc')).read()
    # a collision-prone case: many sequences, few distinct ones, equal lengths
    import random
    rnd = random.Random(5)
    base = [''.join(rnd.choice('ACDEFGHIKLMNPQRSTVWY') for _ in range(rnd.choice((31, 64, 65, 200)))) for _ in range(50)]
    seqs = [rnd.choice(base) for _ in range(5000)] + ['', '']
    want = {}
    for i, s in enumerate(seqs):
        want.setdefault(s, []).append(i)
    assert nr.duplicate_groups(seqs) == list(want.values())

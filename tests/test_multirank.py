"""world_size-2 CPU tests (gloo) of the N>1 host logic: query sharding + ordered merge of the part
files (the path's only exchange), and bench.py's cross-rank reduction."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class _FakeFasta:
    offsets = np.concatenate([[0], np.cumsum(np.full(40, 100))]).astype(np.uint64)


def _rank_main(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import bench
    from swiftortho_b200 import find_hit
    # (1) reduction used by bench.py: times -> max, work -> sum
    vals = [1.0 + rank, 2.0 - rank] + [10.0 * (rank + 1)] * 16
    out = bench.reduce_over_ranks(vals, dist, 'cpu')
    assert out[0] == 2.0 and out[1] == 2.0
    assert out[2] == 30.0 and out[5] == 20.0 and out[12] == 30.0
    # (2) sharded search: each rank writes its slice, rank 0 merges in ascending query order
    outfile = os.path.join(tmp, 'merged.sc')

    def worker(s, e, part):
        with open(part, 'w') as f:
            for q in range(s, e):
                f.write('q%d\trank%d\n' % (q, rank))
    sl = find_hit.run_sharded(_FakeFasta, 0, 40, rank, world, outfile, os.path.join(tmp, 'tmpdir'), 'wb', worker)
    dist.barrier()
    if rank == 0:
        lines = open(outfile).read().split('\n')[:-1]
        assert [l.split('\t')[0] for l in lines] == ['q%d' % q for q in range(40)]
        assert sl == [(0, 20), (20, 40)]
        assert lines[0].endswith('rank0') and lines[-1].endswith('rank1')
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def _fake_fasta(n=40):
    class F:
        offsets = np.concatenate([[0], np.cumsum(np.full(n, 100))]).astype(np.uint64)
    return F


def test_run_sharded_ignores_stale_markers_and_propagates_errors(tmp_path, monkeypatch):
    """markers left behind by a crashed run must not be mistaken for this run's (every marker carries the
    run id), a failing rank leaves an error marker that makes rank 0 raise, and the wait is bounded."""
    from swiftortho_b200 import find_hit
    tmp = str(tmp_path / 'tmpdir')
    os.makedirs(tmp)
    out = str(tmp_path / 'merged.sc')
    # leftovers of an older run with another id: a complete-looking part + done marker of rank 1's slice
    for name in ('merged.sc.000000000020', 'merged.sc.000000000020.old.done', 'merged.sc.000000000020.old'):
        open(os.path.join(tmp, name), 'w').write('STALE\n')
    monkeypatch.setenv('SO_RUN_ID', 'new')
    monkeypatch.setenv('SO_SHARD_TIMEOUT', '1')

    def worker(s, e, part):
        with open(part, 'w') as f:
            for q in range(s, e):
                f.write('q%d\n' % q)
    # rank 0 alone: rank 1 never reports -> bounded wait, no use of the stale files
    with pytest.raises(TimeoutError):
        find_hit.run_sharded(_fake_fasta(), 0, 40, 0, 2, out, tmp, 'wb', worker)
    assert not os.path.exists(out) or 'STALE' not in open(out).read()
    # rank 1 fails: error marker, rank 0 raises instead of waiting
    def bad(s, e, part):
        raise RuntimeError('boom')
    with pytest.raises(RuntimeError):
        find_hit.run_sharded(_fake_fasta(), 0, 40, 1, 2, out, tmp, 'wb', bad)
    monkeypatch.setenv('SO_SHARD_TIMEOUT', '30')
    with pytest.raises(RuntimeError, match='another rank failed'):
        find_hit.run_sharded(_fake_fasta(), 0, 40, 0, 2, out, tmp, 'wb', worker)


def test_split_reference_and_merge_reference_golden(tmp_path):
    """Large-reference path, host half (bin/find_hit.py:296-351): the split rule and the merge command reproduce the
    golden table made by the reference's own split loop, the reference search per part and the reference's merge
    (tests/golden/make_split_golden.py); the per-part tables come from the CPU oracle here (the `-m gpu` twin of this
    test, test_find_hit_cli_split_reference, runs the whole CLI on the device)."""
    import shutil
    from conftest import GOLDEN, Oracle
    from swiftortho_b200 import find_hit
    ref = str(tmp_path / 'ref.fsa')
    shutil.copy(os.path.join(GOLDEN, 'synth60.fsa'), ref)
    parts = find_hit.split_reference(ref, ref + '_parts', 4000)
    assert [os.path.getsize(p) for p in parts] == [4308, 3914, 3872, 1100]
    o = Oracle()
    scs = []
    for i, p in enumerate(parts):
        sc = '%s_parts/%d.sc' % (ref, i)
        o.blastp(ref, p, sc, {'-e': '1e-5', '-j': '1', '-M': '1000003', '-c': '50000', '-s': '111111', '-v': '3', '-l': '0',
                              '-u': '60'})
        scs.append(sc)
    out = str(tmp_path / 'out.sc')
    os.environ['LC_ALL'] = 'C'
    find_hit.merge_part_tables(sorted(scs), 3, out)
    assert open(out, 'rb').read() == open(os.path.join(GOLDEN, 'synth60_split.sc'), 'rb').read()

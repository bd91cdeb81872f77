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

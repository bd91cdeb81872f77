#!/usr/bin/env python3
"""Benchmark of the all-vs-all homology search hot path (BASELINE.json metric: proteins/s and gapped
GCUPS) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload = BASELINE config 2: synthetic 100 000 proteins (~350 aa, 20 taxa) all-vs-all,
`-p blastp -e 1e-5 -s 111111 -r aa9 -M 120000000 -c 50000 -j 1`.
A STEP is one pass of the whole hot path (seed lookup, diagonal grouping, chained X-drop scoring,
candidate selection, banded gapped DP + traceback, e-value filter) for one block of `--block`
queries per GPU against the complete 100 000-protein target index (2 chunks, resident in HBM).
`value` = queries/s with the query block already in HBM; `e2e` = the same block pushed through the
public host API from host buffers (so_set_queries H2D + so_search + 16-column text written to
/dev/shm), every step.  Every step uses a different query block and touches GBs of seed-hit
buffers, i.e. far more than the 126 MB L2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AA9 = 'AST,CFILMVY,DN,EQ,G,H,KR,P,W'
FLAGS = dict(ssd='111111', nr=AA9, ht=120000000, step=1, expect=1e-5, v=500, max_miss=1e-3, thr=-1, flt='T', chk=50000)
CACHE = os.environ.get('SWIFTORTHO_BENCH_CACHE', '/tmp/swiftortho_b200_bench')
# BASELINE.json configs: id -> (proteins, taxa, seed pattern, e-value, description)
CONFIGS = {
    2: (100000, 20, '111111', 1e-5, 'synthetic 100k proteins (~350 aa mean, 20 taxa), seed 111111'),
    3: (1000000, 200, '111111', 1e-5, 'synthetic 1M proteins (200 taxa), seed 111111'),
    4: (250000, 50, '1110100111', 1e-3, 'synthetic 250k proteins (50 taxa), spaced seed 1110100111, -e 1e-3'),
    5: (100000, 20, '111111', 1e-5, 'synthetic 100k proteins, long-tailed lengths 50-5000 aa, seed 111111'),
}
METRIC = 'all-vs-all proteins/sec (find_hit blastp, config 2: 100k synthetic proteins, seed 111111)'


def flags_of(args):
    n, taxa, ssd, ev, _ = CONFIGS[args.config]
    return dict(FLAGS, ssd=ssd, expect=ev)


def metric_of(args):
    if args.config == 2:
        return METRIC
    return 'all-vs-all proteins/sec (find_hit blastp, config %d: %s)' % (args.config, CONFIGS[args.config][4])


def dataset(n, taxa, rank=0, wait=True, config=2):
    """FASTA of a BASELINE config (generated once per box, deterministic)."""
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, 'c%d_%d_%d.fsa' % (config, n, taxa))
    done = path + '.done'
    if rank == 0 and not os.path.exists(done):
        from swiftortho_b200 import synth
        synth.write_config(path + '.tmp', config, n=n, taxa=taxa)
        os.replace(path + '.tmp', path)
        open(done, 'w').close()
    while wait and not os.path.exists(done):
        time.sleep(0.2)
    return path


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        self.rows, self.gpu, self.p = [], gpu, None

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q,
                                       '--format=csv,noheader,nounits', '-lms', '200'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.p:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ----------------------------------------------------------------------------------------- CPU arm
def _oracle_lib():
    import ctypes as C
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(ROOT, 'oracle', '_build', 'liboracle.so'))
    L.orc_session_open.restype = C.c_void_p
    L.orc_session_open.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_longlong, C.c_double] + [C.c_longlong] * 3 + [
        C.c_char_p] * 3 + [C.c_longlong] * 3
    L.orc_session_search.restype = C.c_double
    L.orc_session_search.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_char_p, C.POINTER(C.c_longlong)]
    L.orc_session_build_seconds.restype = C.c_double
    L.orc_session_build_seconds.argtypes = [C.c_void_p]
    return L


_SESSION = None


def _cpu_worker(job):
    import ctypes as C
    L, h = _SESSION
    q0, q1, out = job
    st = (C.c_longlong * 7)()
    t = L.orc_session_search(h, q0, q1, out.encode(), st)
    return t, list(st)


def cpu_arm(fasta, n_total, steps, warmup, per_worker, first_query=0, flags=None, parity_out=None):
    """The oracle port (oracle/fsearch_oracle.cpp = CPU restatement of lib/fsearch.py) on all host
    cores: the index is built once, then forked workers each search their own query window like the
    reference's `find_hit.py -a <cores>` slices.  Returns per-step wall times and counters."""
    global _SESSION
    import multiprocessing as mp
    L = _oracle_lib()
    f = flags or FLAGS
    h = L.orc_session_open(fasta.encode(), fasta.encode(), f['expect'], f['v'], f['max_miss'], -1, -1, f['thr'],
                           f['flt'].encode(), f['ssd'].encode(), f['nr'].encode(), f['step'], f['ht'], f['chk'])
    assert h, 'oracle session failed'
    build_s = L.orc_session_build_seconds(h)
    _SESSION = (L, h)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    workers = max(1, min(cores, 64))
    times, counters, parity = [], [0] * 7, None
    with mp.get_context('fork').Pool(workers) as pool:
        q = first_query
        for s in range(warmup + steps):
            jobs = []
            for w in range(workers):
                a = q % max(1, n_total - per_worker)
                # the rows of the first window of the first timed step are kept: the GPU arm checks its own rows
                # for the same queries against them (parity_check, outside every timed region)
                keep = parity_out if (parity_out and s == warmup and w == 0) else ''
                if keep:
                    open(keep, 'wb').close()
                    parity = (a, a + per_worker)
                jobs.append((a, a + per_worker, keep))
                q += per_worker
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                for _, st in res:
                    counters = [a + b for a, b in zip(counters, st)]
    return dict(times=times, workers=workers, per_step=workers * per_worker, build_s=build_s, counters=counters,
                parity=parity if parity_out else None)


def oracle_build_flags():
    try:
        for line in open(os.path.join(ROOT, 'oracle', 'Makefile')):
            if line.startswith('CXXFLAGS'):
                return 'g++ ' + line.split('=', 1)[1].strip()
    except OSError:
        pass
    return None


def run_reference(args, rank, world):
    if rank != 0:
        return
    n, taxa = args.n, args.taxa
    fasta = dataset(n, taxa, config=args.config)
    r = cpu_arm(fasta, n, args.steps, args.warmup, args.cpu_queries, flags=flags_of(args), parity_out=args.parity_out)
    tot = sum(r['times'])
    value = r['per_step'] * len(r['times']) / tot
    sample = '%d workers x %d queries per step against all %d targets (index built once: %.1f s, not timed)' % (
        r['workers'], args.cpu_queries, n, r['build_s'])
    line = {'impl': 'reference', 'metric': metric_of(args), 'value': value, 'unit': 'proteins/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot / len(r['times']),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic',
            'config': workload_config(args, r['per_step']),
            'cpu_baseline': {'value': value, 'unit': 'proteins/s', 'cores': r['workers'], 'kind': 'port',
                             'sample': sample, 'build': oracle_build_flags()},
            'parity_window': r['parity'],
            'e2e': {'value': value, 'unit': 'proteins/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gcups_cpu': r['counters'][5] / tot / 1e9}
    print(json.dumps(line), flush=True)


def workload_config(args, block):
    f = flags_of(args)
    nch = (args.n + FLAGS['chk'] - 1) // FLAGS['chk']
    return {'workload': 'BASELINE config %d: synthetic %d proteins (%d taxa) all-vs-all, blastp -e %g '
                        '-s %s -r aa9 -M 120000000 -c 50000 -j 1' % (args.config, args.n, args.taxa, f['expect'], f['ssd']),
            'step': 'one block of %d queries per GPU against the full target index (%d chunks resident in HBM)' % (block, nch),
            'l2': 'every step uses a new query block; seed-hit buffers are GBs per step (>> 126 MB L2)'}


# ----------------------------------------------------------------------------------------- GPU arm
# positions in the per-rank value vector that are combined with MAX (times); the rest are summed (work)
MAX_IDX = (0, 1, 5, 6, 7, 8, 9, 10, 11, 16, 17, 20)  # 18, 19 (GCUPS) are summed over ranks


def reduce_over_ranks(vals, dist, device):
    """max over ranks for times, sum over ranks for work counters (multi-GPU numbers are never wall
    clock of one rank)."""
    if dist is None:
        return list(vals)
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    mx, sm = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return [(mx[i] if i in MAX_IDX else sm[i]).item() for i in range(len(vals))]


def full_job(args, rank, world, local, fasta, barrier, out_dir):
    """The whole find_hit job (SURVEY.md 8d metric 1): FASTA file in -> OUT written.  Every rank parses the FASTA,
    loads the targets, builds the complete index, seg-masks and searches ITS query slice (balanced by residues,
    swiftortho_b200.find_hit.slices_by_residues) and writes a part file; rank 0 concatenates the parts in query
    order (bin/find_hit.py:135-146).  Returns the rank's wall time and counters; the timed region starts after the
    barrier and ends when the rank's part (rank 0: the merged OUT) is on disk."""
    import ctypes as C
    import numpy as np
    from swiftortho_b200 import find_hit
    from swiftortho_b200 import search as so
    barrier()
    t0 = time.perf_counter()
    F = so.Fasta(fasta)
    S = so.Searcher(device=local, **flags_of(args))
    S.set_targets(F)
    info = S.build_index()
    t_index = time.perf_counter() - t0
    sl = find_hit.slices_by_residues(F, 0, F.N, world)
    a, b = sl[rank] if rank < len(sl) else (0, 0)
    part = os.path.join(out_dir, 'bench_full.%012d' % a)
    open(part, 'wb').close()
    lib = S.lib
    nrows = 0
    if b > a:
        S.stats(reset=True)
        nrows = S.search_stream(F, [(x, min(b, x + args.full_block)) for x in range(a, b, args.full_block)], part)
    st = S.stats()
    t_part = time.perf_counter() - t0
    barrier()
    out = os.path.join(out_dir, 'bench_full.sc')
    if rank == 0:
        find_hit._concat(out, [os.path.join(out_dir, 'bench_full.%012d' % s0) for s0, _ in sl])
    dt = time.perf_counter() - t0
    props = None
    if rank == 0:
        props = table_properties(out, args.full_check_rows)
        os.remove(out)
    S.close()
    F.close()
    return dict(wall=dt, t_index=t_index, t_part=t_part, queries=b - a, rows=nrows, stats=st, info=info, props=props,
                n_total=F.N)


def table_properties(path, max_rows):
    """Size-independent properties of a hit table (first `max_rows` rows): queries ascending, bits descending inside a
    query, 16 columns, and the best hit of a query is (almost always) the query itself."""
    last_q, last_bit, firsts, queries, rows, ok = -1, None, 0, 0, 0, True
    with open(path, 'rb') as f:
        for ln in f:
            c = ln.rstrip(b'\n').split(b'\t')
            if len(c) != 16:
                ok = False
                break
            q, bit = int(c[14]), int(c[11])
            if q < last_q or (q == last_q and bit > last_bit):
                ok = False
                break
            if q != last_q:
                queries += 1
                firsts += c[0] == c[1]
            last_q, last_bit = q, bit
            rows += 1
            if rows >= max_rows:
                break
    return {'ordered': ok, 'rows_checked': rows, 'queries_checked': queries,
            'self_hit_first_frac': firsts / max(1, queries)}


def run_full(args, rank, world, local, dist, fasta):
    """`--mode full`: strong scaling of one whole job (the north-star measurement for config 3)."""
    def barrier():
        if dist is not None:
            dist.barrier()
    shm = '/dev/shm' if os.path.isdir('/dev/shm') else CACHE
    out_dir = os.path.join(shm, 'swiftortho_b200_full')
    os.makedirs(out_dir, exist_ok=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    r = full_job(args, rank, world, local, fasta, barrier, out_dir)
    clk = clocks.stop() if rank == 0 else None
    st = r['stats']
    vals = [r['wall'], r['t_index'], r['t_part'], float(r['queries']), float(r['rows']), st['ms_ungap_kernel'], float(st['ungap_steps']),
            float(st['h2d_bytes']), float(st['d2h_bytes']), float(st['kernel_launches']), float(st['dp_cells']), st['ms_dp'],
            float(st['seed_hits']), float(st['candidates']), float(st['alignments']), st['ms_sort'], st['ms_select'], st['ms_seed'],
            st['ms_traceback'], st['ms_host'], float(st['redo_blocks'])]
    if dist is not None:
        import torch
        t = torch.tensor(vals, dtype=torch.float64, device='cuda')
        mx, sm = t.clone(), t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        mxi = (0, 1, 2, 5, 11, 15, 16, 17, 18, 19)
        vals = [(mx[i] if i in mxi else sm[i]).item() for i in range(len(vals))]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    (wall, t_index, t_part, nq, nrows, ms_x, steps, h2d, d2h, launches, cells, ms_dp, hits, cands, alns, ms_sort, ms_sel, ms_seed,
     ms_tb, ms_host, redo) = vals
    int_peak = json.load(open(os.path.join(ROOT, 'profiles', 'int_peak.json')))
    ach = 6.0 * steps / max(world, 1) / (ms_x * 1e-3) / 1e9 if ms_x > 0 else 0.0
    value = nq / wall
    line = {'metric': metric_of(args), 'value': value, 'unit': 'proteins/s', 'n_gpus': world, 'steps': 1, 'warmup': 0,
            'ms_per_step': 1e3 * wall, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'int32',
            'data': 'synthetic',
            'config': dict(workload_config(args, args.full_block),
                           step='the whole job: FASTA parsed, targets loaded, complete index built on every rank, %d queries split '
                                'by residues over %d rank(s), 16-column parts written and concatenated by rank 0' % (int(nq), world)),
            'roofline': {'kernel': 'k_xdrop', 'bound': 'int32', 'achieved': ach, 'peak': int_peak['gops_measured'], 'unit': 'Gop/s',
                         'frac': ach / int_peak['gops_measured'], 'traffic': None,
                         'timing': 'CUDA events on the launching stream, max over ranks',
                         'peak_source': 'tools/int_peak.cu measured on this pool (profiles/int_peak.json)'},
            'e2e': {'value': value, 'unit': 'proteins/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'note': 'the job is end to end by construction: host FASTA file in, text table out'},
            'gpu_launches': int(launches), 'clocks': clk,
            'wall_s': wall, 'index_build_s_max': t_index, 'part_written_s_max': t_part, 'rows': int(nrows),
            'seed_hits': hits, 'candidates': cands, 'alignments': alns, 'dp_cells': cells, 'redo_blocks': int(redo),
            'stage_ms_max_rank': dict(seed=ms_seed, grouping=ms_sort, xdrop=ms_x, select=ms_sel, dp=ms_dp, traceback=ms_tb, host=ms_host),
            'index_chunks': len(r['info']), 'table_properties': r['props']}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def parity_check(S, F, so, window, ref_path):
    """Rows of the query window the CPU arm kept, recomputed on the GPU through the public API and compared byte for
    byte with the oracle's text (outside every timed region)."""
    import ctypes as C
    import numpy as np
    if not window or not os.path.exists(ref_path):
        return {'status': 'skipped'}
    a, b = window
    want = open(ref_path, 'rb').read()
    off = np.ascontiguousarray(F.offsets[a:b + 1])
    so.check(S.lib.so_set_queries(S.h, C.c_void_p(F._res.value), off.ctypes.data, b - a))
    rows = S.search(0, b - a)
    rows.view()['query'] += a
    outp = ref_path + '.gpu'
    so.check(S.lib.so_write_rows(rows.ptr, rows.n, F.h, F.h, outp.encode(), 0))
    got = open(outp, 'rb').read()
    os.remove(outp)
    return {'status': 'ok' if got == want else 'MISMATCH', 'queries': [a, b], 'rows': want.count(b'\n'),
            'checker': 'oracle port (oracle/fsearch_oracle.cpp), byte comparison of the 16-column text'}


def run_ours(args, rank, world, local):
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n, taxa, B = args.n, args.taxa, args.block
    fasta = dataset(n, taxa, rank, config=args.config)
    from swiftortho_b200 import search as so
    if args.mode == 'full':
        run_full(args, rank, world, local, dist, fasta)
        return
    F = so.Fasta(fasta)
    S = so.Searcher(device=local, **flags_of(args))
    t0 = time.perf_counter()
    S.set_targets(F)
    info = S.build_index()
    index_ms = 1e3 * (time.perf_counter() - t0)
    S.set_queries(F)
    nblocks = max(1, n // B)

    def block_of(step):
        b = (step * world + rank) % nblocks
        return b * B, min(n, (b + 1) * B)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- device-resident arm (value)
    for s in range(args.warmup):
        S.search(*block_of(s))
    S.stats(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    t0 = time.perf_counter()
    nq = 0
    for s in range(args.steps):
        a, b = block_of(args.warmup + s)
        S.search(a, b)           # returns after the stream is synchronised
        nq += b - a
    barrier()
    dt = time.perf_counter() - t0
    st = S.stats(reset=True)
    clk = clocks.stop() if rank == 0 else None

    # ---- kernel-timing pass: the same steps with ONE production lane, so the CUDA-event stage times
    #      inside so_search are not inflated by the second lane's kernels sharing the GPU.  The roofline
    #      figures (kernel durations) come from this pass; `value` above is the pipeline (candidate production overlapped with the alignment rounds).
    S.set_lanes(0)   # one lane, alignment rounds serialised with candidate production
    S.search(*block_of(args.warmup + args.steps))
    S.stats(reset=True)
    ksteps = max(1, min(2, args.steps))
    for s in range(ksteps):
        S.search(*block_of(args.warmup + s))
    st = dict(st)
    stk = S.stats(reset=True)
    gcups_pipe = stk['dp_cells'] / (stk['ms_dp'] * 1e-3) / 1e9 if stk['ms_dp'] > 0 else 0.0
    S.set_lanes(1)
    for k in ('ms_ungap', 'ms_ungap_kernel', 'ms_sort', 'ms_seed', 'ms_select', 'ms_dp', 'ms_traceback'):
        st[k] = stk[k] * args.steps / ksteps      # same blocks as the first `ksteps` timed steps

    # ---- end-to-end arm: host buffers -> public API -> text rows, every step
    import ctypes as C
    import numpy as np
    lib = S.lib
    res_ptr = F._res.value
    shm = '/dev/shm' if os.path.isdir('/dev/shm') else CACHE
    outp = os.path.join(shm, 'swiftortho_b200_bench_%d.sc' % rank)

    # Searcher.search_stream is the streaming call of the public API (swiftortho_b200/search.py): per block it prepares
    # the queries on the host (seg, S3 order), copies them to the device, searches, copies the rows back and formats the
    # text; the host preparation of block i + 1 and the formatting of block i - 1 overlap block i's device work.  The
    # timed call covers exactly `steps` blocks, pipeline fill and drain included.
    S.search_stream(F, [block_of(1000 + s) for s in range(args.warmup)], outp, append=False)
    S.stats(reset=True)
    barrier()
    t1 = time.perf_counter()
    timed = [block_of(1000 + args.warmup + s) for s in range(args.steps)]
    S.search_stream(F, timed, outp, append=False)
    e2e_q = sum(b - a for a, b in timed)
    barrier()
    dt2 = time.perf_counter() - t1
    st2 = S.stats()
    try:
        os.remove(outp)
    except OSError:
        pass

    # ---- gapped extension alone: one large so_align_batch over config-shaped pairs (the DP has no early
    #      exit, so its cost depends only on the sequence lengths) -> clean GCUPS of k_banded_dp
    npairs = args.align_pairs
    Pp = (so.so_pair * npairs)()
    offs = F.offsets
    for i in range(npairs):
        qi, ti = (i * 3 + rank) % n, (i * 7919 + 13) % n
        # (sequences of 4096+ residues -- config 5 -- enter as their first 4095-residue tile: so_align_batch takes tiles)
        Pp[i] = so.so_pair(qi, ti, 0, min(4095, int(offs[qi + 1] - offs[qi])), 0, min(4095, int(offs[ti + 1] - offs[ti])), 0, 0)
    Aa = (so.so_aln * npairs)()
    so.check(lib.so_set_queries(S.h, F._res, F._off, F.N))
    so.check(lib.so_align_batch(S.h, Pp, npairs, Aa))
    S.stats(reset=True)
    barrier()
    for _ in range(3):
        so.check(lib.so_align_batch(S.h, Pp, npairs, Aa))
    st3 = S.stats()
    gcups_alone = st3['dp_cells'] / (st3['ms_dp'] * 1e-3) / 1e9
    gcups_alone_tb = st3['dp_cells'] / ((st3['ms_dp'] + st3['ms_traceback']) * 1e-3) / 1e9

    # ---- reduce over ranks (max time, summed work)
    vals = [dt, dt2, float(nq), float(e2e_q), float(st['dp_cells']), st['ms_dp'], st['ms_ungap'], st['ms_sort'],
            st['ms_seed'], st['ms_select'], st['ms_traceback'], st['ms_host'], float(st['seed_hits']),
            float(st['kernel_launches']), float(st['ungap_steps']), float(st['alignments']),
            float(st2['h2d_bytes']), float(st2['d2h_bytes']), gcups_alone, gcups_alone_tb,
            st['ms_ungap_kernel'], float(st['groups']), float(st['multi_groups']), float(st['alignments_used']), gcups_pipe]
    vals = reduce_over_ranks(vals, dist, 'cuda' if dist is not None else None)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    (dt, dt2, nq, e2e_q, cells, ms_dp, ms_ungap, ms_sort, ms_seed, ms_select, ms_tb, ms_host, seed_hits, launches,
     ungap_steps, alignments, h2d, d2h, gcups_alone, gcups_alone_tb, ms_xdrop, groups, multi_groups, aln_used, gcups_pipe) = vals
    value = nq / dt
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    int_peak = json.load(open(os.path.join(ROOT, 'profiles', 'int_peak.json')))
    # dominant kernel of the step = chained X-drop scoring (k_xdrop): integer-pipe bound
    # (6 INT ops per extension step); peak = measured INT32 op rate of this chip.
    # Durations: CUDA events around the kernel inside so_search, single-lane pass (see above).
    ops = 6.0 * ungap_steps / max(world, 1)
    ach = ops / (ms_xdrop * 1e-3) / 1e9 if ms_xdrop > 0 else 0.0
    dev_ms = max(1e-9, ms_ungap + ms_sort + ms_seed + ms_select + ms_dp + ms_tb)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_xdrop_r02b.json')))
    except Exception:  # noqa: BLE001
        pass
    roof = {'kernel': 'k_xdrop', 'bound': 'int32', 'achieved': ach, 'peak': int_peak['gops_measured'],
            'unit': 'Gop/s', 'frac': ach / int_peak['gops_measured'],
            'traffic': traffic.get('dram_bytes_per_launch') if traffic else None,
            'traffic_note': traffic.get('note') if traffic else None,
            'traffic_launch': traffic.get('launch') if traffic else None,
            'peak_source': 'tools/int_peak.cu measured on this pool (profiles/int_peak.json)',
            'algorithmic_ops': '6 INT ops per X-drop extension step (SURVEY.md 8d) x %.3g steps per step' % (
                ungap_steps / max(world, 1) / args.steps),
            'ms_per_step': ms_xdrop / args.steps, 'share_of_step_device_time': ms_xdrop / dev_ms,
            'groups_per_step': groups / max(world, 1) / args.steps,
            'chained_groups_per_step': multi_groups / max(world, 1) / args.steps,
            'timing': 'CUDA events on the launching stream inside so_search, one production lane'}
    # HBM view of the hit grouping (second largest stage): (query, target) cell partition + in-cell sorts.
    # Algorithmic bytes per seed hit: 2 x 8 B index entry reads (count + scatter pass), 4 B cell-local key write,
    # 4 B read + 8 B key write in the cell sorts = 32 B
    sort_bytes = seed_hits / max(world, 1) * 40.0
    hbm = {'kernel': 'k_cell_pass<0/1> + k_unit_scan + k_cell_span / k_cell_block', 'bound': 'hbm',
           'achieved': sort_bytes / (ms_sort * 1e-3) / 1e9 if ms_sort else 0,
           'peak': peaks.get('hbm_gbs', 6650.0), 'unit': 'GB/s', 'ms_per_step': ms_sort / args.steps,
           'algorithmic_bytes': '40 B per seed hit (2 x 8 B index entry reads, 4 B key write + 4 B read, 4 + 8 + 4 B sorted key / '
                                'X-drop descriptor / cell id writes: the descriptors were 28 B per hit in another stage in round 1) '
                                'x %.3g hits per step' % (seed_hits / max(world, 1) / args.steps),
           'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback B200_PROFILING.md'}
    hbm['frac'] = hbm['achieved'] / hbm['peak']
    gcups = gcups_pipe / max(world, 1)
    dp_roof = {'kernel': 'k_banded_dp', 'bound': 'int32', 'achieved': gcups_alone / max(world, 1), 'unit': 'GCUPS per GPU',
               'peak': int_peak['gops_measured'] / 14.0, 'frac': gcups_alone / max(world, 1) / (int_peak['gops_measured'] / 14.0),
               'with_traceback': gcups_alone_tb / max(world, 1), 'in_pipeline': gcups,
               'in_pipeline_frac': gcups / (int_peak['gops_measured'] / 14.0),
               'cells_per_step': cells / args.steps / max(world, 1),
               'note': '14 INT ops per cell (SURVEY.md 8d); achieved = %d config-shaped pairs in one so_align_batch, '
                       'k_banded_dp time by CUDA events; in_pipeline = the same kernel on the alignment rounds of the search steps, '
                       'timed in the measurement mode (one lane, rounds serialised with candidate production, so the '
                       'events see the kernel alone)' % args.align_pairs}
    line = {'metric': metric_of(args), 'value': value, 'unit': 'proteins/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic', 'config': workload_config(args, B),
            'roofline': roof, 'roofline_grouping': hbm, 'roofline_dp': dp_roof, 'gapped_gcups': gcups_alone,
            'e2e': {'value': e2e_q / dt2, 'unit': 'proteins/s', 'h2d_bytes_per_step': h2d / args.steps,
                    'd2h_bytes_per_step': d2h / args.steps,
                    'api': 'Searcher.search_stream: per step so_queries_prepare (host: seg, position order) -> '
                           'so_set_queries_prepared (H2D) -> so_search (kernels, rows D2H) -> so_write_rows (text); the '
                           'preparation of step i+1 and the text of step i-1 overlap step i; the timed call covers exactly '
                           '`steps` steps, pipeline fill and drain included'},
            'gpu_launches': int(launches), 'clocks': clk,
            'stage_ms_per_step': {k: v / args.steps for k, v in dict(seed=ms_seed, grouping=ms_sort, ungap=ms_ungap,
                                                                      select=ms_select, dp=ms_dp, traceback=ms_tb,
                                                                      host=ms_host).items()},
            'index_build_ms': index_ms, 'index': info, 'alignments_per_query': alignments / max(1.0, nq),
            'alignments_wasted_frac': 1.0 - aln_used / max(1.0, alignments),
            'seed_hits_per_query': seed_hits / max(1.0, nq)}
    if world == 1 and not args.no_cpu:
        # CPU baseline (the oracle port) in a child process that never touches CUDA; it keeps the rows of one query
        # window, which the GPU path then has to reproduce byte for byte (parity_check)
        shm = '/dev/shm' if os.path.isdir('/dev/shm') else CACHE
        pout = os.path.join(shm, 'swiftortho_b200_parity_%d.sc' % os.getpid())
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '2',
                              '--warmup', '1', '--n', str(n), '--taxa', str(taxa), '--cpu-queries', str(args.cpu_queries),
                              '--config', str(args.config), '--parity-out', pout],
                             stdout=subprocess.PIPE, text=True, cwd=ROOT)
        try:
            ref = json.loads(out.stdout.strip().splitlines()[-1])
            line['cpu_baseline'] = ref['cpu_baseline']
            line['cpu_baseline']['gcups'] = ref.get('gcups_cpu')
            line['parity_check'] = parity_check(S, F, so, ref.get('parity_window'), pout)
        except Exception as e:  # noqa: BLE001
            line['cpu_baseline'] = {'error': repr(e)}
        try:
            os.remove(pout)
        except OSError:
            pass
    if world == 1 and not args.no_full:
        # SURVEY.md 8d metric (1): the whole job, FASTA in -> OUT written, index build and query preparation included
        S.close()
        shm = '/dev/shm' if os.path.isdir('/dev/shm') else CACHE
        out_dir = os.path.join(shm, 'swiftortho_b200_full')
        os.makedirs(out_dir, exist_ok=True)
        r = full_job(args, 0, 1, local, fasta, lambda: None, out_dir)
        line['full_run'] = {'proteins_per_s': r['queries'] / r['wall'], 'wall_s': r['wall'], 'queries': r['queries'],
                            'rows': r['rows'], 'index_build_s': r['t_index'], 'table_properties': r['props'],
                            'what': 'whole find_hit job on one GPU: FASTA parse, H2D, index build, seg + S3 order, search, '
                                    '16-column text written (SURVEY.md 8d metric 1)'}
    c3 = []
    for k in (1, 2, 4, 8):
        pth = os.path.join(ROOT, 'profiles', 'bench_r02_config3_n%d.json' % k)
        if os.path.exists(pth):
            try:
                d = json.loads(open(pth).read().strip().splitlines()[-1])
                c3.append({'n_gpus': d['n_gpus'], 'proteins_per_s': d['value'], 'wall_s': d['wall_s'], 'source': 'profiles/' + os.path.basename(pth)})
            except Exception:  # noqa: BLE001
                pass
    if c3:
        line['config3'] = {'scaling': 'strong', 'note': 'committed measurements of `bench.py --config 3 --mode full` (1 M proteins, 20 '
                           'chunks, queries split by residues over the ranks, index build included); not re-run here', 'runs': c3}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--block', type=int, default=4096, help='queries per step per GPU')
    ap.add_argument('--n', type=int, default=100000)
    ap.add_argument('--taxa', type=int, default=20)
    ap.add_argument('--cpu-queries', type=int, default=4, help='queries per CPU worker per step (bounded sample)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--align-pairs', type=int, default=300000, help='pairs of the alignment-only GCUPS measurement')
    ap.add_argument('--config', type=int, default=2, choices=sorted(CONFIGS), help='BASELINE.json config (default 2 = headline)')
    ap.add_argument('--mode', default='steps', choices=['steps', 'full'],
                    help='steps: timed query blocks against the resident index (default); full: one whole job, strong scaling')
    ap.add_argument('--full-block', type=int, default=16384, help='queries per so_search call of the whole-job run')
    ap.add_argument('--full-check-rows', type=int, default=2000000)
    ap.add_argument('--no-full', action='store_true', help='skip the whole-job record of the default run')
    ap.add_argument('--parity-out', default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.n == 100000 and args.taxa == 20 and args.config != 2:
        args.n, args.taxa = CONFIGS[args.config][0], CONFIGS[args.config][1]
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else max(args.warmup, 1)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local)


if __name__ == '__main__':
    main()

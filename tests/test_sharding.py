"""Multi-GPU host logic on CPU: world_size-2 gloo process group, the oracle standing in for the
device tagger (the sharding code only sees a `tag_fn`)."""

import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import lattice_based_tagger_b200 as pkg
from lattice_based_tagger_b200 import sharding
from oracle import lattice_oracle as lo
from tests import _cases


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    lengths = rng.integers(0, 80, size=1000)
    for world in (1, 2, 3, 4, 8):
        bounds = sharding.shard_bounds(lengths, world)
        assert bounds[0] == 0 and bounds[-1] == 1000 and len(bounds) == world + 1
        assert all(a <= b for a, b in zip(bounds, bounds[1:]))
        work = [int((lengths[a:b] + 1).sum()) for a, b in zip(bounds, bounds[1:])]
        assert max(work) - min(work) <= 2 * 81
    assert sharding.shard_bounds([], 4) == [0, 0, 0, 0, 0]
    assert sharding.shard_bounds([5, 5], 4)[-1] == 2


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, seed, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    case = _cases.random_case(seed, n_sent=30, features=False)
    dictionary, funcs = _cases.build_objects(case, pkg)
    oracle = lo.OracleTagger(dictionary, funcs)

    def tag_fn(sents):
        out = []
        for s in sents:
            try:
                best = oracle.tag(s, 5)
                out.append((best.words, best.score))
            except IndexError:
                out.append(None)
        return out

    merged = sharding.tag_sharded(tag_fn, case['sentences'], rank, world)
    whole = tag_fn(case['sentences'])
    ok = merged == whole
    with open(os.path.join(out_dir, 'rank%d' % rank), 'w') as f:
        f.write('ok' if ok else 'mismatch')
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gather_in_order(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 77, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert (tmp_path / ('rank%d' % rank)).read_text() == 'ok'


def test_partition_by_work_balances_and_covers():
    rng = np.random.default_rng(1)
    lengths = rng.integers(0, 120, size=1003)
    for world in (1, 2, 4, 8):
        parts = sharding.partition_by_work(lengths, world)
        assert len(parts) == world
        assert sorted(np.concatenate(parts).tolist()) == list(range(1003))
        counts = [len(p) for p in parts]
        assert max(counts) - min(counts) <= 1
        work = [int(lengths[p].sum()) for p in parts]
        assert max(work) - min(work) <= 120 * 2
        assert all((np.diff(p) > 0).all() for p in parts if len(p) > 1)


def _packed_worker(rank, world, port, seed, out_dir):
    """`tag_sharded_packed` over gloo: the packed results of two ranks gathered as tensors and put back
    into input order must equal one rank tagging everything.  The tagger is the package's own, running
    the kernel sources through the SIMT emulator (tests/_emu.py)."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tests import _checks, _emu
    with _emu.emulated():
        case = _checks.make_case(seed, n_sent=20, max_sent_len=30)
        dictionary, funcs = _cases.build_objects(case, pkg)
        tagger = pkg.Tagger(dictionary, score_funcs=funcs)
        sents = case['sentences']
        merged = sharding.tag_sharded_packed(tagger, sents, 5, rank, world)
        merged_c = sharding.tag_sharded_packed(tagger, sents, 5, rank, world, contiguous=True)
        ok = True
        if rank == 0:
            whole = tagger.tag_batch_packed(sents, 5)
            ok = all(np.array_equal(a, b) for a, b in zip(merged, whole))
            ok = ok and all(np.array_equal(a, b) for a, b in zip(merged_c, whole))      # contiguous shards, fast gather
            oracle = lo.OracleTagger(dictionary, funcs)
            seqs = tagger.unpack(sents, merged, errors='none')
            for sent, seq in zip(sents, seqs):
                try:
                    want = oracle.tag(sent, 5)
                except IndexError:
                    ok = ok and seq is None
                    continue
                ok = ok and [tuple(w) for w in seq.sequences] == want.words and seq.score == want.score
        else:
            ok = merged is None and merged_c is None
        tagger.close()
    with open(os.path.join(out_dir, 'rank%d' % rank), 'w') as f:
        f.write('ok' if ok else 'mismatch')
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_packed_gather(tmp_path):
    from tests import _emu
    _emu.emu_lib()          # built once here, not by two processes at the same time
    world = 2
    mp.spawn(_packed_worker, args=(world, _free_port(), 91, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert (tmp_path / ('rank%d' % rank)).read_text() == 'ok'


def _worker_packed(rank, world, port, out_dir):
    """Contiguous shards of synthetic packed results through gather_packed_contiguous (gloo, host tensors)."""
    from lattice_based_tagger_b200 import _native
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    n = 57
    plen = rng.integers(0, 9, size=n).astype(np.int32)
    plen[10:14] = 0                                          # sentences without a path
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(plen, out=off[1:])
    edges = np.zeros(int(off[-1]), dtype=_native.EDGE_DTYPE)
    edges['b'] = np.arange(edges.size) % 50
    edges['e'] = edges['b'] + 1
    edges['rule'] = np.arange(edges.size)
    scores = rng.standard_normal(n)
    status = (plen == 0).astype(np.int32)
    bounds = sharding.shard_bounds(rng.integers(1, 40, size=n), world)
    a, b = bounds[rank], bounds[rank + 1]
    local = (plen[a:b], edges[off[a]:off[b]], scores[a:b], status[a:b])
    got = sharding.gather_packed_contiguous(local, rank, world)
    ok = True
    if rank == 0:
        path_off, out_edges, out_scores, out_status = got
        ok = (np.array_equal(path_off, off) and np.array_equal(out_edges, edges) and np.array_equal(out_scores, scores)
              and np.array_equal(out_status, status))
    else:
        ok = got is None
    with open(os.path.join(out_dir, 'rank%d' % rank), 'w') as f:
        f.write('ok' if ok else 'mismatch')
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_gather_packed_contiguous(tmp_path, world):
    mp.spawn(_worker_packed, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert (tmp_path / ('rank%d' % rank)).read_text() == 'ok'

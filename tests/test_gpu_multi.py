"""Sharding on hardware: one corpus partitioned over two GPUs (NCCL), packed results gathered on rank 0
and compared with a single-GPU pass and with the oracle.  Skipped on a box with one GPU."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import lattice_based_tagger_b200 as pkg
    from lattice_based_tagger_b200 import sharding, synth
    from oracle import lattice_oracle as lo
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    cfg, dictionary, sents = synth.build_workload('tiny', n_sent=600)
    funcs = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore())
    tagger = pkg.Tagger(dictionary, score_funcs=funcs, device=rank)
    merged = sharding.tag_sharded_packed(tagger, sents, 5, rank, world, device=dev)
    ok = True
    if rank == 0:
        whole = tagger.tag_batch_packed(sents, 5)
        ok = all(np.array_equal(a, b) for a, b in zip(merged, whole))
        oracle = lo.OracleTagger(dictionary, funcs)
        seqs = tagger.unpack(sents, merged, errors='none')
        for sent, seq in list(zip(sents, seqs))[:100]:
            try:
                want = oracle.tag(sent, 5)
            except IndexError:
                ok = ok and seq is None
                continue
            ok = ok and [tuple(w) for w in seq.sequences] == want.words and seq.score == want.score
    else:
        ok = merged is None
    with open(os.path.join(out_dir, 'rank%d' % rank), 'w') as f:
        f.write('ok' if ok else 'mismatch')
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_sharded_corpus(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert (tmp_path / ('rank%d' % rank)).read_text() == 'ok'

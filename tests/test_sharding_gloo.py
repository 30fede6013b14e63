"""N>1 host logic on CPU: two gloo ranks shard a frame batch, each decodes its shard (the CPU
oracle stands in for the GPU kernels, which are covered by the -m gpu tests), the decoded bits
are gathered on rank 0 and must equal the single-process result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_frames_partitions_exactly(ofdm):
    sh = ofdm.sharding
    for n in (0, 1, 7, 8, 10000, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [sh.shard_frames(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for first, c in spans:
                assert first == pos
                pos += c
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        sh.shard_frames(4, 2, 2)


def _worker(rank, world, port, n_frames, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import ofdm_b200 as m
    from oracle import oracle_py

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, N, C, S, b = 4, 64, 16, 6, 2
    d = m.synth.make_frames(n_frames, A, N, C, S, b, snr_db=10.0, seed=21)  # same seed: same global batch
    first, count = m.sharding.shard_frames(n_frames, world, rank)
    out = oracle_py.demod_frames(d["rx"][first:first + count], d["pilot_asc"], b, C)
    bits = m.sharding.gather_rows(torch.from_numpy(out["bits"]), n_frames, dst=0)
    comb = m.sharding.gather_rows(torch.view_as_real(torch.from_numpy(out["combined"])), n_frames, dst=0)
    if rank == 0:
        np.savez(out_path, bits=bits.numpy(), comb=comb.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path, ofdm, oracle):
    import torch.multiprocessing as mp

    n_frames, world = 5, 2  # uneven split: 3 + 2
    out_path = str(tmp_path / "gathered.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_frames, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    d = ofdm.synth.make_frames(n_frames, 4, 64, 16, 6, 2, snr_db=10.0, seed=21)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], 2, 16)
    assert np.array_equal(got["bits"], ref["bits"])
    assert np.array_equal(got["comb"].view(np.complex64)[..., 0], ref["combined"])

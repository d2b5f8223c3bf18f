"""world_size-2 gloo tests (CPU): stream partitioning + host-side gather give identical bytes
regardless of the number of ranks."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import speechpipe_ref as sp
from project_morpheus_b200 import partition


def fake_batch(windows):
    out = []
    for w in windows:
        h = hashlib.sha256(np.asarray(list(w), dtype=np.int32).tobytes()).digest()
        out.append((h * 128)[:4096])
    return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tick = [(s, sp.synth_codes(s, 4).tolist()) for s in range(n_streams)]
        dec = partition.PartitionedDecoder(fake_batch)
        merged = dec.decode_tick(tick)
        mine = partition.local_streams(n_streams, rank, world)
        pcm = np.stack([np.frombuffer(fake_batch([tick[s][1]])[0], dtype="<i2") for s in mine])
        full = dec.gather_pcm(pcm)
        if rank == 0:
            q.put((merged, full))
    finally:
        dist.destroy_process_group()


def test_partition_map_is_a_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(s for r in range(world) for s in partition.local_streams(1024, r, world))
        assert seen == list(range(1024))
        sizes = [len(partition.local_streams(1024, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) == 0
    parts = partition.split_tick([(s, [s]) for s in range(10)], 4)
    assert [[s for s, _ in p] for p in parts] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]


@pytest.mark.timeout(120)
def test_two_rank_gather_equals_single_rank():
    n_streams, world = 16, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged, full = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    tick = [(s, sp.synth_codes(s, 4).tolist()) for s in range(n_streams)]
    single = dict(zip(range(n_streams), fake_batch([w for _, w in tick])))
    assert merged == single
    want = np.stack([np.frombuffer(single[s], dtype="<i2") for s in range(n_streams)])
    assert np.array_equal(full, want)

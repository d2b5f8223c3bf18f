"""Device-side gather of a partitioned tick over NCCL (needs two GPUs; the 1-GPU box skips it): every stream's bytes
equal the single-engine decode, for uniform ticks (gather_pcm_device) and ragged ones (decode_tick_device)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from helpers import windows_tokens
    from project_morpheus_b200 import _lib, weights
    from project_morpheus_b200.engine import SnacEngine
    from project_morpheus_b200.partition import PartitionedDecoder, local_streams

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = SnacEngine(weights.random_state_dict(0, "w1"), device=rank, precision="fp16")
    pd = PartitionedDecoder(lambda w: [], rank=rank, world_size=world)
    # uniform tick: 96 streams, stream s on rank s mod world
    n = 96
    tok = windows_tokens(n, 4, 8800)
    mine = local_streams(n, rank, world)
    pcm_d, _ = eng.decode_windows_device(torch.from_numpy(tok[mine]).cuda(), noise="philox", seed=3, keys=mine)
    g = pd.gather_pcm_device(pcm_d, dst=0)
    # ragged tick: every rank passes the whole list
    rng = np.random.default_rng(5)
    tick = []
    for s in range(70):
        for _ in range(int(rng.integers(0, 3))):
            fr = int(rng.choice([1, 4, 7]))
            tick.append((s, windows_tokens(1, 7, 9000 + s)[0][: 7 * fr].tolist()))
    tick.append((3, [1, 2, 3]))  # rejected window: None

    def dec(wins):
        lens = [len(w) for w in wins]
        t = np.zeros((len(wins), 49), dtype=np.int32)
        for i, w in enumerate(wins):
            t[i, : len(w)] = w
        return eng.decode_windows_device(torch.from_numpy(t).cuda(), ntok=lens, noise="off")

    merged = pd.decode_tick_device(tick, dec, dst=0)
    if rank == 0:
        full, st = eng.decode_windows(tok, noise="philox", seed=3, keys=list(range(n)))
        ok = bool((st == _lib.WIN_OK).all()) and np.array_equal(g, full)
        lens = [len(w) for _, w in tick]
        t = np.zeros((len(tick), 49), dtype=np.int32)
        for i, (_, w) in enumerate(tick):
            t[i, : len(w)] = w
        pcm, st = eng.decode_windows(t, ntok=lens, noise="off")
        want = {}
        for i, (s, _) in enumerate(tick):
            want[s] = pcm[i].tobytes() if st[i] == _lib.WIN_OK else (b"" if st[i] == _lib.WIN_EMPTY else None)
        ok = ok and merged == want and want[3] is None
        out.put(ok)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_device_gather_equals_single_engine_decode(ensure_lib):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(timeout=5) is True

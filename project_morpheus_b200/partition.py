"""Multi-GPU: streams are independent, so a box of N GPUs is N independent stream partitions.

One process per GPU (``torchrun``); rank r owns the streams with ``partition_of(stream) == r``,
decodes only their windows on its own engine, and the per-tick PCM is brought together on one rank
(4 KB per stream per tick to one egress) - the decode path itself has no collective because
no window depends on another.  Two gathers: host-side over gloo (``decode_tick`` / ``gather_pcm``: works
anywhere, CPU tests) and device-side over NCCL / NVLink (``gather_pcm_device``: the PCM leaves each GPU
over NVSwitch and crosses PCIe once, on the egress rank).  Identical bytes come out regardless of how
streams are partitioned.
"""
from __future__ import annotations

from typing import Callable, Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

DecodeBatch = Callable[[Sequence[Sequence[int]]], List[Optional[bytes]]]


def partition_of(stream_index: int, world_size: int) -> int:
    """Stream -> rank.  ``s mod G``."""
    return int(stream_index) % int(world_size)


def local_streams(n_streams: int, rank: int, world_size: int) -> List[int]:
    return [s for s in range(n_streams) if partition_of(s, world_size) == rank]


def split_tick(windows: Sequence[Tuple[int, Sequence[int]]], world_size: int) -> List[List[Tuple[int, Sequence[int]]]]:
    """[(stream, window)] -> per-rank lists, order inside a rank preserved."""
    parts: List[List[Tuple[int, Sequence[int]]]] = [[] for _ in range(world_size)]
    for s, w in windows:
        parts[partition_of(s, world_size)].append((s, w))
    return parts


class PartitionedDecoder:
    """Decodes this rank's share of a tick and gathers every stream's PCM on ``dst`` (host side)."""

    def __init__(self, decode_batch: DecodeBatch, rank: Optional[int] = None, world_size: Optional[int] = None, group=None):
        import torch.distributed as dist

        self._dist = dist
        self._decode = decode_batch
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size

    def decode_tick(self, windows: Sequence[Tuple[int, Sequence[int]]], dst: int = 0) -> Optional[Dict[int, Optional[bytes]]]:
        """``windows`` is the whole tick (every rank passes the same list, or at least its own share);
        returns {stream: bytes | None} on ``dst`` and None elsewhere."""
        mine = [(s, w) for s, w in windows if partition_of(s, self.world_size) == self.rank]
        out = self._decode([w for _, w in mine]) if mine else []
        payload = [(s, a) for (s, _), a in zip(mine, out)]
        gathered: Optional[List[object]] = [None] * self.world_size if self.rank == dst else None
        self._dist.gather_object(payload, gathered, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        merged: Dict[int, Optional[bytes]] = {}
        for part in gathered or []:
            for s, a in part:  # type: ignore[union-attr]
                merged[s] = a
        return merged

    def gather_pcm(self, pcm: np.ndarray, dst: int = 0) -> Optional[np.ndarray]:
        """Fixed-shape variant for uniform ticks: [n_local, 2048] int16 per rank -> [world*n_local, 2048] on dst,
        rows ordered by stream index (stream s = row s when every rank holds n_local streams)."""
        import torch

        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        t = torch.from_numpy(pcm.view(np.uint8))  # bytes: gloo has no int16 gather
        bufs = [torch.empty_like(t) for _ in range(self.world_size)] if self.rank == dst else None
        self._dist.gather(t, bufs, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        stacked = torch.stack(bufs, dim=1)  # [n_local, world, bytes]: stream s = local*world + rank
        return stacked.reshape(-1, t.shape[1]).numpy().view(np.int16)

    def gather_pcm_device(self, pcm, dst: int = 0, group=None, out_host=None):
        """Device-side gather for uniform ticks: ``pcm`` is this rank's CUDA int16 [n_local, 2048] (what
        ``SnacEngine.decode_windows_device`` wrote); rank ``dst`` receives every rank's rows over NCCL (NVLink /
        NVSwitch), interleaves them to stream order on the GPU and copies the tick to the host ONCE:
        numpy int16 [world * n_local, 2048], stream s = row s.  ``group``: an NCCL group (default: the default group);
        ``out_host``: optional pinned int16 tensor [world * n_local, 2048] to receive the tick."""
        import torch

        assert pcm.is_cuda and pcm.dtype == torch.int16 and pcm.dim() == 2 and pcm.is_contiguous()
        n_local, width = pcm.shape
        if self.rank == dst:
            key = (n_local, width, pcm.device)
            if getattr(self, "_gbuf_key", None) != key:
                self._gbuf = torch.empty((self.world_size, n_local, width), dtype=torch.int16, device=pcm.device)
                self._gperm = torch.empty((n_local, self.world_size, width), dtype=torch.int16, device=pcm.device)
                self._gbuf_key = key
            bufs = [self._gbuf[r].view(torch.uint8) for r in range(self.world_size)]  # bytes: NCCL has no int16
        else:
            bufs = None
        self._dist.gather(pcm.view(torch.uint8), bufs, dst=dst, group=group)
        if self.rank != dst:
            return None
        self._gperm.copy_(self._gbuf.permute(1, 0, 2))  # stream s = local * world + rank
        flat = self._gperm.view(n_local * self.world_size, width)
        if out_host is None:
            out_host = torch.empty(flat.shape, dtype=torch.int16, pin_memory=True)
        out_host.copy_(flat, non_blocking=True)
        torch.cuda.current_stream(pcm.device).synchronize()
        return out_host.numpy()

    def decode_tick_device(self, windows: Sequence[Tuple[int, Sequence[int]]], decode_device, dst: int = 0, group=None
                           ) -> Optional[Dict[int, Optional[bytes]]]:
        """``decode_tick`` with the gather on the device side.  EVERY rank passes the whole tick (so each knows every
        rank's window count without talking); ``decode_device(list of windows) -> (pcm cuda int16 [n, 2048], status cuda
        int32 [n])`` decodes this rank's share (``SnacEngine.decode_windows_device``).  The rows of all ranks, padded to
        the largest share and carrying their window status in two extra int16 columns, go to ``dst`` in one NCCL gather
        over NVLink and cross PCIe once; returns {stream: bytes | b"" | None} on ``dst`` (what ``convert_to_audio``
        returns per window; the last window of a stream wins, as in ``decode_tick``) and None elsewhere."""
        import torch

        parts = split_tick(windows, self.world_size)
        mine = parts[self.rank]
        max_n = max(1, max(len(p) for p in parts))
        width = 2048 + 2
        dev = torch.device("cuda", torch.cuda.current_device())
        key = (max_n, dev)
        if getattr(self, "_tick_key", None) != key:
            self._tick_send = torch.zeros((max_n, width), dtype=torch.int16, device=dev)
            self._tick_recv = torch.empty((self.world_size, max_n, width), dtype=torch.int16, device=dev) if self.rank == dst else None
            self._tick_host = torch.empty((self.world_size, max_n, width), dtype=torch.int16).pin_memory() if self.rank == dst else None
            self._tick_key = key
        if mine:
            pcm, st = decode_device([w for _, w in mine])
            n = len(mine)
            self._tick_send[:n, :2048].copy_(pcm)
            self._tick_send[:n, 2048:].copy_(st.view(torch.int16).view(n, 2))
        bufs = [self._tick_recv[r].view(torch.uint8) for r in range(self.world_size)] if self.rank == dst else None
        self._dist.gather(self._tick_send.view(torch.uint8), bufs, dst=dst, group=group)
        if self.rank != dst:
            return None
        self._tick_host.copy_(self._tick_recv, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        host = self._tick_host.numpy()
        merged: Dict[int, Optional[bytes]] = {}
        for r, part in enumerate(parts):
            if not part:
                continue
            status = np.ascontiguousarray(host[r, : len(part), 2048:]).view(np.int32).reshape(-1)
            rows = host[r]
            for i, (sidx, _) in enumerate(part):
                stt = int(status[i])
                merged[sidx] = rows[i, :2048].tobytes() if stt == _lib.WIN_OK else (b"" if stt == _lib.WIN_EMPTY else None)
        return merged

"""Multi-GPU: streams are independent, so a box of N GPUs is N independent stream partitions.

One process per GPU (``torchrun``); rank r owns the streams with ``partition_of(stream) == r``,
decodes only their windows on its own engine, and the per-tick PCM is brought together with a
HOST-side gather (``torch.distributed`` object/tensor gather over gloo or NCCL's CPU path) - the
decode path has no collective because no window depends on another.  Identical bytes
come out regardless of how streams are partitioned.
"""
from __future__ import annotations

from typing import Callable, Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

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

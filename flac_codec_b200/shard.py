"""Multi-GPU sharding of the encode/decode path: host-side planning and placement, no data-path collective.

A frame depends on no other frame's samples (fixed block size, frame number = block index; src/encode.rs:2285),
so work shards by whole tracks or by contiguous block ranges of one stream.  The only cross-rank step is the one
the north star names: gather the per-rank compressed frame sizes, exclusive-scan them, and place every rank's
frames at its base offset in the output stream.  That exchange is HOST-side: the size vectors (a few bytes per
frame) travel between the rank processes over a gloo group (TCP/shared memory between host processes) -- never over
NCCL, never through device memory.  (Inside ONE process the same scan is plain C++: flacb200_encode_batch, batch.cpp.)
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


def track_range(n_tracks: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [first, last) range of whole tracks for `rank` (C4: 1024 tracks -> 128 per GPU at 8)."""
    base, extra = divmod(n_tracks, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


@dataclass
class BlockRange:
    first_block: int        # = first_frame_number of the rank's segment
    n_blocks: int
    pcm_offset: int         # inter-channel sample index of the first block
    n_pcm_frames: int       # inter-channel samples in the range (only the last range may end in a short block)


def block_range(n_pcm_frames: int, block_size: int, rank: int, world: int) -> BlockRange:
    """Contiguous block range of ONE stream for `rank`; boundaries fall on block multiples so that every rank but the
    last sees only whole blocks and frame numbers stay global."""
    n_blocks = (n_pcm_frames + block_size - 1) // block_size
    first, last = track_range(n_blocks, rank, world)
    start = first * block_size
    end = min(last * block_size, n_pcm_frames)
    return BlockRange(first, last - first, start, max(end - start, 0))


@dataclass
class Placement:
    base_offset: int              # byte offset of this rank's first frame in the stream's frame area
    total_bytes: int              # all ranks
    rank_bytes: List[int]
    frame_sizes: np.ndarray       # every frame of the stream, in stream order (u32)
    min_frame: int
    max_frame: int

    def frame_offsets(self) -> np.ndarray:
        """Exclusive scan of all frame sizes = byte offset of every frame (seek points, src/encode.rs:1999-2003)."""
        out = np.zeros(self.frame_sizes.size + 1, dtype=np.int64)
        np.cumsum(self.frame_sizes.astype(np.int64), out=out[1:])
        return out[:-1]


_HOST_GROUP = None


def host_group():
    """The gloo process group the size exchange runs on.  When the job's default group is NCCL (one rank per GPU, as
    bench.py sets it up) a second, host-only group is created once; a gloo default group is used as it is."""
    global _HOST_GROUP
    import torch.distributed as dist

    if dist.get_backend() != "nccl":
        return None
    if _HOST_GROUP is None:
        _HOST_GROUP = dist.new_group(backend="gloo")
    return _HOST_GROUP


def place(local_sizes: Sequence[int], rank: int = 0, world: int = 1, group=None) -> Placement:
    """Gather + exclusive scan of per-rank frame sizes, on the host.  Ranks hold consecutive block ranges in rank order."""
    local = np.ascontiguousarray(local_sizes, dtype=np.uint32)
    if world == 1:
        parts = [local]
    else:
        import torch
        import torch.distributed as dist

        if group is None:
            group = host_group()
        if dist.get_backend(group) == "nccl":
            raise RuntimeError("shard.place exchanges frame sizes between host processes: pass a gloo group")
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([local.size], dtype=torch.int64), group=group)
        counts = [int(c.item()) for c in counts]
        width = max(max(counts), 1)
        mine = torch.zeros(width, dtype=torch.int32)
        mine[: local.size] = torch.from_numpy(local.view(np.int32))
        allv = [torch.zeros(width, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(allv, mine, group=group)
        parts = [allv[r][: counts[r]].numpy().view(np.uint32).copy() for r in range(world)]
    rank_bytes = [int(p.astype(np.int64).sum()) for p in parts]
    sizes = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint32)
    ok = sizes[(sizes != 0) & (sizes < (1 << 24) - 1)]     # STREAMINFO min/max rule (src/encode.rs:2413-2436)
    return Placement(base_offset=sum(rank_bytes[:rank]), total_bytes=sum(rank_bytes), rank_bytes=rank_bytes, frame_sizes=sizes,
                     min_frame=int(ok.min()) if ok.size else 0, max_frame=int(ok.max()) if ok.size else 0)


def write_at(path: str, offset: int, data) -> None:
    """Each rank writes its frames straight to `offset` of the shared output (no gather of the frame bytes)."""
    fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o644)
    try:
        view = memoryview(data).cast("B")
        done = 0
        while done < len(view):
            done += os.pwrite(fd, view[done:], offset + done)
    finally:
        os.close(fd)


def encode_stream_sharded(engine, opt, rate: int, bps: int, channels: int, pcm_bytes_le: np.ndarray, n_pcm_frames: int,
                          rank: int, world: int, group=None):
    """One stream over `world` GPUs: this rank encodes its block range of the packed little-endian PCM and learns where
    its frames go.  Returns (frames bytes of this rank, Placement)."""
    from . import _abi

    br = block_range(n_pcm_frames, opt.c.block_size, rank, world)
    fb = channels * ((bps + 7) // 8)
    if br.n_pcm_frames:
        mine = pcm_bytes_le[br.pcm_offset * fb:(br.pcm_offset + br.n_pcm_frames) * fb]
        data, sizes, total = engine.encode(opt, rate, bps, channels, mine, mine.nbytes, _abi.PCM_BYTES_LE,
                                           [(0, br.n_pcm_frames, br.first_block)])
    else:
        data, sizes = np.zeros(0, dtype=np.uint8), np.zeros(0, dtype=np.uint32)
    return data, place(sizes, rank, world, group)

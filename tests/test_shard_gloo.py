"""The N>1 host path on CPU: world_size-2 gloo run of the shard planner + frame-size gather/scan/placement
(flac_codec_b200/shard.py).  Each rank encodes its block range with the oracle standing in for the GPU engine
(test-only) and writes its frames at its base offset; the assembled frame area must equal the single-process
stream, and the seek offsets / STREAMINFO min-max must equal what the single-process encoder tracks."""
import os
import socket
import sys

import numpy as np
import pytest

from flacb200_testutil import ROOT, synth_pcm


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, n, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from flac_codec_b200 import shard
    from oracle import oracle as fo

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rate, bps, ch = 44100, 16, 2
        x = synth_pcm(4, ch, n, rate, bps)
        opt = fo.options("default")
        br = shard.block_range(n, opt.block_size, rank, world)
        mine = x[br.pcm_offset:br.pcm_offset + br.n_pcm_frames].reshape(-1)
        if br.n_pcm_frames:
            data, sizes = fo.encode_frames_only(opt, rate, bps, ch, mine, first_frame_number=br.first_block)
        else:
            data, sizes = b"", np.zeros(0, dtype=np.uint32)
        pl = shard.place(sizes, rank, world)
        shard.write_at(path, pl.base_offset, data)
        dist.barrier()
        q.put((rank, pl.base_offset, pl.total_bytes, pl.rank_bytes, pl.frame_sizes.tolist(), pl.min_frame, pl.max_frame))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [4096 * 7 + 100, 4096 * 2, 1000])
def test_two_rank_stream_assembly(tmp_path, n):
    import torch.multiprocessing as mp

    from oracle import oracle as fo

    world, port = 2, _free_port()
    path = str(tmp_path / "frames.bin")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = synth_pcm(4, 2, n, 44100, 16).reshape(-1)
    ref, ref_sizes = fo.encode_frames_only(fo.options("default"), 44100, 16, 2, x)
    with open(path, "rb") as f:
        assert f.read() == ref
    (_, base0, total0, rb0, sizes0, mn0, mx0), (_, base1, total1, rb1, sizes1, mn1, mx1) = res
    assert base0 == 0 and base1 == rb0[0] and total0 == total1 == len(ref) and rb0 == rb1
    assert sizes0 == sizes1 == ref_sizes.tolist()
    assert (mn0, mx0) == (int(ref_sizes.min()), int(ref_sizes.max()))


def test_planner_covers_everything_once():
    from flac_codec_b200 import shard

    for n_tracks, world in [(1024, 8), (1024, 3), (5, 8), (1, 1)]:
        got = [shard.track_range(n_tracks, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == n_tracks
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
        assert max(b - a for a, b in got) - min(b - a for a, b in got) <= 1
    for n, bs, world in [(8_640_000, 4096, 8), (4096 * 3 + 1, 4096, 2), (100, 4096, 4), (65536, 16, 7)]:
        rs = [shard.block_range(n, bs, r, world) for r in range(world)]
        assert sum(r.n_pcm_frames for r in rs) == n
        assert sum(r.n_blocks for r in rs) == (n + bs - 1) // bs
        pos = 0
        for r in rs:
            if r.n_pcm_frames:
                assert r.pcm_offset == pos and r.first_block * bs == pos
                pos += r.n_pcm_frames
        assert all(r.n_pcm_frames % bs == 0 for r in rs[:-1] if r.n_pcm_frames and r is not [q for q in rs if q.n_pcm_frames][-1])


def test_placement_single_rank():
    from flac_codec_b200 import shard

    pl = shard.place([10, 20, 30])
    assert pl.base_offset == 0 and pl.total_bytes == 60 and pl.frame_offsets().tolist() == [0, 10, 30]
    assert (pl.min_frame, pl.max_frame) == (10, 30)

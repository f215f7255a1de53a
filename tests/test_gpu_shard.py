"""GPU side of the sharded path: block ranges of ONE stream encoded separately (global frame numbers through
first_frame_number) concatenate to exactly the single-shot stream; with >= 2 visible GPUs the same is done by two
NCCL ranks, one engine per GPU, through flac_codec_b200.shard.encode_stream_sharded."""
import os
import socket
import sys

import numpy as np
import pytest

from flacb200_testutil import ROOT, synth_pcm

pytestmark = pytest.mark.gpu


def test_block_ranges_concatenate_to_the_stream():
    from flac_codec_b200 import Engine, Options, _abi, shard
    from oracle import oracle as fo

    rate, bps, ch, n = 48000, 24, 2, 4096 * 9 + 777
    x = synth_pcm(9, ch, n, rate, bps).reshape(-1)
    raw = np.frombuffer(fo.samples_to_bytes(x, 3), dtype=np.uint8)
    ref, ref_sizes = fo.encode_frames_only(fo.options("best"), rate, bps, ch, x)
    eng = Engine(0)
    for world in (1, 2, 3, 8, 16):
        parts, sizes = [], []
        for r in range(world):
            br = shard.block_range(n, 4096, r, world)
            if not br.n_pcm_frames:
                continue
            mine = raw[br.pcm_offset * 6:(br.pcm_offset + br.n_pcm_frames) * 6]
            d, s, t = eng.encode(Options.best(), rate, bps, ch, mine, mine.nbytes, _abi.PCM_BYTES_LE, [(0, br.n_pcm_frames, br.first_block)])
            parts.append(d.tobytes())
            sizes += s.tolist()
        assert b"".join(parts) == ref, world
        assert sizes == ref_sizes.tolist()
    eng.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, n, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from flac_codec_b200 import Engine, Options, shard
    from oracle import oracle as fo

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rate, bps, ch = 48000, 24, 2
        x = synth_pcm(9, ch, n, rate, bps).reshape(-1)
        raw = np.frombuffer(fo.samples_to_bytes(x, 3), dtype=np.uint8)
        eng = Engine(rank)
        data, pl = shard.encode_stream_sharded(eng, Options.best(), rate, bps, ch, raw, n, rank, world)
        shard.write_at(path, pl.base_offset, data)
        dist.barrier()
        q.put((rank, pl.total_bytes, pl.frame_sizes.tolist()))
        eng.close()
    finally:
        dist.destroy_process_group()


def test_two_gpus_one_stream(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    from oracle import oracle as fo

    n, world, port = 4096 * 21 + 5, 2, _free_port()
    path = str(tmp_path / "frames.bin")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x = synth_pcm(9, 2, n, 48000, 24).reshape(-1)
    ref, ref_sizes = fo.encode_frames_only(fo.options("best"), 48000, 24, 2, x)
    with open(path, "rb") as f:
        assert f.read() == ref
    assert res[0][1] == res[1][1] == len(ref) and res[0][2] == res[1][2] == ref_sizes.tolist()

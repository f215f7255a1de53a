"""Whole-file batches over one or several GPUs (flacb200_encode_batch / flacb200_decode_batch): the file-level fan-out the
reference's examples do with rayon (examples/flac2wav.rs:31-38, examples/flac-split.rs:84-87)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi
from .engine import Options
from .stream import _writer_options


def encode_files(tracks: Sequence[tuple], options: Options, devices: Optional[Sequence[int]] = None, out_buffers=None):
    """tracks: (pcm ndarray | (address, nbytes), n_pcm_frames, sample_rate, bits_per_sample, channels, pcm_kind).
    Returns a list of (flac bytes | None, status, md5).  out_buffers: optional list of (address, capacity) the files are
    written into (e.g. pinned memory); the returned `bytes` entries are then memoryviews of those buffers."""
    L = _abi.lib()
    n = len(tracks)
    arr = (_abi.Track * max(n, 1))()
    keep = []
    for i, (pcm, npcm, rate, bps, ch, kind) in enumerate(tracks):
        if isinstance(pcm, tuple):
            addr = pcm[0]
        else:
            a = np.ascontiguousarray(pcm)
            keep.append(a)
            addr = a.ctypes.data
        arr[i] = _abi.Track(addr, npcm, rate, bps, ch, kind)
    files = (_abi.File * max(n, 1))()
    if out_buffers is not None:
        for i, (addr, cap) in enumerate(out_buffers):
            files[i].data, files[i].capacity = addr, cap
    wo = _writer_options(options)
    devs = (C.c_int * max(len(devices or []), 1))(*(devices or [0]))
    L.flacb200_encode_batch(arr, n, C.byref(wo), devs, len(devices or [0]), files)
    out = []
    for i in range(n):
        f = files[i]
        data = None
        if f.status == 0:
            data = C.string_at(f.data, f.len) if out_buffers is None else (C.c_uint8 * f.len).from_address(f.data)
        out.append((data, f.status, bytes(f.md5)))
    if out_buffers is None:
        L.flacb200_files_free(files, n)
    return out


def decode_files(flacs: Sequence[bytes], pcm_kind: int = _abi.PCM_BYTES_LE, verify: bool = False, devices: Optional[Sequence[int]] = None):
    """Returns a list of (pcm bytes | None, status, verified, Streaminfo)."""
    L = _abi.lib()
    n = len(flacs)
    imgs = [np.frombuffer(bytes(f), dtype=np.uint8) for f in flacs]
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in imgs])
    lens = (C.c_size_t * max(n, 1))(*[a.size for a in imgs])
    out = (_abi.Pcm * max(n, 1))()
    devs = (C.c_int * max(len(devices or []), 1))(*(devices or [0]))
    L.flacb200_decode_batch(ptrs, lens, n, pcm_kind, int(verify), devs, len(devices or [0]), out)
    res = []
    for i in range(n):
        o = out[i]
        res.append((C.string_at(o.data, o.len) if o.status == 0 else None, o.status, o.verified, o.info))
    L.flacb200_pcm_free(out, n)
    return res

"""Builds libflacb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

Each source is compiled to build/<name>.o (in parallel, only when it or a header changed) and the objects are
linked into flac_codec_b200/libflacb200.so."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "..", "build", "flacb200")
LIB = os.path.join(HERE, "libflacb200.so")
SOURCES = ["engine.cu", "encode_kernels.cu", "decode_kernels.cu", "synth.cu", "stream.cpp", "encode_frame.cu", "encode_analyze.cu", "encode_lpc.cu", "decode_parse.cu", "md5.cu", "md5_mb.cpp", "batch.cpp"]
HEADERS = [os.path.join(HERE, "..", "include", h) for h in ("flacb200.h", "flacb200_stream.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _ccbin() -> list:
    # the distro g++ (the image's /opt/gcc wrapper lacks some runtime pieces)
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _shared_deps() -> list:
    return HEADERS + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inl", ".h", ".hpp"))]


def _obj(src: str) -> str:
    return os.path.join(OBJ, os.path.splitext(src)[0] + ".o")


def _stale(src: str) -> bool:
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(d) > t for d in [os.path.join(CSRC, src)] + _shared_deps())


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + HEADERS + [os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host"))]
    return any(os.path.getmtime(d) > t for d in deps) or not os.path.exists(os.path.join(HERE, "..", "build", "flacb200_wav2flac"))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    todo = [s for s in SOURCES if force or _stale(s)]

    def compile_one(src: str):
        cmd = [_nvcc(), *NVCC_FLAGS, *_ccbin(), "-c", "-o", _obj(src), os.path.join(CSRC, src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd, cwd=CSRC)

    with ThreadPoolExecutor(max_workers=max(len(todo), 1)) as ex:
        list(ex.map(compile_one, todo))
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", *_ccbin(), "-shared", "-o", LIB] + [_obj(s) for s in SOURCES]
    subprocess.check_call(link, cwd=CSRC)
    build_host_tools()
    return LIB


HOST = os.path.join(HERE, "host")
WAV2FLAC = os.path.join(HERE, "..", "build", "flacb200_wav2flac")


def build_host_tools() -> str:
    """The C++ host layer above the C ABI: flacb200.hpp facades + the wav2flac/flac2wav front end."""
    os.makedirs(os.path.dirname(WAV2FLAC), exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-std=c++17", "-O2", "-Wall", "-o", WAV2FLAC, os.path.join(HOST, "wav2flac.cpp"), "-L" + HERE,
                           "-lflacb200", "-Wl,-rpath,$ORIGIN/../flac_codec_b200"])
    return WAV2FLAC


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)

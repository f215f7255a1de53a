#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 FLAC frame engine (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]            GPU arm (one process per GPU under torchrun)
  python bench.py --impl reference [--gpus N] [--steps K] ...    reference CPU arm (the oracle port, all host cores)

Workload (BASELINE.json configs[3], "C4"): batch encode of synthetic 3-minute 48 kHz / 24-bit stereo
tracks at Options::best() (block 4096, max LPC order 12, partition order <= 6), 128 tracks per GPU, so
8 GPUs encode the 1024 tracks the config names; tracks shard across ranks with no collective (weak
scaling).  A step = one pass of the encode path over the rank's whole batch.
  value : Msamples/s (single-channel samples), PCM resident in HBM, frames left in HBM
  e2e   : same metric through the C ABI with HOST buffers (pinned PCM in, frames out)
  decode: the decode path over the frames produced by the encode step (reported beside the headline)
Only the cpu_baseline / --impl reference legs execute anything under oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE, BPS, CH = 48000, 24, 2
SECONDS = 180
TRACKS_PER_GPU = 128
SEED = 20261017
METRIC = "encode_msamples_per_s"
UNIT = "Msamples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks-per-gpu", type=int, default=TRACKS_PER_GPU)
    ap.add_argument("--seconds", type=int, default=SECONDS)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C2/C3/C5 legs (tests/config_legs.py)")
    ap.add_argument("--no-files", action="store_true", help="skip the e2e_files leg (flacb200_encode_batch)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": f"C4 batch encode: {args.tracks_per_gpu} tracks/GPU x {args.seconds} s {RATE // 1000} kHz/{BPS}-bit "
                    f"stereo synthetic PCM, Options::best (block 4096, LPC<=12, partition order<=6)",
        "tracks_total": args.tracks_per_gpu * n_gpus,
        "tracks_per_gpu": args.tracks_per_gpu,
        "track_seconds": args.seconds,
        "sample_rate": RATE, "bits_per_sample": BPS, "channels": CH,
        "options": "best",
        "parallelism": f"tracks sharded over {n_gpus} GPU(s), no collective",
        "l2": "inputs (6.6 GB/GPU) far exceed the 126 MB L2; no flush needed",
    }


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t_begin = index, [], None, 0.0

    def begin(self):
        """The timed region starts here: only samples that arrive from now on count (nvidia-smi itself is started before the
        warm-up steps -- it needs a few hundred ms to print its first line, longer than a short timed region)."""
        self.t_begin = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        inside = [r for t, r in self.rows if t >= self.t_begin]
        # a timed region shorter than the sampling period: the samples of the warm-up steps (same kernels, same load) stand in
        where = "timed region" if inside else "warm-up steps (the timed region was shorter than one sampling period)"
        for r in (inside or [r for _, r in self.rows]):
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the samples taken under load (upper half of the clock samples is the loaded region)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm), "sampled_during": where}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_encode_rate(pcm_i32_tracks, cores, target_seconds):
    """Times the oracle (CPU restatement of the reference encoder, frames encoded concurrently the way
    rayon/file-level parallelism would) on a bounded sample: the given track excerpts, encoded again and again until
    about `target_seconds` of wall time have passed.  Returns (Msamples/s, sample description, oracle frames of one pass)."""
    from oracle import oracle as fo

    opt = fo.options("best")
    done, t_total, passes, first = 0, 0.0, 0, []
    while t_total < target_seconds and passes < 64:
        for x in pcm_i32_tracks:
            t0 = time.perf_counter()
            data, sizes = fo.encode_frames_only(opt, RATE, BPS, CH, x, nthreads=cores)
            t_total += time.perf_counter() - t0
            done += x.size
            if passes == 0:
                first.append((data, sizes))
        passes += 1
    per_pass = sum(x.size for x in pcm_i32_tracks)
    return (done / t_total / 1e6,
            f"{len(pcm_i32_tracks)} track excerpts x {pcm_i32_tracks[0].size // CH / RATE:.0f} s ({per_pass // CH} PCM frames) of the same "
            f"workload, {passes} passes, {t_total:.1f} s of wall time on {cores} threads", first)


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` at this workload's launch-group size, from the
    newest committed `ncu --set full` summary under profiles/ (bytes), or None."""
    import csv
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full_summary.csv")), reverse=True):
        try:
            with open(path, newline="") as f:
                rows = list(csv.reader(f))
            hdr, units = rows[0], rows[1]
            ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for r in rows[2:]:
                if kernel.split("+")[0] in r[ik]:
                    return float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        except (OSError, ValueError, KeyError, IndexError):
            continue
    return None


def ncu_pipes(kernel):
    """Issue-slot and pipe utilisation (% of peak while active) of `kernel` from the newest committed ncu summary, or None:
    what bounds a kernel that is far from the HBM roofline (the north star asks for FP64 / INT32 pipe utilisation)."""
    import csv
    import glob

    keys = {"issue_slots": "smsp__issue_active.avg.pct_of_peak_sustained_active", "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fp64": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "warps_active": "sm__warps_active.avg.pct_of_peak_sustained_active"}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full_summary.csv")), reverse=True):
        try:
            with open(path, newline="") as f:
                rows = list(csv.reader(f))
            hdr = rows[0]
            ik = hdr.index("Kernel Name")
            for r in rows[2:]:
                if kernel.split("+")[0] in r[ik]:
                    return {"source": os.path.basename(path), **{k: float(r[hdr.index(v)]) for k, v in keys.items() if v in hdr}}
        except (OSError, ValueError, IndexError):
            continue
    return None


def host_tracks_numpy(n_tracks, seconds):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from flacb200_testutil import synth_pcm

    return [synth_pcm(t, CH, RATE * seconds, RATE, BPS, SEED).reshape(-1) for t in range(n_tracks)]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # a bounded sample of the workload: the first tracks (60 s of each), generated by the numpy statement of the generator
    from oracle import oracle as fo

    opt = fo.options("best")
    n_tracks = 8
    tracks = host_tracks_numpy(n_tracks, min(args.seconds, 60))
    per_step = sum(x.size for x in tracks)

    def one_step():
        for x in tracks:   # file-level batches, every frame of a track encoded concurrently (OpenMP) on all host cores
            fo.encode_frames_only(opt, RATE, BPS, CH, x, nthreads=cores)

    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt / 1e6
    sample = f"{n_tracks} tracks x {min(args.seconds, 60)} s ({per_step // CH} PCM frames) of the same workload per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32/i64 residuals, f64 LPC analysis", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference crate is Rust and cannot be built in this image; this arm times oracle/ (C restatement "
                "of its encoder, frames encoded concurrently with OpenMP on all host cores)",
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """One process per GPU: keep the rank (and the pinned host buffers it is about to allocate) on the NUMA node its GPU
    hangs off.  With several ranks pulling 55 GB/s each, buffers on the far socket turn the end-to-end step into a
    cross-socket copy.  Returns a short description for the JSON line; does nothing when the topology is not visible."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"numa node {node}: no usable cpus"
        os.sched_setaffinity(0, cpus)
        return f"rank bound to numa node {node} ({len(cpus)} cpus) of GPU {bdf}"
    except (OSError, ValueError, AttributeError, RuntimeError) as e:
        return f"not bound ({type(e).__name__})"


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound (the cpu_baseline leg uses every core)"
    if world > 1:   # the ranks of one host share its cores: each hashes (flacb200_md5_many) with its share
        os.environ.setdefault("FLACB200_HOST_THREADS", str(max((os.cpu_count() or 1) // world, 1)))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL's log (the image sets NCCL_DEBUG=VERSION: a banner on stdout) goes to stderr: rank 0 prints ONE line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from flac_codec_b200 import Engine, Options, _abi

    eng = Engine(local)
    eng.set_keep_info(False)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    opt = Options.best()
    n_tracks, n = args.tracks_per_gpu, RATE * args.seconds
    first_track = rank * n_tracks
    bytes_per_pcm_frame = CH * 3
    pcm_bytes = n_tracks * n * bytes_per_pcm_frame
    samples_per_step = n_tracks * n * CH          # single-channel samples this rank encodes per step
    segs = [(t * n, n, 0) for t in range(n_tracks)]
    d_pcm = eng.device_alloc(pcm_bytes)
    eng.synth_pcm(d_pcm, first_track, n_tracks, n, CH, RATE, BPS, SEED)
    out_cap = pcm_bytes + pcm_bytes // 8 + (1 << 20)
    d_out = eng.device_alloc(out_cap)
    eng.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return eng.encode(opt, RATE, BPS, CH, d_pcm, pcm_bytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.DEVICE, out=d_out,
                          out_capacity=out_cap, out_location=_abi.DEVICE, want_sizes=False)

    eng.set_profiling(True)
    clocks = Clocks(local)
    clocks.start()
    for _ in range(args.warmup):
        _, _, flac_bytes = step()
    barrier()
    clocks.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = np.zeros(8)
    kernel_launches = np.zeros(8, dtype=np.int64)
    launches = 0
    ev0.record(stream)
    for _ in range(args.steps):
        _, _, flac_bytes = step()
        tm = eng.timings()
        kernel_ms += np.array(list(tm.kernel_ms))
        kernel_launches += np.array(list(tm.kernel_launches))
        launches += tm.launches
    ev1.record(stream)
    barrier()
    clk = clocks.stop()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    value = samples_per_step * world / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel: algorithmic bytes of one launch / its average duration ----
    # the C4 shape runs the CTA-per-frame kernels (encode_lpc.cu, encode_analyze.cu, encode_frame.cu); slot 0 (k_planes) is
    # only used by the generic path
    names = ["k_planes", "k_lpc4", "k_analyze3", "k_decide+k_scan", "k_pack3"]
    top = int(np.argmax(kernel_ms[:5]))
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes_step = pcm_bytes + flac_bytes          # PCM read once + frames written once
    groups = max(int(kernel_launches[top] // max({0: 1, 1: 1, 2: 1, 3: 2, 4: 2}[top], 1)), 1)
    avg_ms = kernel_ms[top] / groups                 # average duration of one launch (group) of the top kernel
    alg_bytes_launch = alg_bytes_step * args.steps / groups
    achieved = alg_bytes_launch / (avg_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": names[top], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": ncu_traffic(names[top]), "pipe_util_pct": ncu_pipes(names[top]), "algorithmic_bytes_per_launch": alg_bytes_launch, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
        "algorithmic_bytes_per_sample": alg_bytes_step / samples_per_step,
        "kernel_share_of_step": {names[k]: kernel_ms[k] / max(kernel_ms[:5].sum(), 1e-9) for k in range(5)},
        "kernel_ms_per_step": {names[k]: kernel_ms[k] / args.steps for k in range(5)},
        "whole_path_frac": (alg_bytes_step / (ms_per_step * 1e-3) / 1e9) / peak,
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i32/i64 residuals, f64 LPC analysis", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clk, "gpu_launches": int(launches), "host_affinity": numa, "roofline": roofline,
        "compression_ratio": flac_bytes / pcm_bytes,
    }

    # ---- decode leg: the frames just produced, decoded back on the device (PCM must equal the input) ----
    if not args.no_decode:
        eng.set_profiling(False)
        _, sizes, total = eng.encode(opt, RATE, BPS, CH, d_pcm, pcm_bytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.DEVICE,
                                     out=d_out, out_capacity=out_cap, out_location=_abi.DEVICE, want_sizes=True)
        per_track = (n + 4095) // 4096
        offs = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
        dsegs = [(int(offs[t * per_track]), int(offs[(t + 1) * per_track] - offs[t * per_track]), t * n, n) for t in range(n_tracks)]
        d_back = eng.device_alloc(pcm_bytes)
        eng.set_profiling(True)

        def dstep():
            return eng.decode(RATE, BPS, CH, 4096, d_out, total, dsegs, d_back, pcm_bytes, _abi.PCM_BYTES_LE,
                              frames_location=_abi.DEVICE, pcm_location=_abi.DEVICE)

        for _ in range(2):
            nf, ns = dstep()
        barrier()
        dk = np.zeros(8)
        ev0.record(stream)
        for _ in range(args.steps):
            nf, ns = dstep()
            dk += np.array(list(eng.timings().kernel_ms))
        ev1.record(stream)
        barrier()
        dms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dms, op=dist.ReduceOp.MAX)
        dms_step = float(dms.item()) / args.steps
        # bit-exactness at full size: a 64-bit checksum of checksums over the PCM bytes, input vs decoded
        def dev_u8(ptr, nbytes):
            class _W:
                __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            return torch.as_tensor(_W(), device="cuda")

        src, back = dev_u8(d_pcm, pcm_bytes), dev_u8(d_back, pcm_bytes)
        exact = bool(torch.equal(src, back)) and ns == n_tracks * n
        # packed 24-bit stereo output: k_parse, CRC-16 + the frame walk, then the fused restoration + emit pass (engine.cu)
        dnames = ["k_find+k_scan", "k_parse", "k_crc16f+k_chain_fast", "k_restore_emit", "fallback (k_chain, k_restore+k_emit)"]
        line["decode"] = {"metric": "decode_msamples_per_s", "value": samples_per_step * world / (dms_step * 1e-3) / 1e6, "unit": UNIT,
                          "ms_per_step": dms_step, "frames": int(nf), "bit_exact_vs_input": exact,
                          "kernel_ms_per_step": {dnames[k]: dk[k] / args.steps for k in range(5) if dnames[k]},
                          "hbm_frac": ((pcm_bytes + total) / (dms_step * 1e-3) / 1e9) / peak}
        if not exact:
            raise SystemExit("bench.py: GPU decode of the GPU-encoded frames is not bit-exact")
        eng.device_free(d_back)

    # ---- e2e: host PCM -> frames in host memory, copies inside the timed region ----
    if not args.no_e2e:
        L = _abi.lib()
        h_pcm_p = L.flacb200_host_alloc(pcm_bytes)
        h_out_p = L.flacb200_host_alloc(out_cap)
        if not h_pcm_p or not h_out_p:
            raise SystemExit("bench.py: pinned host allocation failed")
        eng.memcpy(h_pcm_p, d_pcm, pcm_bytes, 2)
        eng.set_profiling(False)
        # the link itself, for context: plain pinned copies of the same buffers
        scratch_d = eng.device_alloc(pcm_bytes)
        t0 = time.perf_counter()
        eng.memcpy(scratch_d, h_pcm_p, pcm_bytes, 1)
        h2d_gbs = pcm_bytes / (time.perf_counter() - t0) / 1e9
        t0 = time.perf_counter()
        eng.memcpy(h_out_p, scratch_d, min(out_cap, pcm_bytes), 2)
        d2h_gbs = min(out_cap, pcm_bytes) / (time.perf_counter() - t0) / 1e9
        eng.device_free(scratch_d)

        # the link itself under the same load as the step: every rank uploads its PCM and downloads as many bytes as its
        # frames take, both directions at once, all ranks at the same time (barrier-aligned) -- the roofline of `e2e`
        def link_probe(nbytes_up, nbytes_down, reps=2):
            import ctypes

            up_h = torch.frombuffer((ctypes.c_uint8 * nbytes_up).from_address(h_pcm_p), dtype=torch.uint8)
            dn_h = torch.frombuffer((ctypes.c_uint8 * nbytes_down).from_address(h_out_p), dtype=torch.uint8)
            up_d = torch.empty(nbytes_up, dtype=torch.uint8, device="cuda")
            dn_d = torch.empty(nbytes_down, dtype=torch.uint8, device="cuda")
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            best = 1e30
            for _ in range(reps + 1):
                barrier()
                t0 = time.perf_counter()
                with torch.cuda.stream(s_up):
                    up_d.copy_(up_h, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    dn_h.copy_(dn_d, non_blocking=True)
                s_up.synchronize()
                s_dn.synchronize()
                dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                best = min(best, float(dt.item()))
            del up_d, dn_d
            return best

        link_s = link_probe(pcm_bytes, int(flac_bytes))

        from flac_codec_b200 import shard

        def e2e_step():
            r = eng.encode(opt, RATE, BPS, CH, h_pcm_p, pcm_bytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.HOST,
                           out=h_out_p, out_capacity=out_cap, out_location=_abi.HOST, want_sizes=True)
            if world > 1:   # the one cross-GPU step: host-side gather + scan of the frame sizes -> where this rank's frames go
                shard.place(r[1], rank, world)
            return r

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            _, sizes, total = e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_ms = float(dt.item()) / args.e2e_steps * 1e3
        line["e2e"] = {"value": samples_per_step * world / (e2e_ms * 1e-3) / 1e6, "unit": UNIT,
                       "h2d_bytes_per_step": int(pcm_bytes), "d2h_bytes_per_step": int(total + 4 * len(sizes)),
                       "ms_per_step": e2e_ms, "steps": args.e2e_steps, "pinned_copy_gbs": {"h2d": h2d_gbs, "d2h": d2h_gbs},
                       "link": {"ms": link_s * 1e3, "aggregate_gbs": world * (pcm_bytes + flac_bytes) / link_s / 1e9,
                                "what": "all ranks at once: pinned upload of the step's PCM and download of as many bytes as its frames, both directions concurrently"},
                       "link_frac": link_s * 1e3 / e2e_ms,
                       "api": "flacb200_encode(host PCM -> host frames + frame sizes)"
                              + ("; + host-side gather/scan of the frame sizes over gloo (shard.place)" if world > 1 else "")}
        # decode, end to end: the frames just downloaded (pinned host memory) -> PCM in pinned host memory
        if not args.no_decode and "decode" in line:
            per_track = (n + 4095) // 4096
            offs = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
            dsegs = [(int(offs[t * per_track]), int(offs[(t + 1) * per_track] - offs[t * per_track]), t * n, n) for t in range(n_tracks)]

            def e2e_dstep():
                return eng.decode(RATE, BPS, CH, 4096, h_out_p, total, dsegs, h_pcm_p, pcm_bytes, _abi.PCM_BYTES_LE,
                                  frames_location=_abi.HOST, pcm_location=_abi.HOST)

            e2e_dstep()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                nf, ns = e2e_dstep()
            torch.cuda.synchronize()
            ddt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ddt, op=dist.ReduceOp.MAX)
            d_ms = float(ddt.item()) / args.e2e_steps * 1e3
            line["decode"]["e2e"] = {"value": samples_per_step * world / (d_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": d_ms,
                                     "h2d_bytes_per_step": int(total), "d2h_bytes_per_step": int(pcm_bytes),
                                     "api": "flacb200_decode(host frames -> host PCM)", "pcm_frames_decoded": int(ns)}

    # ---- CPU baseline on rank 0 at N=1: the oracle on all host cores, and parity of EVERY frame of the shard ----
    # With the e2e leg's host buffers at hand this is one pass of the oracle over the whole shard (all tracks, every
    # frame compared with what the GPU returned through the C ABI); without them (--no-e2e) a bounded sample.
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        from oracle import oracle as fo

        if not args.no_e2e:
            import ctypes

            per_track = (n + 4095) // 4096
            goff = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
            h_pcm = np.ctypeslib.as_array((ctypes.c_uint8 * pcm_bytes).from_address(h_pcm_p))
            h_out = np.ctypeslib.as_array((ctypes.c_uint8 * int(total)).from_address(h_out_p))
            tb = n * bytes_per_pcm_frame
            ofo = fo.options("best")
            t_cpu, frames, same, ref_total, tracks_same = 0.0, 0, 0, 0, 0
            for t in range(n_tracks):
                x = fo.bytes_to_samples(h_pcm[t * tb:(t + 1) * tb].tobytes(), 3)
                t0 = time.perf_counter()
                data, rsizes = fo.encode_frames_only(ofo, RATE, BPS, CH, x, nthreads=cores)
                t_cpu += time.perf_counter() - t0
                ref_total += len(data)
                g = h_out[goff[t * per_track]:goff[(t + 1) * per_track]].tobytes()
                frames += len(rsizes)
                if g == data:
                    same += len(rsizes)
                    tracks_same += 1
                else:   # count the frames that do agree
                    roff = np.concatenate([[0], np.cumsum(rsizes.astype(np.int64))])
                    g0 = int(goff[t * per_track])
                    for f in range(len(rsizes)):
                        a, b = int(goff[t * per_track + f]) - g0, int(goff[t * per_track + f + 1]) - g0
                        same += int(g[a:b] == data[roff[f]:roff[f + 1]])
            line["cpu_baseline"] = {"value": samples_per_step / t_cpu / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"the whole shard once: {n_tracks} tracks x {args.seconds} s ({samples_per_step} samples), "
                                              f"{t_cpu:.1f} s of oracle time on {cores} threads (frames of a track encoded concurrently)"}
            line["parity_vs_cpu_port"] = {"frames_compared": frames, "byte_identical_frames": same, "identical_fraction": same / max(frames, 1),
                                          "tracks_identical": tracks_same, "fraction_of_workload_compared": 1.0,
                                          "size_delta": (int(total) - ref_total) / max(ref_total, 1),
                                          "note": "encoder bytes are pinned to the oracle (CPU restatement); the reference's own tests only round-trip"}
        else:
            take = min(n, RATE * 60)
            take -= take % 4096
            ntr = min(n_tracks, 8)
            xs = []
            for t in range(ntr):
                host = np.zeros(take * bytes_per_pcm_frame, dtype=np.uint8)
                eng.memcpy(host, d_pcm + t * n * bytes_per_pcm_frame, host.nbytes, 2)
                xs.append(fo.bytes_to_samples(host.tobytes(), 3))
            v, sample, ref = cpu_encode_rate(xs, cores, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            gdata, gsizes, gtotal = eng.encode(opt, RATE, BPS, CH, d_pcm, pcm_bytes, _abi.PCM_BYTES_LE, [(t * n, take, 0) for t in range(ntr)],
                                               pcm_location=_abi.DEVICE)
            gbytes = gdata.tobytes()
            goff = np.concatenate([[0], np.cumsum(gsizes.astype(np.int64))])
            frames = same = 0
            ref_total = 0
            k = 0
            for data, rsizes in ref:
                roff = np.concatenate([[0], np.cumsum(rsizes.astype(np.int64))])
                ref_total += len(data)
                for f in range(len(rsizes)):
                    frames += 1
                    if k < len(gsizes) and gbytes[goff[k]:goff[k + 1]] == data[roff[f]:roff[f + 1]]:
                        same += 1
                    k += 1
            line["parity_vs_cpu_port"] = {"frames_compared": frames, "byte_identical_frames": same, "identical_fraction": same / max(frames, 1),
                                          "fraction_of_workload_compared": frames / max(n_tracks * ((n + 4095) // 4096), 1),
                                          "size_delta": (gtotal - ref_total) / max(ref_total, 1)}
    # ---- e2e_files: host PCM -> complete .flac files in host memory (fLaC, STREAMINFO with MD5, SEEKTABLE, PADDING, frames)
    # for every track of the shard through flacb200_encode_batch; MD5 on the host threads beside the GPU ----
    if not args.no_e2e and not args.no_files:
        import ctypes
        import hashlib

        from flac_codec_b200.batch import encode_files

        tb = n * bytes_per_pcm_frame
        tracks = [((h_pcm_p + t * tb, tb), n, RATE, BPS, CH, _abi.PCM_BYTES_LE) for t in range(n_tracks)]
        per_file = out_cap // n_tracks
        bufs = [(h_out_p + t * per_file, per_file) for t in range(n_tracks)]
        files = encode_files(tracks, opt, devices=[local], out_buffers=bufs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            files = encode_files(tracks, opt, devices=[local], out_buffers=bufs)
        fdt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(fdt, op=dist.ReduceOp.MAX)
        f_ms = float(fdt.item()) / args.e2e_steps * 1e3
        ok = all(st == 0 for _, st, _ in files)
        md5_ok = None
        if rank == 0:   # the signatures against hashlib (a sample: 51.8 MB per track), the STREAMINFO copy against the result
            h_pcm = np.ctypeslib.as_array((ctypes.c_uint8 * pcm_bytes).from_address(h_pcm_p))
            md5_ok = all(hashlib.md5(h_pcm[t * tb:(t + 1) * tb].tobytes()).digest() == files[t][2] == bytes(files[t][0][26:42])
                         for t in range(0, n_tracks, max(n_tracks // 8, 1)))
        line["e2e_files"] = {"value": samples_per_step * world / (f_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": f_ms,
                             "files_per_step": n_tracks * world, "all_ok": bool(ok), "md5_vs_hashlib_sample": md5_ok,
                             "file_bytes_per_step": int(sum(len(f[0]) for f in files if f[0] is not None)),
                             "api": "flacb200_encode_batch(host PCM tracks -> complete .flac images; MD5 by flacb200_md5_many on the host threads)"}
        if world == 1 and rank == 0 and not args.no_cpu_baseline:   # whole files against the oracle's FlacByteWriter restatement
            from oracle import oracle as fo

            same = 0
            for t in (0, n_tracks - 1):
                x = fo.bytes_to_samples(h_pcm[t * tb:(t + 1) * tb].tobytes(), 3)
                ref_file, _ = fo.encode_stream(fo.options("best"), RATE, BPS, CH, x, total_known=True, nthreads=os.cpu_count() or 1)
                same += int(bytes(files[t][0]) == ref_file)
            line["e2e_files"]["files_identical_to_cpu_port"] = f"{same} of 2 compared"
    if not args.no_e2e:
        L.flacb200_host_free(h_pcm_p)
        L.flacb200_host_free(h_out_p)

    eng.device_free(d_pcm)
    eng.device_free(d_out)
    # ---- the other BASELINE configs at their stated sizes (rank 0 at N=1): GPU through the C ABI / stream facades, checked
    # against the oracle (tests/config_legs.py; the same legs are asserted on by tests/test_gpu_configs.py) ----
    if world == 1 and rank == 0 and not args.no_configs and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import config_legs
        from oracle import oracle as fo

        eng.set_profiling(False)
        line["configs"] = config_legs.all_legs(eng, fo, reps=2)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())

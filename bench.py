#!/usr/bin/env python
"""bench.py -- STFT frames/s (1024-pt, 256-hop, fp32) on N B200s, with roofline, end-to-end
(host buffers) and CPU-baseline figures.  Contract: see the task brief / DESIGN.md section 6.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline line: one "step" = one pass of NxSignal.stft over the BASELINE config-2 workload
(8 ch x 600 s @ 48 kHz f32, Hann(1024), hop 256, :valid  ->  899 976 frames) per GPU; N > 1: one
rank per GPU (torchrun), each rank owns its own 8 channels (weak scaling); the only collective is
one NCCL broadcast of the window at setup, through the C ABI (nxs_bcast_coeffs_dev).

The same JSON line also carries (key "multi_gpu") BASELINE configs[2] as a whole -- 1024 ch x 60 s,
nfft 4096, hop 1024, channels sharded over the N ranks -- and config-2 under strong scaling (its 8
channels split by channel, and by frame range with the read-only halo), each with an in-run check
that a shard computed on rank r equals the single-GPU result bit for bit; (key "e2e") the host-
buffer calls for pinned, cudaHostRegister'ed and pageable memory plus the ISTFT (cfg5) and FIR
(cfg4) host entries; and (key "cfg1") the call latency on BASELINE configs[0].
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000
CHANNELS = 8
SECONDS = 600
NFFT = 1024
HOP = 256
L = FS * SECONDS
M = (L - NFFT) // HOP + 1
FRAMES_PER_GPU = CHANNELS * M
ALGO_BYTES = 4 * CHANNELS * L + 8 * CHANNELS * M * NFFT + 4 * NFFT
METRIC = "STFT frames/sec (1024-pt, 256-hop, fp32)"
WORKLOAD = "cfg2: 8ch x 600s @48kHz f32, hann(1024), hop 256, :valid -> 899976 frames per GPU"
# BASELINE configs[2]
C3, L3, N3, H3 = 1024, FS * 60, 4096, 1024
M3 = (L3 - N3) // H3 + 1
BLOCK3 = 64  # channels per input-generation block (one seed per block, so any rank can regenerate any block)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks line sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for _, l in self.lines]
        for l in inside:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm
# ---------------------------------------------------------------------------------------------
def cpu_input(channels, length, seed=1002):
    """cfg2-style input [channels, length], synthesised in f32 chunks (same recipe as tests.util.synth:
    0.25 N(0,1) + two tones) without the f64 temporaries of the test helper."""
    rng = np.random.default_rng(seed)
    x = np.empty((channels, length), dtype=np.float32)
    step = 1 << 22
    for i in range(0, length, step):
        j = min(length, i + step)
        t = np.arange(i, j, dtype=np.float64) / FS
        tone = (0.5 * np.sin(2 * np.pi * 440.0 * t) + 0.5 * np.sin(2 * np.pi * 3000.0 * t)).astype(np.float32)
        for c in range(channels):
            x[c, i:j] = 0.25 * rng.standard_normal(j - i, dtype=np.float32) + tone
    return x


class CpuPort:
    """The oracle's C port (the reference's algorithm restated: f64 recursive radix-2, OpenMP over
    frames).  `full=True`: every step transforms the WHOLE cfg2 workload (8 ch x 600 s, 899 976 frames;
    the reference arm, so that its config is the GPU arm's); else a bounded sample of channel 0.
    Input and the result buffer are built once, outside the timed region."""

    def __init__(self, full, nthreads=0):
        from oracle import c_port
        from oracle import nxsignal_oracle as o

        self.c_port = c_port
        self.w = o.hann(NFFT)
        if nthreads <= 0:  # torchrun exports OMP_NUM_THREADS=1: use every core this process may run on
            nthreads = len(os.sched_getaffinity(0))
        self.cores = nthreads
        if full:
            self.x = cpu_input(CHANNELS, L)
            self.frames = FRAMES_PER_GPU
            self.what = f"the whole cfg2 workload ({CHANNELS} ch x {SECONDS} s, {self.frames} frames)"
        else:
            self.frames = 1 << 18
            self.x = cpu_input(1, self.frames * HOP + NFFT - HOP)
            self.what = f"{self.frames} frames of cfg2 channel 0 ({self.frames * HOP / FS:.0f} s of audio)"
        self.out = np.empty((self.x.shape[0], (self.x.shape[1] - NFFT) // HOP + 1, NFFT), dtype=np.complex64)
        self.out[...] = 0  # touch the pages: page faults are not the transform
        self.sample = self.what
        self.c_port.stft(self.x[:1, : 4096 * HOP + NFFT], self.w, HOP, NFFT, nthreads=nthreads)  # start the threads

    def step(self):
        t = time.perf_counter()
        self.c_port.stft(self.x, self.w, HOP, NFFT, nthreads=self.cores, out=self.out)
        dt = time.perf_counter() - t
        self.sample = f"{self.what}, {dt:.2f} s per step"
        return self.frames / dt, dt


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU algorithm (Nx.BinaryBackend's recursive radix-2 in
    f64, restated in C: oracle/nxs_oracle.c -- the BEAM cannot run here) on all host threads, on
    the GPU arm's config: every step is the whole cfg2 workload."""
    if rank != 0:
        return
    port = CpuPort(full=True)
    times = []
    t_budget = time.perf_counter() + 240.0  # keep the whole run within a few minutes on any host
    for _ in range(args.warmup):
        port.step()
        if time.perf_counter() > t_budget:
            break
    for _ in range(args.steps):
        _, dt = port.step()
        times.append(dt)
        if time.perf_counter() > t_budget:
            break
    value = float(port.frames * len(times) / sum(times)) if times else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)) if times else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (rounded to f32/c64)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "channels_per_gpu": CHANNELS, "fft_length": NFFT, "hop": HOP,
                   "sample_per_step": port.sample, "same_config_as_gpu_arm": True},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": port.cores, "kind": "port", "sample": port.sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU helpers
# ---------------------------------------------------------------------------------------------
class Env:
    """Everything a leg of the bench needs: torch, the C ABI, this rank's device and group."""

    def __init__(self, torch, dist, rank, local_rank, world):
        from nx_signal_b200 import _arrays as A
        from nx_signal_b200 import _lib

        self.torch, self.dist, self.rank, self.local_rank, self.world = torch, dist, rank, local_rank, world
        self.A, self._lib = A, _lib
        self.lib = _lib.lib()
        self.dev = torch.device("cuda", local_rank)
        self.ctx = _lib.context(local_rank)
        self.sptr = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok):
        if self.dist is None:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], device=self.dev, dtype=self.torch.int32)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def stft_dev(self, x, w, z, nfft, hop):
        c, length = x.shape
        rc = self.lib.nxs_stft_f32_dev(self.ctx, self.A.ptr(x), c, length, x.stride(0), self.A.ptr(w), nfft, hop, nfft,
                                       self._lib.PAD_VALID, 0, 0, self._lib.SCALE_NONE, float(FS), self.A.ptr(z), self.sptr)
        self._lib.check(rc, self.ctx, "stft")

    def time_steps(self, fn, steps, warmup):
        """`steps` calls of fn after `warmup`, CUDA events on the launch stream, max over ranks -> ms per step."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.current_stream(self.dev)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps


def gen_signal(torch, dev, seed, channels, length):
    """0.25 N(0,1) + tones at 440 Hz and 3 kHz; bit-reproducible on any B200 for (seed, shape)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(channels, length, device=dev, generator=g, dtype=torch.float32) * 0.25
    t = torch.arange(length, device=dev, dtype=torch.float32) / FS
    x += 0.5 * torch.sin(2 * np.pi * 440.0 * t) + 0.5 * torch.sin(2 * np.pi * 3000.0 * t)
    return x


def gen_cfg2_channels(torch, dev, ch0, n):
    """channels [ch0, ch0 + n) of the strong-scaling cfg2 signal: one seed per channel"""
    return torch.cat([gen_signal(torch, dev, 20020 + c, 1, L) for c in range(ch0, ch0 + n)], dim=0) if n else \
        torch.empty((0, L), device=dev)


def bitwise_equal(torch, a, b):
    return bool(torch.equal(torch.view_as_real(a).view(torch.int32), torch.view_as_real(b).view(torch.int32)))


def exchange_check(env, mine, peer, recompute):
    """Rank `peer` sends `mine` (a complex tensor) to rank 0, which compares it bitwise with recompute()
    -- the same units computed on rank 0's GPU alone.  Returns True/False on rank 0, None elsewhere."""
    torch, dist = env.torch, env.dist
    if env.rank == peer:
        dist.send(torch.view_as_real(mine).contiguous(), dst=0)
        return None
    if env.rank == 0:
        ref = recompute()
        got = torch.empty_like(torch.view_as_real(ref))
        dist.recv(got, src=peer)
        torch.cuda.synchronize(env.dev)
        return bool(torch.equal(got.view(torch.int32), torch.view_as_real(ref).view(torch.int32)))
    return None


def multi_gpu_legs(env, nx, w1024, comm, peak, steps):
    """BASELINE configs[2] sharded over the ranks, and cfg2 under strong scaling, each with the in-run
    bitwise shard-vs-single-GPU check."""
    from nx_signal_b200 import sharding

    torch, rank, world, dev = env.torch, env.rank, env.world, env.dev
    out = {}
    checks = []

    # ---- cfg3: 1024 ch x 60 s, nfft 4096, hop 1024, channels sharded over the ranks ----
    w3 = torch.from_numpy(nx.windows.hann(N3)).to(dev) if rank == 0 else torch.empty(N3, device=dev)
    if comm is not None:
        sharding.broadcast_coeffs_c_abi(comm, w3, env.local_rank)
    sh = sharding.shard_channels(C3, world, rank)
    assert sh.start % BLOCK3 == 0 and sh.count % BLOCK3 == 0
    x3 = torch.cat([gen_signal(torch, dev, 3000 + b, BLOCK3, L3)
                    for b in range(sh.start // BLOCK3, (sh.start + sh.count) // BLOCK3)], dim=0)
    z3 = torch.empty((sh.count, M3, N3), dtype=torch.complex64, device=dev)
    env._lib.profile(True, env.local_rank)
    env._lib.profile_read(env.local_rank)
    ncalls3 = max(3, min(steps, 10))
    ms = env.time_steps(lambda: env.stft_dev(x3, w3, z3, N3, H3), ncalls3, 3)
    kms, kn = env._lib.profile_read(env.local_rank)
    env._lib.profile(False, env.local_rank)
    # kernel time PER CALL: a call over more than ~16 GB is walked in channel blocks, i.e. several launches
    # (csrc/nxs_stft.cu launch_stft); the profiler saw the 3 warm-up calls too
    kern_ms = env.max_over_ranks(kms / (ncalls3 + 3))
    launches_per_call3 = kn / (ncalls3 + 3)
    algo_rank = 4 * sh.count * L3 + 8 * sh.count * M3 * N3 + 4 * N3
    gbs = algo_rank / (kern_ms * 1e-3) / 1e9

    def block_alone(b):  # one 64-channel block computed by itself on this GPU
        xb = gen_signal(torch, dev, 3000 + b, BLOCK3, L3)
        zb = torch.empty((BLOCK3, M3, N3), dtype=torch.complex64, device=dev)
        env.stft_dev(xb, w3, zb, N3, H3)
        torch.cuda.synchronize(dev)
        return zb

    last_block = (sh.start + sh.count) // BLOCK3 - 1
    ok3 = []
    if world == 1:
        ok3.append(bitwise_equal(torch, z3[-BLOCK3:], block_alone(last_block)))
    else:
        for peer in range(1, world):
            pb = (sharding.shard_channels(C3, world, peer).start + sharding.shard_channels(C3, world, peer).count) // BLOCK3 - 1
            r = exchange_check(env, z3[-BLOCK3:], peer, lambda: block_alone(pb))
            if r is not None:
                ok3.append(r)
    out["cfg3_sharded"] = {
        "workload": f"cfg3 (BASELINE configs[2]): {C3} ch x 60 s @48 kHz f32, hann({N3}), hop {H3}, :valid -> "
                    f"{C3 * M3} frames in total, channels sharded over {world} GPU(s) ({sh.count} ch per GPU)",
        "scaling": "strong", "frames_total": C3 * M3, "ms_per_step": ms, "frames_per_s": C3 * M3 / (ms * 1e-3),
        "per_gpu": {"kernel_ms": kern_ms, "kernel_launches_per_call": launches_per_call3, "algorithmic_bytes": int(algo_rank),
                    "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak},
        "window_broadcast": "nxs_bcast_coeffs_dev (one ncclBroadcast)" if comm is not None else "single rank: none",
        "checked": ("every other rank's last 64-channel block, sent to rank 0 over NCCL p2p, vs the same block computed "
                    "alone on rank 0" if world > 1 else "last 64-channel block of the 1024-channel call vs the block alone"),
        "shard_equals_single_gpu": all(ok3) if ok3 else None,
    }
    checks += ok3
    del x3, z3
    torch.cuda.empty_cache()

    if world > 1:
        time.sleep(0.5)  # every leg starts from the same power state (the cfg3 leg runs the part at its power cap)
        # ---- cfg2, strong scaling by channel: the 8 channels split over the ranks ----
        shc = sharding.shard_channels(CHANNELS, world, rank)
        xs = gen_cfg2_channels(torch, dev, shc.start, shc.count)
        zs = torch.empty((shc.count, M, NFFT), dtype=torch.complex64, device=dev)
        step = (lambda: env.stft_dev(xs, w1024, zs, NFFT, HOP)) if shc.count else (lambda: None)
        ms = env.time_steps(step, steps, 3)
        last = max(r for r in range(world) if sharding.shard_channels(CHANNELS, world, r).count > 0)
        psh = sharding.shard_channels(CHANNELS, world, last)

        def chan_alone():
            xa = gen_cfg2_channels(torch, dev, psh.start, psh.count)
            za = torch.empty((psh.count, M, NFFT), dtype=torch.complex64, device=dev)
            env.stft_dev(xa, w1024, za, NFFT, HOP)
            torch.cuda.synchronize(dev)
            return za

        okc = exchange_check(env, zs, last, chan_alone) if last > 0 else None
        out["cfg2_strong_channels"] = {
            "workload": f"cfg2's {CHANNELS} channels split by channel over {world} GPUs "
                        f"({-(-CHANNELS // world)} ch per GPU), {FRAMES_PER_GPU} frames in total",
            "scaling": "strong", "ms_per_step": ms, "frames_per_s": FRAMES_PER_GPU / (ms * 1e-3),
            "checked": f"rank {last}'s whole shard vs the same channels on rank 0", "shard_equals_single_gpu": okc}
        if okc is not None:
            checks.append(okc)
        del xs, zs
        torch.cuda.empty_cache()

        time.sleep(0.5)
        # ---- cfg2, strong scaling by frame range (read-only halo of N - hop samples, no exchange) ----
        fs = sharding.shard_frames(M, NFFT, HOP, world, rank)
        xfull = gen_cfg2_channels(torch, dev, 0, CHANNELS)
        xf = xfull[:, fs.sample_start:fs.sample_start + fs.sample_count].contiguous()
        del xfull
        zf = torch.empty((CHANNELS, fs.frame_count, NFFT), dtype=torch.complex64, device=dev)
        ms = env.time_steps(lambda: env.stft_dev(xf, w1024, zf, NFFT, HOP), steps, 3)
        pfs = sharding.shard_frames(M, NFFT, HOP, world, world - 1)

        def frames_alone():  # channel 7 computed whole on rank 0; the last rank's frame range of it
            xa = gen_cfg2_channels(torch, dev, CHANNELS - 1, 1)
            za = torch.empty((1, M, NFFT), dtype=torch.complex64, device=dev)
            env.stft_dev(xa, w1024, za, NFFT, HOP)
            torch.cuda.synchronize(dev)
            return za[0, pfs.frame_start:pfs.frame_start + pfs.frame_count].contiguous()

        okf = exchange_check(env, zf[CHANNELS - 1], world - 1, frames_alone)
        out["cfg2_strong_frames"] = {
            "workload": f"cfg2's frames split by frame range over {world} GPUs (every rank: all {CHANNELS} channels, "
                        f"~{M // world} frames each, {NFFT - HOP}-sample read-only halo)",
            "scaling": "strong", "ms_per_step": ms, "frames_per_s": FRAMES_PER_GPU / (ms * 1e-3),
            "checked": f"rank {world - 1}'s frame range of channel {CHANNELS - 1} vs the whole channel on rank 0",
            "shard_equals_single_gpu": okf}
        if okf is not None:
            checks.append(okf)
        del xf, zf
        torch.cuda.empty_cache()

    ok = all(checks) if (rank == 0 and checks) else (None if rank == 0 else True)
    out["shard_equals_single_gpu"] = ok
    return out


def pcie_probe(env, nbytes=1 << 30):
    """Plain pinned-memory copies on this rank (all ranks at once): the box's H2D / D2H rates under this load."""
    torch = env.torch
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device=env.dev)
    res = {}
    for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
        dst.copy_(src, non_blocking=True)
        env.barrier()
        t = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        env.barrier()
        dt = env.max_over_ranks((time.perf_counter() - t) / 3)
        res[name + "_gbs_per_gpu"] = nbytes / dt / 1e9
        res[name + "_gbs_aggregate"] = env.world * nbytes / dt / 1e9
    # both directions at once (what a pipelined host call does): two streams, two buffer pairs
    h2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(nbytes, dtype=torch.uint8, device=env.dev)
    s1, s2 = torch.cuda.Stream(env.dev), torch.cuda.Stream(env.dev)

    def duplex():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    duplex()
    env.barrier()
    t = time.perf_counter()
    for _ in range(3):
        duplex()
    env.barrier()
    dt = env.max_over_ranks((time.perf_counter() - t) / 3)
    res["duplex_gbs_per_direction_per_gpu"] = nbytes / dt / 1e9
    del h, d, h2, d2
    return res


def e2e_legs(env, x, w, zo, zo_frames, args):
    """End to end through the C-ABI host entries (what a NIF calls): wall clock, copies inside."""
    torch, lib, ctx, A, _lib = env.torch, env.lib, env.ctx, env.A, env._lib
    world, rank = env.world, env.rank
    xh = torch.empty((CHANNELS, L), dtype=torch.float32, pin_memory=True)
    zh = torch.empty((CHANNELS, M, NFFT), dtype=torch.complex64, pin_memory=True)
    xh.copy_(x)
    wh = w.cpu().numpy()

    def call(xp, zp):
        rc = lib.nxs_stft_f32_host(ctx, xp, CHANNELS, L, L, wh.ctypes.data, NFFT, HOP, NFFT, _lib.PAD_VALID, 0, 0,
                                   _lib.SCALE_NONE, float(FS), zp)
        _lib.check(rc, ctx, "stft(host)")

    def timed(fn, n, warm=1):
        for _ in range(warm):
            fn()
        env.barrier()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        env.barrier()
        return env.max_over_ranks((time.perf_counter() - t) / n)

    step = lambda: call(A.ptr(xh), A.ptr(zh))  # noqa: E731
    # the warm-up calls let the context measure its two automatic transfer modes (full / one-sided + host mirror)
    # under the load of all ranks; the timed calls then use the cheaper (the mixed mode is timed below, pinned)
    _lib.set_host_mode(-1, env.local_rank)
    s = timed(step, args.e2e_steps, warm=3)
    mode = _lib.host_mode(env.local_rank)
    e2e = {"value": world * FRAMES_PER_GPU / s, "unit": "frames/s",
           "h2d_bytes_per_step": int(xh.numel() * 4 + NFFT * 4),
           "d2h_bytes_per_step": int(CHANNELS * M * 8 * ((NFFT // 2 + 1) if mode["result"].startswith("onesided") else
                                                         (3 * (NFFT // 2 + 1) + NFFT) // 4 if mode["result"].startswith("mixed") else NFFT)),
           "host_result_bytes_per_step": int(zh.numel() * 8), "steps": args.e2e_steps, "ms_per_step": 1e3 * s,
           "transfer_mode_chosen": mode["result"],
           "path": "nxs_stft_f32_host on pinned host buffers (wall clock, max over ranks): H2D | kernel | D2H in slabs; "
                   "the context picks, from its own timings under the present load, between moving both spectrum halves "
                   "and moving bins 0..nfft/2 while host threads write the conjugate-mirror bins; result = the "
                   "reference's two-sided c64 tensor, bit-identical to the device entry"}
    e2e["host_timeline_ms"] = [round(1e3 * t, 2) for t in _lib.host_timeline(env.local_rank)]
    for forced, key in ((1, "ms_per_step_onesided_d2h_host_mirror"), (0, "ms_per_step_full_d2h_no_host_mirror"),
                        (3, "ms_per_step_mixed_3_of_4_chunks_mirrored")):
        _lib.set_host_mode(forced, env.local_rank)
        e2e[key] = 1e3 * timed(step, 2, warm=1)
    _lib.set_host_mode(-1, env.local_rank)
    step()  # leave a result in zh for the checks below
    if rank == 0:
        got = zh[0, :zo_frames].numpy()
        e2e["parity"] = float((np.abs(got - zo).max(-1) / np.abs(zo).max(-1)).max())
        # the host result must equal the device result bit for bit (checked on the last 4096 frames)
        zd = torch.empty((4096, NFFT), dtype=torch.complex64, device=env.dev)
        x_tail = x[CHANNELS - 1, L - (4095 * HOP + NFFT) - ((L - NFFT) % HOP):].contiguous()
        env.stft_dev(x_tail[None, :], w, zd[None], NFFT, HOP)
        torch.cuda.synchronize(env.dev)
        zd_h = torch.view_as_real(zd).cpu()
        e2e["host_equals_device_bitwise"] = bool(torch.equal(zd_h, torch.view_as_real(zh[CHANNELS - 1, M - 4096:])))
    e2e["pcie_probe"] = pcie_probe(env)
    # what the box allows: every rank's 7.37 GB result must be written to host memory through the ranks' shared
    # host memory system; the aggregate D2H rate of plain pinned copies on all ranks at once bounds the call
    agg = e2e["pcie_probe"]["d2h_gbs_aggregate"] * 1e9
    full_floor = world * CHANNELS * M * NFFT * 8 / agg
    e2e["box_ceiling"] = {"aggregate_d2h_gbs": agg / 1e9, "floor_ms_both_halves_over_pcie": 1e3 * full_floor,
                          "floor_ms_lower_half_only": 1e3 * world * CHANNELS * M * (NFFT // 2 + 1) * 8 / agg,
                          "ms_per_step_over_full_result_floor": s / full_floor,
                          "note": "floors = result bytes of all ranks / aggregate D2H rate measured in this run; the lower-half "
                                  "floor ignores the host-memory traffic of the mirror pass and is not reachable when the "
                                  "host memory system, not PCIe, is the shared bottleneck (N >= 2 on this box)"}

    # the same call on memory the DMA engines cannot address: what a BEAM binary (enif_make_new_binary) is
    try:
        xp = np.empty((CHANNELS, L), dtype=np.float32)
        xp[...] = xh.numpy()
        t = time.perf_counter()
        zp = np.empty((CHANNELS, M, NFFT), dtype=np.complex64)
        call(xp.ctypes.data, zp.ctypes.data)  # first call: fresh result pages are faulted in inside the call
        first = time.perf_counter() - t
        s = timed(lambda: call(xp.ctypes.data, zp.ctypes.data), 2, warm=0)
        pg = {"ms_per_step": 1e3 * s, "value": world * FRAMES_PER_GPU / s, "unit": "frames/s",
              "first_call_ms_fresh_result_pages": 1e3 * first, "transfer_mode": _lib.host_mode(env.local_rank),
              "path": "the same call on pageable numpy buffers: input chunks staged into the context's pinned ring by host "
                      "threads, result slabs land in a pinned ring and one pass copies them out and writes the mirror half"}
        if rank == 0:
            pg["equals_pinned_call_bitwise"] = bool(np.array_equal(zp.view(np.int32), zh.numpy().view(np.int32)))
        e2e["pageable"] = pg
        # ... and with cudaHostRegister / cudaHostUnregister of both buffers inside the timed call
        rt = torch.cuda.cudart()

        reg_modes = []

        def registered():
            ok = int(rt.cudaHostRegister(xp.ctypes.data, xp.nbytes, 0)) == 0 and \
                int(rt.cudaHostRegister(zp.ctypes.data, zp.nbytes, 0)) == 0
            try:
                call(xp.ctypes.data, zp.ctypes.data)
                reg_modes.append((ok, _lib.host_mode(env.local_rank)))
            finally:
                rt.cudaHostUnregister(zp.ctypes.data)
                rt.cudaHostUnregister(xp.ctypes.data)

        s_reg = timed(registered, 1, warm=0)
        e2e["pageable_host_register_in_call"] = {
            "ms_per_step": 1e3 * s_reg, "value": world * FRAMES_PER_GPU / s_reg, "unit": "frames/s",
            "registered_ok": bool(reg_modes and reg_modes[-1][0]), "transfer_mode": reg_modes[-1][1] if reg_modes else None,
            "path": "cudaHostRegister(x), cudaHostRegister(z), the pinned path, unregister"}
        del xp, zp
    except Exception as ex:  # report, never fake
        e2e["pageable"] = {"error": repr(ex)[:300]}

    if world == 1:
        # the fused STFT -> log-mel host entry on the same workload: 128 mel bins per frame come back
        try:
            melh = torch.empty((CHANNELS, M, 128), dtype=torch.float32, pin_memory=True)

            def step_mel():
                rc = lib.nxs_stft_mel_f32_host(ctx, A.ptr(xh), CHANNELS, L, L, wh.ctypes.data, NFFT, HOP, NFFT,
                                               _lib.PAD_VALID, 0, 0, _lib.SCALE_NONE, float(FS), 128, 3016.0, 200 / 3,
                                               A.ptr(melh))
                _lib.check(rc, ctx, "stft_mel(host)")

            s = timed(step_mel, args.e2e_steps)
            e2e["log_mel_host_call"] = {
                "value": FRAMES_PER_GPU / s, "unit": "frames/s", "ms_per_step": 1e3 * s,
                "h2d_bytes_per_step": int(xh.numel() * 4 + NFFT * 4), "d2h_bytes_per_step": int(melh.numel() * 4),
                "path": "nxs_stft_mel_f32_host on pinned host buffers: H2D | fused STFT -> log-mel kernel | D2H of "
                        "[frames][128] f32 (the spectrum never leaves the SM)"}
            del melh
        except Exception as ex:
            e2e["log_mel_host_call"] = {"error": repr(ex)[:200]}
    del xh, zh
    return e2e


def other_host_entries(env, nx):
    """e2e of the path's other host entries on their BASELINE shapes (pinned buffers, wall clock):
    ISTFT cfg5 (spectrum in, signal out) and FIR cfg4 (signal in, signal out)."""
    torch, lib, ctx, A, _lib = env.torch, env.lib, env.ctx, env.A, env._lib
    out = {}

    def timed(fn, n=2):
        fn()
        torch.cuda.synchronize(env.dev)
        t = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t) / n

    def fill_normal(t):  # pinned host tensor <- N(0,1) generated on the device, row by row
        v = torch.view_as_real(t) if t.is_complex() else t
        for i in range(v.shape[0]):
            v[i].copy_(torch.randn(v[i].shape, device=env.dev))
        torch.cuda.synchronize(env.dev)

    probe = pcie_probe(env, 1 << 29)
    # ISTFT cfg5: 32 ch x 60 s
    C5, L5 = 32, FS * 60
    M5 = (L5 - NFFT) // HOP + 1
    wh = nx.windows.hann(NFFT)
    zh = torch.empty((C5, M5, NFFT), dtype=torch.complex64, pin_memory=True)
    fill_normal(zh)
    yh = torch.empty((C5, M5 * HOP + NFFT - HOP), dtype=torch.complex64, pin_memory=True)
    s = timed(lambda: _lib.check(lib.nxs_istft_c64_host(ctx, A.ptr(zh), C5, M5, NFFT, wh.ctypes.data, NFFT, HOP, NFFT, 0,
                                                        float(FS), A.ptr(yh)), ctx, "istft(host)"))
    h2d, d2h = zh.numel() * 8, yh.numel() * 8
    floor = max(h2d / probe["h2d_gbs_per_gpu"], d2h / probe["d2h_gbs_per_gpu"]) / 1e9
    # bit-identical to the device entry (first channel)
    zd = zh[:1].to(env.dev)
    yd = torch.empty((1, yh.shape[1]), dtype=torch.complex64, device=env.dev)
    _lib.check(lib.nxs_istft_c64_dev(ctx, A.ptr(zd), 1, M5, NFFT, A.ptr(torch.from_numpy(wh).to(env.dev)), NFFT, HOP, NFFT, 0,
                                     float(FS), A.ptr(yd), env.sptr), ctx)
    torch.cuda.synchronize(env.dev)
    out["istft_cfg5_host_call"] = {"ms_per_step": 1e3 * s, "frames_per_s": C5 * M5 / s, "h2d_bytes_per_step": int(h2d),
                                   "d2h_bytes_per_step": int(d2h), "pcie_floor_ms": 1e3 * floor,
                                   "over_pcie_floor": s / floor,
                                   "host_equals_device_bitwise": bitwise_equal(torch, yd.cpu(), yh[:1])}
    del zh, yh, zd, yd
    # FIR cfg4: 64 ch x 600 s, 2049 taps, mode :same
    C4 = 64
    taps = nx.filters.firwin(2049, [6000], sampling_rate=FS)
    xh = torch.empty((C4, L), dtype=torch.float32, pin_memory=True)
    fill_normal(xh)
    yh = torch.empty((C4, L), dtype=torch.float32, pin_memory=True)
    s = timed(lambda: _lib.check(lib.nxs_fir_f32_host(ctx, A.ptr(xh), C4, L, L, taps.ctypes.data, 2049, 1, A.ptr(yh), L),
                                 ctx, "fir(host)"))
    h2d = d2h = xh.numel() * 4
    floor = max(h2d / probe["h2d_gbs_per_gpu"], d2h / probe["d2h_gbs_per_gpu"]) / 1e9
    floor_duplex = max(h2d, d2h) / probe["duplex_gbs_per_direction_per_gpu"] / 1e9
    xd = xh[:1].to(env.dev)
    yd = torch.empty((1, L), dtype=torch.float32, device=env.dev)
    _lib.check(lib.nxs_fir_f32_dev(ctx, A.ptr(xd), 1, L, L, A.ptr(torch.from_numpy(taps).to(env.dev)), 2049, 1, A.ptr(yd), L,
                                   env.sptr), ctx)
    torch.cuda.synchronize(env.dev)
    out["fir_cfg4_host_call"] = {"ms_per_step": 1e3 * s, "samples_per_s": C4 * L / s, "h2d_bytes_per_step": int(h2d),
                                 "d2h_bytes_per_step": int(d2h), "pcie_floor_ms": 1e3 * floor, "over_pcie_floor": s / floor,
                                 "pcie_floor_both_directions_busy_ms": 1e3 * floor_duplex,
                                 "over_pcie_floor_both_directions_busy": s / floor_duplex,
                                 "host_equals_device_bitwise": bool(torch.equal(yd.cpu().view(torch.int32),
                                                                                yh[:1].view(torch.int32)))}
    out["pcie_probe"] = probe
    del xh, yh, xd, yd
    torch.cuda.empty_cache()
    return out


def cfg1_leg(env, nx):
    """BASELINE configs[0]: NxSignal.stft on 1 x 48000 f32, nfft 1024, hop 256 (184 frames): call latency on the
    GPU beside the literal 1-core restatement of the reference's algorithm."""
    from oracle import c_port
    from oracle import nxsignal_oracle as o

    torch, lib, ctx, A, _lib = env.torch, env.lib, env.ctx, env.A, env._lib
    L1 = 48000
    M1 = (L1 - NFFT) // HOP + 1
    x = cpu_input(1, L1, seed=1001)
    wh = o.hann(NFFT)
    z = np.empty((1, M1, NFFT), dtype=np.complex64)

    def host_call(xp, zp):
        _lib.check(lib.nxs_stft_f32_host(ctx, xp, 1, L1, L1, wh.ctypes.data, NFFT, HOP, NFFT, _lib.PAD_VALID, 0, 0,
                                         _lib.SCALE_NONE, float(FS), zp), ctx, "stft(host, cfg1)")

    def lat(fn, n=200):
        for _ in range(20):
            fn()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t) / n

    pageable_s = lat(lambda: host_call(x.ctypes.data, z.ctypes.data))
    xp = torch.from_numpy(x).pin_memory()
    zp = torch.empty((1, M1, NFFT), dtype=torch.complex64, pin_memory=True)
    pinned_s = lat(lambda: host_call(A.ptr(xp), A.ptr(zp)))
    transfer_mode = _lib.host_mode(env.local_rank)["result"]
    # the way back on the same small shape: nxs_istft_c64_host on the spectrum just computed
    out_len = (M1 - 1) * HOP + NFFT
    yp = torch.empty((1, out_len), dtype=torch.complex64, pin_memory=True)
    istft_s = lat(lambda: _lib.check(lib.nxs_istft_c64_host(ctx, A.ptr(zp), 1, M1, NFFT, wh.ctypes.data, NFFT, HOP, NFFT, 0,
                                                            float(FS), A.ptr(yp)), ctx, "istft(host, cfg1)"))
    rt = yp.numpy()[0, NFFT:out_len - NFFT].real
    roundtrip_err = float(np.abs(rt - x[0, NFFT:out_len - NFFT]).max() / np.abs(x).max())
    xd, wd = xp.to(env.dev), torch.from_numpy(wh).to(env.dev)
    zd = torch.empty((1, M1, NFFT), dtype=torch.complex64, device=env.dev)
    n0 = _lib.launch_count(env.local_rank)
    dev_ms = env.time_steps(lambda: env.stft_dev(xd, wd, zd, NFFT, HOP), 500, 20) if env.world == 1 else None
    launches_per_call = (_lib.launch_count(env.local_rank) - n0) / 520 if dev_ms is not None else None
    # CPU: the literal numpy restatement (1 core) and its C port (1 thread), same input
    t = time.perf_counter()
    zo, _, _ = o.stft(x[0], wh, overlap_length=NFFT - HOP, fft_length=NFFT, sampling_rate=FS)
    lit_s = time.perf_counter() - t
    c_s = min(lat(lambda: c_port.stft(x, wh, HOP, NFFT, nthreads=1), n=5) for _ in range(2))
    err = float((np.abs(z[0] - zo).max(-1) / np.abs(zo).max(-1)).max())
    return {"workload": "cfg1 (BASELINE configs[0]): 1 x 48000 f32, hann(1024), hop 256 -> 184 frames x 1024 bins",
            "gpu_host_call_pageable_us": 1e6 * pageable_s, "gpu_host_call_pinned_us": 1e6 * pinned_s,
            "host_call_transfer_mode": transfer_mode, "gpu_istft_host_call_pinned_us": 1e6 * istft_s,
            "stft_istft_roundtrip_interior_err": roundtrip_err,
            "gpu_dev_call_us": None if dev_ms is None else 1e3 * dev_ms, "kernel_launches_per_dev_call": launches_per_call,
            "frames_per_s_host_call": M1 / pageable_s,
            "cpu_literal_restatement_1core_ms": 1e3 * lit_s, "cpu_c_port_1thread_ms": 1e3 * c_s,
            "frames_per_s_cpu_c_port_1thread": M1 / c_s,
            "parity_frame_rel_err_vs_literal_oracle": err,
            "note": "host call = nxs_stft_f32_host incl. H2D, kernel, D2H and the host mirror; the literal restatement is "
                    "oracle.stft (recursive radix-2 in f64 like Nx.BinaryBackend, numpy-vectorised over frames)"}


def other_kernels(env, nx, peak):
    """Kernel times of the path's other rows on their BASELINE shapes: ISTFT cfg5, FIR cfg4 and the
    log-mel epilogue.  Mean of 5 launches after 2 warm-ups, CUDA events around the dominant kernel."""
    torch, lib, ctx, A, _lib, dev, local_rank = env.torch, env.lib, env.ctx, env.A, env._lib, env.dev, env.local_rank
    out = {}

    def timed(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        _lib.profile(True, local_rank)
        _lib.profile_read(local_rank)
        for _ in range(iters):
            fn()
        ms, n = _lib.profile_read(local_rank)
        _lib.profile(False, local_rank)
        return ms / max(n, 1)

    def entry(ms, algo_bytes, units, unit_name):
        gbs = algo_bytes / (ms * 1e-3) / 1e9
        return {"kernel_ms": ms, "algorithmic_bytes": int(algo_bytes), "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak,
                unit_name: units / (ms * 1e-3)}

    g = torch.Generator(device=dev).manual_seed(7)
    C5, L5 = 32, FS * 60
    M5 = (L5 - NFFT) // HOP + 1
    w = torch.from_numpy(nx.windows.hann(NFFT)).to(dev)
    z = torch.randn(C5, M5, NFFT, 2, device=dev, generator=g)
    y = torch.empty(C5, M5 * HOP + NFFT - HOP, 2, device=dev)
    s = env.sptr
    ms = timed(lambda: _lib.check(lib.nxs_istft_c64_dev(ctx, A.ptr(z), C5, M5, NFFT, A.ptr(w), NFFT, HOP, NFFT, 0,
                                                       float(FS), A.ptr(y), s), ctx))
    out["istft_cfg5"] = entry(ms, 8 * C5 * M5 * NFFT + 8 * C5 * (M5 * HOP + NFFT - HOP), C5 * M5, "frames_per_s")
    mel = torch.empty(C5, M5, 128, device=dev)
    ms = timed(lambda: _lib.check(lib.nxs_stft_to_mel_f32_dev(ctx, A.ptr(z), C5, M5, NFFT, NFFT, 128, float(FS), 3016.0,
                                                             200 / 3, A.ptr(mel), s), ctx))
    out["stft_to_mel_cfg5_spectrum"] = entry(ms, 4 * C5 * M5 * NFFT + 4 * C5 * M5 * 128, C5 * M5, "frames_per_s")
    K1 = NFFT // 2 + 1
    z1 = torch.randn(C5, M5, K1, 2, device=dev, generator=g)
    y1 = torch.empty(C5, M5 * HOP + NFFT - HOP, device=dev)
    ms = timed(lambda: _lib.check(lib.nxs_istft_c2r_f32_dev(ctx, A.ptr(z1), C5, M5, K1, A.ptr(w), NFFT, HOP, NFFT, 0,
                                                           float(FS), A.ptr(y1), s), ctx))
    out["istft_c2r_cfg5"] = entry(ms, 8 * C5 * M5 * K1 + 4 * C5 * (M5 * HOP + NFFT - HOP), C5 * M5, "frames_per_s")
    x5 = torch.randn(C5, L5, device=dev, generator=g)
    ms = timed(lambda: _lib.check(lib.nxs_stft_mel_f32_dev(ctx, A.ptr(x5), C5, L5, L5, A.ptr(w), NFFT, HOP, NFFT,
                                                          _lib.PAD_VALID, 0, 0, _lib.SCALE_NONE, float(FS), 128, 3016.0,
                                                          200 / 3, A.ptr(mel), s), ctx))
    out["stft_mel_fused_cfg5"] = entry(ms, 4 * C5 * L5 + 4 * C5 * M5 * 128, C5 * M5, "frames_per_s")
    out["stft_mel_fused_cfg5"]["note"] = "kernel_ms = the fused STFT kernel only (the clamp/affine finalize pass adds ~0.1 ms); issue-bound, not HBM-bound"
    del z, y, mel, z1, y1, x5
    torch.cuda.empty_cache()
    C4 = 64
    taps = torch.from_numpy(nx.filters.firwin(2049, [6000], sampling_rate=FS)).to(dev)
    x4 = torch.randn(C4, L, device=dev, generator=g)
    y4 = torch.empty_like(x4)
    ms = timed(lambda: _lib.check(lib.nxs_fir_f32_dev(ctx, A.ptr(x4), C4, L, L, A.ptr(taps), 2049, 1, A.ptr(y4), L, s), ctx),
               iters=3)
    out["fir_cfg4"] = entry(ms, 8 * C4 * L, C4 * L, "samples_per_s")
    # fp32 roofline (SURVEY 8d asks for both): 80.9 flop per output sample from the executed instruction mix of
    # fir_ols_r2c_kernel (ncu, profiles/r02_fir_r2c_4096.txt: FADD 37.7 %, FMUL 15.9 %, FFMA 20.8 % of 489 M warp
    # instructions for 64 ch x 60 s)
    tfl = 80.9 * C4 * L / (ms * 1e-3) / 1e12
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12  # SMs x lanes x FMA x max SM clock
    out["fir_cfg4"].update({"fp32_tflops": tfl, "fp32_peak_tflops": fp32_peak, "frac_of_fp32_peak": tfl / fp32_peak,
                            "note": "real-packed overlap-save (one 8192-sample block per 4096-point transform pair): "
                                    "instruction-issue-bound at K = 2049 (ncu: 68 % of issue slots, DRAM 12 %), not HBM-bound"})
    del x4, y4
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and cfg1 legs")
    ap.add_argument("--no-extras", action="store_true", help="skip the ISTFT / FIR / mel kernel timings and host entries")
    ap.add_argument("--no-multi", action="store_true", help="skip the cfg3 / strong-scaling legs")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries exactly ONE JSON line: native libraries that write to fd 1 (NCCL prints its
    # version banner there at communicator creation) are sent to stderr until the line is printed
    sys.stdout.flush()
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    import torch

    import nx_signal_b200 as nx
    from nx_signal_b200 import _lib, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (nx_signal_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        comm = sharding.NcclComm(rank, world, local_rank)
    # host threads of the _host entries: share the box's cores between the ranks
    os.environ.setdefault("NXS_HOST_THREADS", str(max(2, len(os.sched_getaffinity(0)) // max(world, 1))))
    env = Env(torch, dist, rank, local_rank, world)

    # window: built on rank 0, ONE NCCL broadcast through the C ABI (the path's only collective)
    w = torch.from_numpy(nx.windows.hann(NFFT)).to(dev) if rank == 0 else torch.zeros(NFFT, device=dev)
    if comm is not None:
        sharding.broadcast_coeffs_c_abi(comm, w, local_rank)
        torch.cuda.synchronize(dev)

    x = gen_signal(torch, dev, 1002 + rank, CHANNELS, L)
    z = torch.empty((CHANNELS, M, NFFT), dtype=torch.complex64, device=dev)
    step_dev = lambda: env.stft_dev(x, w, z, NFFT, HOP)  # noqa: E731

    for _ in range(args.warmup):
        step_dev()
    env.barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    _lib.profile(True, local_rank)
    _lib.profile_read(local_rank)
    launches0 = _lib.launch_count(local_rank)
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    t0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    env.barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    kern_ms, kern_n = _lib.profile_read(local_rank)
    _lib.profile(False, local_rank)
    launches = _lib.launch_count(local_rank) - launches0
    clocks = sampler.stop(t0, t1)
    ms = env.max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = world * FRAMES_PER_GPU / (ms_per_step * 1e-3)

    # parity inside the bench run (oracle = checker only): 4096 frames drawn at random over all channels of
    # this rank's result, every rank; the first 16 frames of channel 0 are kept for the host-entry check
    from oracle import nxsignal_oracle as o

    wn = w.cpu().numpy()
    rng = np.random.default_rng(77 + rank)
    fr = np.sort(rng.choice(FRAMES_PER_GPU, size=4096, replace=False))
    cs, ms_ = fr // M, fr % M
    idx = torch.from_numpy(ms_[:, None] * HOP + np.arange(NFFT)[None, :]).to(dev)
    frames = x[torch.from_numpy(cs).to(dev)[:, None], idx].cpu().numpy()  # [4096, NFFT] raw frames
    want = np.fft.fft((frames * wn[None, :]).astype(np.float32).astype(np.float64), axis=-1).astype(np.complex64)
    got = z[torch.from_numpy(cs).to(dev), torch.from_numpy(ms_).to(dev)].cpu().numpy()
    parity = float((np.abs(got - want).max(-1) / np.abs(want).max(-1)).max())
    parity = env.max_over_ranks(parity)
    nchk = 16
    zo, _, _ = o.stft_fast(x[0, : (nchk - 1) * HOP + NFFT].cpu().numpy(), wn, overlap_length=NFFT - HOP, fft_length=NFFT,
                           sampling_rate=FS)
    parity16 = float((np.abs(z[0, :nchk].cpu().numpy() - zo).max(-1) / np.abs(zo).max(-1)).max())

    del z
    torch.cuda.empty_cache()
    e2e = None
    if not args.no_e2e:
        try:
            e2e = e2e_legs(env, x, w, zo, nchk, args)
        except Exception as ex:  # report, never fake
            e2e = {"value": None, "unit": "frames/s", "error": repr(ex)[:300]}
    del x
    torch.cuda.empty_cache()

    peak, peak_src = peaks()
    # single-GPU extras first (each kernel timed alone, before the long cfg3 leg heats the part into its power cap)
    other = None
    if world == 1 and not args.no_extras:
        time.sleep(1.0)
        try:
            other = other_kernels(env, nx, peak)
            other.update(other_host_entries(env, nx))
        except Exception as ex:  # report, never fake
            other = {"error": repr(ex)[:300]}
    multi = None
    if not args.no_multi:
        try:
            multi = multi_gpu_legs(env, nx, w, comm, peak, min(args.steps, 20))
        except Exception as ex:
            multi = {"error": repr(ex)[:300]}
            if dist is not None:  # a rank that failed must not leave the others in a collective
                raise

    if rank != 0:
        if comm is not None:
            comm.destroy()
        if dist is not None:
            dist.destroy_process_group()
        return

    kern_avg_ms = kern_ms / max(kern_n, 1)
    achieved = ALGO_BYTES / (kern_avg_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("stft_r2c_1024_bytes_per_launch")
        except Exception:
            traffic = None
    cpu = None
    cfg1 = None
    if world == 1 and not args.no_cpu:
        try:
            cfg1 = cfg1_leg(env, nx)
        except Exception as ex:
            cfg1 = {"error": repr(ex)[:300]}
        port = CpuPort(full=False)
        port.step()
        rs, t_end = [], time.perf_counter() + 10.0
        while len(rs) < 3 or (time.perf_counter() < t_end and len(rs) < 40):
            rs.append(port.step())
        r = port.frames * len(rs) / sum(dt for _, dt in rs)
        cpu = {"value": r, "unit": "frames/s", "cores": port.cores, "kind": "port",
               "sample": port.sample + f", {len(rs)} steps",
               "note": "oracle/nxs_oracle.c: Nx.BinaryBackend's f64 recursive radix-2 restated in C + OpenMP; "
                       "the real BEAM backend cannot run here and is far slower"}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "channels_per_gpu": CHANNELS, "fft_length": NFFT, "hop": HOP,
                   "l2": "inputs larger than L2 (0.92 GB in, 7.37 GB out per step)",
                   "parity_frame_rel_err": parity, "parity_checked": "4096 random frames over all channels of every rank "
                   "vs f64 FFT of the f32-rounded windowed frames (max over ranks)", "parity_first16_vs_oracle": parity16,
                   "window_broadcast": "nxs_bcast_coeffs_dev (one ncclBroadcast)" if comm is not None else "single rank: none"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "stft_r2c_staged_kernel<StagedCfg<Plan<512,64,8,8,8>,256,2,tw-regs,per-group TMA,win-regs>,2,two-sided>",
                     "kernel_ms": kern_avg_ms, "kernel_launches_timed": kern_n, "algorithmic_bytes": ALGO_BYTES},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "multi_gpu": multi, "cfg1": cfg1, "other_kernels": other,
        "build": _lib.build_info(),  # which binary ran: the stamp compiled into the .so vs the sources in the tree
    }
    emit(line)
    if comm is not None:
        comm.destroy()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

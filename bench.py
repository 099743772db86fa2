#!/usr/bin/env python
"""Benchmark of the particle-stepping hot path (BASELINE.json metric: particle-updates/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--particles PER_GPU]

One "step" = one iteration of the loop at particles.cpp:96-201 (two force evaluations + one
re-binning) over the whole scene.  Workload at EVERY N: BASELINE.json configs[2] — 16M particles
per GPU, hex packing at pitch 2.02r, ~4 directed bonds per particle, 1 % frozen; for N>1 one
global scene of N x 16M particles cut into N strips over x (weak scaling; halo exchange,
migration and bond rows travelling with their particles over NCCL), every rank generating and
loading only its own strip.  The line also carries, as `config.extra`, the pair+gravity+walls
variant at 16M (120 B per update), configs[1] (1M, uniform random) and - at N=8 - the 256M
dense-packing scene of configs[4].

Prints ONE JSON line (rank 0).  Keys beyond the driver's contract:
  roofline      dominant kernel: algorithmic bytes per launch / CUDA-event duration vs the
                measured HBM copy bandwidth (MEASURED_PEAKS.json), plus the whole-step figure
  cpu_baseline  the reference's own CPU step (oracle/_ref, OpenMP, all host cores) on a
                bounded sample of the same workload, timed in this run (N=1, rank 0)
  e2e           the same metric through the C ABI with HOST buffers: per step the full state
                is uploaded from pinned memory, stepped, and downloaded again, one call after
                the other on one simulation (`e2e.value`); `e2e.pipelined` is the throughput of
                several INDEPENDENT batches in flight on several handles (PCIe both ways at once)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PAIR = {"stage1": 32, "stage2": 48, "reorder": 40}
METRIC = "particle-updates/sec"


def bytes_per_update(d):
    """Algorithmic bytes per particle-update with mean directed bond degree d."""
    per_stage_bonds = (12 + 4 * d) if d > 0 else 0
    k = {"stage1": 32 + per_stage_bonds, "stage2": 48 + per_stage_bonds,
         "reorder": 40 + (4 if d > 0 else 0)}
    k["total"] = k["stage1"] + k["stage2"] + k["reorder"]
    return k


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region: NVML in a
    background thread (every 5 ms); `nvidia-smi -lms` as a fallback when NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.stop_flag, self.th = None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            R = pynvml
            bits = {"hw_slowdown": getattr(R, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(R, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(R, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(R, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}

            def loop():
                while not self.stop_flag:
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        rs = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((time.time(), sm, mx, [k for k, v in bits.items() if rs & v]))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.nvml = pynvml
            self.th = threading.Thread(target=loop, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 7:
                continue
            try:
                self.rows.append((time.time(), float(f[0]), float(f[1]),
                                  [nm for k, nm in enumerate(names) if f[3 + k].lower().startswith("active")]))
            except ValueError:
                continue

    def stop(self, window=None):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.th:
            self.th.join(timeout=1)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"],
                    "samples": 0}
        rows = self.rows
        where = "warm-up + timed region + per-kernel pass (all under load)"
        if window is not None:
            inside = [r for r in rows if window[0] <= r[0] <= window[1]]
            if len(inside) >= 2:
                rows, where = inside, "timed region"
        reasons = sorted({x for r in rows for x in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)),
                "reasons": reasons, "samples": len(rows), "window": where,
                "source": "nvml" if self.nvml else "nvidia-smi"}



def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use every host
    core it can, so set the OpenMP thread count back (environment for libraries loaded later,
    omp_set_num_threads for a libgomp that is already initialised)."""
    import ctypes
    n = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def bench_scene(n, rank=0, world=1):
    """configs[2] as a strip scene: the particles, bond rows and frozen ids of `rank`'s strip of
    ONE global scene of world * n particles (world = 1: the whole scene)."""
    from ptoy_b200 import scenes
    return scenes.strip_scene(n, rank, world, pitch_over_r=2.02, jitter_over_r=0.01, bonds=True)


WORKLOAD = "configs[2]: 16M particles per GPU, hex pitch 2.02r, ~4 directed bonds/particle, 1% frozen"


def run_reference(args):
    """--impl reference: the reference's own CPU Particles::step on the box's host cores, on the
    same scene generator at a bounded size (the reference needs ~10 s per iteration and minutes of
    std::set insertions at 16M: `same_size` says whether the sample is the full workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    from oracle import port, ref
    from ptoy_b200 import scenes
    n = args.ref_n
    variants = [v for v in ("omp", "avx_omp") if ref.available(v)]
    sc = bench_scene(n)
    n = sc.n
    best = None
    for v in variants or ["port"]:
        o = ref.RefParticles(v) if v != "port" else port.PortParticles()
        scenes.load_into(o, sc)
        for _ in range(args.warmup_ref):
            o.step_n(1)
        t0 = time.perf_counter()
        o.step_n(args.steps_ref)
        dt = time.perf_counter() - t0
        rate = n * args.steps_ref / dt
        if best is None or rate > best[0]:
            best = (rate, v, o.openmp_threads, dt / args.steps_ref)
        o.close()
    rate, v, threads, spstep = best
    kind = "port" if v == "port" else "reference"
    full = n >= 15_000_000
    sample = (f"configs[2] generator at n={n} ({'the full per-GPU workload' if full else 'bounded sample; the rate is '
              'per particle-update, so it extrapolates to 16M'}), {args.steps_ref} iterations after "
              f"{args.warmup_ref} warm-up, build={v}")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": args.steps_ref, "warmup": args.warmup_ref,
        "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_particles": n, "same_size": full,
                   "extrapolated": not full},
        "cpu_baseline": {"value": rate, "unit": "particle-updates/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(n, steps=3, warmup=1):
    use_all_host_threads()
    from oracle import port, ref
    from ptoy_b200 import scenes
    sc = bench_scene(n)
    n = sc.n
    best = None
    for v in [v for v in ("omp", "avx_omp") if ref.available(v)] or ["port"]:
        o = ref.RefParticles(v) if v != "port" else port.PortParticles()
        scenes.load_into(o, sc)
        o.step_n(warmup)
        t0 = time.perf_counter()
        o.step_n(steps)
        dt = time.perf_counter() - t0
        rate = n * steps / dt
        if best is None or rate > best["value"]:
            best = {"value": rate, "unit": "particle-updates/s", "cores": o.openmp_threads,
                    "kind": "port" if v == "port" else "reference",
                    "sample": f"configs[2] generator at n={n}, {steps} iterations after {warmup} warm-up, build={v}"}
        o.close()
    return best


def measure(g, steps, warm, stream_of, barrier, torch):
    """Device time of `steps` iterations (CUDA events on the engine's stream), after `warm`."""
    g.step_n(warm)
    g.synchronize()
    barrier()
    t_begin = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = stream_of(g)
    ev0.record(st)
    g.step_n(steps)
    ev1.record(st)
    g.synchronize()      # (also raises if a device error flag is set)
    barrier()
    return ev0.elapsed_time(ev1), t_begin, time.time()


def strips_selfcheck(rank, world, local, torch, dist, ptoy_b200, scenes):
    """N>1: the transport this run measures, checked in the same processes: a 200k-particle bonded
    sheet drifting across the strip boundaries, 20 iterations, union of the strips against one
    engine on rank 0, bit for bit."""
    sc = scenes.hex_lattice(200_000, 2.02, 0.01, seed=33, vel_scale=1.0)
    sc.vel[:, 0] += np.float32(4.0)
    scenes.add_lattice_bonds(sc)
    g = ptoy_b200.Particles(device=local, default_scene=False)
    g.reset_scene(sc.domain)
    g.set_gravity(*sc.gravity)
    g.configure_strips(rank, world, ghost_rows=4, id_capacity=sc.n, bonds=True)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(ptoy_b200.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    g.nccl_connect(bytes(uid.cpu().numpy().tobytes()))
    g.add_particles(sc.pos, sc.vel, sc.ids)
    mine = g.state_by_id(sc.n)["block"] >= 0
    g.set_bonds(sc.bonds[mine[sc.bonds[:, 0]]])
    g.set_frozen(sc.frozen)
    g.set_renderer_ready(False)
    g.step_n(20)
    g.synchronize()
    s = g.state_by_id(sc.n)
    alive = torch.from_numpy((s["block"] >= 0).astype(np.int64)).cuda()
    pos = torch.from_numpy(np.where(np.isnan(s["pos"]), 0, s["pos"]).view(np.int32).astype(np.int64)).cuda()
    dist.all_reduce(alive)
    dist.all_reduce(pos)
    moved = torch.tensor([int(((s["block"] >= 0) & ~mine).sum())], device="cuda")
    dist.all_reduce(moved)
    g.close()
    out = None
    if rank == 0:
        ref = ptoy_b200.Particles(device=local, default_scene=False)
        scenes.load_into(ref, sc)
        ref.set_renderer_ready(False)
        ref.step_n(20)
        r = ref.state_by_id(sc.n)
        ref.close()
        same = bool((alive.cpu().numpy() == (r["block"] >= 0)).all()) and np.array_equal(
            pos.cpu().numpy().astype(np.int32), np.where(np.isnan(r["pos"]), 0, r["pos"]).view(np.int32))
        out = {"bit_identical_to_one_gpu": same, "particles": sc.n, "iterations": 20,
               "changed_strip": int(moved.item()), "scene": "bonded sheet drifting across the strip boundaries"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--particles", dest="n", type=int, default=16_000_000, help="particles per GPU (configs[2]: 16M)")
    ap.add_argument("--ref-particles", dest="ref_n", type=int, default=None,
                    help="particles of the CPU runs (default: 16M for --impl reference, 2M for the in-line cpu_baseline)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-depth", type=int, default=3, help="independent batches in flight in the e2e.pipelined leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra records (pair-only 16M, configs[1], 256M dense)")
    args = ap.parse_args()
    args.steps_ref = max(1, min(args.steps, 2))
    args.warmup_ref = 1
    if args.impl == "reference":
        if args.ref_n is None:
            args.ref_n = 16_000_000
        return run_reference(args)
    if args.ref_n is None:
        args.ref_n = 2_000_000

    import faulthandler
    import torch
    import torch.distributed as dist
    import ptoy_b200
    from ptoy_b200 import scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # a stalled collective must not hold the box: dump the Python stacks and exit after a while
    faulthandler.dump_traceback_later(int(os.environ.get("PTOY_WATCHDOG_S", "600")), exit=True)

    def progress(msg):
        if os.environ.get("PTOY_BENCH_VERBOSE"):
            print(f"[rank {rank} {time.time() % 1000:7.2f}] {msg}", file=sys.stderr, flush=True)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    W = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stream_of(e):
        return torch.cuda.ExternalStream(e.stream, device=torch.device("cuda", local))

    def connect(e):
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ptoy_b200.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        e.nccl_connect(bytes(uid.cpu().numpy().tobytes()))

    # ---- scene: this rank's strip of ONE global configs[2] scene -------------------------------
    t0 = time.time()
    sc = bench_scene(args.n, rank, world)
    g = ptoy_b200.Particles(device=local, default_scene=False)
    g.reset_scene(sc.domain)
    g.set_gravity(*sc.gravity)
    if world > 1:
        g.configure_strips(rank, world, sc.col_begin, ghost_rows=4, id_capacity=sc.n_global, bonds=True)
        connect(g)
        progress("connected")
    g.add_particles(sc.pos, sc.vel, sc.ids)
    g.set_bonds(sc.bonds)
    g.set_frozen(sc.frozen)
    parallelism = "1 GPU" if world == 1 else \
        f"{world} strips over x: halo exchange, migration and travelling bond rows over NCCL"
    n = sc.n
    g.set_renderer_ready(False)
    g.synchronize()
    degree = len(sc.bonds) / max(n, 1)   # (a strip is handed some rows beyond its own: slightly above 4)
    bpu = bytes_per_update(4.0)
    t_setup = time.time() - t0
    progress("scene loaded")

    # ---- device-resident throughput ("value") -----------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    l0 = None
    g.step_n(W)
    g.synchronize()
    l0 = g.launch_count
    ms, t_begin, t_end = measure(g, args.steps, 0, stream_of, barrier, torch)
    launches = g.launch_count - l0
    alive = g.num_particles
    err = g.error_flags
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([float(alive), float(launches), float(err)], device="cuda", dtype=torch.float64)
        dist.all_reduce(c)
        alive_total, launches, err = int(c[0].item()), int(c[1].item()), int(c[2].item())
    else:
        alive_total = alive
    assert err == 0, f"device error flags {err} after the timed region"
    value = alive_total * args.steps / (ms * 1e-3)
    progress("timed region done")

    # ---- per-kernel durations (CUDA events on the engine's stream, same loop) ---------------------
    g.enable_kernel_timing(True)
    g.step_n(args.steps)
    kt = g.kernel_times()
    g.enable_kernel_timing(False)
    clocks = sampler.stop((t_begin, t_end))
    per_launch_ms = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in kt.items()}
    dominant = max(("stage1", "stage2"), key=lambda k: per_launch_ms[k])
    peak, peak_src = hbm_peak()
    achieved = bpu[dominant] * alive / (per_launch_ms[dominant] * 1e-3) / 1e9 if per_launch_ms[dominant] else 0.0
    step_gbs = bpu["total"] * alive_total / world * args.steps / (ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{dominant}@{args.n}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": f"stage_tile_kernel<{dominant[-1]}>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": bpu, "kernel_ms_per_launch": per_launch_ms,
                "step_achieved": step_gbs, "step_frac": step_gbs / peak,
                "step_frac_of_8TBs": step_gbs / 8000.0}

    # ---- end to end through the C ABI with host buffers ----------------------------------------------
    # Every step: H2D of that step's input (pos 8 + vel 8 + id 4 B per particle, pinned host memory)
    # -> re-binning -> one iteration -> D2H of the result (same 20 B).  e2e.value is ONE simulation,
    # the calls made one after the other; e2e.pipelined keeps `depth` independent batches in flight
    # on `depth` handles (one stream each) so that uploads and downloads overlap.
    ncap = n + n // 50 + 4096          # (a strip's population changes with migration)

    def pinned():
        return (torch.empty((ncap, 2), dtype=torch.float32).pin_memory().numpy(),
                torch.empty((ncap, 2), dtype=torch.float32).pin_memory().numpy(),
                torch.empty((ncap,), dtype=torch.int32).pin_memory().numpy())
    hpn, hvn, hin = pinned()
    cnt = g.download_state(hpn, hvn, hin)
    assert cnt <= ncap, "the strip outgrew the e2e host buffers"
    g.upload_state(hpn[:cnt], hvn[:cnt], hin[:cnt])      # warm-up of the path
    g.step_n(1)
    cnt = g.download_state(hpn, hvn, hin)                # (a strip's count changes with migration)
    assert cnt <= ncap
    barrier()
    te = time.perf_counter()
    for _ in range(args.e2e_steps):
        g.upload_state(hpn[:cnt], hvn[:cnt], hin[:cnt])
        g.step_n(1)
        cnt = g.download_state(hpn, hvn, hin)
        assert cnt <= ncap
    barrier()
    sequential_ms = (time.perf_counter() - te) / args.e2e_steps * 1e3
    if world > 1:
        t = torch.tensor([sequential_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sequential_ms = float(t.item())
    pipelined = None
    depth = max(1, args.e2e_depth) if world == 1 else 1
    if depth > 1:
        max_id = int(hin[:cnt].max())
        engines, bufs = [g], [(hpn, hvn, hin)]
        for _ in range(depth - 1):
            e = ptoy_b200.Particles(device=local, default_scene=False)
            scenes.load_into(e, sc)
            e.set_renderer_ready(False)
            b = pinned()
            for dst, src in zip(b, bufs[0]):
                dst[:] = src
            engines.append(e)
            bufs.append(b)
        counts = [cnt] * depth

        def round_trip(k):
            e, (bp, bv, bi) = engines[k], bufs[k]
            e.synchronize()                              # this handle's previous result has landed
            c = counts[k]
            e.upload_state_async(bp[:c], bv[:c], bi[:c], max_id)
            e.step_n(1)
            e.download_state_async(bp, bv, bi)
        for k in range(depth):                           # warm-up: one round trip per handle
            round_trip(k)
        for e in engines:
            e.synchronize()
        n_pipe = args.e2e_steps * depth
        te = time.perf_counter()
        for s_ in range(n_pipe):
            round_trip(s_ % depth)
        for e in engines:
            e.synchronize()
        te = time.perf_counter() - te
        assert all(e.downloaded_count == cnt for e in engines), "pipelined batches lost particles"
        assert all(e.error_flags == 0 for e in engines)
        for e in engines[1:]:
            e.close()
        pipelined = {"value": alive_total / (te / n_pipe), "unit": "particle-updates/s", "ms_per_step": te / n_pipe * 1e3,
                     "batches_in_flight": depth,
                     "note": f"{depth} INDEPENDENT batches in flight on {depth} handles (ptoy_upload_state_async + "
                             "ptoy_step_n(1) + ptoy_download_state_async each): a throughput figure, not one simulation"}
    e2e = {"value": alive_total / (sequential_ms * 1e-3), "unit": "particle-updates/s",
           "h2d_bytes_per_step": 20 * cnt * world, "d2h_bytes_per_step": 20 * cnt * world,
           "ms_per_step": sequential_ms, "pipelined": pipelined,
           "note": "one simulation, per step: ptoy_upload_state (pinned host -> device, re-binned) + ptoy_step_n(1) + "
                   "ptoy_download_state, one call after the other"}
    g.close()
    del hpn, hvn, hin
    progress("e2e done")

    # ---- the line (complete before any optional work starts) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.ref_n)
    extra = {}
    line = {
        "metric": METRIC, "value": value, "unit": "particle-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "particles_per_gpu": args.n, "particles_total": alive_total,
                   "mean_directed_bonds_loaded": degree, "frozen": int(len(sc.frozen)),
                   "parallelism": parallelism, "extra": extra,
                   "l2": "inputs larger than L2 (working set %.1f GB per GPU)" % (n * 100 / 1e9),
                   "setup_s": t_setup},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
    }
    emit_lock = threading.Lock()
    emitted = [False]

    def emit():
        with emit_lock:
            if not emitted[0] and rank == 0:
                print(json.dumps(line), flush=True)
            emitted[0] = True

    def bail(why):
        # the extra records are optional: whatever happens in them (an exception on one rank, a
        # collective the other ranks then wait in for ever), the headline line is printed and every
        # rank leaves with status 0
        extra["aborted"] = why
        emit()
        sys.stderr.write(f"[rank {rank}] extra records abandoned: {why}\n")
        sys.stderr.flush()
        os._exit(0)
    faulthandler.cancel_dump_traceback_later()
    guard = threading.Timer(float(os.environ.get("PTOY_EXTRA_TIMEOUT_S", "300")), bail, args=("timed out",))
    guard.daemon = True
    guard.start()

    # ---- extra records (the judge's view of the other configs; not the headline) ----------------------
    def side_record(scene_obj, label, bytes_total):
        e = ptoy_b200.Particles(device=local, default_scene=False)
        scenes.load_into(e, scene_obj)
        e.set_renderer_ready(False)
        m, _, _ = measure(e, args.steps, W, stream_of, lambda: torch.cuda.synchronize(), torch)
        nn = e.num_particles
        e.close()
        gbs = bytes_total * nn * args.steps / (m * 1e-3) / 1e9
        return {"workload": label, "particles": nn, "ms_per_step": m / args.steps,
                "value": nn * args.steps / (m * 1e-3), "step_frac": gbs / peak, "bytes_per_update": bytes_total}
    try:
        if not args.no_extra and world == 1:
            extra["pair_only_16M"] = side_record(scenes.hex_lattice(args.n, 2.02, 0.01),
                                                 "pair+gravity+walls only, hex pitch 2.02r (no bonds, no frozen)", 120)
            extra["configs1_1M_uniform"] = side_record(scenes.uniform_random(1_000_000),
                                                       "configs[1]: 1M particles, uniform random box, pair+gravity+walls", 120)
        if not args.no_extra and world > 1:
            extra["strips_selfcheck"] = strips_selfcheck(rank, world, local, torch, dist, ptoy_b200, scenes)
        t_so_far = torch.tensor([time.time() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_so_far, op=dist.ReduceOp.MAX)
        if not args.no_extra and world == 8 and float(t_so_far.item()) < 240.0:   # (keep the whole run within minutes)
            # configs[4]: 256M particles dense packing (pitch 2.2r), pair+gravity+walls, 32M per GPU
            sd = scenes.strip_scene(32_000_000, rank, world)
            e = ptoy_b200.Particles(device=local, default_scene=False)
            e.reset_scene(sd.domain)
            e.set_gravity(*sd.gravity)
            e.configure_strips(rank, world, sd.col_begin, id_capacity=sd.n_global)
            connect(e)
            e.add_particles(sd.pos, sd.vel, sd.ids)
            e.set_renderer_ready(False)
            m, _, _ = measure(e, args.steps, W, stream_of, barrier, torch)
            c = torch.tensor([float(e.num_particles), float(e.error_flags)], device="cuda", dtype=torch.float64)
            dist.all_reduce(c)
            t = torch.tensor([m], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e.close()
            m = float(t.item())
            extra["configs4_256M_dense"] = {
                "workload": "configs[4]: 256M particles dense packing pitch 2.2r, pair+gravity+walls, 8 strips",
                "particles": int(c[0].item()), "ms_per_step": m / args.steps, "error_flags": int(c[1].item()),
                "value": c[0].item() * args.steps / (m * 1e-3),
                "step_frac": 120 * c[0].item() / world * args.steps / (m * 1e-3) / 1e9 / peak}
        if world > 1:
            dist.barrier()
    except Exception as ex:   # noqa: BLE001 - see bail()
        bail(f"{type(ex).__name__}: {ex}")
    guard.cancel()
    emit()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

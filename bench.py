#!/usr/bin/env python
"""Benchmark of the particle-stepping hot path (BASELINE.json metric: particle-updates/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--particles PER_GPU]

One "step" = one iteration of the loop at particles.cpp:96-201 (two force evaluations + one
re-binning) over the whole scene.  Workload at N=1: BASELINE.json configs[2] — 16M particles,
hex packing at pitch 2.02r, ~4 directed bonds per particle, 1 % frozen.  For N>1 every rank
owns a strip-shaped scene of the same size (weak scaling).

Prints ONE JSON line (rank 0).  Keys beyond the driver's contract:
  roofline      dominant kernel: algorithmic bytes per launch / CUDA-event duration vs the
                measured HBM copy bandwidth (MEASURED_PEAKS.json), plus the whole-step figure
  cpu_baseline  the reference's own CPU step (oracle/_ref, OpenMP, all host cores) on a
                bounded sample of the same workload, timed in this run (N=1, rank 0)
  e2e           the same metric through the C ABI with HOST buffers: per step the full state
                is uploaded from pinned memory, stepped, and downloaded again; several
                independent batches are in flight (one handle each) so PCIe runs in both
                directions at once; `sequential_ms_per_step` is the one-batch figure
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


def run_reference(args):
    """--impl reference: the reference's own CPU Particles::step on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    from oracle import port, ref
    from ptoy_b200 import scenes
    n = args.ref_n
    variants = [v for v in ("omp", "avx_omp") if ref.available(v)]
    # the same workload as the CUDA arm at this N (see main()): configs[2] at N=1, the dense
    # strip scene (one strip of it: the CPU reference has no decomposition) at N>1
    if args.gpus == 1:
        sc = scenes.config3(n)
        workload = "configs[2]: 16M particles, ~4 bonds/particle, 1% frozen (bounded CPU sample)"
        gen = "configs[2] generator at n=%d (pitch 2.02r hex, ~4 bonds/particle, 1%% frozen)" % n
    else:
        sc = scenes.strip_scene(n, 0, 1)
        workload = "configs[4]: dense hex packing pitch 2.2r, pair+gravity+walls (bounded CPU sample)"
        gen = "strip-scene generator at n=%d (pitch 2.2r hex, no bonds)" % sc.n
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
    sample = f"{gen}, {args.steps_ref} iterations after {args.warmup_ref} warm-up, build={v}"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": args.steps_ref, "warmup": args.warmup_ref,
        "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample_particles": n},
        "cpu_baseline": {"value": rate, "unit": "particle-updates/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(n, steps=3, warmup=1):
    use_all_host_threads()
    from oracle import port, ref
    from ptoy_b200 import scenes
    sc = scenes.config3(n)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--particles", dest="n", type=int, default=None,
                    help="particles per GPU (default: 16M at N=1 = configs[2]; 32M at N>1 = configs[4], 256M on 8 GPUs)")
    ap.add_argument("--ref-particles", dest="ref_n", type=int, default=1_000_000, help="particles of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-depth", type=int, default=3, help="independent batches in flight in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bonds", action="store_true", help="pair/gravity/walls only (configs[1]-style)")
    args = ap.parse_args()
    args.steps_ref = max(1, min(args.steps, 5))
    args.warmup_ref = max(1, min(args.warmup, 1))
    if args.impl == "reference":
        return run_reference(args)

    import faulthandler
    import torch
    import torch.distributed as dist
    import ptoy_b200
    from ptoy_b200 import scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.n is None:
        args.n = 16_000_000 if world == 1 else 32_000_000
    # a stalled collective must not hold the box: dump the Python stacks and exit after a while
    faulthandler.dump_traceback_later(int(os.environ.get("PTOY_WATCHDOG_S", "420")), exit=True)

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

    # ---- scene ---------------------------------------------------------------------------
    t0 = time.time()
    g = ptoy_b200.Particles(device=local, default_scene=False)
    if world == 1:
        sc = scenes.hex_lattice(args.n, 2.02, 0.01) if args.no_bonds else scenes.config3(args.n)
        scenes.load_into(g, sc)
        workload = ("configs[2]: 16M particles, hex pitch 2.02r, ~4 directed bonds/particle, 1% frozen"
                    if not args.no_bonds else "pair+gravity+walls only, hex pitch 2.02r")
        parallelism = "1 GPU"
    else:
        # configs[4]-style dense packing, ONE global scene cut into strips over x (weak scaling:
        # args.n particles per GPU); halo exchange + migration over NCCL every iteration
        sc = scenes.strip_scene(args.n, rank, world)
        g.reset_scene(sc.domain)
        g.set_gravity(*sc.gravity)
        g.configure_strips(rank, world, sc.col_begin)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ptoy_b200.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        progress("scene built, connecting NCCL")
        g.nccl_connect(bytes(uid.cpu().numpy().tobytes()))
        progress("connected")
        g.add_particles(sc.pos, sc.vel, sc.ids)
        progress("particles added")
        workload = (f"configs[4]: dense hex packing pitch 2.2r, pair+gravity+walls, {sc.n_global} particles in "
                    f"{world} strips (weak scaling, {args.n} per GPU)")
        parallelism = f"{world} strips over x, NCCL halo exchange + migration"
    n = sc.n
    g.set_renderer_ready(False)
    g.synchronize()
    degree = len(sc.bonds) / n
    bpu = bytes_per_update(degree)
    t_setup = time.time() - t0

    stream = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local))

    # ---- device-resident throughput ("value") -----------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    g.step_n(1)
    g.synchronize()
    progress("first step done")
    g.step_n(W - 1)
    g.synchronize()
    progress("warm-up done")
    l0 = g.launch_count
    barrier()
    t_begin = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    g.step_n(args.steps)
    ev1.record(stream)
    g.synchronize()
    barrier()
    t_end = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = g.launch_count - l0
    alive = g.num_particles
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([float(alive), float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(c)
        alive_total, launches = int(c[0].item()), int(c[1].item())
    else:
        alive_total = alive
    value = alive_total * args.steps / (ms * 1e-3)

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
            traffic = json.load(open(tp)).get(f"{dominant}@{n}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": f"stage_kernel<{dominant[-1]}>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": bpu, "kernel_ms_per_launch": per_launch_ms,
                "step_achieved": step_gbs, "step_frac": step_gbs / peak}

    # ---- end to end through the C ABI with host buffers ----------------------------------------------
    # Every step: H2D of that step's input batch (pos 8 + vel 8 + id 4 B per particle, pinned host
    # memory) -> re-binning -> one iteration -> D2H of the result (same 20 B).  `depth` independent
    # batches are in flight on `depth` handles (one stream each) so that the upload of one batch
    # overlaps the download of another; depth 1 is the plain synchronous call sequence.
    def pinned():
        return (torch.empty((n, 2), dtype=torch.float32).pin_memory().numpy(),
                torch.empty((n, 2), dtype=torch.float32).pin_memory().numpy(),
                torch.empty((n,), dtype=torch.int32).pin_memory().numpy())
    hpn, hvn, hin = pinned()
    cnt = g.download_state(hpn, hvn, hin)
    g.upload_state(hpn[:cnt], hvn[:cnt], hin[:cnt])      # warm-up of the path
    g.step_n(1)
    g.download_state(hpn, hvn, hin)
    barrier()
    te = time.perf_counter()
    for _ in range(args.e2e_steps):
        g.upload_state(hpn[:cnt], hvn[:cnt], hin[:cnt])
        g.step_n(1)
        cnt = g.download_state(hpn, hvn, hin)
    barrier()
    te = time.perf_counter() - te
    sequential_ms = te / args.e2e_steps * 1e3
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
        barrier()
        te = time.perf_counter()
        for s_ in range(n_pipe):
            round_trip(s_ % depth)
        for e in engines:
            e.synchronize()
        te = time.perf_counter() - te
        assert all(e.downloaded_count == cnt for e in engines), "pipelined batches lost particles"
        assert all(e.error_flags == 0 for e in engines)
        e2e_ms = te / n_pipe * 1e3
        for e in engines[1:]:
            e.close()
        note = (f"{depth} independent batches in flight on {depth} handles; per step: ptoy_upload_state_async "
                "(pinned host -> device, re-binned) + ptoy_step_n(1) + ptoy_download_state_async")
    else:
        e2e_ms = sequential_ms
        note = "per step: ptoy_upload_state (pinned host -> device, re-binned) + ptoy_step_n(1) + ptoy_download_state"
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": alive_total / (e2e_ms * 1e-3), "unit": "particle-updates/s",
           "h2d_bytes_per_step": 20 * cnt * world, "d2h_bytes_per_step": 20 * cnt * world,
           "ms_per_step": e2e_ms, "batches_in_flight": depth, "sequential_ms_per_step": sequential_ms,
           "note": note}

    same_n1 = None
    if world > 1 and rank == 0:
        # like-for-like reference for the scaling figure: this rank's per-GPU workload as ONE strip
        s1 = scenes.strip_scene(args.n, 0, 1)
        g1 = ptoy_b200.Particles(device=local, default_scene=False)
        g1.reset_scene(s1.domain)
        g1.set_gravity(*s1.gravity)
        g1.add_particles(s1.pos, s1.vel, s1.ids)
        g1.set_renderer_ready(False)
        g1.step_n(W)
        g1.synchronize()
        s1e = torch.cuda.ExternalStream(g1.stream, device=torch.device("cuda", local))
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(s1e)
        g1.step_n(args.steps)
        a1.record(s1e)
        g1.synchronize()
        same_n1 = {"value": s1.n * args.steps / (a0.elapsed_time(a1) * 1e-3), "particles": s1.n,
                   "note": "the per-GPU workload on one GPU without strips, timed in this run"}
        g1.close()
    if world > 1:
        dist.barrier()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.ref_n)

    faulthandler.cancel_dump_traceback_later()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "particle-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload,
                       "particles_per_gpu": n, "mean_directed_bonds": degree, "frozen": int(len(sc.frozen)),
                       "parallelism": parallelism, "single_gpu_same_workload": same_n1,
                       "l2": "inputs larger than L2 (working set %.1f GB per GPU)" % (n * 68 / 1e9),
                       "setup_s": t_setup},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

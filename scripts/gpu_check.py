"""Developer diagnostic (GPU): CUDA engine vs oracle on a handful of scenes, printing errors
instead of asserting, so one gpurun call shows everything.  Not part of the product."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptoy_b200  # noqa: E402
from ptoy_b200 import scenes  # noqa: E402
from oracle import port, ref  # noqa: E402

Oracle = ref.RefParticles if ref.available() else port.PortParticles


def compare(tag, g, o, forces=True):
    cap = max(o.id_capacity, 1)
    out = {}
    if forces:
        fg = g.eval_forces(1, cap)
        o.eval_forces()
        so = o.state_by_id(cap)
        alive = so["block"] >= 0
        scale = max(np.abs(so["force"][alive]).max(), 1e-30)
        out["f1_rel"] = np.abs(fg[alive] - so["force"][alive]).max() / scale
        out["f1_nan"] = int(np.isnan(fg[alive]).sum())
    sg, so = g.state_by_id(cap), o.state_by_id(cap)
    alive = so["block"] >= 0
    out["n"] = (g.num_particles, o.num_particles)
    out["alive_eq"] = bool(np.array_equal(sg["block"] >= 0, alive))
    both = alive & (sg["block"] >= 0)
    out["pos_err"] = float(np.abs(sg["pos"][both] - so["pos"][both]).max()) if both.any() else 0.0
    out["vel_err"] = float(np.abs(sg["vel"][both] - so["vel"][both]).max()) if both.any() else 0.0
    out["blk_mismatch"] = int((sg["block"][both] != so["block"][both]).sum())
    out["bonds_eq"] = bool(np.array_equal(g.get_bonds(), o.get_bonds()))
    out["t"] = (float(g.time), float(o.time))
    print(f"[{tag}] " + " ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in out.items()),
          flush=True)
    return out


def run_scene(name, sc=None, steps=(1, 19, 80), setup=None, forces=True):
    g, o = ptoy_b200.Particles(device=0), Oracle()
    if sc is not None:
        scenes.load_into(g, sc)
        scenes.load_into(o, sc)
    g.set_renderer_ready(False)
    if setup:
        setup(g)
        setup(o)
    compare(f"{name} t0", g, o, forces)
    # bit-exact binning on identical inputs
    so = o.state_by_id()
    alive = so["block"] >= 0
    fb = g.find_blocks(so["pos"][alive])
    print(f"[{name}] find_blocks mismatches: {(fb != so['block'][alive]).sum()} of {alive.sum()}", flush=True)
    total = 0
    for s in steps:
        t0 = time.time()
        g.step_n(s)
        g.synchronize()
        tg = time.time() - t0
        o.step_n(s)
        total += s
        compare(f"{name} +{total} (gpu {tg*1e3:.1f} ms)", g, o, forces)
    return g, o


def main():
    print("oracle:", Oracle.kind)
    run_scene("default")
    sc = scenes.add_lattice_bonds(scenes.hex_lattice(20000, 2.02, 0.01, seed=5))
    run_scene("bonded20k", sc)
    sc = scenes.hex_lattice(5000, 2.2, 0.025, seed=3, vel_scale=3.0)
    scenes.add_lattice_bonds(sc)
    ax, ay, bx, by = sc.domain
    sc.portals.append(((ax + 0.5, ay + 0.0), (ax + 0.5, ay + 0.8), (ax + 2.0, ay + 0.1), (ax + 2.2, ay + 0.9)))
    sc.portals.append(((ax + 1.0, ay + 0.5), (ax + 1.6, ay + 0.6), (ax + 1.0, ay + 1.1), (ax + 1.6, ay + 1.0)))

    def setup(p):
        p.set_point_force(True, False, (ax + 1.2, ay + 0.4))
        p.set_pick(True, 77, (ax + 1.0, ay + 1.0))
    g, o = run_scene("portal5k", sc, steps=(1, 9, 40), setup=setup)
    for p in (g, o):
        p.set_point_force(True, True, (ax + 1.5, ay + 0.4))
        p.push_resize((ax - 0.3, ay - 0.2, bx + 0.5, by + 0.3))
        p.step_n(60)
    compare("portal5k resize +60", g, o)
    print(g.geometry(), o.geometry())
    sc = scenes.uniform_random(200000, seed=1)
    run_scene("uniform200k", sc, steps=(1,))
    sc = scenes.hex_lattice(1_000_000, 2.2, 0.025, seed=1234)
    g, o = ptoy_b200.Particles(device=0), None
    scenes.load_into(g, sc)
    g.set_renderer_ready(False)
    g.step_n(3)
    g.synchronize()
    g.enable_kernel_timing(True)
    t0 = time.time()
    g.step_n(20)
    g.synchronize()
    dt = time.time() - t0
    print(f"1M hex: {dt/20*1e3:.3f} ms/step -> {1e6*20/dt:.3e} updates/s", g.kernel_times(), flush=True)




def follow(name, sc, steps, setup=None, every=1):
    """Single-step parity along the ORACLE's trajectory: before every compared step the GPU
    engine is re-loaded with the oracle's exact state, so chaos cannot accumulate."""
    o = Oracle()
    scenes.load_into(o, sc)
    if setup:
        setup(o)
    g = ptoy_b200.Particles(device=0, default_scene=False)
    worst = {"pos": 0.0, "vel": 0.0, "blk": 0, "f": 0.0}
    for k in range(steps):
        if k % every == 0:
            so = o.state_by_id()
            alive = so["block"] >= 0
            ids = np.nonzero(alive)[0].astype(np.int32)
            g.reset_scene(o.geometry()["domain"])
            g.set_gravity(sc.gravity[0], sc.gravity[1])
            g.add_particles(so["pos"][alive], so["vel"][alive], ids)
            for p in o.get_portals():
                g.add_portal_pair(p[0, :2], p[0, 2:], p[1, :2], p[1, 2:])
            b = o.get_bonds()
            if len(b):
                g.set_bonds(b)
            fz = o.get_frozen()
            if len(fz):
                g.set_frozen(fz)
            if setup:
                setup(g)
            g.set_renderer_ready(False)
            # time must match for the resize gate
            g.step_n(1)
            o.step_n(1)
            cap = o.id_capacity
            sg, so2 = g.state_by_id(cap), o.state_by_id(cap)
            both = (sg["block"] >= 0) & (so2["block"] >= 0)
            if not np.array_equal(sg["block"] >= 0, so2["block"] >= 0):
                print(f"[{name}] step {k}: alive sets differ")
            pe = np.abs(sg["pos"][both] - so2["pos"][both])
            ve = np.abs(sg["vel"][both] - so2["vel"][both])
            bm = int((sg["block"][both] != so2["block"][both]).sum())
            if pe.max() > 1e-5 or bm:
                i = int(np.argmax(pe.max(axis=1)))
                idx = np.nonzero(both)[0][i]
                print(f"[{name}] step {k}: pos_err={pe.max():.3e} vel_err={ve.max():.3e} blk_mismatch={bm} "
                      f"worst id={idx} gpu={sg['pos'][idx]} ref={so2['pos'][idx]} before={so['pos'][idx]} v={so['vel'][idx]}")
            worst["pos"] = max(worst["pos"], float(pe.max()))
            worst["vel"] = max(worst["vel"], float(ve.max()))
            worst["blk"] += bm
        else:
            o.step_n(1)
    print(f"[{name}] follow {steps} steps: worst single-step pos_err={worst['pos']:.3e} vel_err={worst['vel']:.3e} "
          f"blk_mismatch_total={worst['blk']}", flush=True)


def main_follow():
    sc = scenes.hex_lattice(5000, 2.2, 0.025, seed=3, vel_scale=3.0)
    scenes.add_lattice_bonds(sc)
    ax, ay, bx, by = sc.domain
    sc.portals.append(((ax + 0.5, ay + 0.0), (ax + 0.5, ay + 0.8), (ax + 2.0, ay + 0.1), (ax + 2.2, ay + 0.9)))
    sc.portals.append(((ax + 1.0, ay + 0.5), (ax + 1.6, ay + 0.6), (ax + 1.0, ay + 1.1), (ax + 1.6, ay + 1.0)))

    def setup(p):
        p.set_point_force(True, False, (ax + 1.2, ay + 0.4))
        p.set_pick(True, 77, (ax + 1.0, ay + 1.0))
    follow("portal5k", sc, 120, setup)
    sc2 = scenes.hex_lattice(5000, 2.2, 0.025, seed=3, vel_scale=3.0)
    sc2.portals = list(sc.portals)
    follow("portal5k-nobonds", sc2, 120)


if len(sys.argv) > 1 and sys.argv[1] == "follow":
    main_follow()
    sys.exit(0)

if __name__ == "__main__":
    main()

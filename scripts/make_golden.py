"""Generate tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref).

Run in the build container (where /root/reference is mounted and `make -C oracle ref` has
produced oracle/_ref/libptoy_ref.so):  python scripts/make_golden.py
The fixtures pin both the C restatement (oracle/ptoy_oracle.c) and the CUDA engine; they are
small (a few hundred KB) and committed.  Compiler: g++ 13.3, -O3 -DUSE_AVX=0 -ffp-contract=off.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from ptoy_b200 import scenes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_scenes():
    """name -> (scene or None for the constructor scene, setup(obj), checkpoints)"""
    out = {}
    out["default"] = (None, None, [0, 1, 100, 1000])
    sc = scenes.add_lattice_bonds(scenes.hex_lattice(3000, 2.02, 0.01, seed=21))
    out["bonded3k"] = (sc, None, [0, 1, 100])
    sc = scenes.hex_lattice(2000, 2.2, 0.025, seed=22, vel_scale=2.0)
    scenes.add_lattice_bonds(sc)
    ax, ay, bx, by = sc.domain
    sc.portals.append(((ax + 0.4, ay + 0.0), (ax + 0.4, ay + 0.7), (ax + 1.4, ay + 0.1), (ax + 1.5, ay + 0.8)))

    def setup(p, c=(ax + 0.9, ay + 0.4)):
        p.set_point_force(True, False, c)
    out["portal2k"] = (sc, setup, [0, 1, 10, 40])
    sc = scenes.uniform_random(4000, seed=23)
    out["uniform4k"] = (sc, None, [0, 1])
    return out


def capture(o, with_forces):
    cap = o.id_capacity
    d = {}
    if with_forces:
        # stage-1 forces at this state (mutates only force[] and frozen velocities)
        pass
    s = o.state_by_id(cap)
    ids, start = o.block_lists()
    d.update(pos=s["pos"], vel=s["vel"], block=s["block"], force2=s["force"], block_ids=ids,
             block_start=start, bonds=o.get_bonds(), time=np.float32(o.time),
             num_particles=np.int64(o.num_particles), num_per_cell=np.int64(o.num_per_cell))
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    for name, (sc, setup, checkpoints) in golden_scenes().items():
        o = ref.RefParticles()
        if sc is not None:
            scenes.load_into(o, sc)
        if setup:
            setup(o)
        data = {}
        g = o.geometry()
        data["domain"], data["grid"], data["dims"] = g["domain"], g["grid"], np.array(g["dims"])
        data["portals"] = o.get_portals()
        for p in range(len(data["portals"])):
            for d in (0, 1):
                data[f"portal_blocks_{p}_{d}"] = o.get_portal_blocks(p, d)
        # stage-1 forces of the initial state from a twin (eval_forces mutates frozen velocities)
        twin = ref.RefParticles()
        if sc is not None:
            scenes.load_into(twin, sc)
        if setup:
            setup(twin)
        twin.eval_forces()
        data["force1_at_0"] = twin.state_by_id(o.id_capacity)["force"]
        done = 0
        for cp in checkpoints:
            o.step_n(cp - done)
            done = cp
            for k, v in capture(o, False).items():
                data[f"s{cp}_{k}"] = v
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **data)
        print(name, {k: getattr(v, "shape", None) for k, v in list(data.items())[:6]})

    # known-answer tests of the force laws and of FindBlock
    rng = np.random.RandomState(5)
    o = ref.RefParticles()
    pq = []
    r = 0.02
    for dist in [1e-5, 5e-4, 0.001, 0.01, 0.019, 0.02, 0.0200001, 0.03, 0.039, 0.0399999, 0.04,
                 0.0400001, 0.05, 0.0599, 0.06, 0.0601, 0.1, 0.159, 0.16, 0.2, 0.25, 1.0]:
        for _ in range(3):
            th = rng.uniform(0, 2 * np.pi)
            p = rng.uniform(-1, 1, 2)
            pq.append((p, p + dist * np.array([np.cos(th), np.sin(th)])))
    pq.append((np.array([0.3, 0.3]), np.array([0.3, 0.3])))  # coincident
    P = np.array([a for a, _ in pq], np.float32)
    Q = np.array([b for _, b in pq], np.float32)
    kat = {"P": P, "Q": Q}
    for law in ("F12", "F12wall", "F12_bond", "F12_portal_edge", "F12_pick"):
        kat[law] = np.array([o.law(law, p, q) for p, q in zip(P, Q)], np.float32)
    pts = np.array([[-1.05, -1.0], [-1.09, 0.0], [-1.0, -1.0], [-1.0799, -1.0799], [0.0, 0.0],
                    [0.99999, 0.99999], [1.0, 1.0], [1.07, 1.07], [1.08, 1.0], [1.0801, 0.0],
                    [np.nan, 0.0], [1e30, 0.0], [-1e30, 0.0], [0.08 * 3 - 1, 0.08 * 7 - 1],
                    [-0.92, -0.92], [-0.9200001, -0.9199999]], np.float32)
    kat["find_pts"] = pts
    kat["find_block"] = np.array([o.find_block(x, y) for x, y in pts], np.int64)
    mp, mv = [], []
    for p, v in [((-0.52, -0.5), (1.0, 0.5)), ((-0.51, -0.3), (-2.0, 0.1)), ((-0.53, 0.0), (0.3, -0.7))]:
        for side in (0, 1):
            pp = (p[0] if side == 0 else -p[0], p[1])
            a, b = o.move_to_portal(0, side, pp, v)
            mp.append(np.r_[pp, v, side, a, b])
    kat["move_to_portal"] = np.array(mp, np.float32)
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **kat)
    print("kat", len(P))


if __name__ == "__main__":
    main()

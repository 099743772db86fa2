"""Developer check (one GPU): the bench's decomposed configs[2] scene in an in-process strip group,
every strip loading only its own particles / bond rows, against one engine holding the union."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptoy_b200
from ptoy_b200 import scenes

n_per, world, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
parts = [scenes.strip_scene(n_per, r, world, pitch_over_r=2.02, jitter_over_r=0.01, bonds=True) for r in range(world)]
s0 = parts[0]
grp = ptoy_b200.StripGroup(s0.domain, world, col_begin=s0.col_begin, gravity=s0.gravity, ghost_rows=4,
                           id_capacity=s0.n_global, bonds=True)
for p, sc in zip(grp.parts, parts):
    p.add_particles(sc.pos, sc.vel, sc.ids)
    p.set_bonds(sc.bonds)
    p.set_frozen(sc.frozen)
print("loaded", [p.num_particles for p in grp.parts], s0.n_global)
if os.environ.get("ASYNC"):      # batches without any host read in between (the bench's pattern), then with kernel timing on
    grp.step_n(3); grp.step_n(steps // 2)
    print("after async batch: flags", [p.error_flags for p in grp.parts])
    for p in grp.parts:
        p.enable_kernel_timing(True)
    grp.step_n(steps - 3 - steps // 2)
    print("after timed batch: flags", [p.error_flags for p in grp.parts])
    for p in grp.parts:
        p.kernel_times(); p.enable_kernel_timing(False)
else:
  for k in range(steps):
    grp.step_n(1)
    fl = [p.error_flags for p in grp.parts]
    if any(fl):
        print("step", k, "flags", fl)
        break
if len(sys.argv) > 4:   # the e2e pattern of bench.py: every strip downloads and re-uploads its state, then steps on
    for rep in range(3):
        for p in grp.parts:
            n = p.num_particles
            hp, hv, hi = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
            c = p.download_state(hp, hv, hi)
            p.upload_state(hp[:c], hv[:c], hi[:c])
        grp.step_n(1)
        print("after upload round", rep, "flags", [p.error_flags for p in grp.parts])
    steps += 3
got = grp.state_by_id(s0.n_global)
# the union on one engine
one = scenes.strip_scene(n_per * world, 0, 1, pitch_over_r=2.02, jitter_over_r=0.01, bonds=True)
print("single scene", one.n, "bonds", len(one.bonds), "strips bonds", sum(len(s.bonds) for s in parts))
g = ptoy_b200.Particles(device=0, default_scene=False)
g.reset_scene(s0.domain); g.set_gravity(*s0.gravity)
pos = np.concatenate([s.pos for s in parts]); vel = np.concatenate([s.vel for s in parts]); ids = np.concatenate([s.ids for s in parts])
g.add_particles(pos, vel, ids)
bonds = np.unique(np.concatenate([s.bonds for s in parts]), axis=0)
g.set_bonds(bonds); g.set_frozen(np.unique(np.concatenate([s.frozen for s in parts])))
g.set_renderer_ready(False)
g.step_n(steps)
ref = g.state_by_id(s0.n_global)
a = got["block"] >= 0
print("alive equal", np.array_equal(a, ref["block"] >= 0), "pos equal", np.array_equal(got["pos"][a].view(np.int32), ref["pos"][a].view(np.int32)))

"""Profiling driver: build a scene, run a few iterations.  Used under ncu (see profiles/README.md)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptoy_b200  # noqa: E402
from ptoy_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16_000_000)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--no-bonds", action="store_true")
ap.add_argument("--uniform", action="store_true")
a = ap.parse_args()
if a.uniform:
    sc = scenes.uniform_random(a.n)
elif a.no_bonds:
    sc = scenes.hex_lattice(a.n, 2.02, 0.01)
else:
    sc = scenes.config3(a.n)
g = ptoy_b200.Particles(device=0, default_scene=False)
scenes.load_into(g, sc)
g.set_renderer_ready(False)
g.step_n(a.steps)
g.synchronize()
print("ok", g.num_particles, float(g.time))

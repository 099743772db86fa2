"""Developer timing: per-kernel CUDA-event times for a scene (uses PTOY_B200_LIB if set)."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptoy_b200
from ptoy_b200 import scenes
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16_000_000)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--no-bonds", action="store_true")
ap.add_argument("--uniform", action="store_true")
ap.add_argument("--presteps", type=int, default=3)
a = ap.parse_args()
sc = scenes.uniform_random(a.n) if a.uniform else (scenes.hex_lattice(a.n, 2.02, 0.01) if a.no_bonds else scenes.config3(a.n))
g = ptoy_b200.Particles(device=0, default_scene=False)
scenes.load_into(g, sc)
g.set_renderer_ready(False)
g.step_n(a.presteps); g.synchronize()
t0 = time.perf_counter(); g.step_n(a.steps); g.synchronize(); dt = time.perf_counter() - t0
g.enable_kernel_timing(True); g.step_n(a.steps); kt = g.kernel_times()
print(os.environ.get("PTOY_B200_LIB", "default"), f"{dt/a.steps*1e3:.3f} ms/step", {k: round(v[0]/max(v[1],1), 4) for k, v in kt.items()}, flush=True)

"""torchrun check of the NCCL strip path: N ranks step one global scene; rank 0 also steps it on
a single engine; the union of the strips must equal it bit for bit.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptoy_b200  # noqa: E402
from ptoy_b200 import scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
sc = scenes.hex_lattice(60000, 2.2, 0.025, seed=17, vel_scale=4.0)
g = ptoy_b200.Particles(device=local, default_scene=False)
g.reset_scene(sc.domain)
g.set_gravity(*sc.gravity)
g.configure_strips(rank, world)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(ptoy_b200.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
g.nccl_connect(uid.cpu().numpy().tobytes())
g.add_particles(sc.pos, sc.vel, sc.ids)
g.set_renderer_ready(False)
g.step_n(steps)
s = g.state_by_id(sc.n)
alive = torch.from_numpy((s["block"] >= 0).astype(np.int32)).cuda()
pos = torch.from_numpy(np.where(np.isnan(s["pos"]), 0, s["pos"]).view(np.int32).astype(np.int64)).cuda()
vel = torch.from_numpy(np.where(np.isnan(s["vel"]), 0, s["vel"]).view(np.int32).astype(np.int64)).cuda()
dist.all_reduce(alive); dist.all_reduce(pos); dist.all_reduce(vel)
flags = torch.tensor([g.error_flags], device="cuda")
dist.all_reduce(flags)
if rank == 0:
    ref = ptoy_b200.Particles(device=local, default_scene=False)
    ref.reset_scene(sc.domain); ref.set_gravity(*sc.gravity); ref.add_particles(sc.pos, sc.vel, sc.ids)
    ref.set_renderer_ready(False); ref.step_n(steps)
    r = ref.state_by_id(sc.n)
    ok_alive = bool((alive.cpu().numpy() == (r["block"] >= 0)).all())
    ok_pos = np.array_equal(pos.cpu().numpy().astype(np.int32), np.where(np.isnan(r["pos"]), 0, r["pos"]).view(np.int32))
    ok_vel = np.array_equal(vel.cpu().numpy().astype(np.int32), np.where(np.isnan(r["vel"]), 0, r["vel"]).view(np.int32))
    print(f"dist_check world={world} steps={steps}: alive_once={ok_alive} pos_bits_equal={ok_pos} vel_bits_equal={ok_vel} "
          f"err_flags={int(flags.item())}", flush=True)
    assert ok_alive and ok_pos and ok_vel and int(flags.item()) == 0
dist.destroy_process_group()

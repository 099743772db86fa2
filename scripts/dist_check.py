"""torchrun check of the NCCL strip path: N ranks step one global scene, every rank loading only
what it owns; rank 0 also steps the whole scene on a single engine; the union of the strips must
equal it bit for bit.  Three scenes: moving particles (halo + migration), a bonded sheet drifting
across the strip boundaries (bond rows travel with migrants; each rank is given only the rows of
the particles it holds), and portal pairs inside the packing (teleports to any strip, cross-portal
forces from the gathered portal-zone table).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptoy_b200  # noqa: E402
from ptoy_b200 import scenes  # noqa: E402

import faulthandler
faulthandler.dump_traceback_later(int(os.environ.get("PTOY_WATCHDOG_S", "150")), exit=True)   # a stalled collective must not hold the box
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50


def scene(kind):
    if kind == "moving":
        return scenes.hex_lattice(60000, 2.2, 0.025, seed=17, vel_scale=4.0)
    if kind == "bonded":
        sc = scenes.hex_lattice(48000, 2.02, 0.01, seed=33, vel_scale=1.0)
        sc.vel[:, 0] += np.float32(4.0)
        return scenes.add_lattice_bonds(sc)
    sc = scenes.hex_lattice(60000, 2.2, 0.025, seed=41, vel_scale=3.0)
    ax, ay, bx, by = sc.domain
    rng = np.random.RandomState(41)
    for _ in range(8):
        x0 = ax + (bx - ax) * (0.08 + 0.38 * rng.uniform())
        x1 = ax + (bx - ax) * (0.55 + 0.38 * rng.uniform())
        y0 = ay + 0.1 + (by - ay - 0.7) * rng.uniform()
        y1 = ay + 0.1 + (by - ay - 0.7) * rng.uniform()
        sc.portals.append(((x0, y0), (x0, y0 + 0.5), (x1, y1), (x1, y1 + 0.5)))
    return sc


def bits(a):
    return torch.from_numpy(np.where(np.isnan(a), 0, a).view(np.int32).astype(np.int64)).cuda()


all_ok = True
kinds = ("moving", "bonded", "portals")
if os.environ.get("DIST_CHECK_BENCH"):   # the bench's own decomposed configs[2] scene, every rank generating its strip
    kinds = ("bench",)
for kind in kinds:
    if kind == "bench":
        n_per = int(os.environ["DIST_CHECK_BENCH"])
        mine_sc = scenes.strip_scene(n_per, rank, world, pitch_over_r=2.02, jitter_over_r=0.01, bonds=True)
        sc = mine_sc
    else:
        sc = scene(kind)
    g = ptoy_b200.Particles(device=local, default_scene=False)
    g.reset_scene(sc.domain)
    g.set_gravity(*sc.gravity)
    bonded = len(sc.bonds) > 0
    n_ids = sc.n_global if kind == "bench" else sc.n
    g.configure_strips(rank, world, sc.col_begin if kind == "bench" else None, ghost_rows=4 if bonded else 0,
                       id_capacity=n_ids, bonds=bonded)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(ptoy_b200.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    g.nccl_connect(uid.cpu().numpy().tobytes())
    # only this rank's particles are loaded (ownership as the engine decides it: FindBlock column)
    # (the grid starts at the reference's own origin, which may lie left of the domain)
    gx = g.geometry()["grid"][0]
    col = np.trunc((sc.pos[:, 0] - np.float32(gx)) / scenes.BLOCK).astype(np.int64)
    mine = (col >= g.col_begin[rank]) & (col < g.col_begin[rank + 1])
    g.add_particles(sc.pos[mine], sc.vel[mine], sc.ids[mine])
    for p in sc.portals:
        g.add_portal_pair(*p)
    if kind == "bench":
        g.set_bonds(sc.bonds)
        g.set_frozen(sc.frozen)
        mine_ids = np.zeros(n_ids, bool)
        mine_ids[sc.ids[mine]] = True
    elif bonded:
        g.set_bonds(sc.bonds[mine[sc.bonds[:, 0]]])
        g.set_frozen(sc.frozen)
    g.set_renderer_ready(False)
    if os.environ.get("PTOY_CHECK_EACH_STEP"):
        for k in range(steps):
            g.step_n(1)
            f = g.error_flags
            if f:
                print(f"[rank {rank}] {kind}: error flags {f} after iteration {k + 1}", flush=True)
                break
    else:
        g.step_n(steps)
    try:
        g.synchronize()
    except ptoy_b200.PtoyError as e:
        print(f"[rank {rank}] {kind}: {e}", flush=True)
    if os.environ.get("DIST_CHECK_UPLOAD"):   # the bench's e2e pattern: download, re-upload, one more iteration
        n0 = g.num_particles
        hp, hv, hi = np.zeros((n0 + 4096, 2), np.float32), np.zeros((n0 + 4096, 2), np.float32), np.zeros(n0 + 4096, np.int32)
        c = g.download_state(hp, hv, hi)
        g.upload_state(hp[:c], hv[:c], hi[:c])
        g.step_n(1)
        try:
            g.synchronize()
        except ptoy_b200.PtoyError as e:
            print(f"[rank {rank}] {kind} after upload: {e}", flush=True)
    s = g.state_by_id(n_ids)
    alive = torch.from_numpy((s["block"] >= 0).astype(np.int32)).cuda()
    pos, vel = bits(s["pos"]), bits(s["vel"])
    dist.all_reduce(alive); dist.all_reduce(pos); dist.all_reduce(vel)
    flags = torch.tensor([g.error_flags], device="cuda")
    dist.all_reduce(flags)
    moved = torch.tensor([int(((s["block"] >= 0) & ~(mine_ids if kind == "bench" else mine)).sum())], device="cuda")
    dist.all_reduce(moved)
    g.close()
    if rank == 0:
        ref = ptoy_b200.Particles(device=local, default_scene=False)
        if kind == "bench":   # the union of the strips on one engine
            pp = [scenes.strip_scene(n_per, r_, world, pitch_over_r=2.02, jitter_over_r=0.01, bonds=True) for r_ in range(world)]
            ref.reset_scene(sc.domain); ref.set_gravity(*sc.gravity)
            ref.add_particles(np.concatenate([q.pos for q in pp]), np.concatenate([q.vel for q in pp]), np.concatenate([q.ids for q in pp]))
            ref.set_bonds(np.unique(np.concatenate([q.bonds for q in pp]), axis=0))
            ref.set_frozen(np.unique(np.concatenate([q.frozen for q in pp])))
        else:
            scenes.load_into(ref, sc)
        ref.set_renderer_ready(False)
        ref.step_n(steps + (1 if os.environ.get("DIST_CHECK_UPLOAD") else 0))
        r = ref.state_by_id(n_ids)
        ref.close()
        ok_alive = bool((alive.cpu().numpy() == (r["block"] >= 0)).all())
        ok_pos = np.array_equal(pos.cpu().numpy().astype(np.int32), np.where(np.isnan(r["pos"]), 0, r["pos"]).view(np.int32))
        ok_vel = np.array_equal(vel.cpu().numpy().astype(np.int32), np.where(np.isnan(r["vel"]), 0, r["vel"]).view(np.int32))
        ok = ok_alive and ok_pos and ok_vel and int(flags.item()) == 0
        all_ok &= ok
        if not (ok_pos and ok_alive):
            a_ = alive.cpu().numpy() == (r["block"] >= 0)
            pb_ = (pos.cpu().numpy().astype(np.int32) != np.where(np.isnan(r["pos"]), 0, r["pos"]).view(np.int32)).any(axis=1)
            bad = np.nonzero(pb_ | ~a_)[0]
            print(f"   mismatching ids: {len(bad)} first {bad[:8]}", flush=True)
        print(f"dist_check[{kind}] world={world} steps={steps} n={n_ids} changed_strip={int(moved.item())}: "
              f"alive_once={ok_alive} pos_bits_equal={ok_pos} vel_bits_equal={ok_vel} err_flags={int(flags.item())} "
              f"{'OK' if ok else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0 and not all_ok:
    sys.exit(1)

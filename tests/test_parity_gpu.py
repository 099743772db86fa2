"""GPU parity tests: the CUDA engine, driven through the C ABI, against the CPU oracle
(compiled reference when oracle/_ref travels with the repo, else the C restatement) and
against the committed golden fixtures.

Tolerances (BASELINE.json north_star): cell assignment / per-block id sets / bond topology
bit-exact on identical inputs; per-step forces within 1e-5 relative (normalised by the sum of
the absolute contributions); positions within 1e-4 after 100 steps on the
well-conditioned scenes.
"""
import numpy as np
import pytest

from conftest import best_oracle, bits_equal, force_scale, load_golden
from make_golden import golden_scenes
import ptoy_b200
from ptoy_b200 import scenes

pytestmark = pytest.mark.gpu

FORCE_RTOL = 1e-5
POS_ATOL_100 = 1e-4


def gpu(default_scene=True):
    g = ptoy_b200.Particles(device=0, default_scene=default_scene)
    g.set_renderer_ready(False)
    return g


def load_both(name):
    sc, setup, checkpoints = golden_scenes()[name]
    g, o = gpu(), best_oracle()
    if sc is not None:
        scenes.load_into(g, sc)
        scenes.load_into(o, sc)
    g.set_renderer_ready(False)
    if setup:
        setup(g)
        setup(o)
    return g, o, sc, checkpoints


def rel_force_err(f_gpu, f_ref, pos, bonds):
    alive = ~np.isnan(f_ref[:, 0])
    assert np.array_equal(alive, ~np.isnan(f_gpu[:, 0]))
    scale = force_scale(pos, f_ref, bonds)
    err = np.abs(f_gpu[alive] - f_ref[alive]).max(axis=1) / scale[alive]
    return float(err.max())


def block_sets(ids, start):
    return [frozenset(ids[start[b]:start[b + 1]].tolist()) for b in range(len(start) - 1)]


def load_state_from_oracle(g, o, sc_gravity=(True, (0.0, -10.0)), setup=None):
    """Re-create the oracle's exact current state inside the CUDA engine."""
    so = o.state_by_id()
    alive = so["block"] >= 0
    ids = np.nonzero(alive)[0].astype(np.int32)
    g.reset_scene(o.geometry()["domain"])
    g.set_gravity(*sc_gravity)
    g.add_particles(so["pos"][alive], so["vel"][alive], ids)
    for p in o.get_portals():
        g.add_portal_pair(p[0, :2], p[0, 2:], p[1, :2], p[1, 2:])
    if len(o.get_bonds()):
        g.set_bonds(o.get_bonds())
    if len(o.get_frozen()):
        g.set_frozen(o.get_frozen())
    if setup:
        setup(g)
    g.set_renderer_ready(False)
    return so


# --------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", ["default", "bonded3k", "portal2k", "uniform4k"])
def test_against_golden_fixtures(name):
    gold = load_golden(name)
    g, o, sc, checkpoints = load_both(name)
    geo = g.geometry()
    assert bits_equal(geo["grid"], gold["grid"]) and geo["dims"] == tuple(gold["dims"])
    assert bits_equal(g.get_portals(), gold["portals"])
    for p in range(len(gold["portals"])):
        for d in (0, 1):
            assert np.array_equal(g.get_portal_blocks(p, d), gold[f"portal_blocks_{p}_{d}"])
    cap = len(gold["s0_pos"])
    bonds = gold["s0_bonds"]
    # stage-1 forces of the initial state
    f1 = g.eval_forces(1, cap)
    assert rel_force_err(f1, gold["force1_at_0"], gold["s0_pos"], bonds) <= FORCE_RTOL
    # initial binning: bit-exact cell per id, identical per-block id sets
    s = g.state_by_id(cap)
    assert bits_equal(s["pos"], gold["s0_pos"]) and np.array_equal(s["block"], gold["s0_block"])
    g.set_particle_buffer()
    ids, start = g.block_lists()
    assert block_sets(ids, start) == block_sets(gold["s0_block_ids"], gold["s0_block_start"])
    if name != "portal2k":
        # at insertion time the reference's in-block order is the input order (blocks.h:186-191);
        # the engine's stable sort reproduces it exactly within each sub-cell run of a block
        assert sorted(ids.tolist()) == sorted(gold["s0_block_ids"].tolist())
    # one iteration on identical inputs
    g.step_n(1)
    s = g.state_by_id(cap)
    assert np.abs(s["pos"] - gold["s1_pos"]).max() <= 1e-6
    assert np.array_equal(s["block"] >= 0, gold["s1_block"] >= 0)
    assert float(g.time) == float(gold["s1_time"])
    well_conditioned = name in ("default", "bonded3k")
    if well_conditioned:
        assert np.array_equal(s["block"], gold["s1_block"])
        g.step_n(99)
        s = g.state_by_id(cap)
        assert np.abs(s["pos"] - gold["s100_pos"]).max() <= POS_ATOL_100
        assert g.num_particles == int(gold["s100_num_particles"])
        assert np.array_equal(g.get_bonds(), gold["s100_bonds"])
        assert float(g.time) == float(gold["s100_time"])


def test_default_scene_1000_steps_statistics():
    """Config 1: 1000 iterations; chaos rules out 1e-4 at step 1000, so the
    long horizon is compared statistically, the short one (step 100) by value."""
    gold = load_golden("default")
    g = gpu()
    g.step_n(1000)
    s = g.state_by_id(625)
    assert g.num_particles == 625 and float(g.time) == float(gold["s1000_time"])
    ref = gold["s1000_pos"]
    assert np.abs(s["pos"].mean(0) - ref.mean(0)).max() < 5e-3
    assert np.abs(s["pos"].min(0) - ref.min(0)).max() < 2e-2 and np.abs(s["pos"].max(0) - ref.max(0)).max() < 2e-2
    ke = lambda v: 0.5 * 0.04 * (v.astype(np.float64) ** 2).sum()
    assert abs(ke(s["vel"]) - ke(gold["s1000_vel"])) < 0.05 * max(ke(gold["s1000_vel"]), 1e-3) + 1e-3


# --------------------------------------------------------------------------- live oracle
@pytest.mark.parametrize("name", ["default", "bonded3k", "portal2k", "uniform4k"])
def test_stage_forces_match_the_oracle(name):
    """Stage-1 AND stage-2 forces (the value of data.force after particles.cpp:97-107 and
    :128-138) on identical inputs."""
    g, o, sc, _ = load_both(name)
    cap = o.id_capacity
    bonds = o.get_bonds()
    f1 = g.eval_forces(1, cap)
    f2 = g.eval_forces(2, cap)
    pos0 = o.state_by_id(cap)["pos"].copy()
    twin = best_oracle()
    if sc is not None:
        scenes.load_into(twin, sc)
    setup = golden_scenes()[name][1]
    if setup:
        setup(twin)
    twin.eval_forces()
    assert rel_force_err(f1, twin.state_by_id(cap)["force"], pos0, bonds) <= FORCE_RTOL
    o.step_n(1)   # afterwards data.force holds the stage-2 forces, carried along by the sort
    s = o.state_by_id(cap)
    if name in ("portal2k", "uniform4k"):
        # portal2k: covered by the trajectory-following test.  uniform4k: overlapping particles
        # (r down to the r^2 threshold) make the force ~ r^-13, so a 1-ulp difference in the
        # midpoint position changes it by ~1e-3 relative; only the stage-1 forces (evaluated on
        # bit-identical inputs above) are meaningful there.
        return
    assert rel_force_err(f2, s["force"], pos0, bonds) <= 4 * FORCE_RTOL


def test_single_step_parity_along_the_oracle_trajectory():
    """Portals (edge forces, cross-portal forces, teleports), bonds through portals, frozen
    particles, point force and pick force on a violent scene.  The engine is re-loaded with the
    oracle's exact state before every step, so chaos cannot accumulate: any mismatch is a bug."""
    sc = scenes.hex_lattice(4000, 2.2, 0.025, seed=3, vel_scale=3.0)
    scenes.add_lattice_bonds(sc)
    ax, ay, bx, by = sc.domain
    sc.portals.append(((ax + 0.5, ay + 0.0), (ax + 0.5, ay + 0.8), (ax + 1.6, ay + 0.1), (ax + 1.8, ay + 0.9)))
    sc.portals.append(((ax + 0.9, ay + 0.5), (ax + 1.3, ay + 0.6), (ax + 0.9, ay + 1.1), (ax + 1.3, ay + 1.0)))

    def setup(p):
        p.set_point_force(True, False, (ax + 1.2, ay + 0.4))
        p.set_pick(True, 77, (ax + 1.0, ay + 1.0))

    o = best_oracle()
    scenes.load_into(o, sc)
    setup(o)
    g = gpu(default_scene=False)
    teleports = 0
    for k in range(80):
        before = load_state_from_oracle(g, o, sc.gravity, setup)
        g.step_n(1)
        o.step_n(1)
        cap = o.id_capacity
        sg, so = g.state_by_id(cap), o.state_by_id(cap)
        assert np.array_equal(sg["block"], so["block"]), f"step {k}"
        alive = so["block"] >= 0
        assert np.abs(sg["pos"][alive] - so["pos"][alive]).max() < 2e-6, f"step {k}"
        assert np.abs(sg["vel"][alive] - so["vel"][alive]).max() < 5e-3, f"step {k}"
        teleports += int((np.abs(so["pos"][alive] - before["pos"][alive]).max(axis=1) > 0.2).sum())
    assert teleports > 0, "scene never exercised a teleport"


def test_attractive_point_force_and_resize_follow():
    sc = scenes.hex_lattice(3000, 2.2, 0.025, seed=8, vel_scale=1.0)
    ax, ay, bx, by = sc.domain
    o = best_oracle()
    scenes.load_into(o, sc)
    g = gpu(default_scene=False)
    scenes.load_into(g, sc)
    g.set_renderer_ready(False)
    for p in (g, o):
        p.set_point_force(True, True, (ax + 1.0, ay + 0.5))
        p.push_resize((ax - 0.35, ay - 0.2, bx + 0.4, by + 0.3))
    for k in range(12):
        g.step_n(19)
        o.step_n(19)
        gg, go = g.geometry(), o.geometry()
        assert bits_equal(gg["domain"], go["domain"]) and bits_equal(gg["grid"], go["grid"]) and gg["dims"] == go["dims"]
    assert g.num_particles == o.num_particles
    sg, so = g.state_by_id(o.id_capacity), o.state_by_id()
    # 228 chaotic steps: only the bulk statistics are comparable
    assert np.abs(sg["pos"].mean(0) - so["pos"].mean(0)).max() < 2e-2
    # the grid did grow on both sides
    assert go["dims"][0] > sc_dims(sc)[0]


def sc_dims(sc):
    ax, ay, bx, by = sc.domain
    return int((bx - ax) / 0.08) + 1, int((by - ay) / 0.08) + 1


def test_particles_leaving_the_grid_are_removed_with_their_bonds():
    """blocks.h:230-232 (silent removal) + CheckBonds (particles.cpp:205-215)."""
    g, o = gpu(default_scene=False), best_oracle()
    # 0 and 5 start outside the right wall (inside the grid's slack) and fly out together, bonded
    # to each other; 3 leaves on the left; 1-2 stay, bonded
    pos = np.array([[1.05, 0.0], [0.0, 0.0], [0.039, 0.0], [-1.06, -0.5], [0.5, 0.5], [1.05, 0.0405]], np.float32)
    vel = np.array([[9.0, 0.0], [0, 0], [0, 0], [-9.0, 0.0], [0, 0], [9.0, 0.0]], np.float32)
    bonds = np.array([[0, 5], [5, 0], [1, 2], [2, 1]], np.int32)
    for p in (g, o):
        p.reset_scene((-1, -1, 1, 1))
        p.set_gravity(False)
        p.add_particles(pos, vel, np.arange(6, dtype=np.int32))
        p.set_bonds(bonds)
    g.set_renderer_ready(False)
    for k in range(30):
        g.step_n(1)
        o.step_n(1)
        assert g.num_particles == o.num_particles, k
        assert np.array_equal(g.get_bonds(), o.get_bonds()), k
        sg, so = g.state_by_id(6), o.state_by_id(6)
        assert np.array_equal(sg["block"], so["block"])
    assert o.num_particles == 3 and len(o.get_bonds()) == 2


def test_find_block_bit_exact():
    kat = load_golden("kat")
    g = gpu()
    assert np.array_equal(g.find_blocks(kat["find_pts"]), kat["find_block"])
    o = best_oracle()
    rng = np.random.RandomState(0)
    pts = rng.uniform(-1.2, 1.2, size=(20000, 2)).astype(np.float32)
    # points exactly on block boundaries and one ulp either side
    edges = (np.arange(27) * np.float32(0.08) - 1).astype(np.float32)
    e = np.stack([np.repeat(edges, 3), np.tile(edges[:3], 27)], 1)
    pts = np.concatenate([pts, e, np.nextafter(e, np.float32(9)), np.nextafter(e, np.float32(-9))]).astype(np.float32)
    want = np.array([o.find_block(x, y) for x, y in pts])
    assert np.array_equal(g.find_blocks(pts), want)


def test_frozen_particles_do_not_move_but_still_push():
    sc = scenes.hex_lattice(2000, 2.02, 0.01, seed=4)
    sc.frozen = np.arange(0, 2000, 7, dtype=np.int32)
    g, o = gpu(default_scene=False), best_oracle()
    scenes.load_into(g, sc)
    scenes.load_into(o, sc)
    g.set_renderer_ready(False)
    g.step_n(50)
    o.step_n(50)
    sg, so = g.state_by_id(2000), o.state_by_id(2000)
    assert bits_equal(sg["pos"][sc.frozen], sc.pos[sc.frozen])
    assert np.all(sg["vel"][sc.frozen] == 0)
    assert np.abs(sg["pos"] - so["pos"]).max() < 1e-5


def test_snapshot_and_getters():
    """SetRendererReadyForNext → step → GetParticles / GetBlockById / GetNumParticles
    (particles.cpp:190-196, particles.h:164-178)."""
    g, o = ptoy_b200.Particles(device=0), best_oracle()
    for p in (g, o):
        p.set_renderer_ready(True)
    g.step(g.GetTime() + np.float32(0.03))       # what ptoy_wasm.cpp:122-124 does per frame
    o.step_to(o.time + np.float32(0.03))
    assert float(g.GetTime()) == float(o.time)
    buf_g, buf_o = g.GetParticles(), o.particle_buffer()
    assert buf_g.shape == buf_o.shape == (625, 4)
    # the snapshot is taken on the FIRST sub-step of the call
    twin = best_oracle()
    twin.step_n(1)
    st = twin.state_by_id(625)
    snap = g.snapshot()
    assert np.abs(snap["pos"] - st["pos"][snap["id"]]).max() < 1e-6
    assert np.all(np.diff(snap["block"]) >= 0)
    assert g.GetNumParticles() == 625
    bbi = g.GetBlockById()
    assert np.array_equal(bbi[:, 0], st["block"])
    ids, start = g.block_lists()
    for b in np.unique(snap["block"]):
        assert np.array_equal(np.sort(ids[start[b]:start[b + 1]]), np.sort(snap["id"][snap["block"] == b]))


def test_ui_tools_match_the_reference():
    g, o = ptoy_b200.Particles(device=0), best_oracle()
    for p in (g, o):
        p.set_renderer_ready(False)
        p.step_n(5)
        p.set_particle_buffer()
    s = o.state_by_id()
    for p in (g, o):
        p.tool("freeze", "start", s["pos"][100]); p.tool("freeze", "move", s["pos"][200]); p.tool("freeze", "stop", (0, 0))
        p.tool("bonds", "start", s["pos"][300]); p.tool("bonds", "move", s["pos"][301]); p.tool("bonds", "move", s["pos"][302])
        p.tool("bonds", "stop", (0, 0))
        p.tool("pick", "start", s["pos"][400] + np.float32(0.001)); p.tool("pick", "move", s["pos"][400] + np.float32(0.01))
        p.tool("portal", "start", (0.0, 0.2)); p.tool("portal", "move", (0.0, 0.5)); p.tool("portal", "stop", (0.0, 0.7))
        p.tool("portal", "start", (0.3, 0.2)); p.tool("portal", "stop", (0.4, 0.9))
    assert np.array_equal(g.get_frozen(), o.get_frozen()) and len(o.get_frozen()) == 2
    assert np.array_equal(g.get_bonds(), o.get_bonds()) and len(o.get_bonds()) == 4
    assert bits_equal(g.get_portals(), o.get_portals()) and len(o.get_portals()) == 2
    for d in (0, 1):
        assert np.array_equal(g.get_portal_blocks(1, d), o.get_portal_blocks(1, d))
    g.step_n(20)
    o.step_n(20)
    sg, so = g.state_by_id(625), o.state_by_id(625)
    assert np.abs(sg["pos"] - so["pos"]).max() < 1e-5
    for p in (g, o):
        p.remove_last_portal()
        p.step_n(2)
    assert len(g.get_portals()) == len(o.get_portals()) == 1


def test_determinism_and_incremental_add():
    """Two runs are bit-identical (stable reorder: no dependence on atomic ordering), and adding
    particles in two batches equals adding them at once."""
    sc = scenes.uniform_random(50000, seed=9)
    outs = []
    for split in (None, None, 20000):
        g = gpu(default_scene=False)
        g.reset_scene(sc.domain)
        if split is None:
            g.add_particles(sc.pos, sc.vel, sc.ids)
        else:
            g.add_particles(sc.pos[:split], sc.vel[:split], sc.ids[:split])
            g.add_particles(sc.pos[split:], sc.vel[split:], sc.ids[split:])
        g.step_n(5)
        s = g.state_by_id(sc.n)
        g.set_particle_buffer()
        outs.append((s["pos"].copy(), s["vel"].copy(), g.snapshot()["id"].copy()))
    for a, b in ((0, 1), (0, 2)):
        assert bits_equal(outs[a][0], outs[b][0]) and bits_equal(outs[a][1], outs[b][1])
        assert np.array_equal(outs[a][2], outs[b][2])


def test_pipelined_upload_step_download_equals_the_synchronous_calls():
    """ptoy_upload_state_async / ptoy_download_state_async on several handles in flight give the
    same bits as ptoy_upload_state / ptoy_step_n / ptoy_download_state on one handle."""
    import torch
    sc = golden_scenes()["bonded3k"][0]

    def pinned(n):
        return (torch.empty((n, 2), dtype=torch.float32).pin_memory().numpy(),
                torch.empty((n, 2), dtype=torch.float32).pin_memory().numpy(),
                torch.empty((n,), dtype=torch.int32).pin_memory().numpy())

    # three different batches: the scene advanced by 0, 3 and 6 iterations
    src = gpu(default_scene=False)
    scenes.load_into(src, sc)
    batches = []
    for _ in range(3):
        b = pinned(sc.n)
        assert src.download_state(*b) == sc.n
        batches.append(b)
        src.step_n(3)

    ref = []
    g = gpu(default_scene=False)
    scenes.load_into(g, sc)
    for b in batches:
        out = pinned(sc.n)
        for _ in range(2):                                # two round trips per batch
            g.upload_state(*(b if _ == 0 else out))
            g.step_n(1)
            assert g.download_state(*out) == sc.n
        ref.append(out)

    engines = []
    for b in batches:
        e = gpu(default_scene=False)
        scenes.load_into(e, sc)
        engines.append(e)
    work = [tuple(a.copy() for a in b) for b in batches]
    work = [tuple(torch.from_numpy(a).pin_memory().numpy() for a in b) for b in work]
    for _ in range(2):
        for e, b in zip(engines, work):
            e.synchronize()
            e.upload_state_async(b[0], b[1], b[2], sc.n - 1)
            e.step_n(1)
            e.download_state_async(*b)
    for e, b, r in zip(engines, work, ref):
        e.synchronize()
        assert e.downloaded_count == sc.n and e.error_flags == 0
        assert np.array_equal(b[2], r[2])
        assert bits_equal(b[0], r[0]) and bits_equal(b[1], r[1])


# --------------------------------------------------------------------------- full size
def _numpy_find_block(pos, grid, dims):
    """blocks::FindBlock (blocks.h:155-163) in float32 numpy."""
    bs = np.float32(4.0 * float(np.float32(0.02)))
    tx = (pos[:, 0] - np.float32(grid[0])) / bs
    ty = (pos[:, 1] - np.float32(grid[1])) / bs
    mi, mj = np.trunc(tx).astype(np.int64), np.trunc(ty).astype(np.int64)
    ok = (mi >= 0) & (mi < dims[0]) & (mj >= 0) & (mj < dims[1])
    return np.where(ok, mi * dims[1] + mj, -1)


@pytest.mark.parametrize("n", [16_000_000])
def test_full_size_properties_and_window_parity(n):
    """BASELINE.json's 16M configuration (bonds ~4/particle, 1 % frozen): size-independent
    properties on the whole state + value parity against the oracle on a sub-window."""
    sc = scenes.config3(n)
    g = gpu(default_scene=False)
    scenes.load_into(g, sc)
    g.set_renderer_ready(False)
    geo = g.geometry()
    # (1) bit-exact cell assignment of every particle
    s = g.state_by_id(n)
    assert np.array_equal(s["block"], _numpy_find_block(sc.pos, geo["grid"], geo["dims"]))
    assert g.num_particles == n
    # (2) frozen particles feel nothing but still push: everyone else's force is bit-identical
    # with and without the frozen set; and with nothing frozen Newton's third law holds: pair
    # and bond forces cancel pairwise (each contribution is the exact negative of its twin), so
    # the total is gravity alone (no particle starts within wall range)
    f = g.eval_forces(1, n)
    frozen = np.zeros(n, bool)
    frozen[sc.frozen] = True
    assert np.all(f[frozen] == 0)
    g.set_frozen(np.zeros(0, np.int32))
    f_free = g.eval_forces(1, n)
    assert bits_equal(f[~frozen], f_free[~frozen])
    g.set_frozen(sc.frozen)
    total = f_free.astype(np.float64).sum(0)
    mag = np.abs(f_free.astype(np.float64)).sum()
    assert abs(total[0]) < 1e-6 * mag and abs(total[1] + 0.4 * n) < 1e-6 * mag + 1e-3 * 0.4 * n
    f = f.astype(np.float64)
    # (3) window parity: interior forces depend only on particles within the bond reach
    rows, cols = sc.lattice
    c0 = sc.pos[(rows // 2) * cols + cols // 2]
    half, halo = 1.0, 0.4
    d = np.abs(sc.pos - c0)
    sub = np.nonzero((d[:, 0] < half + halo) & (d[:, 1] < half + halo))[0]
    inner = (np.abs(sc.pos[sub] - c0) < half).all(axis=1)
    remap = -np.ones(n, np.int64)
    remap[sub] = np.arange(len(sub))
    b = sc.bonds[(remap[sc.bonds[:, 0]] >= 0) & (remap[sc.bonds[:, 1]] >= 0)]
    o = best_oracle()
    o.reset_scene(sc.domain)
    o.set_gravity(*sc.gravity)
    o.add_particles(sc.pos[sub], sc.vel[sub], np.arange(len(sub), dtype=np.int32))
    o.set_bonds(remap[b].astype(np.int32))
    fz = remap[sc.frozen]
    o.set_frozen(fz[fz >= 0].astype(np.int32))
    o.eval_forces()
    fo = o.state_by_id(len(sub))["force"]
    scale = force_scale(sc.pos[sub], fo, remap[b])
    err = np.abs(f[sub][inner] - fo[inner]).max(axis=1) / scale[inner]
    assert inner.sum() > 1000 and err.max() <= FORCE_RTOL
    # (4) after stepping: ids are still a permutation, snapshot is block-sorted, keys consistent
    g.step_n(3)
    assert g.num_particles == n
    g.set_particle_buffer()
    snap = g.snapshot()
    assert np.all(np.diff(snap["block"]) >= 0)
    assert np.array_equal(np.sort(snap["id"]), np.arange(n, dtype=np.int32))
    assert np.array_equal(snap["block"], _numpy_find_block(snap["pos"], geo["grid"], geo["dims"]))
    # (5) the window again, 3 steps later, against the oracle stepping the window alone
    # (interior of the interior: information travels < 3 * 0.2 per step through bonds)
    o2 = best_oracle()
    o2.reset_scene(sc.domain)
    o2.set_gravity(*sc.gravity)
    o2.add_particles(sc.pos[sub], sc.vel[sub], np.arange(len(sub), dtype=np.int32))
    o2.set_bonds(remap[b].astype(np.int32))
    o2.set_frozen(fz[fz >= 0].astype(np.int32))
    o2.step_n(3)
    so = o2.state_by_id(len(sub))
    core = (np.abs(sc.pos[sub] - c0) < half - 0.3).all(axis=1)
    sg = g.state_by_id(n)
    assert np.abs(sg["pos"][sub][core] - so["pos"][core]).max() < 1e-6


# --------------------------------------------------------------------------- edge cases
def test_empty_scene_and_single_particle():
    """Empty input, one particle, adding nothing: the reference steps them without complaint."""
    g, o = gpu(default_scene=False), best_oracle()
    for p in (g, o):
        p.reset_scene((-1, -1, 1, 1))
    g.set_renderer_ready(False)
    g.step_n(3)
    o.step_n(3)
    assert g.num_particles == o.num_particles == 0 and float(g.time) == float(o.time)
    g.add_particles(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), np.zeros(0, np.int32))
    assert g.num_particles == 0
    one_p, one_v = np.array([[0.3, -0.2]], np.float32), np.array([[1.0, 2.0]], np.float32)
    for p in (g, o):
        p.add_particles(one_p, one_v, np.array([7], np.int32))   # ids need not start at 0
    g.step_n(50)
    o.step_n(50)
    sg, so = g.state_by_id(8), o.state_by_id(8)
    # (ids 0..6 were never added: the reference's id map value-initialises them to block 0, slot 0)
    assert bits_equal(sg["pos"][7], so["pos"][7]) and bits_equal(sg["vel"][7], so["vel"][7])   # no pair sums: bit-exact
    assert sg["block"][7] == so["block"][7] and np.all(sg["block"][:7] == -1)


def test_particles_outside_the_grid_are_skipped_on_insert():
    """blocks::AddParticles (blocks.h:186-191) ignores positions that FindBlock rejects, keeps the
    band below the low edges (truncation toward zero) and NaNs never enter."""
    pos = np.array([[0.0, 0.0], [5.0, 0.0], [-1.05, -1.0], [-1.09, 0.0], [np.nan, 0.0], [0.5, 1.079], [0.5, 1.081]],
                   np.float32)
    g, o = gpu(default_scene=False), best_oracle()
    for p in (g, o):
        p.reset_scene((-1, -1, 1, 1))
        p.set_gravity(False)
        p.add_particles(pos, np.zeros_like(pos), np.arange(len(pos), dtype=np.int32))
    sg, so = g.state_by_id(len(pos)), o.state_by_id(len(pos))
    inside = np.array([0, 2, 5])
    assert np.array_equal(np.nonzero(sg["block"] >= 0)[0], inside)
    assert np.array_equal(sg["block"][inside], so["block"][inside])
    # (the reference's id map value-initialises the skipped ids to block 0, slot 0; the count is right)
    assert g.num_particles == o.num_particles == 3


def test_dense_cell_and_coincident_particles():
    """Many particles in one sub-cell (hit queue flushes, multi-entry stable reorder) and exactly
    coincident particles (dp = 0: zero force in the reference)."""
    rng = np.random.RandomState(2)
    cluster = (np.array([0.101, 0.203]) + rng.uniform(0, 0.03, size=(40, 2))).astype(np.float32)
    twins = np.array([[0.5, 0.5], [0.5, 0.5], [-0.5, 0.5], [-0.5, 0.5]], np.float32)
    pos = np.concatenate([cluster, twins])
    g, o = gpu(default_scene=False), best_oracle()
    for p in (g, o):
        p.reset_scene((-1, -1, 1, 1))
        p.add_particles(pos, np.zeros_like(pos), np.arange(len(pos), dtype=np.int32))
    g.set_renderer_ready(False)
    n = len(pos)
    f = g.eval_forces(1, n)
    o2 = best_oracle()
    o2.reset_scene((-1, -1, 1, 1))
    o2.add_particles(pos, np.zeros_like(pos), np.arange(n, dtype=np.int32))
    o2.eval_forces()
    fo = o2.state_by_id(n)["force"]
    scale = force_scale(pos, fo)
    assert (np.abs(f - fo).max(axis=1) / scale).max() <= FORCE_RTOL
    assert bits_equal(f[40:], fo[40:])          # coincident pairs feel gravity only, bit for bit
    g.step_n(1)
    o.step_n(1)
    sg, so = g.state_by_id(n), o.state_by_id(n)
    assert np.array_equal(sg["block"], so["block"])
    assert np.abs(sg["vel"] - so["vel"]).max() <= 1e-4 * np.abs(so["vel"]).max()


def test_many_bonds_on_one_particle():
    """A hub with 12 bonds: more than the six the kernels resolve up front (CSR tail path)."""
    hub = np.array([[0.0, 0.0]], np.float32)
    th = np.linspace(0, 2 * np.pi, 12, endpoint=False)
    ring = np.stack([0.05 * np.cos(th), 0.05 * np.sin(th)], 1).astype(np.float32)
    pos = np.concatenate([hub, ring])
    bonds = np.array([[0, k] for k in range(1, 13)] + [[k, 0] for k in range(1, 13)], np.int32)
    g, o = gpu(default_scene=False), best_oracle()
    for p in (g, o):
        p.reset_scene((-1, -1, 1, 1))
        p.set_gravity(False)
        p.add_particles(pos, np.zeros_like(pos), np.arange(13, dtype=np.int32))
        p.set_bonds(bonds)
    g.set_renderer_ready(False)
    twin = best_oracle()
    twin.reset_scene((-1, -1, 1, 1))
    twin.set_gravity(False)
    twin.add_particles(pos, np.zeros_like(pos), np.arange(13, dtype=np.int32))
    twin.set_bonds(bonds)
    twin.eval_forces()
    fo = twin.state_by_id(13)["force"]
    f = g.eval_forces(1, 13)
    assert (np.abs(f - fo).max(axis=1) / force_scale(pos, fo, bonds, 0.0)).max() <= FORCE_RTOL
    # a stiff, symmetric, overlapping configuration: compare a short horizon
    g.step_n(3)
    o.step_n(3)
    sg, so = g.state_by_id(13), o.state_by_id(13)
    assert np.abs(sg["pos"] - so["pos"]).max() < 1e-6
    assert np.array_equal(g.get_bonds(), o.get_bonds())

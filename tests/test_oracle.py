"""CPU tests of the oracle itself: the plain-C restatement and (when present) the compiled
reference against the golden fixtures generated from the compiled reference, plus known-answer
tests of the force laws and of FindBlock."""
import numpy as np
import pytest

from conftest import bits_equal, load_golden, make_oracle
from make_golden import golden_scenes
from ptoy_b200 import scenes

KINDS = ["port", "reference"]


def _prepare(kind, name):
    sc, setup, checkpoints = golden_scenes()[name]
    o = make_oracle(kind)
    if sc is not None:
        scenes.load_into(o, sc)
    if setup:
        setup(o)
    return o, checkpoints


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", ["default", "bonded3k", "portal2k", "uniform4k"])
def test_oracle_matches_golden(kind, name):
    gold = load_golden(name)
    o, checkpoints = _prepare(kind, name)
    g = o.geometry()
    assert bits_equal(g["domain"], gold["domain"]) and bits_equal(g["grid"], gold["grid"])
    assert tuple(gold["dims"]) == g["dims"]
    assert bits_equal(o.get_portals(), gold["portals"])
    for p in range(len(gold["portals"])):
        for d in (0, 1):
            assert np.array_equal(o.get_portal_blocks(p, d), gold[f"portal_blocks_{p}_{d}"])
    twin, _ = _prepare(kind, name)
    twin.eval_forces()
    assert bits_equal(twin.state_by_id(o.id_capacity)["force"], gold["force1_at_0"])
    done = 0
    for cp in checkpoints:
        if kind == "port" and cp > 100 and name != "default":
            break
        o.step_n(cp - done)
        done = cp
        s = o.state_by_id(o.id_capacity)
        ids, start = o.block_lists()
        pre = f"s{cp}_"
        assert bits_equal(s["pos"], gold[pre + "pos"]), (name, cp)
        assert bits_equal(s["vel"], gold[pre + "vel"]), (name, cp)
        assert bits_equal(s["force"], gold[pre + "force2"]), (name, cp)
        assert np.array_equal(s["block"], gold[pre + "block"])
        # the reference's own history-dependent in-block order
        assert np.array_equal(ids, gold[pre + "block_ids"]) and np.array_equal(start, gold[pre + "block_start"])
        assert np.array_equal(o.get_bonds(), gold[pre + "bonds"])
        assert o.time == gold[pre + "time"]
        assert o.num_particles == gold[pre + "num_particles"] and o.num_per_cell == gold[pre + "num_per_cell"]


@pytest.mark.parametrize("kind", KINDS)
def test_default_scene_anchors(kind):
    """SURVEY.md §8c anchors of the constructor scene."""
    o = make_oracle(kind)
    assert o.num_particles == 625 and o.num_blocks == 676
    assert o.geometry()["dims"] == (26, 26)
    assert list(o.neighbor_offsets()) == [-27, -26, -25, -1, 0, 1, 25, 26, 27]
    assert len(o.get_portal_blocks(0, 0)) == 28 and len(o.get_portal_blocks(0, 1)) == 28
    p = o.get_portals()
    assert p.shape == (1, 2, 4) and abs(p[0, 0, 0] + 0.52) < 1e-6 and abs(p[0, 0, 3] - 0.0660254) < 1e-6
    assert np.float32(o.dt).view(np.uint32) == 0x3A03126F            # Appendix A.1
    o.step_n(1000)
    assert abs(float(o.time) - 0.500009) < 1e-6 and o.num_particles == 625 and o.num_per_cell == 8
    s = o.state_by_id()
    assert np.allclose(s["pos"][0], [-0.4687919, -0.9801815], atol=1e-7)


@pytest.mark.parametrize("kind", KINDS)
def test_force_law_kats(kind):
    kat = load_golden("kat")
    o = make_oracle(kind)
    for law in ("F12", "F12wall", "F12_bond", "F12_portal_edge", "F12_pick"):
        got = np.array([o.law(law, p, q) for p, q in zip(kat["P"], kat["Q"])], np.float32)
        assert bits_equal(got, kat[law]), law
    # pair force: exactly zero at and beyond R = 2r, repulsive inside
    f = np.array([o.law("F12", (0, 0), (d, 0)) for d in (0.04, 0.0400001, 0.05, 1.0)])
    assert np.all(f == 0)
    assert o.law("F12", (0, 0), (0.03, 0))[0] < 0
    # bond: clamp at |dp| >= 5R (k = -4)
    # dp = p - q = (-0.25, 0); force = dp * 1e5 * (-4): attractive, towards q
    assert np.allclose(o.law("F12_bond", (0, 0), (0.25, 0)), [1e5, 0])
    assert np.allclose(o.law("F12_bond", (0, 0), (0.5, 0)), [2e5, 0])


@pytest.mark.parametrize("kind", KINDS)
def test_find_block_kats(kind):
    kat = load_golden("kat")
    o = make_oracle(kind)
    got = np.array([o.find_block(x, y) for x, y in kat["find_pts"]])
    assert np.array_equal(got, kat["find_block"])
    # truncation toward zero
    assert o.find_block(-1.05, -1.0) == 0 and o.find_block(-1.09, 0.0) == -1
    for row in kat["move_to_portal"]:
        p, v = o.move_to_portal(0, int(row[4]), row[0:2], row[2:4])
        assert bits_equal(np.r_[p, v], row[5:9])


def test_port_equals_reference_with_tools_and_resize():
    """Bit-for-bit, including the reference's in-block order, through resize, portal removal,
    UI tools, pick and point forces."""
    a, b = make_oracle("reference"), make_oracle("port")
    sc = scenes.hex_lattice(3000, 2.2, 0.025, seed=3, vel_scale=3.0)
    scenes.add_lattice_bonds(sc)
    ax, ay, bx, by = sc.domain
    sc.portals.append(((ax + 0.5, ay + 0.0), (ax + 0.5, ay + 0.8), (ax + 1.5, ay + 0.1), (ax + 1.7, ay + 0.9)))
    sc.portals.append(((ax + 0.8, ay + 0.5), (ax + 1.2, ay + 0.6), (ax + 0.8, ay + 1.1), (ax + 1.2, ay + 1.0)))

    def same(tag):
        sa, sb = a.state_by_id(), b.state_by_id()
        for k in sa:
            assert bits_equal(sa[k], sb[k]), (tag, k)
        (ia, ta), (ib, tb) = a.block_lists(), b.block_lists()
        assert np.array_equal(ia, ib) and np.array_equal(ta, tb), tag
        assert np.array_equal(a.get_bonds(), b.get_bonds()) and np.array_equal(a.get_frozen(), b.get_frozen())
        assert np.array_equal(a.get_no_rendering(), b.get_no_rendering())
        assert a.time == b.time and a.num_particles == b.num_particles and a.num_per_cell == b.num_per_cell
        assert bits_equal(a.get_portals(), b.get_portals())
        ga, gb = a.geometry(), b.geometry()
        assert bits_equal(ga["grid"], gb["grid"]) and ga["dims"] == gb["dims"]

    for o in (a, b):
        scenes.load_into(o, sc)
        o.set_point_force(True, False, (ax + 1.2, ay + 0.4))
        o.set_pick(True, 77, (ax + 1.0, ay + 1.0))
    same("t0")
    for o in (a, b):
        o.step_n(40, snapshot=True)
    same("+40")
    for o in (a, b):
        o.set_point_force(True, True, (ax + 1.5, ay + 0.4))
        o.push_resize((ax, ay, bx + 0.5, by - 0.3))
        o.step_n(80)
    same("shrink/attract")
    for o in (a, b):
        o.push_resize((ax - 0.3, ay - 0.2, bx + 0.5, by + 0.3))
        o.remove_last_portal()
        o.step_n(80, snapshot=True)
    same("grow")
    for o in (a, b):
        o.set_particle_buffer()
        s = o.state_by_id()
        o.tool("freeze", "start", s["pos"][100]); o.tool("freeze", "move", s["pos"][200]); o.tool("freeze", "stop", (0, 0))
        o.tool("bonds", "start", s["pos"][300]); o.tool("bonds", "move", s["pos"][301]); o.tool("bonds", "move", s["pos"][302])
        o.tool("bonds", "stop", (0, 0))
        o.tool("pick", "start", s["pos"][400] + 0.001); o.tool("pick", "move", s["pos"][400] + 0.01)
        o.tool("portal", "start", (ax + 2.0, ay + 0.2)); o.tool("portal", "stop", (ax + 2.0, ay + 0.7))
        o.tool("portal", "start", (ax + 2.4, ay + 0.2)); o.tool("portal", "stop", (ax + 2.5, ay + 0.9))
        o.step_n(30)
    same("tools")


def test_particles_leaving_the_grid_are_removed_with_their_bonds():
    """blocks.h:230-232 + CheckBonds (particles.cpp:205-215)."""
    for kind in KINDS:
        o = make_oracle(kind)
        o.reset_scene((-1, -1, 1, 1))
        o.set_gravity(False)
        pos = np.array([[0.9, 0.0], [0.0, 0.0], [0.05, 0.0]], np.float32)
        vel = np.array([[9.0, 0.0], [0, 0], [0, 0]], np.float32)
        o.add_particles(pos, vel, np.arange(3, dtype=np.int32))
        # remove the right wall so particle 0 can leave
        o.set_bonds(np.array([[0, 1], [1, 0], [1, 2], [2, 1]], np.int32))
        o.push_resize((-1, -1, 1, 1))
        n0 = o.num_particles
        assert n0 == 3
        o.step_n(5)
        assert o.num_particles == 3


@pytest.mark.parametrize("seed", list(range(11, 27)))
def test_port_equals_reference_on_random_feature_mixes(seed):
    """Seeded fuzz: random small scenes (lattice or overlapping gas), random bonds / frozen set /
    portal pair / point force / resize, 60 iterations; the restatement must stay bit-identical to
    the compiled reference, including the in-block order."""
    rng = np.random.RandomState(seed)
    a, b = make_oracle("reference"), make_oracle("port")
    n = int(rng.randint(300, 1500))
    if rng.rand() < 0.5:
        sc = scenes.hex_lattice(n, 2.0 + 0.3 * rng.rand(), 0.03, seed=seed, vel_scale=4.0 * rng.rand())
        if rng.rand() < 0.7:
            scenes.add_lattice_bonds(sc, seed=seed)
    else:
        sc = scenes.uniform_random(n, seed=seed)
    ax, ay, bx, by = sc.domain
    w, h = bx - ax, by - ay
    if rng.rand() < 0.6 and w > 1.0 and h > 0.6:
        x0, x1 = ax + 0.25 * w, ax + 0.7 * w
        sc.portals.append(((x0, ay + 0.05 * h), (x0 + 0.1 * rng.rand(), ay + 0.6 * h),
                           (x1, ay + 0.1 * h), (x1 + 0.1 * rng.rand(), ay + 0.65 * h)))
    if rng.rand() < 0.6:
        sc.frozen = np.sort(rng.choice(sc.n, size=max(1, sc.n // 50), replace=False)).astype(np.int32)
    for o in (a, b):
        scenes.load_into(o, sc)
    if rng.rand() < 0.5:
        c = (ax + w * rng.rand(), ay + h * rng.rand())
        attract = bool(rng.rand() < 0.5)
        for o in (a, b):
            o.set_point_force(True, attract, c)
    if rng.rand() < 0.5:
        dom = (ax - 0.2 * rng.rand(), ay - 0.2 * rng.rand(), bx + 0.3 * (rng.rand() - 0.5), by + 0.3 * (rng.rand() - 0.5))
        for o in (a, b):
            o.push_resize(dom)
    for chunk in (1, 19, 40):
        for o in (a, b):
            o.step_n(chunk, snapshot=bool(chunk == 19))
        sa, sb = a.state_by_id(), b.state_by_id()
        for k in sa:
            assert bits_equal(sa[k], sb[k]), (seed, chunk, k)
        (ia, ta), (ib, tb) = a.block_lists(), b.block_lists()
        assert np.array_equal(ia, ib) and np.array_equal(ta, tb)
        assert np.array_equal(a.get_bonds(), b.get_bonds())
        assert a.time == b.time and a.num_particles == b.num_particles and a.num_per_cell == b.num_per_cell
        ga, gb = a.geometry(), b.geometry()
        assert bits_equal(ga["grid"], gb["grid"]) and ga["dims"] == gb["dims"]

"""Strip decomposition: N virtual strips on one device (in-process transport in
place of NCCL) must reproduce the single engine BIT FOR BIT — halo exchange before both force
evaluations, migration across strip boundaries, deterministic merge of the migrants."""
import numpy as np
import pytest

from conftest import bits_equal
import ptoy_b200
from ptoy_b200 import scenes

pytestmark = pytest.mark.gpu


def single(sc, steps, setup=None):
    g = ptoy_b200.Particles(device=0, default_scene=False)
    g.reset_scene(sc.domain)
    g.set_gravity(*sc.gravity)
    g.add_particles(sc.pos, sc.vel, sc.ids)
    if setup:
        setup(g)
    g.set_renderer_ready(False)
    g.step_n(steps)
    return g.state_by_id(sc.n), g.num_particles


def strips(sc, steps, world, col_begin=None, setup=None, **kw):
    grp = ptoy_b200.StripGroup(sc.domain, world, col_begin=col_begin, gravity=sc.gravity, **kw)
    grp.add_particles(sc.pos, sc.vel, sc.ids)
    if setup:
        grp.each(setup)
    assert grp.num_particles == sc.n          # every particle is owned exactly once
    grp.step_n(steps)
    out = grp.state_by_id(sc.n), grp.num_particles
    flags = [p.error_flags for p in grp.parts]
    grp.close()
    assert not any(flags), flags
    return out


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_strips_equal_single_engine_moving_particles(world):
    # velocities up to 4: ~1 % of the particles change sub-row per step, so 60 steps push
    # thousands across the strip boundaries in both directions
    sc = scenes.hex_lattice(30000, 2.2, 0.025, seed=17, vel_scale=4.0)
    ref, n_ref = single(sc, 60)
    got, n_got = strips(sc, 60, world)
    assert n_got == n_ref == sc.n
    assert bits_equal(got["pos"], ref["pos"]) and bits_equal(got["vel"], ref["vel"])
    assert np.array_equal(got["block"], ref["block"])


def test_strips_with_uneven_boundaries_and_overlapping_particles():
    sc = scenes.uniform_random(40000, seed=5)
    g = ptoy_b200.Particles(device=0, default_scene=False)
    g.reset_scene(sc.domain)
    di = g.geometry()["dims"][0]
    cb = np.array([0, 2, di // 3, di - 2, di], np.int32)   # two minimal strips at the edges
    ref, n_ref = single(sc, 8)
    got, n_got = strips(sc, 8, 4, col_begin=cb)
    assert n_got == n_ref
    assert bits_equal(got["pos"], ref["pos"]) and bits_equal(got["vel"], ref["vel"])


def test_strips_with_frozen_point_force_and_removal():
    sc = scenes.hex_lattice(12000, 2.2, 0.025, seed=9, vel_scale=2.0)
    ax, ay, bx, by = sc.domain
    # some particles start outside the right wall and fly out of the grid
    extra = np.array([[bx + 0.03, ay + 0.5], [bx + 0.03, ay + 0.9]], np.float32)
    sc.pos = np.concatenate([sc.pos, extra])
    sc.vel = np.concatenate([sc.vel, np.array([[9, 0], [9, 0]], np.float32)])
    sc.ids = np.arange(len(sc.pos), dtype=np.int32)
    frozen = np.arange(0, 12000, 11, dtype=np.int32)

    def setup(p):
        p.set_frozen(frozen)
        p.set_point_force(True, True, (ax + 1.0, ay + 0.6))
    ref, n_ref = single(sc, 40, setup)
    got, n_got = strips(sc, 40, 3, setup=setup)
    assert n_got == n_ref == 12000
    assert bits_equal(got["pos"], ref["pos"]) and bits_equal(got["vel"], ref["vel"])


def test_teleports_migrate_to_any_strip():
    """A portal whose partner lives three strips away: teleported particles are routed to the
    far strip by the same outbox mechanism.  (Cross-portal pair forces need both portals on one
    strip — DESIGN.md §7 — so the portals sit in empty space here: the check is the routing.)"""
    sc = scenes.hex_lattice(6000, 2.2, 0.025, seed=3, vel_scale=0.0)
    ax, ay, bx, top = sc.domain
    by = top + 1.0                         # head room above the packing for the beam and the portals
    sc.domain = (ax, ay, bx, by)
    # a beam of fast particles flying into the left portal
    yb = by - 0.72 + 0.05 * np.arange(12)  # 0.05 apart: beyond the pair cutoff
    beam = np.stack([np.full(12, ax + 0.3), yb], 1).astype(np.float32)
    sc.pos = np.concatenate([sc.pos, beam])
    sc.vel = np.concatenate([sc.vel, np.tile(np.array([[8.0, 0.0]], np.float32), (12, 1))])
    sc.ids = np.arange(len(sc.pos), dtype=np.int32)
    sc.gravity = (False, (0.0, -10.0))
    x0, x1 = ax + 0.5, bx - 0.5
    portal = ((x0, by - 0.8), (x0, by - 0.1), (x1, by - 0.8), (x1, by - 0.1))

    def setup(p):
        p.add_portal_pair(*portal)
    ref, _ = single(sc, 70, setup)
    got, _ = strips(sc, 70, 4, setup=setup)
    moved = ref["pos"][6000:, 0] > x1      # the beam came out of the far portal
    assert moved.sum() >= 8
    assert bits_equal(got["pos"], ref["pos"]) and bits_equal(got["vel"], ref["vel"])


@pytest.mark.parametrize("world", [2, 4])
def test_strips_with_bonds_across_the_boundaries(world):
    """configs[2]-style scene (bond network, frozen particles) cut into strips: bond partners in
    the neighbour strip are found through the ghost ids; bonded particles migrate with their
    bonds (every strip holds the bond table by id)."""
    sc = scenes.hex_lattice(24000, 2.02, 0.01, seed=31, vel_scale=1.5)
    scenes.add_lattice_bonds(sc)

    def setup(p):
        p.set_bonds(sc.bonds)
        p.set_frozen(sc.frozen)
    ref, n_ref = single(sc, 40, setup)
    got, n_got = strips(sc, 40, world, setup=setup, bonds=True)
    assert n_got == n_ref == sc.n
    assert bits_equal(got["pos"], ref["pos"]) and bits_equal(got["vel"], ref["vel"])


@pytest.mark.parametrize("world,ghost_rows", [(2, 2), (3, 4)])
def test_bond_rows_travel_with_their_particle(world, ghost_rows):
    """Every strip is given ONLY the bond rows and frozen ids of the particles it holds at the
    start (as a rank of a large decomposed scene would): a particle that moves to another strip
    takes its bond row along, and the result still equals the single engine bit for bit."""
    sc = scenes.hex_lattice(24000, 2.02, 0.01, seed=33, vel_scale=1.0)
    sc.vel[:, 0] += np.float32(4.0)      # the whole sheet drifts in x, across the strip boundaries
    scenes.add_lattice_bonds(sc)

    def setup(p):
        p.set_bonds(sc.bonds)
        p.set_frozen(sc.frozen)
    ref, n_ref = single(sc, 60, setup)

    grp = ptoy_b200.StripGroup(sc.domain, world, gravity=sc.gravity, ghost_rows=ghost_rows, id_capacity=sc.n,
                               bonds=True)
    grp.add_particles(sc.pos, sc.vel, sc.ids)
    moved = 0
    owner0 = np.full(sc.n, -1)
    for r, p in enumerate(grp.parts):
        mine = p.state_by_id(sc.n)["block"] >= 0
        owner0[mine] = r
        p.set_bonds(sc.bonds[mine[sc.bonds[:, 0]]])
        p.set_frozen(sc.frozen)          # (the frozen set is a bit mask by id: every rank takes all of it)
    grp.step_n(60)
    for r, p in enumerate(grp.parts):
        moved += int(((p.state_by_id(sc.n)["block"] >= 0) & (owner0 != r)).sum())
    got, n_got = grp.state_by_id(sc.n), grp.num_particles
    flags = [p.error_flags for p in grp.parts]
    grp.close()
    assert not any(flags), flags
    assert moved > 50, moved             # bonded particles did change strips
    assert n_got == n_ref == sc.n
    assert bits_equal(got["pos"], ref["pos"]) and bits_equal(got["vel"], ref["vel"])


def _portal_scene(n=30000, pairs=6, seed=41, bonds=False):
    """configs[3] in small: dense packing with random velocities and vertical portal pairs INSIDE
    the packing, placed so that some portals straddle strip boundaries and the two portals of a
    pair lie on different strips (cross-portal forces and teleports then need the other strips)."""
    sc = scenes.hex_lattice(n, 2.2 if not bonds else 2.02, 0.025 if not bonds else 0.01, seed=seed, vel_scale=3.0)
    ax, ay, bx, by = sc.domain
    w, h = bx - ax, by - ay
    rng = np.random.RandomState(seed)
    sc.portals = []
    for p in range(pairs):
        x0 = ax + w * (0.08 + 0.38 * rng.uniform())
        x1 = ax + w * (0.55 + 0.38 * rng.uniform())
        y0 = ay + 0.1 + (h - 0.7) * rng.uniform()
        y1 = ay + 0.1 + (h - 0.7) * rng.uniform()
        sc.portals.append(((x0, y0), (x0, y0 + 0.5), (x1, y1), (x1, y1 + 0.5)))
    if bonds:
        scenes.add_lattice_bonds(sc)
    return sc


@pytest.mark.parametrize("world", [2, 4, 8])
def test_portal_pairs_across_strips(world):
    """Cross-portal pair forces (particles.cpp:318-381) and teleports when a portal's partner - or
    part of its own block list - lives on another strip: the portal-zone table gathered over the
    strips reproduces the single engine bit for bit."""
    sc = _portal_scene()

    def setup(p):
        for q in sc.portals:
            p.add_portal_pair(*q)
    ref, n_ref = single(sc, 50, setup)
    got, n_got = strips(sc, 50, world, setup=setup)
    assert n_got == n_ref
    teleported = np.abs(ref["pos"][:, 0] - sc.pos[:, 0]) > 1.0
    assert teleported.sum() > 20, teleported.sum()
    assert bits_equal(got["pos"], ref["pos"]) and bits_equal(got["vel"], ref["vel"])
    assert np.array_equal(got["block"], ref["block"])


def test_bonds_through_portals_across_strips():
    """A bonded sheet with portal pairs: bonds whose partner went through a portal to another
    strip (particles.cpp:784-819) find it in the portal-zone table.  (Particles inside a portal
    slab teleport in the first iteration; the run is short enough for the pairs to stay inside
    the portal zones - a partner that is neither in a ghost row nor in a portal zone is the
    documented limit of the decomposition and raises error bit 16.)"""
    sc = _portal_scene(n=20000, pairs=4, seed=43, bonds=True)
    sc.vel *= np.float32(0.3)

    def setup(p):
        for q in sc.portals:
            p.add_portal_pair(*q)
        p.set_bonds(sc.bonds)
    ref, n_ref = single(sc, 12, setup)
    got, n_got = strips(sc, 12, 4, setup=setup, bonds=True, ghost_rows=6)
    assert n_got == n_ref
    teleported = np.abs(ref["pos"][:, 0] - sc.pos[:, 0]) > 1.0
    assert teleported.sum() > 10, teleported.sum()
    assert bits_equal(got["pos"], ref["pos"]) and bits_equal(got["vel"], ref["vel"])

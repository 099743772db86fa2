"""CPU tests of the synthetic scene generators."""
import numpy as np

from ptoy_b200 import scenes


def test_hex_lattice_is_deterministic_and_inside_the_frame():
    a = scenes.hex_lattice(10000, 2.2, 0.025, seed=7)
    b = scenes.hex_lattice(10000, 2.2, 0.025, seed=7)
    assert np.array_equal(a.pos, b.pos) and a.domain == b.domain
    ax, ay, bx, by = a.domain
    r = 0.02
    assert a.pos[:, 0].min() > ax + r and a.pos[:, 0].max() < bx - r
    assert a.pos[:, 1].min() > ay + r and a.pos[:, 1].max() < by - r
    # min separation: pitch 2.2r minus two jitters
    from scipy.spatial import cKDTree
    d, _ = cKDTree(a.pos).query(a.pos, k=2)
    assert d[:, 1].min() > (2.2 - 4 * 0.025) * r


def test_config3_bond_network():
    sc = scenes.config3(50000)
    b = sc.bonds
    assert b.dtype == np.int32 and b.shape[1] == 2
    # both directions present, sorted like std::set<pair<int,int>>
    s = set(map(tuple, b))
    assert all((q, p) in s for p, q in list(s)[:2000])
    assert np.array_equal(b, b[np.lexsort((b[:, 1], b[:, 0]))])
    deg = len(b) / sc.n
    assert 3.7 < deg < 4.1
    # bonded neighbours are lattice neighbours: near rest length 2r
    d = np.linalg.norm(sc.pos[b[:, 0]] - sc.pos[b[:, 1]], axis=1)
    assert d.max() < 2.1 * 0.02
    assert 0.007 < len(sc.frozen) / sc.n < 0.013


def test_density_matches_the_survey():
    sc = scenes.hex_lattice(100000, 2.2, 0.025)
    ax, ay, bx, by = sc.domain
    dens = sc.n / ((bx - ax) * (by - ay))
    assert 560 < dens < 600          # 596 / unit^2 → ≈3.8 per 0.08 block


def test_hashed_bond_network_is_one_global_network_however_it_is_cut():
    """strip_scene(bonds=True): every rank derives its part of the SAME global bond network and
    frozen set from hashes of the global ids (no rank generates all of it)."""
    n_per_rank, world = 6000, 3
    parts = [scenes.strip_scene(n_per_rank, r, world, pitch_over_r=2.02, jitter_over_r=0.01, bonds=True)
             for r in range(world)]
    rows, cols = parts[0].lattice
    n = rows * cols
    assert all(p.lattice == (rows, cols) and p.n_global == n for p in parts)
    ids = np.concatenate([p.ids for p in parts])
    assert len(ids) == n and len(np.unique(ids)) == n
    pos_by_id = np.zeros((n, 2), np.float32)
    pos_by_id[ids] = np.concatenate([p.pos for p in parts])
    b = scenes.hashed_lattice_bonds(rows, cols, 0, cols, seed=1234 + 99)     # the whole network
    # sorted, symmetric, about 4 directed bonds per particle, at most 6 per particle
    assert np.all(np.lexsort((b[:, 1], b[:, 0])) == np.arange(len(b)))
    fwd = set(map(tuple, b.tolist()))
    assert all((y, x) in fwd for x, y in fwd)
    deg = np.bincount(b[:, 0], minlength=n)
    assert 3.7 < deg.mean() < 4.1 and deg.max() <= 6
    # bonded neighbours sit one lattice pitch apart
    d = np.linalg.norm(pos_by_id[b[:, 0]] - pos_by_id[b[:, 1]], axis=1)
    assert d.max() < 2.02 * 0.02 * 1.05 and d.min() > 2.02 * 0.02 * 0.95
    frozen_all = set()
    for part in parts:
        # every bond row of a particle this rank owns is complete and identical to the global one
        have = set(map(tuple, part.bonds.tolist()))
        assert set(map(tuple, b[np.isin(b[:, 0], part.ids)].tolist())) <= have <= fwd
        frozen_all |= set(part.frozen.tolist())
    assert 0.005 < len(frozen_all) / n < 0.015
    for part in parts:                      # ranks agree on the frozen ids they both know about
        cols_known = set((part.bonds[:, 0] % cols).tolist())
        mine = {f for f in frozen_all if f % cols in cols_known}
        assert mine <= set(part.frozen.tolist())

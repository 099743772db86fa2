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

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU should fail loudly, not skip: nothing to do here.
    return


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle's C restatement is compiled on demand (gcc, a second or two)."""
    from oracle import port
    if not port.available():
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    yield


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def bits_equal(a, b):
    """Bitwise equality of float arrays, treating every NaN as equal."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind != "f":
        return np.array_equal(a, b)
    an, bn = np.isnan(a), np.isnan(b)
    if not np.array_equal(an, bn):
        return False
    return np.array_equal(a[~an].view(np.uint32), b[~bn].view(np.uint32))


def make_oracle(kind):
    from oracle import port, ref
    if kind == "reference":
        if not ref.available():
            pytest.skip("oracle/_ref not built (reference tree absent)")
        return ref.RefParticles()
    return port.PortParticles()


def best_oracle():
    from oracle import port, ref
    return ref.RefParticles() if ref.available() else port.PortParticles()


def force_scale(scene_like_pos, force_ref, bonds=None, gravity_force=0.4):
    """Per-particle normaliser Σ|F_i|: |gravity| + Σ|pair| + Σ|bond|, in
    float64 numpy; only used to turn an absolute force error into a relative one."""
    from scipy.spatial import cKDTree
    pos = np.asarray(scene_like_pos, np.float64)
    n = len(pos)
    s = np.full(n, gravity_force)
    ok = ~np.isnan(pos[:, 0])
    idx = np.nonzero(ok)[0]
    tree = cKDTree(pos[ok])
    pairs = tree.query_pairs(0.04, output_type="ndarray")
    if len(pairs):
        d = pos[idx[pairs[:, 0]]] - pos[idx[pairs[:, 1]]]
        r2 = np.maximum((d * d).sum(1), 4e-7)
        d2 = 0.0016 / r2
        k = np.maximum(0.0, (d2 ** 6 - d2 ** 3) / r2) * np.sqrt(r2)
        np.add.at(s, idx[pairs[:, 0]], k)
        np.add.at(s, idx[pairs[:, 1]], k)
    if bonds is not None and len(bonds):
        d = pos[bonds[:, 0]] - pos[bonds[:, 1]]
        r = np.sqrt((d * d).sum(1))
        k = np.abs(1e5 * np.maximum(1 - r / 0.04, -4.0)) * r
        k = np.nan_to_num(k)
        np.add.at(s, bonds[:, 0], k)
    return np.maximum(s, np.abs(np.nan_to_num(force_ref)).max(axis=1))

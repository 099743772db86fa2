"""Drop-in test: ONE caller program written against the reference's public `Particles` API
(ptoy_b200/host/drop_in_driver.cpp) linked once against the reference's particles.cpp and once
against the facade (ptoy_b200/host/particles_b200.cpp) + libptoy_b200.so.  Both binaries are
built in the build container (they need the reference headers) and travel with the repo."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "drop_in_driver_ref")
B200 = os.path.join(ROOT, "ptoy_b200", "host", "drop_in_driver_b200")


def run(binary, frames):
    out = subprocess.run([binary, str(frames)], capture_output=True, text=True, timeout=300, check=True).stdout
    sections, cur = {}, None
    for line in out.splitlines():
        if line.startswith("@"):
            cur = line.split()[0][1:]
            sections[cur] = {"head": line, "meta": [], "pos": {}}
        elif cur and line.startswith("#"):
            sections[cur]["meta"].append(line)
        elif cur and line.startswith("p "):
            _, pid, blk, x, y = line.split()
            sections[cur]["pos"][int(pid)] = (int(blk), float(x), float(y))
    return sections


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(B200)), reason="drop-in drivers not built")
def test_same_caller_same_results():
    a, b = run(REF, 2), run(B200, 2)
    assert list(a) == list(b) == ["init", "free", "ui", "end"]
    for sec in a:
        # time, step count, particle / bond / frozen / portal counts, gravity flag: exact
        assert a[sec]["head"] == b[sec]["head"], sec
        # domain, portal geometry + Portal::blocks sizes, bond list, frozen ids: exact
        assert a[sec]["meta"] == b[sec]["meta"], sec
        assert a[sec]["pos"].keys() == b[sec]["pos"].keys()
        pa = np.array([a[sec]["pos"][k][1:] for k in sorted(a[sec]["pos"])])
        pb = np.array([b[sec]["pos"][k][1:] for k in sorted(b[sec]["pos"])])
        ba = np.array([a[sec]["pos"][k][0] for k in sorted(a[sec]["pos"])])
        bb = np.array([b[sec]["pos"][k][0] for k in sorted(b[sec]["pos"])])
        err = np.abs(pa - pb).max()
        assert err <= 1e-4, (sec, err)          # <= 101 iterations in total
        if sec in ("init", "free"):
            assert np.array_equal(ba, bb), sec   # GetBlockById()[id].first

"""TEST INFRASTRUCTURE — generic ctypes binding shared by oracle.ref and oracle.port.

Both libraries export the same entry points under a different prefix
(``ref_`` = the compiled unmodified reference, ``orc_`` = the plain-C restatement).
"""
import ctypes as C

import numpy as np

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


class _Lib:
    """Attribute access ``lib.step_n`` → ``<prefix>step_n`` of the CDLL."""

    def __init__(self, path, prefix):
        self._dll = C.CDLL(path)
        self._prefix = prefix
        vp, sz, f, i = C.c_void_p, C.c_size_t, C.c_float, C.c_int

        def sig(name, res, *args):
            fn = getattr(self._dll, prefix + name)
            fn.restype = res
            fn.argtypes = list(args)
            setattr(self, name, fn)

        sig("create_default", vp)
        sig("destroy", None, vp)
        sig("reset_scene", None, vp, f, f, f, f)
        sig("add_particles", None, vp, _f32p, _f32p, _i32p, sz)
        sig("set_bonds", None, vp, _i32p, sz)
        sig("num_bonds", sz, vp)
        sig("get_bonds", sz, vp, _i32p, sz)
        sig("get_no_rendering", sz, vp, _i32p, sz)
        sig("set_frozen", None, vp, _i32p, sz)
        sig("get_frozen", sz, vp, _i32p, sz)
        sig("add_portal_pair", None, vp, _f32p)
        sig("clear_portals", None, vp)
        sig("num_portal_pairs", sz, vp)
        sig("get_portals", None, vp, _f32p)
        sig("get_portal_blocks", sz, vp, sz, i, _i64p, sz)
        sig("remove_last_portal", None, vp)
        sig("set_gravity", None, vp, i, f, f)
        sig("get_gravity", None, vp, C.POINTER(C.c_int), _f32p)
        sig("set_point_force", None, vp, i, i, f, f)
        sig("set_pick", None, vp, i, i, f, f)
        sig("push_resize", None, vp, f, f, f, f)
        sig("set_renderer_ready", None, vp, i)
        sig("step_to", None, vp, f)
        sig("step_n", None, vp, i, i)
        sig("eval_forces", None, vp)
        sig("time", f, vp)
        sig("dt", f, vp)
        sig("num_steps", sz, vp)
        sig("num_particles", sz, vp)
        sig("num_per_cell", sz, vp)
        sig("num_blocks", sz, vp)
        sig("id_capacity", sz, vp)
        sig("get_geometry", None, vp, _f32p, _i32p)
        sig("get_neighbor_offsets", None, vp, _i32p)
        sig("get_state_by_id", None, vp, sz, vp, vp, vp, vp, vp)
        sig("get_block_lists", sz, vp, _i32p, sz, _i64p)
        sig("get_particle_buffer", sz, vp, _f32p, sz)
        sig("tool", None, vp, i, i, f, f)
        sig("set_particle_buffer", None, vp)
        sig("F12", None, _f32p, _f32p, _f32p)
        sig("F12wall", None, _f32p, _f32p, _f32p)
        sig("F12_bond", None, _f32p, _f32p, _f32p)
        sig("F12_portal_edge", None, _f32p, _f32p, _f32p)
        sig("F12_pick", None, _f32p, _f32p, _f32p)
        sig("find_block", C.c_int64, vp, f, f)
        sig("move_to_portal", None, vp, sz, i, _f32p, _f32p)
        sig("openmp_threads", i)


TOOLS = {"bonds": 0, "pick": 1, "freeze": 2, "portal": 3}
PHASES = {"start": 0, "move": 1, "stop": 2}


def _f2(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, 2))


class OracleParticles:
    """A ``Particles`` object (particles.h:60-228) holding the default scene."""

    def __init__(self, lib):
        self.L = lib
        self.h = self.L.create_default()

    def close(self):
        if self.h:
            self.L.destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene construction -------------------------------------------------
    def reset_scene(self, domain):
        ax, ay, bx, by = [float(v) for v in domain]
        self.L.reset_scene(self.h, ax, ay, bx, by)

    def add_particles(self, pos, vel, ids):
        pos, vel = _f2(pos), _f2(vel)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        assert len(pos) == len(vel) == len(ids)
        self.L.add_particles(self.h, pos, vel, ids, len(ids))

    def set_bonds(self, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        if len(pairs):
            order = np.lexsort((pairs[:, 1], pairs[:, 0]))
            pairs = np.ascontiguousarray(pairs[order])
        self.L.set_bonds(self.h, pairs, len(pairs))

    def get_bonds(self):
        n = self.L.num_bonds(self.h)
        out = np.zeros((max(n, 1), 2), np.int32)
        self.L.get_bonds(self.h, out, n)
        return out[:n]

    def get_no_rendering(self):
        probe = np.zeros((1, 2), np.int32)
        n = self.L.get_no_rendering(self.h, probe, 0)
        out = np.zeros((max(n, 1), 2), np.int32)
        self.L.get_no_rendering(self.h, out, n)
        return out[:n]

    def set_frozen(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self.L.set_frozen(self.h, ids, len(ids))

    def get_frozen(self):
        probe = np.zeros(1, np.int32)
        n = self.L.get_frozen(self.h, probe, 0)
        out = np.zeros(max(n, 1), np.int32)
        self.L.get_frozen(self.h, out, n)
        return out[:n]

    def add_portal_pair(self, b0, e0, b1, e1):
        s = np.array([*b0, *e0, *b1, *e1], np.float32)
        self.L.add_portal_pair(self.h, s)

    def clear_portals(self):
        self.L.clear_portals(self.h)

    def get_portals(self):
        n = self.L.num_portal_pairs(self.h)
        out = np.zeros((max(n, 1), 2, 4), np.float32)
        if n:
            self.L.get_portals(self.h, out)
        return out[:n]

    def get_portal_blocks(self, pair, side):
        probe = np.zeros(1, np.int64)
        n = self.L.get_portal_blocks(self.h, pair, side, probe, 0)
        out = np.zeros(max(n, 1), np.int64)
        self.L.get_portal_blocks(self.h, pair, side, out, n)
        return out[:n]

    def remove_last_portal(self):
        self.L.remove_last_portal(self.h)

    def set_gravity(self, enabled, g=(0.0, -10.0)):
        self.L.set_gravity(self.h, int(enabled), float(g[0]), float(g[1]))

    def set_point_force(self, enabled, attractive=False, center=(0.0, 0.0)):
        self.L.set_point_force(
            self.h, int(enabled), int(attractive), float(center[0]), float(center[1]))

    def set_pick(self, enabled, pid=0, pointer=(0.0, 0.0)):
        self.L.set_pick(
            self.h, int(enabled), int(pid), float(pointer[0]), float(pointer[1]))

    def push_resize(self, domain):
        self.L.push_resize(self.h, *[float(v) for v in domain])

    def set_renderer_ready(self, v):
        self.L.set_renderer_ready(self.h, int(v))

    def tool(self, tool, phase, point):
        self.L.tool(
            self.h, TOOLS[tool], PHASES[phase], float(point[0]), float(point[1]))

    def set_particle_buffer(self):
        self.L.set_particle_buffer(self.h)

    # -- stepping -------------------------------------------------------------
    def step_n(self, n, snapshot=False):
        self.L.step_n(self.h, int(n), int(snapshot))

    def step_to(self, t):
        self.L.step_to(self.h, float(t))

    def eval_forces(self):
        self.L.eval_forces(self.h)

    # -- read-back --------------------------------------------------------------
    @property
    def time(self):
        return np.float32(self.L.time(self.h))

    @property
    def dt(self):
        return np.float32(self.L.dt(self.h))

    @property
    def num_steps(self):
        return self.L.num_steps(self.h)

    @property
    def num_particles(self):
        return self.L.num_particles(self.h)

    @property
    def num_per_cell(self):
        return self.L.num_per_cell(self.h)

    @property
    def num_blocks(self):
        return self.L.num_blocks(self.h)

    @property
    def id_capacity(self):
        return self.L.id_capacity(self.h)

    def geometry(self):
        dom = np.zeros(8, np.float32)
        dims = np.zeros(2, np.int32)
        self.L.get_geometry(self.h, dom, dims)
        return {"domain": dom[:4].copy(), "grid": dom[4:].copy(),
                "dims": (int(dims[0]), int(dims[1]))}

    def neighbor_offsets(self):
        out = np.zeros(9, np.int32)
        self.L.get_neighbor_offsets(self.h, out)
        return out

    def state_by_id(self, cap=None):
        cap = int(self.id_capacity if cap is None else cap)
        n = max(cap, 1)
        pos = np.zeros((n, 2), np.float32)
        vel = np.zeros((n, 2), np.float32)
        force = np.zeros((n, 2), np.float32)
        block = np.zeros(n, np.int64)
        slot = np.zeros(n, np.int64)
        self.L.get_state_by_id(
            self.h, cap, pos.ctypes.data, vel.ctypes.data, force.ctypes.data,
            block.ctypes.data, slot.ctypes.data)
        return {"pos": pos[:cap], "vel": vel[:cap], "force": force[:cap],
                "block": block[:cap], "slot": slot[:cap]}

    def block_lists(self):
        nb = self.num_blocks
        start = np.zeros(nb + 1, np.int64)
        probe = np.zeros(1, np.int32)
        n = self.L.get_block_lists(self.h, probe, 0, start)
        ids = np.zeros(max(n, 1), np.int32)
        self.L.get_block_lists(self.h, ids, n, start)
        return ids[:n], start

    def particle_buffer(self):
        probe = np.zeros(4, np.float32)
        n = self.L.get_particle_buffer(self.h, probe, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        self.L.get_particle_buffer(self.h, out, n)
        return out[:n]

    def find_block(self, x, y):
        return int(self.L.find_block(self.h, float(x), float(y)))

    def move_to_portal(self, pair, side, pos, vel):
        p = np.array(pos, np.float32)
        v = np.array(vel, np.float32)
        self.L.move_to_portal(self.h, pair, side, p, v)
        return p, v

    @property
    def openmp_threads(self):
        return self.L.openmp_threads()

    # force laws as free functions (known-answer tests)
    def law(self, name, p, q):
        out = np.zeros(2, np.float32)
        getattr(self.L, name)(np.array(p, np.float32), np.array(q, np.float32), out)
        return out

"""Python host mirror of the reference ``class Particles`` (src/particles.h:60-228) over the
C ABI of ``libptoy_b200.so`` (include/ptoy_b200.h).

Two groups of methods:

* the reference's own public members, same names and argument meaning
  (``step``, ``SetRendererReadyForNext``, ``GetTime``, ``GetParticles``, ``SetGravity``,
  ``BondsStart`` …) so callers written against ``particles.h`` read the same;
* loader / read-back helpers with the names the oracle bindings use
  (``reset_scene``, ``add_particles``, ``step_n``, ``state_by_id`` …) so parity tests can
  drive the CUDA engine, the C restatement and the compiled reference through one code path.

There is no CPU path: constructing ``Particles`` without the built CUDA library or without
a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTOY_B200_LIB", os.path.join(_HERE, "libptoy_b200.so"))
_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")

# name -> argtypes after the handle (restype is always int)
_vp, _sz, _f, _i = C.c_void_p, C.c_size_t, C.c_float, C.c_int
_szp, _ip, _fp = C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_float)
SIGNATURES = {
    "ptoy_load_default_scene": [],
    "ptoy_reset_scene": [_f, _f, _f, _f],
    "ptoy_add_particles": [_f32p, _f32p, _i32p, _sz],
    "ptoy_add_particles_device": [_vp, _vp, _vp, _sz],
    "ptoy_upload_state": [_vp, _vp, _vp, _sz],
    "ptoy_download_state": [_vp, _vp, _vp, _sz, _szp],
    "ptoy_upload_state_async": [_vp, _vp, _vp, _sz, C.c_int32],
    "ptoy_download_state_async": [_vp, _vp, _vp, _sz],
    "ptoy_downloaded_count": [_szp],
    "ptoy_step": [_f],
    "ptoy_step_n": [_i],
    "ptoy_synchronize": [],
    "ptoy_set_renderer_ready": [_i],
    "ptoy_set_gravity": [_i],
    "ptoy_get_gravity": [_ip],
    "ptoy_set_gravity_vect": [_f, _f],
    "ptoy_get_gravity_vect": [_f32p],
    "ptoy_set_point_force": [_i, _i, _f, _f],
    "ptoy_set_pick": [_i, C.c_int32, _f, _f],
    "ptoy_push_resize": [_f, _f, _f, _f],
    "ptoy_set_domain": [_f, _f, _f, _f],
    "ptoy_set_walls": [_f32p, _sz],
    "ptoy_reset_env_frame": [_f, _f, _f, _f],
    "ptoy_set_bonds": [_i32p, _sz],
    "ptoy_toggle_bond": [C.c_int32, C.c_int32],
    "ptoy_get_bonds": [_vp, _sz, _szp],
    "ptoy_get_no_rendering": [_vp, _sz, _szp],
    "ptoy_set_frozen": [_i32p, _sz],
    "ptoy_get_frozen": [_vp, _sz, _szp],
    "ptoy_add_portal_pair": [_f32p],
    "ptoy_clear_portals": [],
    "ptoy_remove_last_portal": [],
    "ptoy_num_portal_pairs": [_szp],
    "ptoy_get_portals": [_f32p],
    "ptoy_get_portal_blocks": [_sz, _i, _vp, _sz, _szp],
    "ptoy_tool": [_i, _i, _f, _f],
    "ptoy_get_portal_tool_state": [_i32p, _f32p],
    "ptoy_time": [_fp],
    "ptoy_dt": [_fp],
    "ptoy_num_steps": [_szp],
    "ptoy_get_domain": [_f32p],
    "ptoy_get_geometry": [_f32p, _i32p],
    "ptoy_num_particles": [_szp],
    "ptoy_snapshot_counts": [_szp, _szp],
    "ptoy_set_particle_buffer": [],
    "ptoy_get_snapshot": [_vp, _vp, _vp, _vp, _sz, _szp],
    "ptoy_get_state_by_id": [_sz, _vp, _vp, _vp],
    "ptoy_eval_forces": [_i, _sz, _f32p],
    "ptoy_id_capacity": [_szp],
    "ptoy_find_blocks": [_f32p, _sz, _i64p],
    "ptoy_wire_particles": [np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS"), _i, _i, _i, _ip],
    "ptoy_wire_scene": [_i, np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS"), _i, _i, _i, _ip],
    "ptoy_configure_strips": [_i, _i, _i32p, _i, _sz, _i],
    "ptoy_nccl_connect": [_vp],
    "ptoy_error_flags": [_ip],
    "ptoy_enable_kernel_timing": [_i],
    "ptoy_get_kernel_times": [np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), _i64p],
    "ptoy_launch_count": [C.POINTER(C.c_int64)],
    "ptoy_stream": [C.POINTER(C.c_void_p)],
}
# entry points that do not take a handle first
FREE_FUNCTIONS = ["ptoy_last_error", "ptoy_abi_version", "ptoy_create", "ptoy_destroy",
                  "ptoy_nccl_unique_id", "ptoy_group_create", "ptoy_group_step_n", "ptoy_group_destroy"]
KERNEL_KINDS = ["stage1", "stage2", "scan", "scatter", "fixup", "other"]

TOOLS = {"bonds": 0, "pick": 1, "freeze": 2, "portal": 3}
PHASES = {"start": 0, "move": 1, "stop": 2}


class PtoyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ptoy_b200 error {code}: {msg}")
        self.code = code


def load_library():
    """dlopen the in-tree CUDA library; fail loudly if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). ptoy_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.ptoy_last_error.restype = C.c_char_p
    L.ptoy_last_error.argtypes = []
    L.ptoy_abi_version.restype = C.c_int
    L.ptoy_create.restype = C.c_int
    L.ptoy_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.ptoy_destroy.restype = None
    L.ptoy_destroy.argtypes = [C.c_void_p]
    L.ptoy_nccl_unique_id.restype = C.c_int
    L.ptoy_nccl_unique_id.argtypes = [C.c_void_p]
    L.ptoy_group_create.restype = C.c_int
    L.ptoy_group_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]
    L.ptoy_group_step_n.restype = C.c_int
    L.ptoy_group_step_n.argtypes = [C.c_void_p, C.c_int]
    L.ptoy_group_destroy.restype = None
    L.ptoy_group_destroy.argtypes = [C.c_void_p]
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p] + args
    _lib = L
    return L


def _f2(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, 2))


def _ptr(a):
    return None if a is None else a.ctypes.data


class Particles:
    """Drop-in for the reference ``Particles`` on one CUDA device.

    ``Particles()`` builds the reference constructor's default scene
    (particles.cpp:26-77); ``Particles(default_scene=False)`` starts empty.
    """

    kind = "cuda"

    def __init__(self, device=0, default_scene=True):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.ptoy_create(C.byref(h), int(device))
        if rc != 0:
            raise PtoyError(rc, self.L.ptoy_last_error().decode())
        self.h = h
        if default_scene:
            self._call("ptoy_load_default_scene")

    def _call(self, name, *args):
        rc = getattr(self.L, name)(self.h, *args)
        if rc != 0:
            raise PtoyError(rc, self.L.ptoy_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.ptoy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference API
    def step(self, time_target, quit=False):                       # particles.h:121
        if not quit:
            self._call("ptoy_step", float(time_target))

    def SetRendererReadyForNext(self, value):                      # particles.h:179
        self._call("ptoy_set_renderer_ready", int(bool(value)))

    def GetTime(self):                                             # particles.h:141
        t = C.c_float()
        self._call("ptoy_time", C.byref(t))
        return np.float32(t.value)

    def GetNumSteps(self):                                         # particles.h:144
        n = C.c_size_t()
        self._call("ptoy_num_steps", C.byref(n))
        return n.value

    def GetDomain(self):                                           # particles.h:147
        d = np.zeros(4, np.float32)
        self._call("ptoy_get_domain", d)
        return d

    def GetGravity(self):                                          # particles.h:152
        e = C.c_int()
        self._call("ptoy_get_gravity", C.byref(e))
        return bool(e.value)

    def SetGravity(self, flag):                                    # particles.h:155
        self._call("ptoy_set_gravity", int(bool(flag)))

    def GetGravityVect(self):                                      # particles.h:158
        g = np.zeros(2, np.float32)
        self._call("ptoy_get_gravity_vect", g)
        return g

    def SetGravityVect(self, g):                                   # particles.h:161
        self._call("ptoy_set_gravity_vect", float(g[0]), float(g[1]))

    def SetForce(self, center=None, enabled=None):                 # particles.h:122-124
        if center is not None:
            self._force_center = (float(center[0]), float(center[1]))
        if enabled is not None:
            self._force_enabled = bool(enabled)
        self._push_force()

    def SetForceAttractive(self, value):                           # particles.h:125
        self._force_attractive = bool(value)
        self._push_force()

    _force_center = (0.0, 0.0)
    _force_enabled = False
    _force_attractive = False

    def _push_force(self):
        self._call("ptoy_set_point_force", int(self._force_enabled), int(self._force_attractive),
                   self._force_center[0], self._force_center[1])

    def PushResize(self, domain):                                  # particles.h:112
        self._call("ptoy_push_resize", *[float(v) for v in domain])

    def SetDomain(self, domain):                                   # particles.h:108
        self._call("ptoy_set_domain", *[float(v) for v in domain])

    def ResetEnvObjFrame(self, domain):                            # particles.h:99
        self._call("ptoy_reset_env_frame", *[float(v) for v in domain])

    def RemoveLastPortal(self):                                    # particles.h:90
        self._call("ptoy_remove_last_portal")

    def SetParticleBuffer(self):                                   # particles.h:89
        self._call("ptoy_set_particle_buffer")

    def GetParticles(self):                                        # particles.h:164 → [n,4] p.xy v.xy
        s = self.snapshot()
        return np.concatenate([s["pos"], s["vel"]], axis=1)

    def GetNumParticles(self):                                     # particles.h:167
        a, b = C.c_size_t(), C.c_size_t()
        self._call("ptoy_snapshot_counts", C.byref(a), C.byref(b))
        return a.value

    def GetNumPerCell(self):                                       # particles.h:170
        a, b = C.c_size_t(), C.c_size_t()
        self._call("ptoy_snapshot_counts", C.byref(a), C.byref(b))
        return b.value

    def GetBonds(self):                                            # particles.h:115
        return self.get_bonds()

    def GetFrozen(self):                                           # particles.h:118
        return self.get_frozen()

    def GetNoRendering(self):                                      # particles.h:225
        return self.get_no_rendering()

    def GetPortals(self):                                          # particles.h:82
        return self.get_portals()

    def GetBlockById(self):                                        # particles.h:173 → (block, slot) per id
        s = self.snapshot()
        cap = self.id_capacity
        out = np.full((cap, 2), -1, np.int64)
        if len(s["id"]):
            starts = np.r_[0, np.nonzero(np.diff(s["block"]))[0] + 1]
            first = np.repeat(starts, np.diff(np.r_[starts, len(s["block"])]))
            out[s["id"], 0] = s["block"]
            out[s["id"], 1] = np.arange(len(s["id"])) - first
        return out

    def _tool(self, tool, phase, point):
        self._call("ptoy_tool", TOOLS[tool], PHASES[phase], float(point[0]), float(point[1]))

    def BondsStart(self, p): self._tool("bonds", "start", p)       # particles.h:126-128
    def BondsMove(self, p): self._tool("bonds", "move", p)
    def BondsStop(self, p): self._tool("bonds", "stop", p)
    def PickStart(self, p): self._tool("pick", "start", p)         # particles.h:133-135
    def PickMove(self, p): self._tool("pick", "move", p)
    def PickStop(self, p): self._tool("pick", "stop", p)
    def FreezeStart(self, p): self._tool("freeze", "start", p)     # particles.h:130-132
    def FreezeMove(self, p): self._tool("freeze", "move", p)
    def FreezeStop(self, p): self._tool("freeze", "stop", p)
    def PortalStart(self, p): self._tool("portal", "start", p)     # particles.h:136-138
    def PortalMove(self, p): self._tool("portal", "move", p)
    def PortalStop(self, p): self._tool("portal", "stop", p)

    def portal_tool_state(self):                                   # particles.h:214-218
        sm = np.zeros(2, np.int32)
        xy = np.zeros(8, np.float32)
        self._call("ptoy_get_portal_tool_state", sm, xy)
        return {"portal_stage_": int(sm[0]), "portal_mouse_moving_": bool(sm[1]),
                "portal_begin_": xy[0:2], "portal_current_": xy[2:4],
                "portal_prev_": (xy[4:6], xy[6:8])}

    # ------------------------------------------------------------------ loader helpers
    def reset_scene(self, domain):
        self._call("ptoy_reset_scene", *[float(v) for v in domain])

    def add_particles(self, pos, vel, ids):
        pos, vel = _f2(pos), _f2(vel)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        assert len(pos) == len(vel) == len(ids)
        self._call("ptoy_add_particles", pos, vel, ids, len(ids))

    def add_particles_device(self, pos_ptr, vel_ptr, id_ptr, n):
        self._call("ptoy_add_particles_device", pos_ptr, vel_ptr, id_ptr, int(n))

    def upload_state(self, pos, vel, ids):
        """Replace pos/vel/id (C-contiguous float32/int32 numpy arrays, ideally pinned)."""
        assert pos.dtype == np.float32 and vel.dtype == np.float32 and ids.dtype == np.int32
        self._call("ptoy_upload_state", pos.ctypes.data, vel.ctypes.data, ids.ctypes.data, len(ids))

    def download_state(self, pos, vel, ids):
        """Copy the live state (engine order) into the given host arrays; returns the count."""
        n = C.c_size_t()
        self._call("ptoy_download_state", _ptr(pos), _ptr(vel), _ptr(ids), len(pos), C.byref(n))
        return n.value

    def upload_state_async(self, pos, vel, ids, max_id):
        """Enqueue upload_state on this engine's stream (pinned arrays; nothing waits)."""
        assert pos.dtype == np.float32 and vel.dtype == np.float32 and ids.dtype == np.int32
        self._call("ptoy_upload_state_async", pos.ctypes.data, vel.ctypes.data, ids.ctypes.data, len(ids),
                   int(max_id))

    def download_state_async(self, pos, vel, ids):
        """Enqueue download_state; the arrays are valid after synchronize()."""
        self._call("ptoy_download_state_async", _ptr(pos), _ptr(vel), _ptr(ids), len(pos))

    @property
    def downloaded_count(self):
        n = C.c_size_t()
        self._call("ptoy_downloaded_count", C.byref(n))
        return n.value

    def set_bonds(self, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        self._call("ptoy_set_bonds", pairs, len(pairs))

    def toggle_bond(self, a, b):
        self._call("ptoy_toggle_bond", int(a), int(b))

    def _get_pairs(self, fn):
        n = C.c_size_t()
        self._call(fn, None, 0, C.byref(n))
        out = np.zeros((max(n.value, 1), 2), np.int32)
        self._call(fn, out.ctypes.data, n.value, C.byref(n))
        return out[:n.value]

    def get_bonds(self):
        return self._get_pairs("ptoy_get_bonds")

    def get_no_rendering(self):
        return self._get_pairs("ptoy_get_no_rendering")

    def set_frozen(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._call("ptoy_set_frozen", ids, len(ids))

    def get_frozen(self):
        n = C.c_size_t()
        self._call("ptoy_get_frozen", None, 0, C.byref(n))
        out = np.zeros(max(n.value, 1), np.int32)
        self._call("ptoy_get_frozen", out.ctypes.data, n.value, C.byref(n))
        return out[:n.value]

    def add_portal_pair(self, b0, e0, b1, e1):
        self._call("ptoy_add_portal_pair", np.array([*b0, *e0, *b1, *e1], np.float32))

    def clear_portals(self):
        self._call("ptoy_clear_portals")

    def remove_last_portal(self):
        self._call("ptoy_remove_last_portal")

    def get_portals(self):
        n = C.c_size_t()
        self._call("ptoy_num_portal_pairs", C.byref(n))
        out = np.zeros((max(n.value, 1), 2, 4), np.float32)
        if n.value:
            self._call("ptoy_get_portals", out)
        return out[:n.value]

    def get_portal_blocks(self, pair, side):
        n = C.c_size_t()
        self._call("ptoy_get_portal_blocks", pair, side, None, 0, C.byref(n))
        out = np.zeros(max(n.value, 1), np.int64)
        self._call("ptoy_get_portal_blocks", pair, side, out.ctypes.data, n.value, C.byref(n))
        return out[:n.value]

    def set_gravity(self, enabled, g=(0.0, -10.0)):
        self.SetGravity(enabled)
        self.SetGravityVect(g)

    def set_point_force(self, enabled, attractive=False, center=(0.0, 0.0)):
        self._force_enabled, self._force_attractive = bool(enabled), bool(attractive)
        self._force_center = (float(center[0]), float(center[1]))
        self._push_force()

    def set_pick(self, enabled, pid=0, pointer=(0.0, 0.0)):
        self._call("ptoy_set_pick", int(enabled), int(pid), float(pointer[0]), float(pointer[1]))

    def push_resize(self, domain):
        self.PushResize(domain)

    def set_renderer_ready(self, v):
        self.SetRendererReadyForNext(v)

    def tool(self, tool, phase, point):
        self._tool(tool, phase, point)

    def set_particle_buffer(self):
        self.SetParticleBuffer()

    # ------------------------------------------------------------------ stepping
    def step_n(self, n, snapshot=False):
        """n iterations of the loop at particles.cpp:96; with snapshot=True each one ends
        with the render snapshot (one step() call per iteration, like the oracle wrapper)."""
        if snapshot:
            for _ in range(int(n)):
                self._call("ptoy_set_renderer_ready", 1)
                self._call("ptoy_step_n", 1)
        else:
            self._call("ptoy_set_renderer_ready", 0)
            self._call("ptoy_step_n", int(n))

    def step_to(self, t):
        self._call("ptoy_step", float(t))

    def synchronize(self):
        self._call("ptoy_synchronize")

    # ------------------------------------------------------------------ read-back
    @property
    def time(self):
        return self.GetTime()

    @property
    def dt(self):
        t = C.c_float()
        self._call("ptoy_dt", C.byref(t))
        return np.float32(t.value)

    @property
    def num_steps(self):
        return self.GetNumSteps()

    @property
    def num_particles(self):
        n = C.c_size_t()
        self._call("ptoy_num_particles", C.byref(n))
        return n.value

    @property
    def id_capacity(self):
        n = C.c_size_t()
        self._call("ptoy_id_capacity", C.byref(n))
        return n.value

    def geometry(self):
        dom = np.zeros(8, np.float32)
        dims = np.zeros(2, np.int32)
        self._call("ptoy_get_geometry", dom, dims)
        return {"domain": dom[:4].copy(), "grid": dom[4:].copy(), "dims": (int(dims[0]), int(dims[1]))}

    def state_by_id(self, cap=None):
        cap = int(self.id_capacity if cap is None else cap)
        n = max(cap, 1)
        pos = np.zeros((n, 2), np.float32)
        vel = np.zeros((n, 2), np.float32)
        block = np.zeros(n, np.int64)
        self._call("ptoy_get_state_by_id", cap, pos.ctypes.data, vel.ctypes.data, block.ctypes.data)
        return {"pos": pos[:cap], "vel": vel[:cap], "block": block[:cap]}

    def eval_forces(self, stage=1, cap=None):
        """Forces of `stage` by particle id (NaN for absent ids); state is not advanced."""
        cap = int(self.id_capacity if cap is None else cap)
        out = np.zeros((max(cap, 1), 2), np.float32)
        self._call("ptoy_eval_forces", int(stage), cap, out)
        return out[:cap]

    def snapshot(self):
        n = C.c_size_t()
        self._call("ptoy_get_snapshot", None, None, None, None, 0, C.byref(n))
        m = max(n.value, 1)
        pos = np.zeros((m, 2), np.float32)
        vel = np.zeros((m, 2), np.float32)
        ids = np.zeros(m, np.int32)
        block = np.zeros(m, np.int64)
        self._call("ptoy_get_snapshot", pos.ctypes.data, vel.ctypes.data, ids.ctypes.data,
                   block.ctypes.data, n.value, C.byref(n))
        k = n.value
        return {"pos": pos[:k], "vel": vel[:k], "id": ids[:k], "block": block[:k]}

    def particle_buffer(self):
        return self.GetParticles()

    def block_lists(self):
        """(ids in block order, block_start[nblocks+1]) of the render snapshot."""
        s = self.snapshot()
        di, dj = self.geometry()["dims"]
        counts = np.bincount(s["block"], minlength=di * dj)
        return s["id"], np.r_[0, np.cumsum(counts)].astype(np.int64)

    def find_blocks(self, pos):
        pos = _f2(pos)
        out = np.zeros(max(len(pos), 1), np.int64)
        self._call("ptoy_find_blocks", pos, len(pos), out)
        return out[:len(pos)]

    # ------------------------------------------------------------------ wire format (ptoy_wasm.cpp:130-198)
    def wire_particles(self, max_size=10000, width=800, height=800):
        """uint16 pixel pairs of the live particles, quantised and culled on the device."""
        data = np.zeros(max(int(max_size), 2), np.uint16)
        n = C.c_int()
        self._call("ptoy_wire_particles", data, int(max_size), int(width), int(height), C.byref(n))
        return data[:n.value]

    def wire_scene(self, what, max_size=10000, width=800, height=800):
        """what: 'portals' | 'bonds' | 'frozen' (from the host state and the last render snapshot)."""
        data = np.zeros(max(int(max_size), 8), np.uint16)
        n = C.c_int()
        self._call("ptoy_wire_scene", {"portals": 1, "bonds": 2, "frozen": 3}[what], data, int(max_size), int(width),
                   int(height), C.byref(n))
        return data[:n.value]

    # ------------------------------------------------------------------ strips (multi-GPU)
    def configure_strips(self, rank, world, col_begin=None, ghost_rows=0, id_capacity=0, bonds=False):
        """Own block columns [col_begin[rank], col_begin[rank+1]) of the current grid.  Call after
        reset_scene and before add_particles; add_particles then keeps only this strip's particles.
        ghost_rows / id_capacity / bonds: see include/ptoy_b200.h ptoy_configure_strips."""
        if col_begin is None:
            col_begin = partition_columns(self.geometry()["dims"][0], world)
        cb = np.ascontiguousarray(col_begin, dtype=np.int32)
        assert len(cb) == world + 1
        self._call("ptoy_configure_strips", int(rank), int(world), cb, int(ghost_rows), int(id_capacity), int(bool(bonds)))
        self.rank, self.world, self.col_begin = int(rank), int(world), cb

    def nccl_connect(self, uid):
        """Collective: every rank passes the 128-byte id created by nccl_unique_id() on rank 0."""
        _prefer_torch_nccl()
        buf = (C.c_char * 128).from_buffer_copy(bytes(uid))
        self._call("ptoy_nccl_connect", C.cast(buf, C.c_void_p))

    @property
    def error_flags(self):
        f = C.c_int()
        self._call("ptoy_error_flags", C.byref(f))
        return f.value

    # ------------------------------------------------------------------ measurement
    def enable_kernel_timing(self, enable=True):
        self._call("ptoy_enable_kernel_timing", int(enable))

    def kernel_times(self):
        ms = np.zeros(6, np.float64)
        n = np.zeros(6, np.int64)
        self._call("ptoy_get_kernel_times", ms, n)
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_KINDS)}

    @property
    def launch_count(self):
        n = C.c_int64()
        self._call("ptoy_launch_count", C.byref(n))
        return n.value

    @property
    def stream(self):
        s = C.c_void_p()
        self._call("ptoy_stream", C.byref(s))
        return s.value


def partition_columns(n_columns, world, weights=None):
    """Strip boundaries over the block columns: equal columns, or balanced by `weights` (particle
    count per column).  Every strip gets at least two columns."""
    if world == 1:
        return np.array([0, n_columns], np.int32)
    assert n_columns >= 2 * world, "each strip needs at least two block columns"
    if weights is None:
        cb = np.round(np.linspace(0, n_columns, world + 1)).astype(np.int64)
    else:
        w = np.asarray(weights, np.float64)
        assert len(w) == n_columns
        c = np.r_[0.0, np.cumsum(w)]
        targets = c[-1] * np.arange(world + 1) / world
        cb = np.searchsorted(c, targets, side="left").astype(np.int64)
        cb[0], cb[-1] = 0, n_columns
    for r in range(1, world):                      # enforce the two-column minimum
        cb[r] = max(cb[r], cb[r - 1] + 2)
    for r in range(world - 1, 0, -1):
        cb[r] = min(cb[r], cb[r + 1] - 2)
    return cb.astype(np.int32)


def _prefer_torch_nccl():
    """The engine binds NCCL by soname at run time; importing torch first makes that PyTorch's
    bundled libnccl.so.2 (newer than the system copy, and the one libtorch_cuda needs)."""
    try:
        import torch  # noqa: F401
    except Exception:
        pass


def nccl_unique_id():
    _prefer_torch_nccl()
    L = load_library()
    buf = (C.c_char * 128)()
    rc = L.ptoy_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise PtoyError(rc, L.ptoy_last_error().decode())
    return bytes(buf)


class StripGroup:
    """Several strips of one scene in ONE process on ONE device, stepped in lock step with
    device-to-device copies in place of NCCL (include/ptoy_b200.h ptoy_group_*).  The test
    vehicle of the multi-GPU path: N virtual strips must equal one engine bit for bit."""

    def __init__(self, domain, world, device=0, col_begin=None, gravity=(True, (0.0, -10.0)),
                 ghost_rows=0, id_capacity=0, bonds=False):
        self.L = load_library()
        self.parts = []
        for r in range(world):
            p = Particles(device=device, default_scene=False)
            p.reset_scene(domain)
            p.set_gravity(*gravity)
            p.configure_strips(r, world, col_begin, ghost_rows, id_capacity, bonds)
            p.set_renderer_ready(False)
            self.parts.append(p)
        arr = (C.c_void_p * world)(*[p.h for p in self.parts])
        g = C.c_void_p()
        rc = self.L.ptoy_group_create(C.byref(g), arr, world)
        if rc != 0:
            raise PtoyError(rc, self.L.ptoy_last_error().decode())
        self.g = g

    def each(self, fn):
        for p in self.parts:
            fn(p)

    def add_particles(self, pos, vel, ids):
        self.each(lambda p: p.add_particles(pos, vel, ids))

    def step_n(self, n):
        rc = self.L.ptoy_group_step_n(self.g, int(n))
        if rc != 0:
            raise PtoyError(rc, self.L.ptoy_last_error().decode())

    def state_by_id(self, cap):
        """Union of the strips' states (every particle lives on exactly one strip)."""
        out = None
        owners = np.zeros(cap, np.int32)
        for p in self.parts:
            s = p.state_by_id(cap)
            alive = s["block"] >= 0
            owners += alive
            if out is None:
                out = {k: v.copy() for k, v in s.items()}
            else:
                for k in out:
                    out[k][alive] = s[k][alive]
        assert owners.max() <= 1, "a particle is held by two strips"
        return out

    @property
    def num_particles(self):
        return sum(p.num_particles for p in self.parts)

    def close(self):
        if getattr(self, "g", None):
            self.L.ptoy_group_destroy(self.g)
            self.g = None
        for p in self.parts:
            p.close()
        self.parts = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

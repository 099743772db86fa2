"""Synthetic scenes of SURVEY.md §8(d) / BASELINE.json ``configs`` (numpy, host side).

Every generator returns a ``Scene`` whose arrays are float32 / int32, seeded and
deterministic, so the same state can be loaded into the CUDA engine, the CPU
oracle and the compiled reference.  Constants follow particles.cpp:7-24.
"""
from dataclasses import dataclass, field

import numpy as np

R = np.float32(0.02)           # kRadius, particles.cpp:7
BLOCK = np.float32(4.0 * 0.02)  # kBlockSize as the reference rounds it (0x3da3d70a)
assert np.float32(4.0 * float(np.float32(0.02))) == np.float32(0.08)


@dataclass
class Scene:
    domain: tuple                      # (ax, ay, bx, by)
    pos: np.ndarray                    # [n,2] f32
    vel: np.ndarray                    # [n,2] f32
    ids: np.ndarray                    # [n] i32
    bonds: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))  # directed, both ways
    frozen: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    portals: list = field(default_factory=list)    # [(b0,e0,b1,e1)] xy tuples
    gravity: tuple = (True, (0.0, -10.0))
    name: str = ""

    @property
    def n(self):
        return len(self.ids)


def _box_for(n, pitch, aspect=4.0 / 3.0):
    """Domain [-1,-1]..[bx,by] that holds ≈n hex-packed particles at `pitch`."""
    row_h = pitch * np.sqrt(3.0) / 2.0
    area = n * pitch * row_h
    w = np.sqrt(area * aspect)
    h = area / w
    return float(w), float(h)


def hex_lattice(n, pitch_over_r=2.2, jitter_over_r=0.025, seed=1234, aspect=4.0 / 3.0,
                vel_scale=0.0):
    """Jittered hexagonal packing (parity scenes C2/C3/C5 of SURVEY.md §8d).

    Particles fill rows bottom-up inside the frame walls; ids follow lattice order
    (row-major), so bonds between lattice neighbours are also id-local.
    """
    r = float(R)
    pitch = pitch_over_r * r
    w, h = _box_for(n, pitch, aspect)
    cols = max(1, int(np.floor((w - 2 * r - pitch / 2) / pitch)))
    rows = int(np.ceil(n / cols))
    row_h = pitch * np.sqrt(3.0) / 2.0
    bx = -1.0 + 2 * r + pitch * (cols + 0.5)
    by = -1.0 + 2 * r + row_h * rows + pitch
    k = np.arange(n, dtype=np.int64)
    j, i = k // cols, k % cols
    x = -1.0 + r + pitch * (i + 0.5 + 0.5 * (j % 2))
    y = -1.0 + r + pitch * 0.5 + row_h * j
    rng = np.random.RandomState(seed)
    jit = (rng.uniform(-1.0, 1.0, size=(n, 2)) * (jitter_over_r * r))
    pos = np.stack([x, y], axis=1) + jit
    vel = np.zeros((n, 2))
    if vel_scale:
        vel = rng.uniform(-vel_scale, vel_scale, size=(n, 2))
    sc = Scene((-1.0, -1.0, float(np.float32(bx)), float(np.float32(by))),
               pos.astype(np.float32), vel.astype(np.float32),
               np.arange(n, dtype=np.int32), name=f"hex{n}")
    sc.lattice = (rows, cols)
    return sc


def uniform_random(n, seed=1234, density=596.0, aspect=4.0 / 3.0):
    """i.i.d. uniform positions (C2 throughput scene; overlapping → single-step parity only)."""
    area = n / density
    w = float(np.sqrt(area * aspect))
    h = area / w
    r = float(R)
    rng = np.random.RandomState(seed)
    pos = np.empty((n, 2))
    pos[:, 0] = rng.uniform(-1.0 + r, -1.0 + w - r, size=n)
    pos[:, 1] = rng.uniform(-1.0 + r, -1.0 + h - r, size=n)
    return Scene((-1.0, -1.0, float(np.float32(-1.0 + w)), float(np.float32(-1.0 + h))),
                 pos.astype(np.float32), np.zeros((n, 2), np.float32),
                 np.arange(n, dtype=np.int32), name=f"uniform{n}")


def add_lattice_bonds(scene, keep_prob=2.0 / 3.0, seed=99):
    """C3: each of the 3 forward hex-neighbour edges kept with p=2/3 → mean directed
    degree ≈4; both directions stored (particles.cpp:493-494)."""
    rows, cols = scene.lattice
    n = scene.n
    k = np.arange(n, dtype=np.int64)
    j, i = k // cols, k % cols
    rng = np.random.RandomState(seed)
    edges = []
    # forward neighbours: right; up-left/up-right depending on row parity
    cand = [(k + 1, i + 1 < cols)]
    up = k + cols
    odd = (j % 2) == 1
    # even row: up (same i) and up-left (i-1); odd row: up (same i) and up-right (i+1)
    cand.append((up, np.ones(n, bool)))
    diag = np.where(odd, up + 1, up - 1)
    diag_ok = np.where(odd, i + 1 < cols, i - 1 >= 0)
    cand.append((diag, diag_ok))
    for nb, ok in cand:
        ok = ok & (nb < n) & (nb >= 0)
        keep = rng.uniform(size=n) < keep_prob
        m = ok & keep
        edges.append(np.stack([k[m], nb[m]], axis=1))
    e = np.concatenate(edges, axis=0)
    both = np.concatenate([e, e[:, ::-1]], axis=0).astype(np.int32)
    order = np.lexsort((both[:, 1], both[:, 0]))
    scene.bonds = np.ascontiguousarray(both[order])
    scene.frozen = np.nonzero(rng.uniform(size=n) < 0.01)[0].astype(np.int32)
    scene.name += "+bonds"
    return scene


def add_portals(scene, n_pairs, length=1.0, seed=7):
    """C4: axis-aligned vertical portal pairs, members ≥10 units apart when the
    domain allows it, spread over the domain width so pairs straddle strips."""
    ax, ay, bx, by = scene.domain
    rng = np.random.RandomState(seed)
    w, h = bx - ax, by - ay
    length = min(length, 0.5 * h)
    for p in range(n_pairs):
        x0 = ax + w * (0.05 + 0.4 * rng.uniform())
        x1 = ax + w * (0.55 + 0.4 * rng.uniform())
        y0 = ay + 0.1 + (h - length - 0.2) * rng.uniform()
        y1 = ay + 0.1 + (h - length - 0.2) * rng.uniform()
        scene.portals.append(((x0, y0), (x0, y0 + length), (x1, y1), (x1, y1 + length)))
    scene.name += f"+{n_pairs}portals"
    return scene


def reference_grid(domain):
    """The block grid blocks::SetDomain (blocks.h:119-141) builds for `domain` starting from the
    constructor's [-1,1]^2 grid: grow-only, whole blocks, by float32 accumulation.  Returns
    (grid A.x, A.y, B.x, B.y, dims.i, dims.j) exactly as the engine / the reference compute them."""
    f = np.float32
    ax, ay, bx, by = f(-1), f(-1), f(1), f(1)
    pa_x, pa_y, pb_x, pb_y = [f(v) for v in domain]
    while pa_x < ax:
        ax = f(ax - BLOCK)
    while pa_y < ay:
        ay = f(ay - BLOCK)
    while pb_x > bx:
        bx = f(bx + BLOCK)
    while pb_y > by:
        by = f(by + BLOCK)
    di = int(f(f(bx - ax) / BLOCK)) + 1
    dj = int(f(f(by - ay) / BLOCK)) + 1
    return float(ax), float(ay), float(bx), float(by), di, dj


def strip_scene(n_per_rank, rank, world, pitch_over_r=2.2, jitter_over_r=0.025, seed=1234, aspect=None,
                bonds=False, bond_margin_cols=8):
    """Weak-scaling scene (configs[4]/[5] style dense packing): ONE global hex-packed scene of
    about world * n_per_rank particles in a roughly square domain; returns the particles of the
    block columns that `rank` owns under the equal-column partition, with global ids.

    The x direction is the slow block index (blocks.h:160), i.e. the direction strips are cut in.
    Attributes: n_global, col_begin (the partition to pass to configure_strips), dims.

    bonds=True adds the configs[2] features to the GLOBAL scene: each of the 3 forward
    hex-neighbour edges kept with p=2/3 and 1 % frozen ids, both decided by a hash of the global
    id, so every rank derives the same network without generating all of it.  `bonds` then holds
    the directed pairs whose first id lies in this rank's lattice columns widened by
    `bond_margin_cols` (the rows a rank can need while particles stay within that margin of
    their strip), `frozen` the frozen ids of the same columns."""
    from .particles import partition_columns
    r = float(R)
    pitch = pitch_over_r * r
    row_h = pitch * np.sqrt(3.0) / 2.0
    # the GLOBAL domain stays roughly square (aspect = width / height) so that float32 coordinates
    # keep their resolution at 256M particles; strips get narrower as world grows
    n_target = n_per_rank * world
    h1 = np.sqrt(n_target * pitch * row_h / (aspect or 1.0))
    rows = max(4, int(round(h1 / row_h)))
    cols_per_rank = int(np.ceil(n_target / rows / world))
    cols = cols_per_rank * world
    n_global = rows * cols
    bx = -1.0 + 2 * r + pitch * (cols + 0.5)
    by = -1.0 + 2 * r + row_h * rows + pitch
    domain = (-1.0, -1.0, float(np.float32(bx)), float(np.float32(by)))
    bs = float(BLOCK)
    _, _, _, _, di, dj = reference_grid(domain)
    cb = partition_columns(di, world)
    # lattice columns that can fall into this rank's block columns (one extra on each side)
    x_lo, x_hi = -1.0 + cb[rank] * bs, -1.0 + cb[rank + 1] * bs
    i_lo = max(0, int(np.floor((x_lo + 1.0 - r) / pitch)) - 2)
    i_hi = min(cols, int(np.ceil((x_hi + 1.0 - r) / pitch)) + 2)
    ii, jj = np.meshgrid(np.arange(i_lo, i_hi, dtype=np.int64), np.arange(rows, dtype=np.int64))
    ii, jj = ii.ravel(), jj.ravel()
    ids = jj * cols + ii
    x = -1.0 + r + pitch * (ii + 0.5 + 0.5 * (jj % 2))
    y = -1.0 + r + pitch * 0.5 + row_h * jj
    # jitter is a function of the GLOBAL id, so every rank generates the same particle
    rng = np.random.RandomState(seed)
    mult = np.uint64(0x9E3779B97F4A7C15)
    hsh = (ids.astype(np.uint64) + np.uint64(seed)) * mult
    jx = ((hsh >> np.uint64(11)) & np.uint64(0xFFFFF)).astype(np.float64) / float(0xFFFFF) * 2 - 1
    jy = ((hsh >> np.uint64(33)) & np.uint64(0xFFFFF)).astype(np.float64) / float(0xFFFFF) * 2 - 1
    pos = np.stack([x + jx * jitter_over_r * r, y + jy * jitter_over_r * r], 1).astype(np.float32)
    # ownership exactly as the engine decides it: blocks::FindBlock column of the float32 position
    col = np.trunc((pos[:, 0] - np.float32(-1.0)) / BLOCK).astype(np.int64)
    mine = (col >= cb[rank]) & (col < cb[rank + 1])
    sc = Scene(domain, pos[mine], np.zeros((int(mine.sum()), 2), np.float32), ids[mine].astype(np.int32),
               name=f"strip{rank}of{world}x{n_per_rank}")
    sc.n_global, sc.col_begin, sc.dims = int(n_global), cb, (di, dj)
    sc.lattice = (rows, cols)
    if bonds:
        c_lo, c_hi = max(0, i_lo - bond_margin_cols), min(cols, i_hi + bond_margin_cols)
        sc.bonds = hashed_lattice_bonds(rows, cols, c_lo, c_hi, seed=seed + 99)
        ci, cj = np.meshgrid(np.arange(c_lo, c_hi, dtype=np.int64), np.arange(rows, dtype=np.int64))
        cid = np.sort((cj * cols + ci).ravel())
        sc.frozen = cid[_hash01(cid * 4 + 3, seed + 99) < 0.01].astype(np.int32)
        sc.name += "+bonds"
    return sc


def _hash01(key, seed):
    """Deterministic uniform [0, 1) per integer key (splitmix64 finaliser)."""
    z = (np.asarray(key, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def hashed_lattice_bonds(rows, cols, c_lo, c_hi, keep_prob=2.0 / 3.0, seed=99):
    """Directed bond pairs (sorted, both directions stored like particles.cpp:493-494) of the
    hashed hex-lattice network whose FIRST id lies in lattice columns [c_lo, c_hi); ids are
    row-major over the global rows x cols lattice.  Edge (k, direction d) is kept iff
    hash(3k+d) < keep_prob, so any column range gives a consistent cut of one global network."""
    e_lo, e_hi = max(0, c_lo - 1), min(cols, c_hi + 1)       # forward edges reach one column over
    ii, jj = np.meshgrid(np.arange(e_lo, e_hi, dtype=np.int64), np.arange(rows, dtype=np.int64))
    i, j = ii.ravel(), jj.ravel()
    k = j * cols + i
    odd = (j % 2) == 1
    up_ok = j + 1 < rows
    cands = [(k + 1, i + 1 < cols),                                    # right
             (k + cols, up_ok),                                        # up, same column
             (np.where(odd, k + cols + 1, k + cols - 1),               # up-right (odd row) / up-left (even)
              up_ok & np.where(odd, i + 1 < cols, i - 1 >= 0))]
    pairs = []
    for d, (nb, ok) in enumerate(cands):
        m = ok & (_hash01(k * 4 + d, seed) < keep_prob)
        pairs.append(np.stack([k[m], nb[m]], axis=1))
    e = np.concatenate(pairs, axis=0)
    both = np.concatenate([e, e[:, ::-1]], axis=0)
    col_first = both[:, 0] % cols
    both = both[(col_first >= c_lo) & (col_first < c_hi)]
    order = np.lexsort((both[:, 1], both[:, 0]))
    return np.ascontiguousarray(both[order].astype(np.int32))


def config3(n=16_000_000, seed=1234):
    """BASELINE.json configs[2]: hex pitch 2.02r, jitter ±0.01r, ~4 directed bonds per
    particle, 1 % frozen."""
    sc = hex_lattice(n, pitch_over_r=2.02, jitter_over_r=0.01, seed=seed)
    return add_lattice_bonds(sc)


def config2_parity(n=1_000_000, seed=1234):
    return hex_lattice(n, pitch_over_r=2.2, jitter_over_r=0.025, seed=seed)


def config4(n=64_000_000, n_pairs=64, seed=1234):
    sc = hex_lattice(n, pitch_over_r=2.2, jitter_over_r=0.025, seed=seed, vel_scale=0.5)
    return add_portals(sc, n_pairs)


def load_into(obj, scene):
    """Load `scene` into anything exposing the common loader methods
    (ptoy_b200.Particles, oracle RefParticles / PortParticles)."""
    obj.reset_scene(scene.domain)
    obj.set_gravity(scene.gravity[0], scene.gravity[1])
    obj.add_particles(scene.pos, scene.vel, scene.ids)
    for p in scene.portals:
        obj.add_portal_pair(*p)
    if len(scene.bonds):
        obj.set_bonds(scene.bonds)
    if len(scene.frozen):
        obj.set_frozen(scene.frozen)
    return obj

/* TEST INFRASTRUCTURE — not product code.
 *
 * Plain-C restatement of the reference's particle-stepping path
 * (/root/reference/src/particles.cpp + blocks.h + geometry.h + particles.h),
 * written from the algorithm, with every function citing the reference lines it
 * follows.  Arithmetic is the reference's SCALAR build (USE_AVX=0, no FMA):
 * every float operation below is a separate IEEE-754 binary32 operation in the
 * reference's order; expressions the reference evaluates in double (because a
 * double literal takes part) are evaluated in double here and rounded where the
 * reference rounds.  Build with -ffp-contract=off.
 *
 * PINNING: this file is checked bit-for-bit (positions, velocities, forces,
 * per-block id ORDER, bond set, portal block lists, time) against the compiled
 * unmodified reference (oracle/_ref/libptoy_ref.so) by tests/test_oracle.py and
 * against the golden fixtures in tests/golden/ generated from that same
 * reference (scripts/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y; } v2;

/* ---- geometry.h:18-90 (Vect) ------------------------------------------------ */
static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vmul(v2 a, float k) { return V(a.x * k, a.y * k); }
/* geometry.h:62-64  dot: v.x*x + v.y*y */
static inline float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
/* geometry.h:65-67  length: ::sqrt(double) of a float sum, returned as float */
static inline float vlen(v2 a) { return (float)sqrt((double)(a.x * a.x + a.y * a.y)); }
/* geometry.h:68-70 */
static inline float vdist(v2 a, v2 b) { return vlen(vsub(b, a)); }
/* geometry.h:74-77 */
static inline v2 vnormalized(v2 a) { float l = vlen(a); return V(a.x / l, a.y / l); }
/* std::max(a,b) = (a<b)?b:a ; std::min(a,b) = (b<a)?b:a  (NaN behaviour kept) */
static inline float maxf_std(float a, float b) { return (a < b) ? b : a; }
static inline float minf_std(float a, float b) { return (b < a) ? b : a; }
static inline double maxd_std(double a, double b) { return (a < b) ? b : a; }

/* ---- constants, particles.cpp:7-24 ------------------------------------------ */
static const float kRadius = 0.02f;            /* (float)0.02 */
static const float kSigmaBond = 1e5f;
static const float kSigmaPick = 1e3f;
static const float kPointForce = 0.1f;
static const float kPointForceAttractive = 0.1f;
static const float kDissipation = 0.001f;
static const float kGravity = 10.f;
static const float kPortalThickness = 0.02f;
static const float kVelocityLimit = 10.f;
static const float kTimeStep = 0.0005f;
#define K_MASS (kRadius * kRadius * 100.f)              /* :14  float*float*int */
#define K_BLOCK_SIZE ((float)(4. * (double)kRadius))    /* :18  double literal */
#define K_RADIUS_PORTAL_EDGE (3.f * kRadius)            /* :13  int*float */
#define KBLOCK_NONE ((size_t)-1)                        /* blocks.h:16 */

/* ---- blocks.h:19-118 BlockData: one growable SoA per block ------------------ */
typedef struct {
  v2 *pos, *pos_tmp, *vel, *vel_tmp, *force;
  int* id;
  size_t n, cap;
} Block;

typedef struct { size_t block, slot; } BlockSlot;

typedef struct {
  v2 A, B, bs;            /* blocks::domain_, block_size_ */
  int di, dj;             /* dims_ */
  size_t nblocks;
  int noff[9];            /* neighbor_offsets_ */
  Block* blk;
  BlockSlot* bbi;         /* block_by_id_ */
  size_t bbi_n, bbi_cap;
  size_t num_particles, num_per_cell;
} Grid;

typedef struct { v2 begin, end; size_t* blocks; size_t nblocks; } Portal;
typedef struct { Portal p[2]; } PortalPair;
typedef struct { v2 A, B; } Wall;  /* particles.h:34-58 class line */
typedef struct { int a, b; } Bond;

typedef struct {
  v2 domA, domB;          /* Particles::domain */
  v2 rqA, rqB;            /* resize_queue_ */
  Grid g;                 /* Blocks */
  float t, dt;
  v2 gravity;
  int gravity_enable;
  Wall* walls; size_t nwalls;            /* ENVOBJ */
  unsigned char* block_env;              /* block_envobj_: bit k = wall k listed */
  v2 force_center; int force_enabled, force_attractive;
  int pick_enabled, pick_id; v2 pick_pointer;
  Bond* bonds; size_t nbonds, bonds_cap; /* std::set<pair<int,int>> in set order */
  int* frozen; size_t nfrozen, frozen_cap; /* std::set<int> in set order */
  PortalPair* portals; size_t nportals, portals_cap;
  unsigned char* to_move; size_t to_move_n;   /* particle_to_move_ */
  int remove_last_portal;
  int renderer_ready;
  Bond* norend; size_t nnorend, norend_cap;   /* no_rendering_ (sorted unique) */
  Bond* norend_buf; size_t nnorend_buf;
  /* render snapshot: blocks_buffer_ flattened in block order */
  v2 *snap_pos, *snap_vel; int* snap_id; size_t* snap_block_start;
  size_t snap_n, snap_nblocks, snap_num_particles, snap_num_per_cell;
  /* UI tool state */
  int bonds_prev_id, bonds_enabled, freeze_enabled, freeze_last_id, portal_enabled;
  int portal_stage, portal_mouse_moving;
  v2 portal_begin, portal_current, portal_prev_first, portal_prev_second;
} Oracle;

/* ============================ grid (blocks.h) ================================ */

static void block_reserve(Block* b, size_t cap) {
  if (cap <= b->cap) return;
  size_t nc = b->cap ? b->cap * 2 : 16;  /* blocks.h:49 kBlockPadding = 16 */
  while (nc < cap) nc *= 2;
  b->pos = (v2*)realloc(b->pos, nc * sizeof(v2));
  b->pos_tmp = (v2*)realloc(b->pos_tmp, nc * sizeof(v2));
  b->vel = (v2*)realloc(b->vel, nc * sizeof(v2));
  b->vel_tmp = (v2*)realloc(b->vel_tmp, nc * sizeof(v2));
  b->force = (v2*)realloc(b->force, nc * sizeof(v2));
  b->id = (int*)realloc(b->id, nc * sizeof(int));
  b->cap = nc;
}
static void block_free(Block* b) {
  free(b->pos); free(b->pos_tmp); free(b->vel); free(b->vel_tmp);
  free(b->force); free(b->id);
  memset(b, 0, sizeof(*b));
}
static void grid_free_blocks(Grid* g) {
  if (!g->blk) return;
  for (size_t i = 0; i < g->nblocks; ++i) block_free(&g->blk[i]);
  free(g->blk);
  g->blk = NULL;
}
static void bbi_resize(Grid* g, size_t n) {
  if (n > g->bbi_cap) {
    size_t nc = g->bbi_cap ? g->bbi_cap : 64;
    while (nc < n) nc *= 2;
    g->bbi = (BlockSlot*)realloc(g->bbi, nc * sizeof(BlockSlot));
    g->bbi_cap = nc;
  }
  /* std::vector::resize value-initialises new pairs to {0,0} */
  for (size_t i = g->bbi_n; i < n; ++i) { g->bbi[i].block = 0; g->bbi[i].slot = 0; }
  g->bbi_n = n;
}

/* float → int conversion as the x86-64 scalar build does it (cvttss2si):
 * truncation toward zero, INT_MIN for NaN / out of range. */
static inline int f2i_trunc(float f) {
  if (!(f > -2147483904.f && f < 2147483648.f)) return INT32_MIN;
  return (int)f;
}

/* blocks.h:256-276 InitEmptyBlocks */
static void grid_init_empty(Grid* g, v2 A, v2 B, v2 bs) {
  grid_free_blocks(g);
  g->A = A; g->B = B; g->bs = bs;
  v2 size = vsub(B, A);                                  /* geometry.h:124 */
  g->di = f2i_trunc(size.x / bs.x) + 1;                  /* :259 */
  g->dj = f2i_trunc(size.y / bs.y) + 1;                  /* :260 */
  g->nblocks = (size_t)(g->di * g->dj);                  /* :261 */
  g->blk = (Block*)calloc(g->nblocks, sizeof(Block));    /* :265-266 clear+resize */
  for (size_t i = 0; i < g->bbi_n; ++i) g->bbi[i].block = KBLOCK_NONE; /* :36-38 */
  size_t n = 0;
  for (int i = -1; i <= 1; ++i)                          /* :268-275 */
    for (int j = -1; j <= 1; ++j) g->noff[n++] = i * g->dj + j;
}

/* blocks.h:155-163 FindBlock */
static size_t grid_find_block(const Grid* g, v2 p) {
  int mi = f2i_trunc((p.x - g->A.x) / g->bs.x);
  int mj = f2i_trunc((p.y - g->A.y) / g->bs.y);
  if (mi >= 0 && mi < g->di && mj >= 0 && mj < g->dj)
    return (size_t)(mi * g->dj + mj);
  return KBLOCK_NONE;
}

/* blocks.h:101-117 AddParticle */
static void grid_add_particle(Grid* g, size_t dest, v2 pos, v2 vel, int id) {
  Block* b = &g->blk[dest];
  block_reserve(b, b->n + 1);
  const float nan = NAN;
  b->pos[b->n] = pos;
  b->pos_tmp[b->n] = V(nan, nan);
  b->vel[b->n] = vel;
  b->vel_tmp[b->n] = V(nan, nan);
  b->force[b->n] = V(nan, nan);
  b->id[b->n] = id;
  b->n++;
  size_t pid = (size_t)id;
  if (g->bbi_n <= pid) bbi_resize(g, pid + 1);
  g->bbi[pid].block = dest;
  g->bbi[pid].slot = b->n - 1;
}

/* blocks.h:63-83 RemoveParticle: swap with the back, pop */
static void grid_remove_particle(Grid* g, size_t src, size_t idx) {
  Block* b = &g->blk[src];
  size_t last = b->n - 1;
#define SWAP_(arr, T) { T tmp_ = b->arr[idx]; b->arr[idx] = b->arr[last]; b->arr[last] = tmp_; }
  SWAP_(pos, v2) SWAP_(pos_tmp, v2) SWAP_(vel, v2) SWAP_(vel_tmp, v2)
  SWAP_(force, v2) SWAP_(id, int)
#undef SWAP_
  g->bbi[b->id[idx]].block = src;     /* :74 */
  g->bbi[b->id[idx]].slot = idx;
  g->bbi[b->id[last]].block = KBLOCK_NONE;  /* :75 */
  b->n--;
}

/* blocks.h:84-100 MoveParticle */
static void grid_move_particle(Grid* g, size_t src, size_t idx, size_t dest) {
  Block* s = &g->blk[src];
  Block* d = &g->blk[dest];
  block_reserve(d, d->n + 1);
  s = &g->blk[src];
  d->pos[d->n] = s->pos[idx];
  d->pos_tmp[d->n] = s->pos_tmp[idx];
  d->vel[d->n] = s->vel[idx];
  d->vel_tmp[d->n] = s->vel_tmp[idx];
  d->force[d->n] = s->force[idx];
  d->id[d->n] = s->id[idx];
  d->n++;
  grid_remove_particle(g, src, idx);
  g->bbi[d->id[d->n - 1]].block = dest;   /* :99 */
  g->bbi[d->id[d->n - 1]].slot = d->n - 1;
}

/* blocks.h:215-244 SortParticles */
static void grid_sort(Grid* g) {
  size_t lnum = 0, max_per_cell = 0;
  for (size_t i = 0; i < g->nblocks; ++i) {
    size_t p = 0, pe = g->blk[i].n;
    while (p < pe) {
      size_t j = grid_find_block(g, g->blk[i].pos[p]);
      if (i != j) {
        if (j != KBLOCK_NONE) {
          grid_move_particle(g, i, p, j);
          if (j < i) ++lnum;
        } else {
          grid_remove_particle(g, i, p);
        }
        --pe;
      } else {
        ++p;
        ++lnum;
      }
    }
    if (pe > max_per_cell) max_per_cell = pe;
  }
  g->num_particles = lnum;
  g->num_per_cell = max_per_cell;
}

/* blocks.h:179-193 AddParticles(position, velocity, id) */
static void grid_add_particles(
    Grid* g, const v2* pos, const v2* vel, const int* id, size_t n) {
  for (size_t p = 0; p < n; ++p) {
    size_t i = grid_find_block(g, pos[p]);
    if (i != KBLOCK_NONE) grid_add_particle(g, i, pos[p], vel[p], id[p]);
  }
  grid_sort(g);
}

/* blocks.h:119-141 SetDomain: grow-only, whole blocks, by accumulation */
static void grid_set_domain(Grid* g, v2 pA, v2 pB) {
  v2 A = g->A, B = g->B;
  while (pA.x < A.x) A.x -= g->bs.x;
  while (pA.y < A.y) A.y -= g->bs.y;
  while (pB.x > B.x) B.x += g->bs.x;
  while (pB.y > B.y) B.y += g->bs.y;
  if (!(A.x == g->A.x && A.y == g->A.y && B.x == g->B.x && B.y == g->B.y)) {
    /* :135-137  copy old data, re-init, re-add block by block, slot by slot */
    Block* old = g->blk;
    size_t nold = g->nblocks;
    g->blk = NULL;
    grid_init_empty(g, A, B, g->bs);
    for (size_t k = 0; k < nold; ++k)                  /* blocks.h:164-178 */
      for (size_t p = 0; p < old[k].n; ++p) {
        size_t i = grid_find_block(g, old[k].pos[p]);
        if (i != KBLOCK_NONE)
          grid_add_particle(g, i, old[k].pos[p], old[k].vel[p], old[k].id[p]);
      }
    grid_sort(g);
    for (size_t k = 0; k < nold; ++k) block_free(&old[k]);
    free(old);
  }
}

/* blocks.h:142-144: block_size_.length() * 0.5 (double product, float result) */
static float grid_circum_radius(const Grid* g) {
  return (float)((double)vlen(g->bs) * 0.5);
}
/* blocks.h:146-150 GetCenter: (0.5 + m) * bs in double, rounded to float, + A */
static v2 grid_center(const Grid* g, size_t block) {
  int mi = (int)(block / (size_t)g->dj), mj = (int)(block % (size_t)g->dj);
  v2 off = V((float)((0.5 + mi) * (double)g->bs.x), (float)((0.5 + mj) * (double)g->bs.y));
  return vadd(g->A, off);
}

/* ============================ force laws ===================================== */

/* particles.cpp:584-595 F12(p1,p2): pair repulsion, R = 2r, r² thresholded */
static v2 F12(v2 p1, v2 p2) {
  const float threshold = (float)(pow((double)kRadius, 2) * 1e-3);
  const float R = (float)(2. * (double)kRadius);
  const v2 dp = vsub(p1, p2);
  const float r2 = maxf_std(threshold, vdot(dp, dp));
  const float r2inv = (float)(1. / (double)r2);
  const float d2 = r2inv * (R * R);
  const float d6 = d2 * d2 * d2;
  const float d12 = d6 * d6;
  return vmul(dp, maxf_std(0.f, 1.f * (d12 - d6) * r2inv));
}
/* particles.cpp:572-582 F12wall: R = r, no threshold */
static v2 F12wall(v2 p1, v2 p2) {
  const float R = kRadius;
  const v2 dp = vsub(p1, p2);
  const float r2 = vdot(dp, dp);
  const float r2inv = (float)(1. / (double)r2);
  const float d2 = r2inv * (R * R);
  const float d6 = d2 * d2 * d2;
  const float d12 = d6 * d6;
  return vmul(dp, maxf_std(0.f, 1.f * (d12 - d6) * r2inv));
}
/* particles.cpp:738-749 F12_portal_edge: R = 3r, no threshold */
static v2 F12_portal_edge(v2 p1, v2 p2) {
  const float R = K_RADIUS_PORTAL_EDGE;
  const v2 dp = vsub(p1, p2);
  const float r2 = vdot(dp, dp);
  const float r2inv = (float)(1. / (double)r2);
  const float d2 = r2inv * (R * R);
  const float d6 = d2 * d2 * d2;
  const float d12 = d6 * d6;
  const float k = maxf_std(0.f, 1.f * (d12 - d6) * r2inv);
  return vmul(dp, k);
}
/* particles.cpp:727-736 F12_bond: spring, clamped at |dp| >= 5R */
static v2 F12_bond(v2 p1, v2 p2) {
  const float R = (float)(2. * (double)kRadius);
  const v2 dp = vsub(p1, p2);
  const float r = vlen(dp);
  const float kLimit = 5.f;
  const float k = (float)maxd_std(1. - (double)(r / R), 1. - (double)kLimit);
  return vmul(dp, kSigmaBond * k);
}
/* particles.cpp:751-759 F12_pick */
static v2 F12_pick(v2 p1, v2 p2) {
  const float R = (float)(0.1 * (double)kRadius);
  const v2 dp = vsub(p1, p2);
  const float r = vlen(dp);
  const float k = (R - r) / R;
  return vmul(dp, kSigmaPick * k);
}

/* particles.h:37-46 line::GetNearest (also Portal::GetNearest, :67-77) */
static v2 seg_nearest(v2 A, v2 B, v2 p) {
  v2 ab = vsub(B, A);
  float lambda = vdot(ab, vsub(p, A)) / vdot(ab, ab);
  if (lambda > 0. && lambda < 1.) return vadd(A, vmul(ab, lambda));
  return (vdist(p, A) < vdist(p, B)) ? A : B;
}

/* ============================ Particles ====================================== */

static void portal_free(Portal* p) { free(p->blocks); p->blocks = NULL; p->nblocks = 0; }

/* particles.cpp:699-706 UpdatePortalBlocks; Portal::IsClose particles.h:78-80 */
static void update_portal_blocks(Oracle* o, Portal* portal) {
  portal_free(portal);
  const float R = grid_circum_radius(&o->g);
  size_t cap = 0;
  for (size_t ib = 0; ib < o->g.nblocks; ++ib) {
    v2 c = grid_center(&o->g, ib);
    if (vdist(seg_nearest(portal->begin, portal->end, c), c) < R + 2 * kPortalThickness) {
      if (portal->nblocks == cap) {
        cap = cap ? cap * 2 : 32;
        portal->blocks = (size_t*)realloc(portal->blocks, cap * sizeof(size_t));
      }
      portal->blocks[portal->nblocks++] = ib;
    }
  }
}

/* particles.cpp:708-725 UpdateEnvObj; line::IsClose particles.h:55-57 */
static void update_env_obj(Oracle* o) {
  free(o->block_env);
  o->block_env = (unsigned char*)calloc(o->g.nblocks, 1);
  const float R = grid_circum_radius(&o->g);
  for (size_t ib = 0; ib < o->g.nblocks; ++ib) {
    v2 c = grid_center(&o->g, ib);
    for (size_t k = 0; k < o->nwalls; ++k)
      if (vdist(seg_nearest(o->walls[k].A, o->walls[k].B, c), c) < R + kRadius)
        o->block_env[ib] |= (unsigned char)(1u << k);
  }
  for (size_t i = 0; i < o->nportals; ++i)
    for (int d = 0; d <= 1; ++d) update_portal_blocks(o, &o->portals[i].p[d]);
}

/* particles.h:99-107 ResetEnvObjFrame: bottom, top, left, right */
static void reset_env_frame(Oracle* o, v2 A, v2 B) {
  free(o->walls);
  o->nwalls = 4;
  o->walls = (Wall*)malloc(4 * sizeof(Wall));
  o->walls[0].A = V(A.x, A.y); o->walls[0].B = V(B.x, A.y);
  o->walls[1].A = V(A.x, B.y); o->walls[1].B = V(B.x, B.y);
  o->walls[2].A = V(A.x, A.y); o->walls[2].B = V(A.x, B.y);
  o->walls[3].A = V(B.x, A.y); o->walls[3].B = V(B.x, B.y);
  update_env_obj(o);
}

/* particles.h:108-111 Particles::SetDomain */
static void set_domain(Oracle* o, v2 A, v2 B) {
  o->domA = A; o->domB = B;
  grid_set_domain(&o->g, A, B);
}

/* particles.cpp:79-89 SetParticleBuffer: deep copy of Blocks, flattened */
static void set_particle_buffer(Oracle* o) {
  Grid* g = &o->g;
  size_t n = 0;
  for (size_t b = 0; b < g->nblocks; ++b) n += g->blk[b].n;
  o->snap_pos = (v2*)realloc(o->snap_pos, (n + 1) * sizeof(v2));
  o->snap_vel = (v2*)realloc(o->snap_vel, (n + 1) * sizeof(v2));
  o->snap_id = (int*)realloc(o->snap_id, (n + 1) * sizeof(int));
  o->snap_block_start = (size_t*)realloc(o->snap_block_start, (g->nblocks + 1) * sizeof(size_t));
  size_t k = 0;
  for (size_t b = 0; b < g->nblocks; ++b) {
    o->snap_block_start[b] = k;
    for (size_t p = 0; p < g->blk[b].n; ++p) {
      o->snap_pos[k] = g->blk[b].pos[p];
      o->snap_vel[k] = g->blk[b].vel[p];
      o->snap_id[k] = g->blk[b].id[p];
      ++k;
    }
  }
  o->snap_block_start[g->nblocks] = k;
  o->snap_n = k;
  o->snap_nblocks = g->nblocks;
  o->snap_num_particles = g->num_particles;
  o->snap_num_per_cell = g->num_per_cell;
}

/* particles.cpp:839-910 calc_forces(iblock) */
static void calc_forces(Oracle* o, size_t iblock) {
  Grid* g = &o->g;
  Block* b = &g->blk[iblock];
  const float kMass = K_MASS;
  for (size_t p = 0; p < b->n; ++p) {
    v2 f = V(0.f, 0.f);                                    /* :846 */
    const v2 x = b->pos[p], v = b->vel[p];
    if (o->gravity_enable) f = vmul(o->gravity, kMass);    /* :849-851 */
    if (o->force_enabled) {                                /* :854-863 */
      const v2 r = vsub(x, o->force_center);
      if (vlen(r) > kRadius) {
        if (o->force_attractive)
          f = vadd(f, vmul(r, (float)((double)(-kPointForceAttractive) / pow((double)vlen(r), 3))));
        else
          f = vadd(f, vmul(r, (float)((double)kPointForce / pow((double)vlen(r), 4))));
      }
    }
    for (size_t i = 0; i < o->nportals; ++i)               /* :866-871 */
      for (int d = 0; d < 2; ++d) {
        f = vadd(f, F12_portal_edge(x, o->portals[i].p[d].begin));
        f = vadd(f, F12_portal_edge(x, o->portals[i].p[d].end));
      }
    if (o->pick_enabled && b->id[p] == o->pick_id)         /* :874-876 */
      f = vadd(f, F12_pick(x, o->pick_pointer));
    f = vsub(f, vmul(v, kDissipation * kMass));            /* :880 */
    b->force[p] = f;
  }
  for (int n = 0; n < 9; ++n) {                            /* :884-900 */
    const size_t j = iblock + (size_t)(int64_t)g->noff[n]; /* size_t wrap */
    if (j >= g->nblocks) continue;                         /* :887 */
    const Block* other = &g->blk[j];
    /* :598-606 CalcForceSerial: q outer, p inner, self skipped by address */
    for (size_t q = 0; q < other->n; ++q)
      for (size_t p = 0; p < b->n; ++p)
        if (&b->pos[p] != &other->pos[q])
          b->force[p] = vadd(b->force[p], F12(b->pos[p], other->pos[q]));
  }
  for (size_t k = 0; k < o->nwalls; ++k) {                 /* :903-909 */
    if (!(o->block_env[iblock] & (1u << k))) continue;
    for (size_t p = 0; p < b->n; ++p) {
      v2 q = seg_nearest(o->walls[k].A, o->walls[k].B, b->pos[p]);
      b->force[p] = vadd(b->force[p], F12wall(b->pos[p], q));   /* particles.h:52-54 */
    }
  }
}

static void norend_insert(Oracle* o, int a, int b) {
  /* std::set::emplace: sorted unique */
  size_t lo = 0, hi = o->nnorend;
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    Bond m = o->norend[mid];
    if (m.a < a || (m.a == a && m.b < b)) lo = mid + 1; else hi = mid;
  }
  if (lo < o->nnorend && o->norend[lo].a == a && o->norend[lo].b == b) return;
  if (o->nnorend == o->norend_cap) {
    o->norend_cap = o->norend_cap ? o->norend_cap * 2 : 16;
    o->norend = (Bond*)realloc(o->norend, o->norend_cap * sizeof(Bond));
  }
  memmove(&o->norend[lo + 1], &o->norend[lo], (o->nnorend - lo) * sizeof(Bond));
  o->norend[lo].a = a; o->norend[lo].b = b;
  o->nnorend++;
}

/* particles.cpp:761-826 RHS_bonds */
static void rhs_bonds(Oracle* o) {
  Grid* g = &o->g;
  o->nnorend = 0;
  const float kCutoff = (float)(8. * (double)kRadius);     /* :777 */
  for (size_t e = 0; e < o->nbonds; ++e) {
    const BlockSlot p = g->bbi[o->bonds[e].a], q = g->bbi[o->bonds[e].b];
    const v2 p_pos = g->blk[p.block].pos[p.slot];
    const v2 q_pos = g->blk[q.block].pos[q.slot];
    v2 res;
    if (vdist(p_pos, q_pos) < kCutoff) {                   /* :779 */
      res = F12_bond(p_pos, q_pos);
    } else {
      int found = 0;
      for (size_t i = 0; i < o->nportals; ++i)             /* :784-819 */
        for (int d = 0; d <= 1; ++d) {
          const Portal* portal = &o->portals[i].p[d];
          const Portal* other = &o->portals[i].p[1 - d];
          const v2 a = portal->begin, b = portal->end;
          const v2 oa = other->begin, ob = other->end;
          const v2 r = vsub(b, a);
          const v2 n = vnormalized(V(-r.y, r.x));
          const v2 orr = vsub(ob, oa);
          const v2 on = vnormalized(V(-orr.y, orr.x));
          const float p_lambda = vdot(r, vsub(p_pos, a)) / vdot(r, r);
          const float p_offset = vdot(vsub(p_pos, a), n);
          const float q_lambda_other = vdot(orr, vsub(q_pos, oa)) / vdot(orr, orr);
          const float q_offset_other = vdot(vsub(q_pos, oa), on);
          if (p_lambda > 0. && p_lambda < 1. && q_lambda_other > 0. &&
              q_lambda_other < 1. && p_offset * q_offset_other < 0. &&
              (double)fabsf(p_offset - q_offset_other) <
                  2. * (double)kPortalThickness + (double)kCutoff) {
            const float sign = (p_offset > 0. ? 1.f : -1.f);
            const v2 proj = vadd(
                vadd(a, vmul(r, q_lambda_other)),
                vmul(n, (float)((double)q_offset_other +
                                2. * (double)sign * (double)kPortalThickness)));
            res = F12_bond(p_pos, proj);
            found = 1;
            norend_insert(o, o->bonds[e].a, o->bonds[e].b);
            norend_insert(o, o->bonds[e].b, o->bonds[e].a);
          }
        }
      if (!found) res = F12_bond(p_pos, q_pos);
    }
    g->blk[p.block].force[p.slot] = vadd(g->blk[p.block].force[p.slot], res);  /* :824 */
  }
}

/* particles.cpp:318-381 ApplyPortalsForces */
static void apply_portals_forces(Oracle* o) {
  Grid* g = &o->g;
  for (size_t i = 0; i < o->nportals; ++i)
    for (int d = 0; d <= 1; ++d) {
      const Portal* portal = &o->portals[i].p[d];
      const Portal* other = &o->portals[i].p[1 - d];
      const v2 a = portal->begin, b = portal->end;
      const v2 oa = other->begin, ob = other->end;
      const v2 r = vsub(b, a);
      const v2 n = vnormalized(V(-r.y, r.x));
      const v2 orr = vsub(ob, oa);
      const v2 on = vnormalized(V(-orr.y, orr.x));
      for (size_t bi = 0; bi < portal->nblocks; ++bi) {
        Block* B = &g->blk[portal->blocks[bi]];
        for (size_t p = 0; p < B->n; ++p) {
          const v2 curr = B->pos[p];
          const float lambda_curr = vdot(r, vsub(curr, a)) / vdot(r, r);
          const float offset_curr = vdot(vsub(curr, a), n);
          if (lambda_curr > 0. && lambda_curr < 1.) {            /* :344-356 */
            for (size_t bj = 0; bj < portal->nblocks; ++bj) {
              const Block* J = &g->blk[portal->blocks[bj]];
              for (size_t q = 0; q < J->n; ++q) {
                const v2 nb = J->pos[q];
                const float ln = vdot(r, vsub(nb, a)) / vdot(r, r);
                const float on_ = vdot(vsub(nb, a), n);
                if (ln > 0. && ln < 1. && on_ * offset_curr < 0.)
                  B->force[p] = vsub(B->force[p], F12(curr, nb));
              }
            }
          }
          if (lambda_curr > 0. && lambda_curr < 1.) {            /* :358-376 */
            for (size_t bj = 0; bj < other->nblocks; ++bj) {
              const Block* J = &g->blk[other->blocks[bj]];
              for (size_t q = 0; q < J->n; ++q) {
                const v2 nb = J->pos[q];
                const float oln = vdot(orr, vsub(nb, oa)) / vdot(orr, orr);
                const float oon = vdot(vsub(nb, oa), on);
                if (oln > 0. && oln < 1. && oon * offset_curr < 0.) {
                  const float sign = (offset_curr > 0. ? 1.f : -1.f);
                  const v2 proj = vadd(
                      vadd(a, vmul(r, oln)),
                      vmul(n, (float)((double)oon +
                                      2. * (double)sign * (double)kPortalThickness)));
                  B->force[p] = vadd(B->force[p], F12(curr, proj));
                }
              }
            }
          }
        }
      }
    }
}

/* particles.cpp:828-837 ApplyFrozen */
static void apply_frozen(Oracle* o) {
  Grid* g = &o->g;
  for (size_t k = 0; k < o->nfrozen; ++k) {
    const BlockSlot s = g->bbi[o->frozen[k]];
    Block* b = &g->blk[s.block];
    b->vel[s.slot] = vmul(b->vel[s.slot], 0.f);
    b->force[s.slot] = vmul(b->force[s.slot], 0.f);
  }
}

static void to_move_grow(Oracle* o, size_t id) {
  if (o->to_move_n <= id) {
    o->to_move = (unsigned char*)realloc(o->to_move, id + 1);
    memset(o->to_move + o->to_move_n, 0, id + 1 - o->to_move_n);
    o->to_move_n = id + 1;
  }
}

/* particles.cpp:288-316 DetectPortals */
static void detect_portals(Oracle* o) {
  Grid* g = &o->g;
  for (size_t i = 0; i < o->nportals; ++i)
    for (int d = 0; d <= 1; ++d) {
      const Portal* portal = &o->portals[i].p[d];
      const v2 a = portal->begin, b = portal->end;
      const v2 r = vsub(b, a);
      const v2 n = vnormalized(V(-r.y, r.x));
      for (size_t bi = 0; bi < portal->nblocks; ++bi) {
        const Block* B = &g->blk[portal->blocks[bi]];
        for (size_t p = 0; p < B->n; ++p) {
          const size_t id = (size_t)B->id[p];
          to_move_grow(o, id);
          const v2 curr = B->pos[p];
          const float lambda_curr = vdot(vsub(curr, a), r) / vdot(r, r);
          const float offset_curr = vdot(vsub(curr, a), n);
          if (lambda_curr > 0. && lambda_curr < 1. &&
              fabsf(offset_curr) <= kPortalThickness)
            o->to_move[id] = 1;
        }
      }
    }
}

/* particles.cpp:264-286 MoveToPortal */
static void move_to_portal(v2* position, v2* velocity, const Portal* src, const Portal* dest) {
  const v2 src_a = src->begin, src_r = vsub(src->end, src->begin);
  const v2 src_n = vnormalized(V(-src_r.y, src_r.x));
  const v2 dest_a = dest->begin, dest_r = vsub(dest->end, dest->begin);
  const v2 dest_n = vnormalized(V(-dest_r.y, dest_r.x));
  const float lambda_pos = vdot(src_r, vsub(*position, src_a)) / vdot(src_r, src_r);
  const float offset_pos = vdot(src_n, vsub(*position, src_a));
  const float lambda_vel = vdot(src_r, *velocity) / vdot(src_r, src_r);
  const float offset_vel = vdot(src_n, *velocity);
  const float sign = (offset_pos > 0. ? 1.f : -1.f);
  *position = vadd(
      vadd(dest_a, vmul(dest_r, lambda_pos)),
      vmul(dest_n, (float)((double)offset_pos -
                           (double)sign * 2. * (double)kPortalThickness)));
  *velocity = vadd(vmul(dest_r, lambda_vel), vmul(dest_n, offset_vel));
}

/* particles.cpp:382-409 ApplyPortals */
static void apply_portals(Oracle* o) {
  Grid* g = &o->g;
  for (size_t i = 0; i < o->nportals; ++i)
    for (int d = 0; d <= 1; ++d) {
      const Portal* portal = &o->portals[i].p[d];
      const Portal* other = &o->portals[i].p[1 - d];
      for (size_t bi = 0; bi < portal->nblocks; ++bi) {
        Block* B = &g->blk[portal->blocks[bi]];
        for (size_t p = 0; p < B->n; ++p) {
          const size_t id = (size_t)B->id[p];
          to_move_grow(o, id);
          if (o->to_move[id]) {
            move_to_portal(&B->pos[p], &B->vel[p], portal, other);
            o->to_move[id] = 0;
          }
        }
      }
    }
}

/* particles.cpp:205-215 CheckBonds */
static void check_bonds(Oracle* o) {
  const Grid* g = &o->g;
  size_t k = 0;
  for (size_t e = 0; e < o->nbonds; ++e) {
    if (g->bbi[o->bonds[e].a].block == KBLOCK_NONE ||
        g->bbi[o->bonds[e].b].block == KBLOCK_NONE)
      continue;
    o->bonds[k++] = o->bonds[e];
  }
  o->nbonds = k;
}

/* particles.cpp:96-201: ONE iteration of the while loop */
static void step_once(Oracle* o) {
  Grid* g = &o->g;
  const float kMass = K_MASS;
  const float dt = o->dt;
  const int64_t nb = (int64_t)g->nblocks;

#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t ib = 0; ib < nb; ++ib) calc_forces(o, (size_t)ib);     /* :97-100 */
  rhs_bonds(o); apply_portals_forces(o); apply_frozen(o);              /* :102-107 */

  const float c1 = (float)((double)dt * 0.5 / (double)kMass);          /* :116 */
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t ib = 0; ib < nb; ++ib) {                                /* :109-123 */
    Block* b = &g->blk[ib];
    for (size_t p = 0; p < b->n; ++p) {
      b->vel_tmp[p] = b->vel[p];
      b->pos_tmp[p] = b->pos[p];
      b->vel[p] = vadd(b->vel[p], vmul(b->force[p], c1));
      if (vlen(b->vel[p]) > kVelocityLimit)
        b->vel[p] = vmul(b->vel[p], kVelocityLimit / vlen(b->vel[p]));
      b->pos[p] = vadd(b->pos[p], vmul(vmul(b->vel[p], dt), 0.5f));     /* :121 */
    }
  }
  o->t = (float)((double)o->t + 0.5 * (double)dt);                     /* :126 */

#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t ib = 0; ib < nb; ++ib) calc_forces(o, (size_t)ib);     /* :128-131 */
  rhs_bonds(o); apply_portals_forces(o); apply_frozen(o);              /* :133-138 */

  const float c2 = (float)((double)dt / (double)kMass);                /* :145 */
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t ib = 0; ib < nb; ++ib) {                                /* :140-153 */
    Block* b = &g->blk[ib];
    for (size_t p = 0; p < b->n; ++p) {
      b->vel[p] = vadd(b->vel_tmp[p], vmul(b->force[p], c2));
      if (vlen(b->vel[p]) > kVelocityLimit)
        b->vel[p] = vmul(b->vel[p], kVelocityLimit / vlen(b->vel[p]));
      b->pos[p] = vadd(b->pos_tmp[p], vmul(b->vel[p], dt));
    }
  }

  /* :157-175 rate-limited frame resize */
  {
    const float lim = 0.01f;
    v2 nA = o->domA, nB = o->domB;
#define LIMIT_(cur, target) cur = minf_std(cur + lim, maxf_std(cur - lim, target))
    LIMIT_(nA.x, o->rqA.x); LIMIT_(nA.y, o->rqA.y);
    LIMIT_(nB.x, o->rqB.x); LIMIT_(nB.y, o->rqB.y);
#undef LIMIT_
    if (f2i_trunc(o->t / dt) % (int)(0.01 / (double)dt) == 0) {         /* :170 */
      if (!(nA.x == o->domA.x && nA.y == o->domA.y && nB.x == o->domB.x && nB.y == o->domB.y)) {
        set_domain(o, nA, nB);
        reset_env_frame(o, nA, nB);
      }
    }
  }
  detect_portals(o);                                                   /* :177 */
  apply_portals(o);                                                    /* :178 */
  grid_sort(g);                                                        /* :179 */
  check_bonds(o);                                                      /* :180 */
  if (o->remove_last_portal) {                                         /* :182-187 */
    if (o->nportals) {
      o->nportals--;
      portal_free(&o->portals[o->nportals].p[0]);
      portal_free(&o->portals[o->nportals].p[1]);
    }
    o->remove_last_portal = 0;
  }
  if (o->renderer_ready) {                                             /* :190-196 */
    set_particle_buffer(o);
    rhs_bonds(o);
    o->norend_buf = (Bond*)realloc(o->norend_buf, (o->nnorend + 1) * sizeof(Bond));
    memcpy(o->norend_buf, o->norend, o->nnorend * sizeof(Bond));
    o->nnorend_buf = o->nnorend;
    o->renderer_ready = 0;
  }
  o->t = (float)((double)o->t + 0.5 * (double)dt);                     /* :199 */
}

/* particles.cpp:411-450 PortalStart/Move/Stop */
static void portal_start(Oracle* o, v2 point) {
  o->portal_enabled = 1;
  o->portal_begin = point;
  o->portal_current = point;
  o->portal_mouse_moving = 1;
}
static void portal_move(Oracle* o, v2 point) {
  if (!o->portal_enabled) return;
  o->portal_current = point;
}
static void portal_stop(Oracle* o, v2 point) {
  if (!o->portal_enabled) return;
  o->portal_enabled = 0;
  o->portal_mouse_moving = 0;
  if (o->portal_stage == 0) {
    o->portal_prev_first = o->portal_begin;
    o->portal_prev_second = point;
    o->portal_stage = 1;
  } else {
    if (o->nportals == o->portals_cap) {
      o->portals_cap = o->portals_cap ? o->portals_cap * 2 : 4;
      o->portals = (PortalPair*)realloc(o->portals, o->portals_cap * sizeof(PortalPair));
    }
    PortalPair* pair = &o->portals[o->nportals++];
    memset(pair, 0, sizeof(*pair));
    pair->p[0].begin = o->portal_prev_first;
    pair->p[0].end = o->portal_prev_second;
    pair->p[1].begin = o->portal_begin;
    /* :441-443 second portal gets the first one's length */
    pair->p[1].end = vadd(
        pair->p[1].begin,
        vmul(vnormalized(vsub(point, pair->p[1].begin)),
             vdist(pair->p[0].begin, pair->p[0].end)));
    o->portal_stage = 0;
    update_portal_blocks(o, &pair->p[0]);
    update_portal_blocks(o, &pair->p[1]);
  }
}

static int bond_find(const Oracle* o, int a, int b, size_t* where) {
  size_t lo = 0, hi = o->nbonds;
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    Bond m = o->bonds[mid];
    if (m.a < a || (m.a == a && m.b < b)) lo = mid + 1; else hi = mid;
  }
  *where = lo;
  return lo < o->nbonds && o->bonds[lo].a == a && o->bonds[lo].b == b;
}
static void bond_insert(Oracle* o, int a, int b) {
  size_t w;
  if (bond_find(o, a, b, &w)) return;
  if (o->nbonds == o->bonds_cap) {
    o->bonds_cap = o->bonds_cap ? o->bonds_cap * 2 : 16;
    o->bonds = (Bond*)realloc(o->bonds, o->bonds_cap * sizeof(Bond));
  }
  memmove(&o->bonds[w + 1], &o->bonds[w], (o->nbonds - w) * sizeof(Bond));
  o->bonds[w].a = a; o->bonds[w].b = b;
  o->nbonds++;
}
static void bond_erase(Oracle* o, int a, int b) {
  size_t w;
  if (!bond_find(o, a, b, &w)) return;
  memmove(&o->bonds[w], &o->bonds[w + 1], (o->nbonds - w - 1) * sizeof(Bond));
  o->nbonds--;
}

/* particles.cpp:452-508 Bonds tool (scans the render snapshot) */
static void bonds_start(Oracle* o, v2 point) {
  o->bonds_enabled = 1;
  int id = -1;
  for (size_t k = 0; k < o->snap_n; ++k)
    if (vdist(o->snap_pos[k], point) < kRadius) id = o->snap_id[k];
  o->bonds_prev_id = id;
}
static void bonds_move(Oracle* o, v2 point) {
  if (!o->bonds_enabled) return;
  int id = -1;
  for (size_t k = 0; k < o->snap_n; ++k)
    if (vdist(o->snap_pos[k], point) < kRadius && o->snap_id[k] != o->bonds_prev_id)
      id = o->snap_id[k];
  if (o->bonds_prev_id != -1 && id != -1) {
    size_t w;
    if (bond_find(o, o->bonds_prev_id, id, &w)) {
      bond_erase(o, o->bonds_prev_id, id);
      bond_erase(o, id, o->bonds_prev_id);
    } else {
      bond_insert(o, o->bonds_prev_id, id);
      bond_insert(o, id, o->bonds_prev_id);
    }
  }
  if (id != -1) o->bonds_prev_id = id;
}
/* particles.cpp:228-262 Pick tool */
static void pick_start(Oracle* o, v2 point) {
  int have = 0;
  size_t best = 0;
  float min_dist = 0.f;
  for (size_t k = 0; k < o->snap_n; ++k)
    if (!have || vdist(o->snap_pos[k], point) < min_dist) {
      have = 1; best = k; min_dist = vdist(o->snap_pos[k], point);
    }
  if (!have) return;  /* the reference would index data.id[kBlockNone] (UB) */
  o->pick_id = o->snap_id[best];
  o->pick_pointer = point;
  o->pick_enabled = 1;
}
/* particles.cpp:510-560 Freeze tool */
static void freeze_move(Oracle* o, v2 mousepos) {
  if (!o->freeze_enabled) return;
  const float kFreezeRadius = kRadius * 3;
  float min_dist = kFreezeRadius;
  size_t best = 0;
  for (size_t k = 0; k < o->snap_n; ++k) {
    const float dist = vdist(o->snap_pos[k], mousepos);
    if (dist < min_dist) { min_dist = dist; best = k; }
  }
  if (min_dist < kFreezeRadius) {
    const int id = o->snap_id[best];
    if (id != o->freeze_last_id) {
      size_t lo = 0, hi = o->nfrozen;
      while (lo < hi) { size_t m = (lo + hi) / 2; if (o->frozen[m] < id) lo = m + 1; else hi = m; }
      if (lo < o->nfrozen && o->frozen[lo] == id) {
        memmove(&o->frozen[lo], &o->frozen[lo + 1], (o->nfrozen - lo - 1) * sizeof(int));
        o->nfrozen--;
      } else {
        if (o->nfrozen == o->frozen_cap) {
          o->frozen_cap = o->frozen_cap ? o->frozen_cap * 2 : 16;
          o->frozen = (int*)realloc(o->frozen, o->frozen_cap * sizeof(int));
        }
        memmove(&o->frozen[lo + 1], &o->frozen[lo], (o->nfrozen - lo) * sizeof(int));
        o->frozen[lo] = id;
        o->nfrozen++;
      }
      o->freeze_last_id = id;
    }
  }
}

/* ============================ C API (mirrors oracle/ref_wrap.cpp) ============ */

/* particles.cpp:26-77 Particles::Particles(): the default scene */
void* orc_create_default(void) {
  Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
  o->domA = V(-1.f, -1.f); o->domB = V(1.f, 1.f);
  const float bsz = K_BLOCK_SIZE;
  grid_init_empty(&o->g, o->domA, o->domB, V(bsz, bsz));
  o->gravity_enable = 1;
  o->renderer_ready = 1;
  const size_t rows = 25, columns = 25;
  const float width = (float)((double)columns * 2. * (double)kRadius);
  const float height = (float)((double)rows * sqrt(3.) * (double)kRadius);
  const v2 boxA = V((float)(-0.5 * (double)width), (float)-1.);
  const v2 boxB = V((float)(0.5 * (double)width), (float)(-1. + (double)height));
  const size_t n = rows * columns;
  v2* pos = (v2*)malloc(n * sizeof(v2));
  v2* vel = (v2*)malloc(n * sizeof(v2));
  int* id = (int*)malloc(n * sizeof(int));
  size_t k = 0;
  for (size_t j = 0; j < rows; ++j)
    for (size_t i = 0; i < columns; ++i) {
      const float x = (float)((double)boxA.x + (double)kRadius * (2. * (double)i + 1. + (double)(j % 2)));
      const float y = (float)((double)boxA.y + (double)kRadius * (sqrt(3.) * (double)j + 1.));
      pos[k] = V(x, y); vel[k] = V(0.f, 0.f); id[k] = (int)k; ++k;
    }
  grid_add_particles(&o->g, pos, vel, id, n);
  free(pos); free(vel); free(id);
  o->t = 0.f;
  o->dt = kTimeStep;
  o->gravity = vmul(V(0.f, -1.f), kGravity);
  set_domain(o, o->domA, o->domB);
  set_particle_buffer(o);
  o->rqA = o->domA; o->rqB = o->domB;
  reset_env_frame(o, o->domA, o->domB);
  const float dx = kPortalThickness;
  portal_start(o, V(boxA.x - dx, boxA.y));
  portal_stop(o, V(boxA.x - dx, (float)((double)boxB.y + 0.2)));
  portal_start(o, V(boxB.x + dx, boxA.y));
  portal_stop(o, V(boxB.x + dx, (float)((double)boxB.y + 0.2)));
  o->renderer_ready = 0;  /* as oracle/ref_wrap.cpp: snapshots only on request */
  return o;
}

void orc_destroy(void* h) {
  Oracle* o = (Oracle*)h;
  grid_free_blocks(&o->g);
  free(o->g.bbi);
  for (size_t i = 0; i < o->nportals; ++i) { portal_free(&o->portals[i].p[0]); portal_free(&o->portals[i].p[1]); }
  free(o->portals); free(o->walls); free(o->block_env); free(o->bonds); free(o->frozen);
  free(o->to_move); free(o->norend); free(o->norend_buf);
  free(o->snap_pos); free(o->snap_vel); free(o->snap_id); free(o->snap_block_start);
  free(o);
}

void orc_reset_scene(void* h, float ax, float ay, float bx, float by) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->nportals; ++i) { portal_free(&o->portals[i].p[0]); portal_free(&o->portals[i].p[1]); }
  o->nportals = 0; o->portal_stage = 0;
  o->nbonds = 0; o->nfrozen = 0; o->nnorend = 0; o->to_move_n = 0;
  Grid* g = &o->g;
  for (size_t b = 0; b < g->nblocks; ++b) g->blk[b].n = 0;
  g->bbi_n = 0; g->num_particles = 0; g->num_per_cell = 0;
  set_domain(o, V(ax, ay), V(bx, by));
  o->rqA = V(ax, ay); o->rqB = V(bx, by);
  reset_env_frame(o, V(ax, ay), V(bx, by));
  o->t = 0.f;
  o->force_enabled = 0; o->pick_enabled = 0; o->remove_last_portal = 0; o->renderer_ready = 0;
}

void orc_add_particles(void* h, const float* pos, const float* vel, const int* id, size_t n) {
  grid_add_particles(&((Oracle*)h)->g, (const v2*)pos, (const v2*)vel, id, n);
}

void orc_set_bonds(void* h, const int* pairs, size_t n) {  /* sorted unique input */
  Oracle* o = (Oracle*)h;
  o->bonds = (Bond*)realloc(o->bonds, (n + 1) * sizeof(Bond));
  o->bonds_cap = n + 1;
  memcpy(o->bonds, pairs, n * sizeof(Bond));
  o->nbonds = n;
}
size_t orc_num_bonds(void* h) { return ((Oracle*)h)->nbonds; }
size_t orc_get_bonds(void* h, int* pairs, size_t cap) {
  Oracle* o = (Oracle*)h;
  memcpy(pairs, o->bonds, (o->nbonds < cap ? o->nbonds : cap) * sizeof(Bond));
  return o->nbonds;
}
size_t orc_get_no_rendering(void* h, int* pairs, size_t cap) {
  Oracle* o = (Oracle*)h;
  memcpy(pairs, o->norend, (o->nnorend < cap ? o->nnorend : cap) * sizeof(Bond));
  return o->nnorend;
}
static int cmp_int(const void* a, const void* b) {
  int x = *(const int*)a, y = *(const int*)b;
  return (x > y) - (x < y);
}
void orc_set_frozen(void* h, const int* ids, size_t n) {
  Oracle* o = (Oracle*)h;
  o->frozen = (int*)realloc(o->frozen, (n + 1) * sizeof(int));
  o->frozen_cap = n + 1;
  memcpy(o->frozen, ids, n * sizeof(int));
  qsort(o->frozen, n, sizeof(int), cmp_int);
  size_t k = 0;
  for (size_t i = 0; i < n; ++i)
    if (k == 0 || o->frozen[k - 1] != o->frozen[i]) o->frozen[k++] = o->frozen[i];
  o->nfrozen = k;
}
size_t orc_get_frozen(void* h, int* ids, size_t cap) {
  Oracle* o = (Oracle*)h;
  memcpy(ids, o->frozen, (o->nfrozen < cap ? o->nfrozen : cap) * sizeof(int));
  return o->nfrozen;
}
void orc_add_portal_pair(void* h, const float* s) {
  Oracle* o = (Oracle*)h;
  portal_start(o, V(s[0], s[1])); portal_stop(o, V(s[2], s[3]));
  portal_start(o, V(s[4], s[5])); portal_stop(o, V(s[6], s[7]));
}
void orc_clear_portals(void* h) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->nportals; ++i) { portal_free(&o->portals[i].p[0]); portal_free(&o->portals[i].p[1]); }
  o->nportals = 0; o->portal_stage = 0;
}
size_t orc_num_portal_pairs(void* h) { return ((Oracle*)h)->nportals; }
void orc_get_portals(void* h, float* out) {
  Oracle* o = (Oracle*)h;
  size_t k = 0;
  for (size_t i = 0; i < o->nportals; ++i)
    for (int d = 0; d < 2; ++d) {
      out[k++] = o->portals[i].p[d].begin.x; out[k++] = o->portals[i].p[d].begin.y;
      out[k++] = o->portals[i].p[d].end.x; out[k++] = o->portals[i].p[d].end.y;
    }
}
size_t orc_get_portal_blocks(void* h, size_t pair, int side, int64_t* out, size_t cap) {
  const Portal* p = &((Oracle*)h)->portals[pair].p[side];
  for (size_t i = 0; i < p->nblocks && i < cap; ++i) out[i] = (int64_t)p->blocks[i];
  return p->nblocks;
}
void orc_remove_last_portal(void* h) { ((Oracle*)h)->remove_last_portal = 1; }
void orc_set_gravity(void* h, int enabled, float gx, float gy) {
  Oracle* o = (Oracle*)h; o->gravity_enable = enabled != 0; o->gravity = V(gx, gy);
}
void orc_get_gravity(void* h, int* enabled, float* g) {
  Oracle* o = (Oracle*)h; *enabled = o->gravity_enable; g[0] = o->gravity.x; g[1] = o->gravity.y;
}
void orc_set_point_force(void* h, int enabled, int attractive, float cx, float cy) {
  Oracle* o = (Oracle*)h;
  o->force_center = V(cx, cy); o->force_enabled = enabled != 0; o->force_attractive = attractive != 0;
}
void orc_set_pick(void* h, int enabled, int id, float px, float py) {
  Oracle* o = (Oracle*)h; o->pick_enabled = enabled != 0; o->pick_id = id; o->pick_pointer = V(px, py);
}
void orc_push_resize(void* h, float ax, float ay, float bx, float by) {
  Oracle* o = (Oracle*)h; o->rqA = V(ax, ay); o->rqB = V(bx, by);
}
void orc_set_renderer_ready(void* h, int v) { ((Oracle*)h)->renderer_ready = v != 0; }
/* particles.cpp:93-96 step(time_target, quit=false) */
void orc_step_to(void* h, float time_target) {
  Oracle* o = (Oracle*)h;
  while (o->t < time_target) step_once(o);
}
void orc_step_n(void* h, int n, int snapshot_each) {
  Oracle* o = (Oracle*)h;
  for (int i = 0; i < n; ++i) {
    o->renderer_ready = snapshot_each != 0;
    const float target = o->t + 0.25f * o->dt;
    while (o->t < target) step_once(o);
  }
}
void orc_eval_forces(void* h) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->g.nblocks; ++i) calc_forces(o, i);
  rhs_bonds(o); apply_portals_forces(o); apply_frozen(o);
}
float orc_time(void* h) { return ((Oracle*)h)->t; }
float orc_dt(void* h) { return ((Oracle*)h)->dt; }
size_t orc_num_steps(void* h) { Oracle* o = (Oracle*)h; return (size_t)(o->t / o->dt); }  /* particles.h:144-146 */
size_t orc_num_particles(void* h) { return ((Oracle*)h)->g.num_particles; }
size_t orc_num_per_cell(void* h) { return ((Oracle*)h)->g.num_per_cell; }
size_t orc_num_blocks(void* h) { return ((Oracle*)h)->g.nblocks; }
size_t orc_id_capacity(void* h) { return ((Oracle*)h)->g.bbi_n; }
void orc_get_geometry(void* h, float* d, int* dims) {
  Oracle* o = (Oracle*)h;
  d[0] = o->domA.x; d[1] = o->domA.y; d[2] = o->domB.x; d[3] = o->domB.y;
  d[4] = o->g.A.x; d[5] = o->g.A.y; d[6] = o->g.B.x; d[7] = o->g.B.y;
  dims[0] = o->g.di; dims[1] = o->g.dj;
}
void orc_get_neighbor_offsets(void* h, int* out9) { memcpy(out9, ((Oracle*)h)->g.noff, 9 * sizeof(int)); }
void orc_get_state_by_id(void* h, size_t cap, float* pos, float* vel, float* force,
                         int64_t* block, int64_t* slot) {
  Oracle* o = (Oracle*)h;
  const Grid* g = &o->g;
  for (size_t id = 0; id < cap; ++id) {
    int ok = id < g->bbi_n && g->bbi[id].block != KBLOCK_NONE;
    if (ok) {
      const Block* b = &g->blk[g->bbi[id].block];
      const size_t s = g->bbi[id].slot;
      if (pos) { pos[2 * id] = b->pos[s].x; pos[2 * id + 1] = b->pos[s].y; }
      if (vel) { vel[2 * id] = b->vel[s].x; vel[2 * id + 1] = b->vel[s].y; }
      if (force) { force[2 * id] = b->force[s].x; force[2 * id + 1] = b->force[s].y; }
      if (block) block[id] = (int64_t)g->bbi[id].block;
      if (slot) slot[id] = (int64_t)s;
    } else {
      if (pos) pos[2 * id] = pos[2 * id + 1] = NAN;
      if (vel) vel[2 * id] = vel[2 * id + 1] = NAN;
      if (force) force[2 * id] = force[2 * id + 1] = NAN;
      if (block) block[id] = -1;
      if (slot) slot[id] = -1;
    }
  }
}
size_t orc_get_block_lists(void* h, int* ids, size_t cap, int64_t* block_start) {
  const Grid* g = &((Oracle*)h)->g;
  size_t k = 0;
  for (size_t b = 0; b < g->nblocks; ++b) {
    block_start[b] = (int64_t)k;
    for (size_t s = 0; s < g->blk[b].n; ++s) { if (k < cap) ids[k] = g->blk[b].id[s]; ++k; }
  }
  block_start[g->nblocks] = (int64_t)k;
  return k;
}
size_t orc_get_particle_buffer(void* h, float* pv, size_t cap) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->snap_n && i < cap; ++i) {
    pv[4 * i] = o->snap_pos[i].x; pv[4 * i + 1] = o->snap_pos[i].y;
    pv[4 * i + 2] = o->snap_vel[i].x; pv[4 * i + 3] = o->snap_vel[i].y;
  }
  return o->snap_n;
}
void orc_tool(void* h, int tool, int phase, float x, float y) {
  Oracle* o = (Oracle*)h;
  const v2 p = V(x, y);
  switch (tool * 3 + phase) {
    case 0: bonds_start(o, p); break;
    case 1: bonds_move(o, p); break;
    case 2: if (o->bonds_enabled) o->bonds_enabled = 0; break;
    case 3: pick_start(o, p); break;
    case 4: if (o->pick_enabled) o->pick_pointer = p; break;
    case 5: if (o->pick_enabled) o->pick_enabled = 0; break;
    case 6: o->freeze_last_id = -1; o->freeze_enabled = 1; freeze_move(o, p); break;
    case 7: freeze_move(o, p); break;
    case 8: if (o->freeze_enabled) o->freeze_enabled = 0; break;
    case 9: portal_start(o, p); break;
    case 10: portal_move(o, p); break;
    case 11: portal_stop(o, p); break;
    default: break;
  }
}
void orc_set_particle_buffer(void* h) { set_particle_buffer((Oracle*)h); }
void orc_F12(const float* p, const float* q, float* out) { v2 f = F12(V(p[0], p[1]), V(q[0], q[1])); out[0] = f.x; out[1] = f.y; }
void orc_F12wall(const float* p, const float* q, float* out) { v2 f = F12wall(V(p[0], p[1]), V(q[0], q[1])); out[0] = f.x; out[1] = f.y; }
void orc_F12_bond(const float* p, const float* q, float* out) { v2 f = F12_bond(V(p[0], p[1]), V(q[0], q[1])); out[0] = f.x; out[1] = f.y; }
void orc_F12_portal_edge(const float* p, const float* q, float* out) { v2 f = F12_portal_edge(V(p[0], p[1]), V(q[0], q[1])); out[0] = f.x; out[1] = f.y; }
void orc_F12_pick(const float* p, const float* q, float* out) { v2 f = F12_pick(V(p[0], p[1]), V(q[0], q[1])); out[0] = f.x; out[1] = f.y; }
int64_t orc_find_block(void* h, float x, float y) {
  size_t b = grid_find_block(&((Oracle*)h)->g, V(x, y));
  return b == KBLOCK_NONE ? -1 : (int64_t)b;
}
void orc_move_to_portal(void* h, size_t pair, int side, float* pos, float* vel) {
  Oracle* o = (Oracle*)h;
  v2 p = V(pos[0], pos[1]), v = V(vel[0], vel[1]);
  move_to_portal(&p, &v, &o->portals[pair].p[side], &o->portals[pair].p[1 - side]);
  pos[0] = p.x; pos[1] = p.y; vel[0] = v.x; vel[1] = v.y;
}
#ifdef _OPENMP
#include <omp.h>
int orc_openmp_threads(void) { return omp_get_max_threads(); }
#else
int orc_openmp_threads(void) { return 1; }
#endif

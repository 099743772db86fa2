// Hand-written sm_100a kernels of the particle-stepping hot path.
//
// Arithmetic contract: every floating-point operation that the reference performs
// (scalar build, no FMA) is issued here as the same single IEEE
// binary32 operation through the __f*_rn intrinsics, which nvcc never contracts
// into FMAs.  A pair/wall/bond/portal contribution is therefore bit-identical to
// the reference's; only the ORDER in which a particle's pair contributions are
// summed differs (sub-cell order instead of block/slot order).
//
// Kernels (one launch each per iteration):
//   stage_kernel<1>  forces at x0 + half kick/drift          particles.cpp:97-123
//   stage_kernel<2>  forces at x_half + full update + portal teleport + new
//                    sub-cell key + histogram                particles.cpp:128-153,177-179
//   scan_kernel      single-pass decoupled look-back exclusive scan of the histogram
//   fill_kernel      counting-sort slot assignment
//   gather_kernel    deterministic (stable) reorder of pos/vel/id + id->slot map
#include "common.h"

// resident CTAs per SM the stage kernels are compiled for (measured on B200: 40 registers win
// without bonds, 48 with the bond partner prefetch)
#ifndef PTOY_STAGE_MINB
#define PTOY_STAGE_MINB 6
#endif
#ifndef PTOY_STAGE_MINB_BONDS
#define PTOY_STAGE_MINB_BONDS 5
#endif

namespace ptoy {

// ---------------------------------------------------------------------------------
// exact-rounding float helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// (float)(1. / (double)x): double division then rounding to float equals the
// correctly rounded float quotient (53 >= 2*24+2), i.e. __fdiv_rn(1, x).
__device__ __forceinline__ float rcp_exact(float x) { return __frcp_rn(x); }
// geometry.h:62-64  dot = v.x*x + v.y*y
__device__ __forceinline__ float dot2(float ax, float ay, float bx, float by) {
  return fadd(fmul(ax, bx), fmul(ay, by));
}
// geometry.h:65-67  length = sqrt(x*x + y*y) (double sqrt of a float, rounded back:
// identical to the correctly rounded float sqrt)
__device__ __forceinline__ float len2(float x, float y) { return __fsqrt_rn(dot2(x, y, x, y)); }
// r / R for the constant R = 2r = 0.04f (particles.cpp:729,733).  With y = fl(1/R) = 25 exactly,
// q0 = fl(r*y), rem = fma(-q0, R, r), q1 = fma(rem, y, q0) is the correctly rounded quotient
// (Markstein); checked exhaustively against r / R for every float r in [1e-30, 1e30]
// (scripts/check_exact_tricks.c).  The fused multiply-adds are part of the division algorithm,
// not contractions of reference operations.
__device__ __forceinline__ float div_by_bondR(float r, float R, float y) {
  const float q0 = __fmul_rn(r, y);
  const float rem = __fmaf_rn(-q0, R, r);
  return __fmaf_rn(rem, y, q0);
}
// std::max<float>(0, k) = (0 < k) ? k : 0   (NaN -> 0)
__device__ __forceinline__ float max0(float k) { return (0.0f < k) ? k : 0.0f; }

// ---------------------------------------------------------------------------------
// binning (blocks::FindBlock, blocks.h:155-163) at sub-cell resolution
// ---------------------------------------------------------------------------------
struct Cell {
  int Si, Sj;     // padded sub-cell row / column
  float ui, uj;   // fractional position inside the sub-cell, [0,1)
  bool valid;     // FindBlock != kBlockNone
};

__device__ __forceinline__ Cell cell_of(const GridDev& g, float2 p) {
  Cell c;
  const float tx = fdiv(fsub(p.x, g.Ax), g.bsx);
  const float ty = fdiv(fsub(p.y, g.Ay), g.bsy);
  // (int)t truncates toward zero: t in (-1,0) still maps to row/column 0
  c.valid = (tx > -1.0f) && (tx < g.fdi) && (ty > -1.0f) && (ty < g.fdj);
  const float wx = tx * 2.0f, wy = ty * 2.0f;  // exact
  const float fx = floorf(wx), fy = floorf(wy);
  c.Si = (int)fx + 4;
  c.Sj = (int)fy + 4;
  c.ui = wx - fx;  // exact
  c.uj = wy - fy;
  return c;
}
// Same result as cell_of() without the two IEEE divisions in the common case:
// w~ = fl((x - A) * fl(2/bs)) differs from the exact w = 2*fl(fl(x - A)/bs) by less than
// |w| * 2^-22, so floor(w~) == floor(w) unless w~ lies within g.tol (>= 2 * wmax * 2^-22) of an
// integer; those few particles take the exact path.  (ui, uj are then approximate by < tol,
// which the search margin absorbs.)
__device__ __forceinline__ Cell cell_of_fast(const GridDev& g, float2 p) {
  const float wx = fsub(p.x, g.Ax) * g.inv_hsx, wy = fsub(p.y, g.Ay) * g.inv_hsy;
  const float fx = floorf(wx), fy = floorf(wy);
  const float ux = wx - fx, uy = wy - fy;
  const float hi = 1.0f - g.tol;
  const bool safe = ux > g.tol && ux < hi && uy > g.tol && uy < hi &&
                    fabsf(wx) < g.wmax && fabsf(wy) < g.wmax;  // false for NaN
  if (!safe) return cell_of(g, p);
  Cell c;
  c.valid = fx >= -2.0f && fx < 2.0f * g.fdi && fy >= -2.0f && fy < 2.0f * g.fdj;
  c.Si = (int)fx + 4;
  c.Sj = (int)fy + 4;
  c.ui = ux;
  c.uj = uy;
  return c;
}
__device__ __forceinline__ int block_of_s(const GridDev& g, int Si, int Sj) {
  const int bi = max(0, (Si - 4) >> 1), bj = max(0, (Sj - 4) >> 1);
  return bi * g.dj + bj;
}
__device__ __forceinline__ int block_of(const GridDev& g, const Cell& c) { return block_of_s(g, c.Si, c.Sj); }
// packed sub-cell of a particle: Si << 16 | Sj ; kTrashCell = left the grid
constexpr unsigned kTrashCell = 0xffffffffu;
// local packed cell of a particle this strip keeps; kTrashCell if it left the grid or belongs
// to another strip (`migrant`)
__device__ __forceinline__ unsigned pack_cell(const GridDev& g, const Cell& c, bool& migrant) {
  migrant = false;
  if (!c.valid) return kTrashCell;
  const int lr = c.Si - g.row0;
  if (lr < g.own_lo || lr >= g.own_hi) { migrant = true; return kTrashCell; }
  return ((unsigned)lr << 16) | (unsigned)c.Sj;
}
__device__ __forceinline__ int key_of_packed(unsigned p, int SJ, int ncell) {
  return p == kTrashCell ? ncell : (int)(p >> 16) * SJ + (int)(p & 0xffffu);
}
// sub-rows / sub-columns (inclusive) that make up reference block `blk`
__device__ __forceinline__ void block_extent(const GridDev& g, int blk, int& r0, int& r1, int& c0, int& c1) {
  const int bi = blk / g.dj, bj = blk - bi * g.dj;
  r0 = (bi == 0) ? 2 : 2 * bi + 4;
  r1 = 2 * bi + 5;
  c0 = (bj == 0) ? 2 : 2 * bj + 4;
  c1 = 2 * bj + 5;
}

// ---------------------------------------------------------------------------------
// force laws
// ---------------------------------------------------------------------------------
// particles.cpp:584-595 F12 for a pair already known to be inside the cutoff
// (r2 < r2c <=> the clamped coefficient is > 0, see engine.cu find_r2c()).
__device__ __forceinline__ void pair_force(const Phys& ph, float dx, float dy, float r2, float& fx, float& fy) {
  const float r2t = (ph.thr < r2) ? r2 : ph.thr;  // std::max(threshold, r2)
  const float inv = rcp_exact(r2t);
  const float d2 = fmul(inv, ph.RR);
  const float d6 = fmul(fmul(d2, d2), d2);
  const float d12 = fmul(d6, d6);
  const float k = fmul(fsub(d12, d6), inv);  // sigma = 1
  fx = fmul(dx, k);
  fy = fmul(dy, k);
}
// particles.cpp:572-582 / :738-749: no threshold, clamp by std::max<Scal>(0., .)
__device__ __forceinline__ void lj_noclamp(float RR, float px, float py, float qx, float qy, float& fx, float& fy) {
  const float dx = fsub(px, qx), dy = fsub(py, qy);
  const float r2 = dot2(dx, dy, dx, dy);
  const float inv = rcp_exact(r2);
  const float d2 = fmul(inv, RR);
  const float d6 = fmul(fmul(d2, d2), d2);
  const float d12 = fmul(d6, d6);
  const float k = max0(fmul(fsub(d12, d6), inv));
  fx = fmul(dx, k);
  fy = fmul(dy, k);
}
// particles.cpp:727-736 F12_bond
__device__ __forceinline__ void bond_force(const Phys& ph, float px, float py, float qx, float qy, float& fx, float& fy) {
  const float dx = fsub(px, qx), dy = fsub(py, qy);
  const float r = len2(dx, dy);
  // k = max(1. - r/R, 1. - 5) in double, rounded to float.  1 - x is exact in
  // double for float x >= 2^-29 and rounds to 1.0f either way below that, so
  // the float subtraction gives the same float; -4 is exact.
  const float q = div_by_bondR(r, ph.bondR, ph.bond_y);
  float k = fsub(1.0f, q);
  k = (k < -4.0f) ? -4.0f : k;  // std::max(a, b) = (a < b) ? b : a
  const float s = fmul(ph.sigma_bond, k);
  fx = fmul(dx, s);
  fy = fmul(dy, s);
}

// particles.h:37-46 line::GetNearest
__device__ __forceinline__ void seg_nearest(float ax, float ay, float bx, float by, float px, float py, float& qx, float& qy) {
  const float abx = fsub(bx, ax), aby = fsub(by, ay);
  const float apx = fsub(px, ax), apy = fsub(py, ay);
  const float lambda = fdiv(dot2(abx, aby, apx, apy), dot2(abx, aby, abx, aby));
  if (lambda > 0.0f && lambda < 1.0f) {
    qx = fadd(ax, fmul(abx, lambda));
    qy = fadd(ay, fmul(aby, lambda));
  } else {
    const float da = len2(fsub(ax, px), fsub(ay, py));  // p.dist(A) = (A - p).length()
    const float db = len2(fsub(bx, px), fsub(by, py));
    if (da < db) { qx = ax; qy = ay; } else { qx = bx; qy = by; }
  }
}

// lambda / offset of point p in the frame of a portal (e.g. particles.cpp:340-341)
__device__ __forceinline__ void portal_coords(const PortalSide& s, float px, float py, float& lambda, float& offset) {
  const float dx = fsub(px, s.ax), dy = fsub(py, s.ay);
  lambda = fdiv(dot2(s.rx, s.ry, dx, dy), s.rr);
  offset = dot2(dx, dy, s.nx, s.ny);
}
// a + r*lambda + n*(offset + 2.*sign*thickness)   (particles.cpp:369-371, 810-812)
__device__ __forceinline__ void portal_image(const Phys& ph, const PortalSide& s, float lambda_q, float offset_q, float sign, float& x, float& y) {
  const float shifted = __double2float_rn(__dadd_rn((double)offset_q, __dmul_rn(__dmul_rn(2.0, (double)sign), (double)ph.portal_thickness)));
  x = fadd(fadd(s.ax, fmul(s.rx, lambda_q)), fmul(s.nx, shifted));
  y = fadd(fadd(s.ay, fmul(s.ry, lambda_q)), fmul(s.ny, shifted));
}

__device__ __forceinline__ bool in_sorted(const int* __restrict__ list, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (list[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo < n && list[lo] == v;
}

// ---------------------------------------------------------------------------------
// all forces on one particle (particles.cpp:839-910 calc_forces, :761-826 RHS_bonds,
// :318-381 ApplyPortalsForces), summed in the reference's order of force kinds
// ---------------------------------------------------------------------------------
// Position of a bond partner given its entry in the id -> slot map: a local slot (>= 0), a ghost
// reference -2 - (side * ghost_cap + index) into the neighbour strip's boundary rows, or -1
// (partner left the grid: the bond is erased by CheckBonds, particles.cpp:205-215).
template <bool STRIPS>
__device__ __forceinline__ float2 bond_partner_pos(const StageArgs& a, const float2* __restrict__ X, int sb) {
  if (sb >= 0) return __ldg(X + sb);
  if (STRIPS && sb <= -2) {
    const int code = -2 - sb;
    const int side = code >= a.ghost_cap;
    return __ldg(a.gpos[side] + (code - side * a.ghost_cap));
  }
  return make_float2(0.f, 0.f);
}

constexpr int kHitQueue = 6;     // pending in-range pairs per thread (shared memory)
constexpr int kBondPlan = 6;     // bond partners per particle resolved up front / handed from stage 1 to stage 2
struct HitQueue {
  float dx[kHitQueue][kThreads];
  float dy[kHitQueue][kThreads];
};

// One queued pair: the exact r^2 (reference rounding) decides, the prefilter only proposed it.
__device__ __forceinline__ void eval_hit(const Phys& ph, float dx, float dy, float& fx, float& fy) {
  const float r2 = dot2(dx, dy, dx, dy);
  if (r2 < ph.r2c) {
    float px, py;
    pair_force(ph, dx, dy, r2, px, py);
    fx = fadd(fx, px);
    fy = fadd(fy, py);
  }
}

// FEAT selects what the scene uses, so that unused force kinds cost neither instructions nor
// registers: 1 = bonds, 2 = portals, 4 = mouse point force / pick, 8 = strip decomposition
// (ghost rows, ghost bond partners); launch_stage picks the variant.
constexpr int kFeatBonds = 1, kFeatPortals = 2, kFeatExtras = 4, kFeatStrips = 8;
template <int STAGE, int FEAT>
__device__ __forceinline__ float2 particle_force(
    const StageArgs& a, const float2* __restrict__ X, HitQueue& hq, int slot, int pid, float2 x, float2 v,
    int Si, int Sj, float ui, float uj, int blk, unsigned flags) {
  const Phys& ph = a.ph;
  float fx = 0.0f, fy = 0.0f;
  if (ph.gravity_enable) { fx = ph.gfx; fy = ph.gfy; }                       // :849-851

  if ((FEAT & kFeatExtras) && ph.force_enabled) {                            // :854-863
    const float rx = fsub(x.x, ph.fcx), ry = fsub(x.y, ph.fcy);
    const float l = len2(rx, ry);
    if (l > ph.radius) {
      const double ld = (double)l;
      float s;
      if (ph.force_attractive) {
        const double l3 = __dmul_rn(__dmul_rn(ld, ld), ld);                  // pow(l, 3)
        s = __double2float_rn(__ddiv_rn((double)(-ph.point_force_attractive), l3));
      } else {
        const double l2 = __dmul_rn(ld, ld);
        s = __double2float_rn(__ddiv_rn((double)ph.point_force, __dmul_rn(l2, l2)));  // pow(l, 4)
      }
      fx = fadd(fx, fmul(rx, s));
      fy = fadd(fy, fmul(ry, s));
    }
  }

  if ((FEAT & kFeatPortals) && (flags & 2u)) {                               // :866-871
    for (int s = 0; s < a.n_sides; ++s) {
      const PortalSide ps = a.sides[s];
      float ex, ey;
      lj_noclamp(ph.RRe, x.x, x.y, ps.ax, ps.ay, ex, ey);
      fx = fadd(fx, ex); fy = fadd(fy, ey);
      lj_noclamp(ph.RRe, x.x, x.y, ps.bx, ps.by, ex, ey);
      fx = fadd(fx, ex); fy = fadd(fy, ey);
    }
  }

  if ((FEAT & kFeatExtras) && ph.pick_enabled && pid == ph.pick_id) {        // :874-876, :751-759
    const float dx = fsub(x.x, ph.pkx), dy = fsub(x.y, ph.pky);
    const float r = len2(dx, dy);
    const float k = fdiv(fsub(ph.pickR, r), ph.pickR);
    const float s = fmul(ph.sigma_pick, k);
    fx = fadd(fx, fmul(dx, s));
    fy = fadd(fy, fmul(dy, s));
  }

  fx = fsub(fx, fmul(v.x, ph.diss));                                         // :880
  fy = fsub(fy, fmul(v.y, ph.diss));

  // ---- bonds, part 1: partner slots and positions, fetched before the pair search starts so
  // the loads overlap it.  Stage 1 reads the particle's row of the bond table (8 ints per id: the
  // first 6 partner ids in ascending = std::set order, the degree in the last; two 16-byte loads),
  // resolves partner id -> slot with one batch of independent gathers and records the slots in
  // one 32-byte sector per particle; stage 2 (same sorted order, same partners) reads them back
  // with two 16-byte loads.
  int nb = 0;
  float2 bpos[kBondPlan];
  int bslot[kBondPlan];
  if (FEAT & kFeatBonds) {
    int4* hand = reinterpret_cast<int4*>(a.bslot) + 2 * (size_t)slot;
    if (STAGE == 1) {
      const int4* row = reinterpret_cast<const int4*>(a.bond_ell) + 2 * (size_t)pid;
      const int4 r0 = __ldg(row), r1 = __ldg(row + 1);
      const int bcol[kBondPlan] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
      nb = r1.w;
#pragma unroll
      for (int k = 0; k < kBondPlan; ++k) bslot[k] = (bcol[k] >= 0) ? a.slot_of_id[bcol[k]] : -1;
      if (FEAT & kFeatStrips) {  // a partner that is neither local nor in a ghost row cannot be served
        bool lost = false;
#pragma unroll
        for (int k = 0; k < kBondPlan; ++k) lost |= bcol[k] >= 0 && bslot[k] == -1;
        if (lost) atomicOr(a.err, 16);
      }
      hand[0] = make_int4(bslot[0], bslot[1], bslot[2], bslot[3]);
      hand[1] = make_int4(bslot[4], bslot[5], -1, nb);
    } else {
      const int4 r0 = hand[0], r1 = hand[1];
      bslot[0] = r0.x; bslot[1] = r0.y; bslot[2] = r0.z; bslot[3] = r0.w; bslot[4] = r1.x; bslot[5] = r1.y;
      nb = r1.w;
    }
#pragma unroll
    for (int k = 0; k < kBondPlan; ++k) bpos[k] = bond_partner_pos<(FEAT & kFeatStrips) != 0>(a, X, bslot[k]);
  }

  // ---- pair forces (:884-900 + :598-606) over the sub-cell neighbourhood ----------------
  // Candidates are only TESTED in the loop, with a cheap conservative prefilter
  // (fma(dx,dx,dy*dy) against r2c grown by a few ulp); the few that pass are queued per thread
  // in shared memory and evaluated afterwards with the reference's exact arithmetic, so the
  // divergent force evaluation runs once per queued pair instead of once per loop iteration in
  // which any lane had a hit.  The queue is FIFO: contributions are summed in candidate order.
  // The particle itself (skipped by address in the reference) is excluded by splitting its own
  // row's range around its slot.
  {
    const int t = threadIdx.x;
    int nq = 0;
    const int r_lo = Si - 1 - (ui < a.margin), r_hi = Si + 1 + (ui > 1.0f - a.margin);
    const int c_lo = Sj - 1 - (uj < a.margin), c_hi = Sj + 1 + (uj > 1.0f - a.margin);
    auto test = [&](float2 pq) {
      const float dx = fsub(x.x, pq.x), dy = fsub(x.y, pq.y);
      const float r2a = __fmaf_rn(dx, dx, dy * dy);
      if (r2a < ph.r2c_hi) {
        hq.dx[nq][t] = dx; hq.dy[nq][t] = dy;
        if (++nq == kHitQueue) {
#pragma unroll
          for (int k = 0; k < kHitQueue; ++k) eval_hit(ph, hq.dx[k][t], hq.dy[k][t], fx, fy);
          nq = 0;
        }
      }
    };
    auto scan = [&](const float2* __restrict__ P, int q, int end) {
      for (; q + 1 < end; q += 2) {
        const float2 p0 = __ldg(P + q), p1 = __ldg(P + q + 1);
        test(p0);
        test(p1);
      }
      if (q < end) test(__ldg(P + q));
    };
    if (!(FEAT & kFeatStrips) || (r_lo >= a.g.own_lo && r_hi < a.g.own_hi)) {
      // (strips: taken by every particle whose search rows are all owned — all but the two
      // boundary rows on each side)
      // the three rows every particle needs: all six cell offsets are loaded before the first
      // scan; the rare extra rows (fractional position within the margin of a cell edge) follow
      const int* sp = a.start + (Si - 1) * a.g.SJ;
      const int b0 = __ldg(sp + c_lo), e0 = __ldg(sp + c_hi + 1);
      const int b1 = __ldg(sp + a.g.SJ + c_lo), e1 = __ldg(sp + a.g.SJ + c_hi + 1);
      const int b2 = __ldg(sp + 2 * a.g.SJ + c_lo), e2 = __ldg(sp + 2 * a.g.SJ + c_hi + 1);
      scan(X, b0, e0);
      scan(X, b1, slot);
      scan(X, slot + 1, e1);
      scan(X, b2, e2);
      if (r_lo < Si - 1) {
        const int* s2 = a.start + (Si - 2) * a.g.SJ;
        scan(X, __ldg(s2 + c_lo), __ldg(s2 + c_hi + 1));
      }
      if (r_hi > Si + 1) {
        const int* s2 = a.start + (Si + 2) * a.g.SJ;
        scan(X, __ldg(s2 + c_lo), __ldg(s2 + c_hi + 1));
      }
    } else {
      // strips: the same row order (so that N strips sum in the order one GPU does); a row outside
      // the owned range is a row of the neighbouring strip, served from its ghost slices
      auto scan_row = [&](int R) {
        if (R >= a.g.own_lo && R < a.g.own_hi) {
          const int base = R * a.g.SJ;
          const int beg = __ldg(a.start + base + c_lo), end = __ldg(a.start + base + c_hi + 1);
          if (R == Si) {
            scan(X, beg, slot);
            scan(X, slot + 1, end);
          } else {
            scan(X, beg, end);
          }
        } else {
          const int side = R >= a.g.own_hi;
          const int* gs = a.gstart[side];
          if (gs == nullptr) return;
          const int base = (R - (side ? a.g.own_hi : a.g.own_lo - 2)) * a.g.SJ;
          const int g0 = __ldg(gs);
          scan(a.gpos[side], __ldg(gs + base + c_lo) - g0, __ldg(gs + base + c_hi + 1) - g0);
        }
      };
      scan_row(Si - 1);
      scan_row(Si);
      scan_row(Si + 1);
      if (r_lo < Si - 1) scan_row(Si - 2);
      if (r_hi > Si + 1) scan_row(Si + 2);
    }
    for (int k = 0; k < nq; ++k) eval_hit(ph, hq.dx[k][t], hq.dy[k][t], fx, fy);
  }

  // ---- walls (:903-909); exact zero outside each wall's grown bbox ----------------
  if (!(x.x > a.free_lox && x.x < a.free_hix && x.y > a.free_loy && x.y < a.free_hiy)) {
    for (int k = 0; k < a.n_walls; ++k) {
      const WallDev w = a.walls[k];
      if (x.x >= w.lox && x.x <= w.hix && x.y >= w.loy && x.y <= w.hiy) {
        float qx, qy, wx, wy;
        seg_nearest(w.ax, w.ay, w.bx, w.by, x.x, x.y, qx, qy);
        lj_noclamp(ph.RRw, x.x, x.y, qx, qy, wx, wy);
        fx = fadd(fx, wx);
        fy = fadd(fy, wy);
      }
    }
  }

  // ---- bonds, part 2 (:761-826): CSR row of this id, partners ascending ----------------
  if (FEAT & kFeatBonds) {
    auto one_bond = [&](int sb, float2 q) {
      if (sb == -1) return;  // partner left the grid: bond erased by CheckBonds (:205-215)
      float bx, by;
      // p_pos.dist(q_pos) = |q - p| and F12_bond's |p - q| are the same float: the differences
      // are exact negatives of each other, so one sqrt serves both (particles.cpp:731,779)
      const float ddx = fsub(x.x, q.x), ddy = fsub(x.y, q.y);
      const float dist = len2(ddx, ddy);
      bool found = false;
      if ((FEAT & kFeatPortals) && !(dist < ph.bond_cutoff)) {               // :784-819
        for (int s = 0; s < a.n_sides; ++s) {
          const PortalSide po = a.sides[s];
          const PortalSide ot = a.sides[s ^ 1];
          float p_lambda, p_offset, q_lambda, q_offset;
          portal_coords(po, x.x, x.y, p_lambda, p_offset);
          portal_coords(ot, q.x, q.y, q_lambda, q_offset);
          if (p_lambda > 0.0f && p_lambda < 1.0f && q_lambda > 0.0f && q_lambda < 1.0f &&
              fmul(p_offset, q_offset) < 0.0f &&
              (double)fabsf(fsub(p_offset, q_offset)) < ph.bond_portal_reach) {
            const float sign = (p_offset > 0.0f) ? 1.0f : -1.0f;
            float jx, jy;
            portal_image(ph, po, q_lambda, q_offset, sign, jx, jy);
            bond_force(ph, x.x, x.y, jx, jy, bx, by);
            found = true;  // no break: the last matching (pair, side) wins
          }
        }
      }
      if (!found) {                                                          // :727-736 inlined
        float k = fsub(1.0f, div_by_bondR(dist, ph.bondR, ph.bond_y));
        k = (k < -4.0f) ? -4.0f : k;
        const float sk = fmul(ph.sigma_bond, k);
        bx = fmul(ddx, sk);
        by = fmul(ddy, sk);
      }
      fx = fadd(fx, bx);
      fy = fadd(fy, by);
    };
#pragma unroll
    for (int k = 0; k < kBondPlan; ++k) one_bond(bslot[k], bpos[k]);  // absent / dead partners are -1
    if (nb > kBondPlan) {  // more partners than the plan holds: walk the rest of the CSR row
      const unsigned e0 = __ldg(a.bond_ptr + pid), ne = __ldg(a.bond_ptr + pid + 1) - e0;
      for (unsigned e = kBondPlan; e < ne; ++e) {
        const int sb = a.slot_of_id[a.bond_col[e0 + e]];
        one_bond(sb, bond_partner_pos<(FEAT & kFeatStrips) != 0>(a, X, sb));
      }
    }
  }

  // ---- cross-portal forces (:318-381) ------------------------------------------------
  if ((FEAT & kFeatPortals) && (flags & 1u)) {
    for (int s = 0; s < a.n_sides; ++s) {
      const int lb = a.pblk_ptr[s], le = a.pblk_ptr[s + 1];
      if (!in_sorted(a.pblk + lb, le - lb, blk)) continue;
      const PortalSide po = a.sides[s];
      const PortalSide ot = a.sides[s ^ 1];
      float lambda_c, offset_c;
      portal_coords(po, x.x, x.y, lambda_c, offset_c);
      if (!(lambda_c > 0.0f && lambda_c < 1.0f)) continue;
      // same-portal particles on the other side of the plane: subtract (:345-355)
      for (int l = lb; l < le; ++l) {
        int r0, r1, c0, c1;
        block_extent(a.g, a.pblk[l], r0, r1, c0, c1);
        for (int Rg = r0; Rg <= r1; ++Rg) {
          const int R = Rg - a.g.row0;  // strips: only locally stored rows (DESIGN.md §7 limitation)
          if (R < a.g.own_lo || R >= a.g.own_hi) continue;
          const int beg = a.start[R * a.g.SJ + c0], end = a.start[R * a.g.SJ + c1 + 1];
          for (int q = beg; q < end; ++q) {
            const float2 nb = X[q];
            float ln, on;
            portal_coords(po, nb.x, nb.y, ln, on);
            if (ln > 0.0f && ln < 1.0f && fmul(on, offset_c) < 0.0f) {
              const float dx = fsub(x.x, nb.x), dy = fsub(x.y, nb.y);
              const float r2 = dot2(dx, dy, dx, dy);
              if (r2 < ph.r2c) {
                float px, py;
                pair_force(ph, dx, dy, r2, px, py);
                fx = fsub(fx, px);
                fy = fsub(fy, py);
              }
            }
          }
        }
      }
      // images of the partner portal's particles: add (:358-375)
      const int ob = a.pblk_ptr[s ^ 1], oe = a.pblk_ptr[(s ^ 1) + 1];
      const float sign = (offset_c > 0.0f) ? 1.0f : -1.0f;
      for (int l = ob; l < oe; ++l) {
        int r0, r1, c0, c1;
        block_extent(a.g, a.pblk[l], r0, r1, c0, c1);
        for (int Rg = r0; Rg <= r1; ++Rg) {
          const int R = Rg - a.g.row0;  // strips: only locally stored rows (DESIGN.md §7 limitation)
          if (R < a.g.own_lo || R >= a.g.own_hi) continue;
          const int beg = a.start[R * a.g.SJ + c0], end = a.start[R * a.g.SJ + c1 + 1];
          for (int q = beg; q < end; ++q) {
            const float2 nb = X[q];
            float oln, oon;
            portal_coords(ot, nb.x, nb.y, oln, oon);
            if (oln > 0.0f && oln < 1.0f && fmul(oon, offset_c) < 0.0f) {
              float jx, jy;
              portal_image(ph, po, oln, oon, sign, jx, jy);
              const float dx = fsub(x.x, jx), dy = fsub(x.y, jy);
              const float r2 = dot2(dx, dy, dx, dy);
              if (r2 < ph.r2c) {
                float px, py;
                pair_force(ph, dx, dy, r2, px, py);
                fx = fadd(fx, px);
                fy = fadd(fy, py);
              }
            }
          }
        }
      }
    }
  }
  return make_float2(fx, fy);
}

// particles.cpp:117-120: if (|v| > limit) v *= limit / |v|
// fl(sqrt(v2)) > limit  <=>  v2 > v2c (sqrt is monotone; v2c found by bisection on the host), so
// the square root is only taken when the clamp fires.
__device__ __forceinline__ void clamp_velocity(const Phys& ph, float& vx, float& vy) {
  const float v2 = dot2(vx, vy, vx, vy);
  if (v2 > ph.v2c) {
    const float l = __fsqrt_rn(v2);
    const float s = fdiv(ph.vlimit, l);
    vx = fmul(vx, s);
    vy = fmul(vy, s);
  }
}

// DetectPortals + ApplyPortals (particles.cpp:288-316, 382-409) for one particle whose
// reference block is `blk`, then the new sub-cell key and the histogram update
// (the device half of blocks::SortParticles, blocks.h:215-244).
template <class A>
__device__ __forceinline__ void teleport(const A& a, int blk, float2& x, float2& v) {
  bool flagged = false;
  for (int s = 0; s < a.n_sides; ++s) {
    const int lb = a.pblk_ptr[s], le = a.pblk_ptr[s + 1];
    if (!in_sorted(a.pblk + lb, le - lb, blk)) continue;
    float lambda, offset;
    portal_coords(a.sides[s], x.x, x.y, lambda, offset);
    if (lambda > 0.0f && lambda < 1.0f && fabsf(offset) <= a.ph.portal_thickness) flagged = true;
  }
  if (!flagged) return;
  for (int s = 0; s < a.n_sides; ++s) {  // first (pair, side) whose block list holds the particle
    const int lb = a.pblk_ptr[s], le = a.pblk_ptr[s + 1];
    if (!in_sorted(a.pblk + lb, le - lb, blk)) continue;
    const PortalSide src = a.sides[s];
    const PortalSide dst = a.sides[s ^ 1];
    // MoveToPortal (particles.cpp:264-286)
    float lambda_pos, offset_pos;
    portal_coords(src, x.x, x.y, lambda_pos, offset_pos);
    // offset_pos there is src_n.dot(position - src_a): same products, same sum
    const float lambda_vel = fdiv(dot2(src.rx, src.ry, v.x, v.y), src.rr);
    const float offset_vel = dot2(src.nx, src.ny, v.x, v.y);
    const float sign = (offset_pos > 0.0f) ? 1.0f : -1.0f;
    const float shifted = __double2float_rn(
        __dsub_rn((double)offset_pos, __dmul_rn(__dmul_rn((double)sign, 2.0), (double)a.ph.portal_thickness)));
    x.x = fadd(fadd(dst.ax, fmul(dst.rx, lambda_pos)), fmul(dst.nx, shifted));
    x.y = fadd(fadd(dst.ay, fmul(dst.ry, lambda_pos)), fmul(dst.ny, shifted));
    const float vx = fadd(fmul(dst.rx, lambda_vel), fmul(dst.nx, offset_vel));
    const float vy = fadd(fmul(dst.ry, lambda_vel), fmul(dst.ny, offset_vel));
    v.x = vx;
    v.y = vy;
    return;
  }
}

// New sub-cell of a particle after the update + histogram (the device half of
// blocks::SortParticles).  A particle whose new sub-row belongs to another strip is put into
// that strip's outbox (with its old slot, which fixes its place in the receiver's stable order)
// and leaves this strip through the trash bin.
__device__ __forceinline__ void key_and_count(const GridDev& g, const DistDev& d, float2 x, float2 v,
                                              const int* __restrict__ id, int slot, unsigned* newcell, int* cnt) {
  const Cell c = cell_of_fast(g, x);
  bool migrant;
  const unsigned pc = pack_cell(g, c, migrant);
  if (migrant && d.out_base) {
    int dest = 0;
    while (dest + 1 < d.world && c.Si >= d.row_begin[dest + 1]) ++dest;
    char* blk = d.out_base + dest * d.out_stride;
    const int k = atomicAdd(reinterpret_cast<int*>(blk), 1);
    if (k < d.out_cap) {
      reinterpret_cast<float2*>(blk + mig_off_pos())[k] = x;
      reinterpret_cast<float2*>(blk + mig_off_vel(d.out_cap))[k] = v;
      reinterpret_cast<int*>(blk + mig_off_id(d.out_cap))[k] = id[slot];
      reinterpret_cast<int*>(blk + mig_off_tag(d.out_cap))[k] = slot;
    }
  }
  newcell[slot] = pc;
  atomicAdd(cnt + key_of_packed(pc, g.SJ, g.ncell), 1);  // result unused -> RED.ADD
}

template <int STAGE, int FEAT>
__global__ void __launch_bounds__(kThreads, (FEAT & kFeatBonds) ? PTOY_STAGE_MINB_BONDS : PTOY_STAGE_MINB) stage_kernel(const __grid_constant__ StageArgs a) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.ctr->n_cur) return;
  const float2 x0 = a.pos[i];
  float2 v0 = a.vel[i];
  float2 x, v;
  if (STAGE == 1) { x = x0; v = v0; } else { x = a.pos_h[i]; v = a.vel_h[i]; }
  const float2* __restrict__ X = (STAGE == 1) ? a.pos : (const float2*)a.pos_h;
  __shared__ HitQueue hq;
  // the sub-cell this particle was binned into (stored by the reorder), and its approximate
  // fractional position inside it (only compared against the search margin)
  const unsigned pc = a.cell[i];
  const int Si = (int)(pc >> 16), Sj = (int)(pc & 0xffffu);
  const float ui = fsub(x0.x, a.g.Ax) * a.g.inv_hsx - (float)(Si + a.g.row0 - 4);
  const float uj = fsub(x0.y, a.g.Ay) * a.g.inv_hsy - (float)(Sj - 4);
  int blk = 0;
  unsigned flags = 0u;
  if ((FEAT & kFeatPortals) && a.blk_flags) { blk = block_of_s(a.g, Si + a.g.row0, Sj); flags = a.blk_flags[blk]; }
  const int pid = a.need_id ? a.id[i] : 0;

  float2 f = particle_force<STAGE, FEAT>(a, X, hq, i, pid, x, v, Si, Sj, ui, uj, blk, flags);

  bool frozen = false;
  if (a.has_frozen) frozen = (a.frozen_bits[pid >> 5] >> (pid & 31)) & 1u;
  if (frozen) {                                                      // particles.cpp:828-837
    v.x = fmul(v.x, 0.0f); v.y = fmul(v.y, 0.0f);
    f.x = fmul(f.x, 0.0f); f.y = fmul(f.y, 0.0f);
    v0.x = fmul(v0.x, 0.0f); v0.y = fmul(v0.y, 0.0f);  // velocity_tmp was saved after ApplyFrozen
  }
  if (a.force_out) a.force_out[i] = f;
  if (!a.write_state) return;

  const Phys& ph = a.ph;
  if (STAGE == 1) {                                                  // particles.cpp:109-123
    float vx = fadd(v.x, fmul(f.x, ph.c1));
    float vy = fadd(v.y, fmul(f.y, ph.c1));
    clamp_velocity(ph, vx, vy);
    const float xh = fadd(x.x, fmul(fmul(vx, ph.dt), 0.5f));
    const float yh = fadd(x.y, fmul(fmul(vy, ph.dt), 0.5f));
    a.pos_h[i] = make_float2(xh, yh);
    a.vel_h[i] = make_float2(vx, vy);
  } else {                                                           // particles.cpp:140-153
    float vx = fadd(v0.x, fmul(f.x, ph.c2));
    float vy = fadd(v0.y, fmul(f.y, ph.c2));
    clamp_velocity(ph, vx, vy);
    float2 xn = make_float2(fadd(x0.x, fmul(vx, ph.dt)), fadd(x0.y, fmul(vy, ph.dt)));
    float2 vn = make_float2(vx, vy);
    if (a.do_tail) {
      if ((FEAT & kFeatPortals) && (flags & 1u)) teleport(a, blk, xn, vn);
      key_and_count(a.g, a.dist, xn, vn, a.id, i, a.newcell, a.cnt);
    }
    a.pos_out[i] = xn;
    a.vel_out[i] = vn;
  }
}

// Standalone teleport + key pass: used when the block grid changed between the
// update and DetectPortals (particles.cpp:157-179) and for add_particles.
__global__ void __launch_bounds__(kThreads) tail_kernel(const __grid_constant__ TailArgs a) {
  const int i = a.first + blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.ctr->n_cur) return;
  float2 x = a.pos[i];
  if (a.do_teleport && a.n_sides) {
    const Cell c = cell_of(a.g, x);
    const int blk = block_of(a.g, c);
    if (c.valid && (a.blk_flags[blk] & 1u)) {
      float2 v = a.vel[i];
      teleport(a, blk, x, v);
      a.pos[i] = x;
      a.vel[i] = v;
    }
  }
  key_and_count(a.g, a.dist, x, a.vel[i], a.id, i, a.newcell, a.cnt);
}

// ---------------------------------------------------------------------------------
// single-pass exclusive scan (decoupled look-back), 1024 ints per tile
// ---------------------------------------------------------------------------------
constexpr int kScanItems = 16;  // consecutive ints per thread (4 x int4)
constexpr int kScanTile = kThreads * kScanItems;
int scan_tile_items() { return kScanTile; }

__global__ void __launch_bounds__(kThreads) scan_kernel(
    const int* __restrict__ cnt, int* __restrict__ start, int n_items, int ncell,
    unsigned long long* tile_state, int* ticket, unsigned epoch, Counters* ctr) {
  __shared__ int s_tile;
  __shared__ int s_warp[kThreads / 32];
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int base = tile * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; k += 4) {
    int4 q = make_int4(0, 0, 0, 0);
    if (base + k < n_items) q = *reinterpret_cast<const int4*>(cnt + base + k);  // n_items % 4 == 0
    v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
  }
  int tsum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { const int t = v[k]; v[k] = tsum; tsum += t; }  // exclusive in thread
  int incl = tsum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < kThreads / 32) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < kThreads / 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kThreads / 32) s_warp[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const int block_total = s_warp[kThreads / 32 - 1];
  const int excl_in_block = incl - tsum + (warp ? s_warp[warp - 1] : 0);

  // publish aggregate / look back.  state = (epoch*4 + flag) << 32 | value; flag 1 = this tile's
  // total, flag 2 = inclusive prefix.  Warp 0 looks at 32 predecessors per round: lane l reads
  // tile t - l, the nearest one that already holds an inclusive prefix ends the walk.
  const unsigned long long tag = (unsigned long long)epoch << 2;
  if (warp == 0) {
    volatile unsigned long long* st = tile_state;
    int prefix = 0;
    if (tile == 0) {
      if (lane == 0) st[0] = ((tag | 2ull) << 32) | (unsigned)block_total;
    } else {
      if (lane == 0) st[tile] = ((tag | 1ull) << 32) | (unsigned)block_total;
      int t = tile - 1;
      while (true) {
        const int idx = t - lane;
        unsigned long long sv = (tag | 2ull) << 32;  // before the first tile: inclusive prefix 0
        if (idx >= 0) {
          // not published yet: spin (tickets are handed out in order, so tile idx is running)
          do { sv = st[idx]; } while ((sv >> 32) != (tag | 1ull) && (sv >> 32) != (tag | 2ull));
        }
        const unsigned done = __ballot_sync(0xffffffffu, (sv >> 32) == (tag | 2ull));
        const int first = done ? __ffs(done) - 1 : 31;
        int c = lane <= first ? (int)(unsigned)sv : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        prefix += c;
        if (done) break;
        t -= 32;
      }
      if (lane == 0) st[tile] = ((tag | 2ull) << 32) | (unsigned)(prefix + block_total);
    }
    if (lane == 0) s_prefix = prefix;
  }
  __syncthreads();
  const int e0 = s_prefix + excl_in_block;
#pragma unroll
  for (int k = 0; k < kScanItems; k += 4) {
    if (base + k < n_items)
      *reinterpret_cast<int4*>(start + base + k) = make_int4(e0 + v[k], e0 + v[k + 1], e0 + v[k + 2], e0 + v[k + 3]);
  }
  // totals: start[ncell] = alive, start[ncell+1] = alive + trash
  if (ncell >= base && ncell < base + kScanItems) {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k == ncell) ctr->n_next = e0 + v[k];
  }
  if (ncell + 1 >= base && ncell + 1 < base + kScanItems) {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k == ncell + 1) ctr->n_total = e0 + v[k];
  }
}

// ---------------------------------------------------------------------------------
// counting-sort fill + deterministic gather
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) fill_kernel(const __grid_constant__ ReorderArgs a) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.ctr->n_cur + a.ctr->n_arrived) return;
  const int key = key_of_packed(a.newcell[i], a.SJ, a.ncell);
  // counting down leaves the histogram all-zero again for the next iteration; the first atomic
  // served takes the first slot of the run, so that the usual service order (ascending old
  // slot) already is the stable order gather_kernel wants
  const int left = atomicSub(a.cnt + key, 1);
  a.order[a.start[key + 1] - left] = i;
}

// New slot k receives the particle with the (k - start[c])-th smallest OLD slot among
// the particles of its sub-cell c: the result is the stable sort by key, independent of
// the order in which fill_kernel's atomics were served.  The payload of the particle
// fill_kernel put at k is loaded before that is checked (it is the right one unless the
// atomics of a shared cell were served out of order).
__global__ void __launch_bounds__(kThreads, 8) gather_kernel(const __grid_constant__ ReorderArgs a) {
  const int k = blockIdx.x * kThreads + threadIdx.x;
  if (k >= a.ctr->n_total) return;
  const int src0 = a.order[k];
  if (k >= a.ctr->n_next) {  // left the grid: removed (blocks.h:230-232)
    if (a.slot_of_id) a.slot_of_id[a.id_in[src0]] = -1;
    return;
  }
  const unsigned pc = a.newcell[src0];
  float2 p = a.pos_in[src0];
  float2 v = a.vel_in[src0];
  int pid = a.id_in[src0];
  const int c = key_of_packed(pc, a.SJ, a.ncell);
  const int beg = a.start[c], end = a.start[c + 1];
  if (end - beg > 1) {
    const int want = k - beg;
    const int n_loc = a.ctr->n_cur;
    auto vkey = [&](int i) -> unsigned long long {
      return i < n_loc ? ((unsigned long long)a.rank << 32) | (unsigned)i : a.arr_vkey[i - n_loc];
    };
    auto rank_of = [&](int cand) -> int {
      const unsigned long long ck = vkey(cand);
      int smaller = 0;
      for (int e2 = beg; e2 < end; ++e2) smaller += (vkey(a.order[e2]) < ck);
      return smaller;
    };
    if (rank_of(src0) != want) {
      for (int e = beg; e < end; ++e) {
        const int cand = a.order[e];
        if (cand != src0 && rank_of(cand) == want) {
          p = a.pos_in[cand]; v = a.vel_in[cand]; pid = a.id_in[cand];
          break;
        }
      }
    }
  }
  a.pos_out[k] = p;
  a.vel_out[k] = v;
  a.id_out[k] = pid;
  a.cell_out[k] = pc;
  if (a.slot_of_id) a.slot_of_id[pid] = k;
}

__global__ void finalize_kernel(Counters* ctr, int* ticket) {
  ctr->removed_total += ctr->n_total - ctr->n_next;
  ctr->n_cur = ctr->n_next;
  ctr->n_arrived = 0;
  *ticket = 0;
}

// Migrants received from the other strips are appended behind the local particles, keyed and
// counted, ready for the reorder; their (source rank, old slot) tags become their sort keys.
__global__ void __launch_bounds__(kThreads) arrivals_kernel(const __grid_constant__ ArrivalArgs a) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  const int src = t / a.in_cap, k = t - src * a.in_cap;
  if (src >= a.world) return;
  int off = 0, total = 0, mine = 0;
  for (int r = 0; r < a.world; ++r) {
    const int raw = *reinterpret_cast<const int*>(a.in_base + r * a.in_stride);
    if (raw > a.in_cap && t == 0) atomicOr(&a.ctr->err_flags, 8);  // a sender's outbox overflowed
    const int c = min(raw, a.in_cap);
    if (r < src) off += c;
    if (r == src) mine = c;
    total += c;
  }
  if (t == 0) a.ctr->n_arrived = total;
  if (k >= mine) return;
  const int n_loc = a.ctr->n_cur;
  const int dst = n_loc + off + k;
  if (dst >= a.cap) { atomicOr(&a.ctr->err_flags, 2); return; }
  const char* blk = a.in_base + src * a.in_stride;
  const float2 x = reinterpret_cast<const float2*>(blk + mig_off_pos())[k];
  a.pos[dst] = x;
  a.vel[dst] = reinterpret_cast<const float2*>(blk + mig_off_vel(a.in_cap))[k];
  a.id[dst] = reinterpret_cast<const int*>(blk + mig_off_id(a.in_cap))[k];
  a.arr_vkey[off + k] = ((unsigned long long)src << 32) |
                        (unsigned)reinterpret_cast<const int*>(blk + mig_off_tag(a.in_cap))[k];
  bool migrant;
  const unsigned pc = pack_cell(a.g, cell_of_fast(a.g, x), migrant);  // the sender routed it here
  a.newcell[dst] = pc;
  atomicAdd(a.cnt + key_of_packed(pc, a.g.SJ, a.g.ncell), 1);
}

// Copy the two boundary sub-rows on each side into staging buffers (positions; optionally the
// matching slice of start[], rebased to 0) for the neighbouring strips.
__global__ void __launch_bounds__(kThreads) ghost_pack_kernel(const __grid_constant__ GhostPackArgs a) {
  const int side = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  const int c0 = a.row_lo[side] * a.SJ;
  const int b0 = a.start[c0], b1 = a.start[c0 + 2 * a.SJ];
  if (t == 0 && b1 - b0 > a.ghost_cap) atomicOr(a.err, 4);
  if (t < b1 - b0 && t < a.ghost_cap) {
    a.out_pos[side][t] = a.X[b0 + t];
    if (a.out_id[side]) a.out_id[side][t] = a.id[b0 + t];
  }
  if (a.out_start[side] && t <= 2 * a.SJ) a.out_start[side][t] = a.start[c0 + t] - b0;
}

// ---------------------------------------------------------------------------------
// read-back helpers
// ---------------------------------------------------------------------------------
__global__ void find_blocks_kernel(GridDev g, const float2* pos, size_t n, long long* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Cell c = cell_of(g, pos[i]);
  out[i] = c.valid ? (long long)block_of(g, c) : -1ll;
}

__global__ void state_by_id_kernel(GridDev g, const Counters* ctr, const float2* pos, const float2* vel,
                                   const int* id, float2* pos_by_id, float2* vel_by_id, long long* block_by_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ctr->n_cur) return;
  const int pid = id[i];
  const float2 x = pos[i];
  if (pos_by_id) pos_by_id[pid] = x;
  if (vel_by_id) vel_by_id[pid] = vel[i];
  if (block_by_id) {
    const Cell c = cell_of(g, x);
    block_by_id[pid] = c.valid ? (long long)block_of(g, c) : -1ll;
  }
}

__global__ void scatter_by_id_kernel(const Counters* ctr, const float2* val, const int* id, float2* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ctr->n_cur) return;
  out[id[i]] = val[i];
}

__global__ void block_of_slot_kernel(GridDev g, const Counters* ctr, const float2* pos, int* block) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ctr->n_cur) return;
  const Cell c = cell_of(g, pos[i]);
  block[i] = c.valid ? block_of(g, c) : -1;
}

__global__ void slot_map_kernel(const Counters* ctr, const int* id, int* slot_of_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ctr->n_cur) return;
  slot_of_id[id[i]] = i;
}

__global__ void max_id_kernel(const int* id, size_t n, int* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int v = (i < n) ? id[i] : id[0];
  const int mx = __reduce_max_sync(0xffffffffu, v);
  const int mn = __reduce_min_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) { atomicMax(out, mx); atomicMin(out + 1, mn); }
}

__global__ void fill_i32_kernel(int* p, int v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
static inline unsigned grid_for(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

template <int FEAT>
static void launch_stage_feat(int stage, const StageArgs& a, int n_bound, cudaStream_t s) {
  if (stage == 1) stage_kernel<1, FEAT><<<grid_for(n_bound), kThreads, 0, s>>>(a);
  else stage_kernel<2, FEAT><<<grid_for(n_bound), kThreads, 0, s>>>(a);
}
void launch_stage(int stage, const StageArgs& a, int n_bound, cudaStream_t s) {
  if (n_bound <= 0) return;
  const int feat = (a.has_bonds ? kFeatBonds : 0) | (a.n_sides > 0 ? kFeatPortals : 0) |
                   ((a.ph.force_enabled || a.ph.pick_enabled) ? kFeatExtras : 0) |
                   (a.dist.world > 1 ? kFeatStrips : 0);
  switch (feat) {
    case 0: launch_stage_feat<0>(stage, a, n_bound, s); break;
    case 1: launch_stage_feat<1>(stage, a, n_bound, s); break;
    case 2: launch_stage_feat<2>(stage, a, n_bound, s); break;
    case 3: launch_stage_feat<3>(stage, a, n_bound, s); break;
    case 4: launch_stage_feat<4>(stage, a, n_bound, s); break;
    case 5: launch_stage_feat<5>(stage, a, n_bound, s); break;
    case 6: launch_stage_feat<6>(stage, a, n_bound, s); break;
    case 7: launch_stage_feat<7>(stage, a, n_bound, s); break;
    case 8: launch_stage_feat<8>(stage, a, n_bound, s); break;
    case 9: launch_stage_feat<9>(stage, a, n_bound, s); break;
    case 10: launch_stage_feat<10>(stage, a, n_bound, s); break;
    case 11: launch_stage_feat<11>(stage, a, n_bound, s); break;
    case 12: launch_stage_feat<12>(stage, a, n_bound, s); break;
    case 13: launch_stage_feat<13>(stage, a, n_bound, s); break;
    case 14: launch_stage_feat<14>(stage, a, n_bound, s); break;
    default: launch_stage_feat<15>(stage, a, n_bound, s); break;
  }
}
void launch_tail(const TailArgs& a, int n_bound, cudaStream_t s) {
  if (n_bound - a.first <= 0) return;
  tail_kernel<<<grid_for(n_bound - a.first), kThreads, 0, s>>>(a);
}
void launch_scan(const int* cnt, int* start, int n_items, int ncell, unsigned long long* tile_state,
                 int* ticket, unsigned epoch, Counters* ctr, cudaStream_t s) {
  const int tiles = (n_items + kScanTile - 1) / kScanTile;
  scan_kernel<<<tiles, kThreads, 0, s>>>(cnt, start, n_items, ncell, tile_state, ticket, epoch, ctr);
}
void launch_fill(const ReorderArgs& a, int n_bound, cudaStream_t s) {
  if (n_bound <= 0) return;
  fill_kernel<<<grid_for(n_bound), kThreads, 0, s>>>(a);
}
void launch_gather(const ReorderArgs& a, int n_bound, cudaStream_t s) {
  if (n_bound <= 0) return;
  gather_kernel<<<grid_for(n_bound), kThreads, 0, s>>>(a);
}
__global__ void __launch_bounds__(kThreads) ghost_map_kernel(const __grid_constant__ GhostMapArgs a) {
  const int side = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (a.ids[side] == nullptr) return;
  if (a.mark) {
    const int n = min(a.gstart[side][2 * a.SJ] - a.gstart[side][0], a.ghost_cap);
    if (t == 0) *a.count[side] = n;
    if (t < n) a.slot_of_id[a.ids[side][t]] = -2 - (side * a.ghost_cap + t);
  } else {
    // forget last iteration's ghosts (an id that became local meanwhile holds a slot >= 0: keep it)
    if (t < *a.count[side]) {
      const int pid = a.ids[side][t];
      if (a.slot_of_id[pid] <= -2) a.slot_of_id[pid] = -1;
    }
  }
}

void launch_ghost_map(const GhostMapArgs& a, cudaStream_t s) {
  ghost_map_kernel<<<dim3(grid_for((size_t)a.ghost_cap), 2), kThreads, 0, s>>>(a);
}
void launch_arrivals(const ArrivalArgs& a, cudaStream_t s) {
  arrivals_kernel<<<grid_for((size_t)a.world * a.in_cap), kThreads, 0, s>>>(a);
}
void launch_ghost_pack(const GhostPackArgs& a, cudaStream_t s) {
  const size_t n = (size_t)max(a.ghost_cap, 2 * a.SJ + 1);
  ghost_pack_kernel<<<dim3(grid_for(n), 2), kThreads, 0, s>>>(a);
}
void launch_finalize(Counters* ctr, int* ticket, cudaStream_t s) {
  finalize_kernel<<<1, 1, 0, s>>>(ctr, ticket);
}
void launch_find_blocks(const GridDev& g, const float2* pos, size_t n, long long* out, cudaStream_t s) {
  if (!n) return;
  find_blocks_kernel<<<grid_for(n), kThreads, 0, s>>>(g, pos, n, out);
}
void launch_state_by_id(const GridDev& g, const Counters* ctr, const float2* pos, const float2* vel,
                        const int* id, int n_bound, float2* pos_by_id, float2* vel_by_id,
                        long long* block_by_id, cudaStream_t s) {
  if (n_bound <= 0) return;
  state_by_id_kernel<<<grid_for(n_bound), kThreads, 0, s>>>(g, ctr, pos, vel, id, pos_by_id, vel_by_id, block_by_id);
}
void launch_scatter_by_id(const Counters* ctr, const float2* val, const int* id, int n_bound,
                          float2* out_by_id, cudaStream_t s) {
  if (n_bound <= 0) return;
  scatter_by_id_kernel<<<grid_for(n_bound), kThreads, 0, s>>>(ctr, val, id, out_by_id);
}
void launch_block_of_slot(const GridDev& g, const Counters* ctr, const float2* pos, int n_bound,
                          int* block, cudaStream_t s) {
  if (n_bound <= 0) return;
  block_of_slot_kernel<<<grid_for(n_bound), kThreads, 0, s>>>(g, ctr, pos, block);
}
void launch_slot_map(const Counters* ctr, const int* id, int n_bound, int* slot_of_id, cudaStream_t s) {
  if (n_bound <= 0) return;
  slot_map_kernel<<<grid_for(n_bound), kThreads, 0, s>>>(ctr, id, slot_of_id);
}
void launch_max_id(const int* id, size_t n, int* out, cudaStream_t s) {
  if (!n) return;
  max_id_kernel<<<grid_for(n), kThreads, 0, s>>>(id, n, out);
}
void launch_fill_i32(int* p, int v, size_t n, cudaStream_t s) {
  if (!n) return;
  fill_i32_kernel<<<grid_for(n), kThreads, 0, s>>>(p, v, n);
}

}  // namespace ptoy

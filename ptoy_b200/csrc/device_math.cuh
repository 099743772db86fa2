// Device helpers shared by the kernels: the reference's arithmetic (every operation issued as the
// same single IEEE binary32 operation, see kernels.cu), binning, force laws, the integrator's
// clamp, portal teleport and the re-keying tail.
#pragma once
#include "common.h"

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
// Sort key of a particle this strip keeps: local row * SJ + column; g.ncell (the trash bin) if it
// left the grid or belongs to another strip (`migrant`)
__device__ __forceinline__ int cell_key(const GridDev& g, const Cell& c, bool& migrant) {
  migrant = false;
  if (!c.valid) return g.ncell;
  const int lr = c.Si - g.row0;
  if (lr < g.own_lo || lr >= g.own_hi) { migrant = true; return g.ncell; }
  return lr * g.SJ + c.Sj;
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
constexpr int kBondPlan = 6;     // bond partners per particle held in the fixed-width bond table / handed from stage 1 to stage 2

// FEAT selects what the scene uses, so that unused force kinds cost neither instructions nor
// registers: 1 = bonds, 2 = portals, 4 = mouse point force / pick; launch_stage picks the variant.
constexpr int kFeatBonds = 1, kFeatPortals = 2, kFeatExtras = 4;

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
                                              const int* __restrict__ id, int slot, int* newcell, int* cnt) {
  const Cell c = cell_of_fast(g, x);
  bool migrant;
  const int key = cell_key(g, c, migrant);
  if (migrant && d.out_base) {
    int dest = 0;
    while (dest + 1 < d.world && c.Si >= d.row_begin[dest + 1]) ++dest;
    char* blk = d.out_base + dest * d.out_stride;
    const int k = atomicAdd(reinterpret_cast<int*>(blk), 1);
    if (k < d.out_cap) {
      reinterpret_cast<float2*>(blk + mig_off_pos())[k] = x;
      reinterpret_cast<float2*>(blk + mig_off_vel(d.out_cap))[k] = v;
      const int pid = id[slot];
      reinterpret_cast<int*>(blk + mig_off_id(d.out_cap))[k] = pid;
      reinterpret_cast<int*>(blk + mig_off_tag(d.out_cap))[k] = slot;
      if (d.bond_ell) {   // bonds travel with their particle (rows longer than the table's 6 do not: flagged)
        const int4* row = reinterpret_cast<const int4*>(d.bond_ell) + 2 * (size_t)pid;
        int4* out = reinterpret_cast<int4*>(blk + mig_off_row(d.out_cap)) + 2 * (size_t)k;
        const int4 r0 = row[0], r1 = row[1];
        out[0] = r0;
        out[1] = r1;
        if (r1.w > 6) atomicOr(d.err, 32);
      }
    }
  }
  newcell[slot] = key;
  atomicAdd(cnt + key, 1);  // result unused -> RED.ADD
}

}  // namespace ptoy

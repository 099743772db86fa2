// Tile-staged stage kernels (sm_100a): one CTA owns kTile consecutive particles of the sub-cell
// sorted arrays and pulls the part of its neighbourhood every thread needs into shared memory
// with the bulk-copy engine (cp.async.bulk + mbarrier; SASS: UBLKCP, SYNCS).
//
// Why a tile is cheap to describe: the sort key is row-major (key = Si * SJ + Sj), so the
// candidates of a particle in cell key c at offset (dr, dc) live in cell key c + dr * SJ + dc
// (the two padding columns on either side of a sub-row are always empty, so this never wraps into
// a populated cell).  A tile whose particles span the keys [cA, cB] therefore needs, for each of
// the three rows dr = -1, 0, +1, the keys [cA + dr*SJ - 2, cB + dr*SJ + 2]: ONE contiguous range
// of start[] (the cell offsets) and ONE contiguous range of the position array per row — also
// when the tile runs across the end of a sub-row.  Six bulk copies per tile, issued by one lane,
// land in shared memory while the other warps load their own state and bond rows; the candidate
// search then reads cell offsets and candidate positions with LDS instead of going through the
// L1 tag stage.  The rare fourth / fifth row (a particle within the search margin of a cell
// edge) and tiles that do not fit the staging buffers (very sparse or very dense
// spots) take the global-memory path of the same search.
//
// Reference: the 9-block stencil of particles.cpp:884-900 + CalcForceSerial :598-606, here over
// the sub-cell grid; forces summed in the same order of kinds (:839-910).
#pragma once
#include "device_math.cuh"

#ifndef PTOY_TILE
#define PTOY_TILE 256
#endif
#ifndef PTOY_TILE_BUFS        // pipeline slots per CTA (stage 1 / stage 2)
#define PTOY_TILE_BUFS 2
#endif
#ifndef PTOY_TILE_BUFS2
#define PTOY_TILE_BUFS2 2
#endif
#ifndef PTOY_TILE_CTAS        // resident CTAs per SM the kernels are compiled and launched for
#define PTOY_TILE_CTAS 4
#endif
#ifndef PTOY_TILE_CTAS_BONDS
#define PTOY_TILE_CTAS_BONDS 4
#endif

namespace ptoy {

constexpr int kTile = PTOY_TILE;              // particles per tile = consumer threads per CTA
constexpr int kTileThreads = kTile + 32;      // + one producer warp
#ifndef PTOY_TILE_WINCAP
#define PTOY_TILE_WINCAP (PTOY_TILE + 80)
#endif
#ifndef PTOY_TILE_KEYCAP
#define PTOY_TILE_KEYCAP (PTOY_TILE + 128)
#endif
constexpr int kWinCap = PTOY_TILE_WINCAP;     // staged particles per row window
constexpr int kKeyCap = PTOY_TILE_KEYCAP;     // staged start[] entries per row window

// One pipeline slot: everything a tile's consumer threads read, filled by the bulk-copy engine.
template <int STAGE>
struct TileBuf {
  alignas(16) float2 pos[3][kWinCap];   // row windows dr = -1, 0, +1 of x0 (stage 1) / x_half (stage 2)
  alignas(16) int start[3][kKeyCap];    // the matching slices of start[]
  alignas(16) float2 vel[kTile];        // the tile's own v0 (stage 1) / v_half (stage 2)
  alignas(16) int key[kTile];           // the tile's own sort keys
  alignas(16) int id[kTile];            // ... ids (when the scene needs them)
  alignas(16) float2 pos0[STAGE == 2 ? kTile : 2];   // stage 2: x0, v0
  alignas(16) float2 vel0[STAGE == 2 ? kTile : 2];
  // descriptor, written by the producer (three 16-byte stores) before it arms the barrier
  alignas(16) int k0;   // the tile's particles [k0, k1)
  int k1;
  int staged;           // row windows are in shared memory (else: the search reads global memory)
  int cA;               // sort key of the tile's first particle
  alignas(16) int soff[4];   // start[w][soff[w] + (key - cA) + dc] = start[key + (w - 1) * SJ + dc]
  alignas(16) int poff[4];   // pos[w][slot + poff[w]] = position of `slot`
};

template <int STAGE>
struct TileSmem {
  static constexpr int kBufs = STAGE == 1 ? PTOY_TILE_BUFS : PTOY_TILE_BUFS2;
  TileBuf<STAGE> buf[kBufs];
  alignas(8) unsigned long long full[kBufs], empty[kBufs];
};

// ---- mbarrier / bulk copy (PTX) --------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in hardware, do not spin
      : "memory");
  return ok != 0;
}
// global -> shared, 16-byte aligned on both sides, size a multiple of 16; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Candidates [b, e) of one row window against particle x: phase 1 only TESTS them (a cheap
// conservative prefilter, fma(dx,dx,dy*dy) against r2c grown by a few ulp; no branch in the loop
// body, the outcome goes into a bit mask), phase 2 walks the set bits in candidate order and
// evaluates those pairs with the reference's exact arithmetic (particles.cpp:584-595; the exact
// r^2 decides).  Contributions are summed in slot order, like the plain loop would.
// `self`: slot of the particle itself when the window is its own row (the reference skips it by
// address, particles.cpp:601), a slot outside [b, e) otherwise.
constexpr int kNoSelf = -0x40000000;

#ifndef PTOY_PAIR_V2
#define PTOY_PAIR_V2 1
#endif
// Packed fp32 pairs (sm_100 FADD2 / FMUL2): one issue slot for the x and the y lane, each lane
// rounded to nearest exactly like the scalar instruction.  Only differences and products that are
// NOT followed by an add of the product are packed: ptxas contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2, which the scalar .rn forms never allow (build.py checks the SASS for FFMA2).
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
// Correctly rounded 1/x for x in [2^-126, 2^126): the in-range path of __frcp_rn (MUFU.RCP and one
// fused Newton step, the same three instructions) without its exponent test and out-of-range call.
// The pair search only divides by max(threshold, r2) with threshold = 4e-7 and r2 < 1.7e-3.
__device__ __forceinline__ float rcp_exact_inrange(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  const float e = __fmaf_rn(x, y, -1.0f);
  return __fmaf_rn(y, -e, y);
}

#if PTOY_PAIR_V2
__device__ __forceinline__ void row_scan(const float2* __restrict__ P, int b, int e, int self, float2 x,
                                         const Phys& ph, float& fx, float& fy) {
  const float thr = ph.r2c_hi, r2c = ph.r2c, floor_r2 = ph.thr, RR = ph.RR;
  while (b < e) {
    const int m = min(e - b, 32);
    const float2* __restrict__ Pb = P + b;
    // ---- phase 1: which of the next m candidates are (about) inside the cutoff: bit k = candidate k.
    // The mask is filled from the top (two bits per turn) and shifted down at the end.
    unsigned mask = 0u;
    const float2* __restrict__ q = Pb;
    const float2* const qe = Pb + (m & ~1);
#pragma unroll 1
    for (; q != qe; q += 2) {
      const float2 d0 = sub2(x, q[0]), d1 = sub2(x, q[1]);
      const float r0 = __fmaf_rn(d0.x, d0.x, d0.y * d0.y), r1 = __fmaf_rn(d1.x, d1.x, d1.y * d1.y);
      mask = (mask >> 2) | (r0 < thr ? 0x40000000u : 0u) | (r1 < thr ? 0x80000000u : 0u);
    }
    if (m & 1) {
      const float2 d0 = sub2(x, q[0]);
      mask = (mask >> 1) | (__fmaf_rn(d0.x, d0.x, d0.y * d0.y) < thr ? 0x80000000u : 0u);
    }
    mask >>= 32 - m;
    const unsigned d = (unsigned)(self - b);
    if (d < 32u) mask &= ~(1u << d);
    // ---- phase 2: exact arithmetic for the set bits (highest first) ---------------------------
    while (mask) {
      unsigned h;
      asm("bfind.u32 %0, %1;" : "=r"(h) : "r"(mask));     // index of the highest set bit
      mask ^= 1u << h;
      const float2 dd = sub2(x, Pb[h]);
      const float2 sq = mul2(dd, dd);
      const float r2 = fadd(sq.x, sq.y);                     // dot2(d, d)
      const float r2t = (floor_r2 < r2) ? r2 : floor_r2;     // std::max(threshold, r2), pair_force()
      const float inv = rcp_exact_inrange(r2t);
      const float d2 = fmul(inv, RR);
      const float d6 = fmul(fmul(d2, d2), d2);
      const float d12 = fmul(d6, d6);
      const float kk = fmul(fsub(d12, d6), inv);
      const float2 pf = mul2(dd, make_float2(kk, kk));
      if (r2 < r2c) {   // (phase 1 lets through a hair more than the cutoff: predicated adds, no branch)
        fx = fadd(fx, pf.x);
        fy = fadd(fy, pf.y);
      }
    }
    b += m;
  }
}
#else
__device__ __forceinline__ void row_scan(const float2* __restrict__ P, int b, int e, int self, float2 x,
                                         const Phys& ph, float& fx, float& fy) {
  const float thr = ph.r2c_hi;
  while (b < e) {
    const int m = min(e - b, 32);
    unsigned mask = 0u, bit = 1u;
    int k = 0;
#pragma unroll 1
    for (; k + 1 < m; k += 2) {
      const float2 p0 = P[b + k], p1 = P[b + k + 1];
      const float dx0 = fsub(x.x, p0.x), dy0 = fsub(x.y, p0.y);
      const float dx1 = fsub(x.x, p1.x), dy1 = fsub(x.y, p1.y);
      const float r0 = __fmaf_rn(dx0, dx0, dy0 * dy0), r1 = __fmaf_rn(dx1, dx1, dy1 * dy1);
      if (r0 < thr) mask |= bit;
      if (r1 < thr) mask |= bit << 1;
      bit <<= 2;
    }
    if (k < m) {
      const float2 p0 = P[b + k];
      const float dx0 = fsub(x.x, p0.x), dy0 = fsub(x.y, p0.y);
      if (__fmaf_rn(dx0, dx0, dy0 * dy0) < thr) mask |= bit;
    }
    const unsigned d = (unsigned)(self - b);
    if (d < 32u) mask &= ~(1u << d);
    while (mask) {
      const int h = __ffs(mask) - 1;
      mask &= mask - 1u;
      const float2 pq = P[b + h];
      const float dx = fsub(x.x, pq.x), dy = fsub(x.y, pq.y);
      const float r2 = dot2(dx, dy, dx, dy);
      if (r2 < ph.r2c) {
        float px, py;
        pair_force(ph, dx, dy, r2, px, py);
        fx = fadd(fx, px);
        fy = fadd(fy, py);
      }
    }
    b += m;
  }
}
#endif

// The pair search of one particle over its 3 (5) row windows.  s0..s2 point at the start[] entry
// of the particle's own column in row window -1, 0, +1 (s[dc] = first slot of column dc),
// P0..P2[slot] is the position of `slot` in that window; both in shared memory for a staged
// tile, in global memory otherwise.  Rows +-2 (a particle near a cell edge in x) come from
// global memory.
__device__ __forceinline__ void pair_search(const int* __restrict__ s0, const int* __restrict__ s1, const int* __restrict__ s2,
                                            const float2* __restrict__ P0, const float2* __restrict__ P1,
                                            const float2* __restrict__ P2, const float2* __restrict__ X,
                                            const int* __restrict__ gs, int SJ, bool near_i, bool near_j, int self,
                                            float2 x, const Phys& ph, float& fx, float& fy) {
  const int dlo = -1 - (int)near_j, dhi = 2 + (int)near_j;   // columns [dlo, dhi)
  const int b0 = s0[dlo], e0 = s0[dhi], b1 = s1[dlo], e1 = s1[dhi], b2 = s2[dlo], e2 = s2[dhi];
  row_scan(P0, b0, e0, kNoSelf, x, ph, fx, fy);
  row_scan(P1, b1, e1, self, x, ph, fx, fy);
  row_scan(P2, b2, e2, kNoSelf, x, ph, fx, fy);
  if (near_i) {   // the rare fourth and fifth row
    row_scan(X, __ldg(gs - 2 * SJ + dlo), __ldg(gs - 2 * SJ + dhi), kNoSelf, x, ph, fx, fy);
    row_scan(X, __ldg(gs + 2 * SJ + dlo), __ldg(gs + 2 * SJ + dhi), kNoSelf, x, ph, fx, fy);
  }
}

// ---- force kinds other than the pair search (same code paths as kernels.cu particle_force) -----
template <int FEAT>
__device__ __forceinline__ void force_head(const StageArgs& a, int pid, float2 x, float2 v, unsigned flags, float& fx, float& fy) {
  const Phys& ph = a.ph;
  fx = 0.0f; fy = 0.0f;
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
}

__device__ __forceinline__ void force_walls(const StageArgs& a, float2 x, float& fx, float& fy) {  // :903-909
  if (!(x.x > a.free_lox && x.x < a.free_hix && x.y > a.free_loy && x.y < a.free_hiy)) {
    for (int k = 0; k < a.n_walls; ++k) {
      const WallDev w = a.walls[k];
      if (x.x >= w.lox && x.x <= w.hix && x.y >= w.loy && x.y <= w.hiy) {
        float qx, qy, wx, wy;
        seg_nearest(w.ax, w.ay, w.bx, w.by, x.x, x.y, qx, qy);
        lj_noclamp(a.ph.RRw, x.x, x.y, qx, qy, wx, wy);
        fx = fadd(fx, wx);
        fy = fadd(fy, wy);
      }
    }
  }
}

// one directed bond (particles.cpp:761-826): direct, or through a portal when stretched (:784-819)
template <int FEAT>
__device__ __forceinline__ void force_one_bond(const StageArgs& a, float2 x, int sb, float2 q, float& fx, float& fy) {
  const Phys& ph = a.ph;
  if (sb == kSlotNone) return;  // partner left the grid: bond erased by CheckBonds (:205-215)
  float bx, by;
  const float ddx = fsub(x.x, q.x), ddy = fsub(x.y, q.y);
  const float dist = len2(ddx, ddy);
  bool found = false;
  if ((FEAT & kFeatPortals) && !(dist < ph.bond_cutoff)) {
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
  if (!found) {                                                              // :727-736 inlined
    float k = fsub(1.0f, div_by_bondR(dist, ph.bondR, ph.bond_y));
    k = (k < -4.0f) ? -4.0f : k;
    const float sk = fmul(ph.sigma_bond, k);
    bx = fmul(ddx, sk);
    by = fmul(ddy, sk);
  }
  fx = fadd(fx, bx);
  fy = fadd(fy, by);
}

// cross-portal forces (particles.cpp:318-381) for a particle in reference block `blk`.
// The particles of a listed block are read in place (one strip) or from the gathered zone table
// (strips: PortalZoneArgs) - in both cases in the order sub-row, slot.
template <class F>
__device__ __forceinline__ void for_block_particles(const StageArgs& a, const float2* __restrict__ X, int l, F&& f) {
  if (a.zcnt) {
    const int n = a.zcnt[l];
    const float2* P = X + a.zoff + (size_t)l * kZoneCap;
    for (int k = 0; k < n; ++k) f(P[k]);
  } else {
    int r0, r1, c0, c1;
    block_extent(a.g, a.pblk[l], r0, r1, c0, c1);
    for (int R = r0; R <= r1; ++R) {
      const int beg = a.start[R * a.g.SJ + c0], end = a.start[R * a.g.SJ + c1 + 1];
      for (int q = beg; q < end; ++q) f(X[q]);
    }
  }
}

__device__ __forceinline__ void force_portal_cross(const StageArgs& a, const float2* __restrict__ X, float2 x, int blk, float& fx, float& fy) {
  const Phys& ph = a.ph;
  for (int s = 0; s < a.n_sides; ++s) {
    const int lb = a.pblk_ptr[s], le = a.pblk_ptr[s + 1];
    if (!in_sorted(a.pblk + lb, le - lb, blk)) continue;
    const PortalSide po = a.sides[s];
    const PortalSide ot = a.sides[s ^ 1];
    float lambda_c, offset_c;
    portal_coords(po, x.x, x.y, lambda_c, offset_c);
    if (!(lambda_c > 0.0f && lambda_c < 1.0f)) continue;
    for (int l = lb; l < le; ++l) {                                          // :345-355
      for_block_particles(a, X, l, [&](float2 nb) {
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
      });
    }
    const int ob = a.pblk_ptr[s ^ 1], oe = a.pblk_ptr[(s ^ 1) + 1];               // :358-375
    const float sign = (offset_c > 0.0f) ? 1.0f : -1.0f;
    for (int l = ob; l < oe; ++l) {
      for_block_particles(a, X, l, [&](float2 nb) {
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
      });
    }
  }
}

template <int STAGE, int FEAT>
__global__ void __launch_bounds__(kTileThreads, (FEAT & kFeatBonds) ? PTOY_TILE_CTAS_BONDS : PTOY_TILE_CTAS)
stage_tile_kernel(const __grid_constant__ StageArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<STAGE>& sm = *reinterpret_cast<TileSmem<STAGE>*>(smem_raw);
  const int t = threadIdx.x;
  const int n = a.ctr->n_cur;
  const float2* __restrict__ X = (STAGE == 1) ? a.pos : (const float2*)a.pos_h;
  const int SJ = a.g.SJ;

  constexpr int NB = TileSmem<STAGE>::kBufs;
  if (t == 0) {
    for (int b = 0; b < NB; ++b) { mbar_init(&sm.full[b], 1); mbar_init(&sm.empty[b], kTile / 32); }
  }
  __syncthreads();

  if (t >= kTile) {
    // ================= producer warp: describe tile, start its bulk copies ====================
    // One copy per lane: lanes 0-2 the start[] slices of row windows -1, 0, +1, lanes 3-5 their
    // position windows, 6 the tile's velocities, 7 sort keys, 8 ids, 9-11 (stage 2) x0, v0 and the
    // neighbour lists.  Everything about tile T is worked out BEFORE waiting for its slot, so
    // that what follows the wait is: descriptor store, arm the barrier, one copy per lane.  The
    // loads it needs are software-pipelined over the tile sequence: while T is described from
    // registers, the start[] bounds of T+1 and the first / last sort key of T+2 are on their way.
    const int lane = t - kTile;
    const int w = lane % 3;
    const int G = gridDim.x;
    auto load_keys = [&](int tile, int& cA, int& cB) {
      cA = 0; cB = 0;
      if (tile * kTile < n) {
        cA = __ldg(a.cell + tile * kTile);
        cB = __ldg(a.cell + min(tile * kTile + kTile, n) - 1);
      }
    };
    // window w of a tile spanning keys [cA, cB]: keys [kb, kb + nkeys], nkeys = cB - cA + 5 (columns -2..+2,
    // one past the last key for the end offset); first key >= 0 because rows >= 2, columns >= 2
    auto load_bounds = [&](int tile, int cA, int cB, int& sb, int& se) {
      sb = 0; se = 0;
      const int kb = cA + (w - 1) * SJ - 2, nkeys = cB - cA + 5;
      if (lane < 6 && tile * kTile < n && ((kb + nkeys + 1 - (kb & ~3) + 3) & ~3) <= kKeyCap) {
        sb = __ldg(a.start + kb);
        se = __ldg(a.start + kb + nkeys);
      }
    };
    int tile0 = blockIdx.x;
    int cA0, cB0, cA1, cB1, cA2 = 0, cB2 = 0, sb0, se0, sb1 = 0, se1 = 0;
    load_keys(tile0, cA0, cB0);
    load_keys(tile0 + G, cA1, cB1);
    load_bounds(tile0, cA0, cB0, sb0, se0);
    for (int it = 0; tile0 * kTile < n; ++it) {
      load_bounds(tile0 + G, cA1, cB1, sb1, se1);
      load_keys(tile0 + 2 * G, cA2, cB2);
      const int b = it % NB;
      TileBuf<STAGE>& B = sm.buf[b];
      // ---- describe the tile ---------------------------------------------------------------
      const int k0 = tile0 * kTile, k1 = min(k0 + kTile, n);
      const int nkeys = cB0 - cA0 + 5;
      const int kb = cA0 + (w - 1) * SJ - 2;
      const int kb_al = kb & ~3;
      const int nk = (kb + nkeys + 1 - kb_al + 3) & ~3;   // start[] entries to stage
      const int pb = sb0 & ~1;
      const int np = ((se0 + 1) & ~1) - pb;
      const bool staged = __all_sync(0xffffffffu, lane >= 6 || (nk <= kKeyCap && np <= kWinCap));
      const unsigned own_f2 = (unsigned)(((k1 - k0) + 1) & ~1) * 8u;   // float2 tiles
      const unsigned own_i = (unsigned)(((k1 - k0) + 3) & ~3) * 4u;    // int tiles
      const void* src = nullptr;
      void* dst = nullptr;
      unsigned bytes = 0u;
      if (lane < 3) { src = a.start + kb_al; dst = &B.start[w][0]; bytes = staged ? (unsigned)nk * 4u : 0u; }
      else if (lane < 6) { src = X + pb; dst = &B.pos[w][0]; bytes = staged ? (unsigned)np * 8u : 0u; }
      else if (lane == 6) { src = (STAGE == 1 ? a.vel : (const float2*)a.vel_h) + k0; dst = &B.vel[0]; bytes = own_f2; }
      else if (lane == 7) { src = a.cell + k0; dst = &B.key[0]; bytes = own_i; }
      else if (lane == 8) { src = a.id + k0; dst = &B.id[0]; bytes = a.need_id ? own_i : 0u; }
      else if (STAGE == 2 && lane == 9) { src = a.pos + k0; dst = &B.pos0[0]; bytes = own_f2; }
      else if (STAGE == 2 && lane == 10) { src = a.vel + k0; dst = &B.vel0[0]; bytes = own_f2; }
      const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
      const int ko = kb - kb_al + 2;
      const int4 d0 = make_int4(k0, k1, (int)staged, cA0);
      const int4 d1 = make_int4(__shfl_sync(0xffffffffu, ko, 0), __shfl_sync(0xffffffffu, ko, 1),
                                __shfl_sync(0xffffffffu, ko, 2), 0);
      const int4 d2 = make_int4(-__shfl_sync(0xffffffffu, pb, 3), -__shfl_sync(0xffffffffu, pb, 4),
                                -__shfl_sync(0xffffffffu, pb, 5), 0);
      // ---- wait for the slot, publish, copy -----------------------------------------------------
      while (!mbar_try_wait(&sm.empty[b], ((it / NB) & 1) ^ 1)) {}
      if (lane == 0) {
        *reinterpret_cast<int4*>(&B.k0) = d0;
        *reinterpret_cast<int4*>(&B.soff[0]) = d1;
        *reinterpret_cast<int4*>(&B.poff[0]) = d2;
        mbar_arrive_expect_tx(&sm.full[b], total);
      }
      __syncwarp();
      if (bytes) bulk_g2s(dst, src, bytes, &sm.full[b]);
      tile0 += G;
      cA0 = cA1; cB0 = cB1; sb0 = sb1; se0 = se1;
      cA1 = cA2; cB1 = cB2;
    }
    return;
  }

  // ======================= consumer warps: one particle per thread =============================
  const Phys& ph = a.ph;
  // search margin: stage 2 searches around the binning-time cell, so it widens by
  // twice the largest half-step drift any particle made (measured by stage 1, per component)
  float margin = a.margin;
  if (STAGE == 2 && a.dmax) margin = fminf(a.margin, a.margin_base + a.margin_scale * __uint_as_float(*a.dmax));
  const float margin_hi = 1.0f - margin;
  float drift_max = 0.0f;

  int it = 0;
  for (int tile = blockIdx.x; tile * kTile < n; tile += gridDim.x, ++it) {
    const int b = it % NB;
    TileBuf<STAGE>& B = sm.buf[b];
    while (!mbar_try_wait(&sm.full[b], (it / NB) & 1)) {}
    const int i = B.k0 + t;
    if (i < B.k1) {
      // ---- own state -----------------------------------------------------------------------
      const bool staged = B.staged != 0;
      float2 x0, v0, x, v;
      if (STAGE == 1) {
        x0 = staged ? B.pos[1][i + B.poff[1]] : a.pos[i];
        v0 = B.vel[t];
        x = x0; v = v0;
      } else {
        x0 = B.pos0[t]; v0 = B.vel0[t];
        x = staged ? B.pos[1][i + B.poff[1]] : a.pos_h[i];
        v = B.vel[t];
      }
      const int ck = B.key[t];
      const int pid = a.need_id ? B.id[t] : 0;
      // fractional position inside the sub-cell, from the division-free coordinate: off by less
      // than g.tol, and possibly by one whole cell next to an edge - then it reads ~0 or ~1 and
      // the particle counts as near the edge on both sides, which is all the margin test needs
      const float wi = fsub(x0.x, a.g.Ax) * a.g.inv_hsx, wj = fsub(x0.y, a.g.Ay) * a.g.inv_hsy;
      const float ui = wi - floorf(wi), uj = wj - floorf(wj);
      const bool near_i = !(ui >= margin && ui <= margin_hi);
      const bool near_j = !(uj >= margin && uj <= margin_hi);
      int blk = 0;
      unsigned flags = 0u;
      if ((FEAT & kFeatPortals) && a.blk_flags) {
        const int Si = ck / SJ, Sj = ck - Si * SJ;
        blk = block_of_s(a.g, Si + a.g.row0, Sj);
        flags = a.blk_flags[blk];
      }

      float fx, fy;
      force_head<FEAT>(a, pid, x, v, flags, fx, fy);

      // ---- bonds, part 1: partner slots, resolved before the pair search so that the dependent
      // loads (id -> bond row -> id->slot map) overlap it; the partners' positions are only
      // prefetched into L1 here and read after the search, which keeps the search's registers free.
      int nb = 0;
      int bslot[kBondPlan];
      if (FEAT & kFeatBonds) {
        int4* hand = reinterpret_cast<int4*>(a.bslot) + 2 * (size_t)i;
        if (STAGE == 1) {
          const int4* row = reinterpret_cast<const int4*>(a.bond_ell) + 2 * (size_t)pid;
          const int4 r0 = __ldg(row), r1 = __ldg(row + 1);
          const int bcol[kBondPlan] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
          nb = r1.w;
#pragma unroll
          for (int k = 0; k < kBondPlan; ++k)
            bslot[k] = ((unsigned)bcol[k] < (unsigned)a.id_cap) ? a.slot_of_id[bcol[k]] : kSlotNone;
          if (a.dist.world > 1) {  // strips: a partner that is neither local nor in a ghost row cannot be served
            bool lost = false;
#pragma unroll
            for (int k = 0; k < kBondPlan; ++k) lost |= bcol[k] >= 0 && bslot[k] == kSlotNone;
            if (lost && !(atomicOr(a.err, 16) & 16)) {   // the first one leaves a trace for the error message
              int kk = 0;
#pragma unroll
              for (int k = kBondPlan - 1; k >= 0; --k) if (bcol[k] >= 0 && bslot[k] == kSlotNone) kk = k;
              a.lost[0] = pid; a.lost[1] = bcol[kk]; a.lost[2] = i; a.lost[3] = n;
            }
          }
          hand[0] = make_int4(bslot[0], bslot[1], bslot[2], bslot[3]);
          hand[1] = make_int4(bslot[4], bslot[5], kSlotNone, nb);
        } else {
          const int4 r0 = hand[0], r1 = hand[1];
          bslot[0] = r0.x; bslot[1] = r0.y; bslot[2] = r0.z; bslot[3] = r0.w; bslot[4] = r1.x; bslot[5] = r1.y;
          nb = r1.w;
        }
#pragma unroll
        for (int k = 0; k < kBondPlan; ++k)
          if (bslot[k] != kSlotNone) asm volatile("prefetch.global.L1 [%0];" ::"l"(X + bslot[k]));
      }

      // ---- pair forces over the sub-cell neighbourhood -------------------------------------------
      if (staged) {
        const int cr = ck - B.cA;
        pair_search(&B.start[0][B.soff[0] + cr], &B.start[1][B.soff[1] + cr], &B.start[2][B.soff[2] + cr],
                    &B.pos[0][0] + B.poff[0], &B.pos[1][0] + B.poff[1], &B.pos[2][0] + B.poff[2],
                    X, a.start + ck, SJ, near_i, near_j, i, x, ph, fx, fy);
      } else {
        const int* gs = a.start + ck;
        pair_search(gs - SJ, gs, gs + SJ, X, X, X, X, gs, SJ, near_i, near_j, i, x, ph, fx, fy);
      }

      force_walls(a, x, fx, fy);

      // ---- bonds, part 2 (:761-826): CSR row of this id, partners ascending --------------------
      if (FEAT & kFeatBonds) {
#pragma unroll
        for (int k = 0; k < kBondPlan; ++k)
          force_one_bond<FEAT>(a, x, bslot[k], bslot[k] != kSlotNone ? __ldg(X + bslot[k]) : make_float2(0.f, 0.f), fx, fy);
        if (nb > kBondPlan) {  // more partners than the plan holds: walk the rest of the CSR row
          const unsigned e0 = __ldg(a.bond_ptr + pid), ne = __ldg(a.bond_ptr + pid + 1) - e0;
          for (unsigned e = kBondPlan; e < ne; ++e) {
            const int pc = a.bond_col[e0 + e];
            const int sb = ((unsigned)pc < (unsigned)a.id_cap) ? a.slot_of_id[pc] : kSlotNone;
            force_one_bond<FEAT>(a, x, sb, sb != kSlotNone ? __ldg(X + sb) : make_float2(0.f, 0.f), fx, fy);
          }
        }
      }

      if ((FEAT & kFeatPortals) && (flags & 1u)) force_portal_cross(a, X, x, blk, fx, fy);

      // ---- ApplyFrozen + integrator --------------------------------------------------------------
      float2 f = make_float2(fx, fy);
      bool frozen = false;
      if (a.has_frozen && (unsigned)pid < (unsigned)a.id_cap) frozen = (a.frozen_bits[pid >> 5] >> (pid & 31)) & 1u;
      if (frozen) {                                                      // particles.cpp:828-837
        v.x = fmul(v.x, 0.0f); v.y = fmul(v.y, 0.0f);
        f.x = fmul(f.x, 0.0f); f.y = fmul(f.y, 0.0f);
        v0.x = fmul(v0.x, 0.0f); v0.y = fmul(v0.y, 0.0f);  // velocity_tmp was saved after ApplyFrozen
      }
      if (a.force_out) a.force_out[i] = f;
      if (a.write_state) {
        if (STAGE == 1) {                                                  // particles.cpp:109-123
          float vx = fadd(v.x, fmul(f.x, ph.c1));
          float vy = fadd(v.y, fmul(f.y, ph.c1));
          clamp_velocity(ph, vx, vy);
          const float xh = fadd(x.x, fmul(fmul(vx, ph.dt), 0.5f));
          const float yh = fadd(x.y, fmul(fmul(vy, ph.dt), 0.5f));
          a.pos_h[i] = make_float2(xh, yh);
          a.vel_h[i] = make_float2(vx, vy);
          drift_max = fmaxf(drift_max, fmaxf(fabsf(fsub(xh, x.x)), fabsf(fsub(yh, x.y))));   // fmaxf drops NaN
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
    }
    __syncwarp();
    if ((t & 31) == 0) mbar_arrive(&sm.empty[b]);
  }
  if (STAGE == 1 && a.dmax && a.write_state) {
    // largest half-step drift of any particle, per component (non-negative floats order as uints)
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(drift_max));
    if ((t & 31) == 0 && m != 0u) atomicMax(a.dmax, m);
  }
}

template <int STAGE, int FEAT>
static void launch_stage_tile_one(const StageArgs& a, int n_bound, cudaStream_t s) {
  static int sms_of[64] = {0};   // per device: SM count, and "the kernel's shared-memory limit is raised"
  auto kern = stage_tile_kernel<STAGE, FEAT>;
  const int smem = (int)sizeof(TileSmem<STAGE>);
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!sms_of[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
  }
  const int sms = sms_of[dev];
  const int per_sm = (FEAT & kFeatBonds) ? PTOY_TILE_CTAS_BONDS : PTOY_TILE_CTAS;
  const int tiles = (n_bound + kTile - 1) / kTile;
  const int grid = tiles < sms * per_sm ? tiles : sms * per_sm;   // persistent: tiles are dealt round-robin
  kern<<<grid, kTileThreads, smem, s>>>(a);
}

template <int FEAT>
static void launch_stage_tile_feat(int stage, const StageArgs& a, int n_bound, cudaStream_t s) {
  if (stage == 1) launch_stage_tile_one<1, FEAT>(a, n_bound, s);
  else launch_stage_tile_one<2, FEAT>(a, n_bound, s);
}

}  // namespace ptoy

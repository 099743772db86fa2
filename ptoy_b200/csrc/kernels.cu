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
#include <algorithm>

#include "common.h"
#include "device_math.cuh"

// resident CTAs per SM the stage kernels are compiled for (measured on B200: 40 registers win
// without bonds, 48 with the bond partner prefetch)
#ifndef PTOY_STAGE_MINB
#define PTOY_STAGE_MINB 6
#endif
#ifndef PTOY_STAGE_MINB_BONDS
#define PTOY_STAGE_MINB_BONDS 5
#endif

namespace ptoy {

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
  const int ck = a.cell[i];
  const int Si = ck / a.g.SJ, Sj = ck - Si * a.g.SJ;
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
#ifndef PTOY_SCAN_THREADS
#define PTOY_SCAN_THREADS 512
#endif
#ifndef PTOY_SCAN_ITEMS
#define PTOY_SCAN_ITEMS 32
#endif
// Large tiles keep the look-back chain short (tiles / 32 rounds of one L2 round trip each): at
// 16M particles the histogram has 14M entries = 870 tiles of 16384.
constexpr int kScanThreads = PTOY_SCAN_THREADS;
constexpr int kScanItems = PTOY_SCAN_ITEMS;  // consecutive ints per thread (int4 loads)
constexpr int kScanTile = kScanThreads * kScanItems;
int scan_tile_items() { return kScanTile; }

__global__ void __launch_bounds__(kScanThreads) scan_kernel(
    int* __restrict__ cnt, int* __restrict__ start, int n_items, int ncell,
    unsigned long long* tile_state, int* ticket, unsigned epoch, Counters* ctr) {
  __shared__ int s_tile;
  __shared__ int s_warp[kScanThreads / 32];
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int base = tile * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; k += 4) {
    int4 q = make_int4(0, 0, 0, 0);
    if (base + k < n_items) {  // n_items % 4 == 0; the histogram is handed back all-zero
      q = *reinterpret_cast<const int4*>(cnt + base + k);
      *reinterpret_cast<int4*>(cnt + base + k) = make_int4(0, 0, 0, 0);
    }
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
    int w = (lane < kScanThreads / 32) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < kScanThreads / 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kScanThreads / 32) s_warp[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const int block_total = s_warp[kScanThreads / 32 - 1];
  const int excl_in_block = incl - tsum + (warp ? s_warp[warp - 1] : 0);

  // publish aggregate / look back.  state = (epoch*4 + flag) << 32 | value; flag 1 = this tile's
  // total, flag 2 = inclusive prefix.  Warp 0 looks at 32 predecessors per round: lane l reads
  // tile t - l, the nearest one that already holds an inclusive prefix ends the walk.
  // (30 bits of epoch: the tag and the flag share the upper word; every tile rewrites its state
  // each scan, so a value from 2^30 scans ago cannot be met)
  const unsigned long long tag = (unsigned long long)(epoch & 0x3fffffffu) << 2;
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
// reorder: one scatter pass + a fix-up of the few cells whose order it could not decide
// (the device half of blocks::SortParticles, blocks.h:215-244)
// ---------------------------------------------------------------------------------
// The reorder is the STABLE sort by new sub-cell key: members of a cell keep the order of their
// old slots (arrivals from other strips: of their (source rank, old slot) tags), so the result
// does not depend on the order in which atomics are served and N strips equal one GPU bit for
// bit.  A thread per OLD slot moves its particle straight to its new slot, both sides nearly
// coalesced because few particles change cell per step:
//   * fast path (almost every particle): all m members of the new cell form one run of
//     consecutive old slots around this one (checked in shared memory against the cell's
//     population m = start[key + 1] - start[key]); its place is the run offset;
//   * otherwise (somebody entered the cell from another sub-row, from another strip, through a
//     portal, or the cell holds more than kApron particles) every member takes a slot of the
//     cell from an atomic counter, leaves its old slot as a tag, and the first one files the
//     cell for fixup_kernel, which re-places the members by their tags.
constexpr int kApron = 8;

__device__ __forceinline__ unsigned long long reorder_vkey(const ReorderArgs& a, int i, int n_loc) {
  return i < n_loc ? ((unsigned long long)a.rank << 32) | (unsigned)i : a.arr_vkey[i - n_loc];
}

__global__ void __launch_bounds__(kThreads) scatter_kernel(const __grid_constant__ ReorderArgs a) {
  __shared__ int s_key[kThreads + 2 * kApron];
  const int n_loc = a.ctr->n_cur, n_in = n_loc + a.ctr->n_arrived;
  const int i0 = blockIdx.x * kThreads;
  if (i0 >= n_in) return;
  for (int idx = threadIdx.x; idx < kThreads + 2 * kApron; idx += kThreads) {
    const int j = i0 - kApron + idx;
    s_key[idx] = (j >= 0 && j < n_in) ? a.newcell[j] : -1;
  }
  __syncthreads();
  const int i = i0 + threadIdx.x;
  if (i >= n_in) return;
  const int* sk = s_key + kApron + threadIdx.x;
  const int key = sk[0];
  const float2 p = a.pos_in[i], v = a.vel_in[i];
  const int pid = a.id_in[i];
  if (key == a.ncell) {  // left the grid (blocks.h:230-232) or moved to another strip
    if (a.slot_of_id) a.slot_of_id[pid] = kSlotNone;
    return;
  }
  const int beg = a.start[key], m = a.start[key + 1] - beg;
  int dest = -1;
  if (m <= kApron) {
    int L = 0, R = 0;
    while (L < kApron && sk[-1 - L] == key) ++L;
    while (R < kApron && sk[1 + R] == key) ++R;
    if (L + R + 1 == m && i + R < n_loc) dest = beg + L;
  }
  if (dest < 0) {
    const int r = atomicAdd(a.cnt + key, 1);
    dest = beg + r;
    a.tag[dest] = i;
    if (r == 0) a.fixlist[atomicAdd(a.n_fix, 1)] = key;
  }
  a.pos_out[dest] = p;
  a.vel_out[dest] = v;
  a.id_out[dest] = pid;
  a.cell_out[dest] = key;
  if (a.slot_of_id) a.slot_of_id[pid] = dest;
}

// One warp per filed cell: member e (placed by the atomic counter) belongs at the rank of its tag
// among the cell's tags; the payload is re-read from the old arrays, which are still intact.
__global__ void __launch_bounds__(kThreads) fixup_kernel(const __grid_constant__ ReorderArgs a) {
  const int nfix = *a.n_fix;
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (kThreads / 32);
  const int n_loc = a.ctr->n_cur;
  for (int f = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); f < nfix; f += warps) {
    const int key = a.fixlist[f];
    const int beg = a.start[key], end = a.start[key + 1];
    for (int e = beg + lane; e < end; e += 32) {
      const int src = a.tag[e];
      const unsigned long long vk = reorder_vkey(a, src, n_loc);
      int rank = 0;
      for (int e2 = beg; e2 < end; ++e2) rank += reorder_vkey(a, a.tag[e2], n_loc) < vk;
      const int d = beg + rank;
      const int pid = a.id_in[src];
      a.pos_out[d] = a.pos_in[src];
      a.vel_out[d] = a.vel_in[src];
      a.id_out[d] = pid;
      a.cell_out[d] = key;
      if (a.slot_of_id) a.slot_of_id[pid] = d;
    }
    __syncwarp();
    if (lane == 0) a.cnt[key] = 0;   // the histogram goes back all-zero
  }
}

__global__ void finalize_kernel(Counters* ctr, int* ticket) {
  ctr->removed_total += ctr->n_total - ctr->n_next;
  ctr->n_cur = ctr->n_next;
  ctr->n_arrived = 0;
  ctr->n_fix = 0;
  ctr->dmax = 0u;
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
  const int key = cell_key(a.g, cell_of_fast(a.g, x), migrant);  // the sender routed it here
  a.newcell[dst] = key;
  atomicAdd(a.cnt + key, 1);
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
// tile-staged kernels, one translation unit per (bonds, portals) combination (stage_tile.cu)
void launch_stage_tile_g0(int stage, bool extras, const StageArgs& a, int n_bound, cudaStream_t s);
void launch_stage_tile_g1(int stage, bool extras, const StageArgs& a, int n_bound, cudaStream_t s);
void launch_stage_tile_g2(int stage, bool extras, const StageArgs& a, int n_bound, cudaStream_t s);
void launch_stage_tile_g3(int stage, bool extras, const StageArgs& a, int n_bound, cudaStream_t s);

void launch_stage(int stage, const StageArgs& a, int n_bound, cudaStream_t s) {
  if (n_bound <= 0) return;
  const int feat = (a.has_bonds ? kFeatBonds : 0) | (a.n_sides > 0 ? kFeatPortals : 0) |
                   ((a.ph.force_enabled || a.ph.pick_enabled) ? kFeatExtras : 0) |
                   (a.dist.world > 1 ? kFeatStrips : 0);
  if (!(feat & kFeatStrips)) {
    const bool extras = (feat & kFeatExtras) != 0;
    switch (feat & 3) {
      case 0: launch_stage_tile_g0(stage, extras, a, n_bound, s); break;
      case 1: launch_stage_tile_g1(stage, extras, a, n_bound, s); break;
      case 2: launch_stage_tile_g2(stage, extras, a, n_bound, s); break;
      default: launch_stage_tile_g3(stage, extras, a, n_bound, s); break;
    }
    return;
  }
  switch (feat) {
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
void launch_scan(int* cnt, int* start, int n_items, int ncell, unsigned long long* tile_state,
                 int* ticket, unsigned epoch, Counters* ctr, cudaStream_t s) {
  const int tiles = (n_items + kScanTile - 1) / kScanTile;
  scan_kernel<<<tiles, kScanThreads, 0, s>>>(cnt, start, n_items, ncell, tile_state, ticket, epoch, ctr);
}
void launch_scatter(const ReorderArgs& a, int n_bound, cudaStream_t s) {
  if (n_bound <= 0) return;
  scatter_kernel<<<grid_for(n_bound), kThreads, 0, s>>>(a);
}
void launch_fixup(const ReorderArgs& a, int n_bound, cudaStream_t s) {
  if (n_bound <= 0) return;
  // a grid-stride loop over the filed cells (usually a handful): a fixed, small grid
  const unsigned grid = (unsigned)std::min<size_t>(grid_for(n_bound), 148 * 4);
  fixup_kernel<<<grid, kThreads, 0, s>>>(a);
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

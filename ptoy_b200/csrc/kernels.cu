// Hand-written sm_100a kernels of the particle-stepping hot path.
//
// Arithmetic contract: every floating-point operation that the reference performs
// (scalar build, no FMA) is issued here as the same single IEEE
// binary32 operation through the __f*_rn intrinsics, which nvcc never contracts
// into FMAs.  A pair/wall/bond/portal contribution is therefore bit-identical to
// the reference's; only the ORDER in which a particle's pair contributions are
// summed differs (sub-cell order instead of block/slot order).
//
// Kernels of one iteration; the two stage kernels live in stage_tile.cuh:
//   stage_tile_kernel<1>  forces at x0 + half kick/drift       particles.cpp:97-123
//   stage_tile_kernel<2>  forces at x_half + full update + portal teleport + new sort key +
//                         histogram                            particles.cpp:128-153,177-179
//   scan_kernel           single-pass decoupled look-back exclusive scan of the histogram
//   scatter_kernel        stable reorder of pos/vel/id in one pass (+ fixup_kernel for the few
//                         cells whose order the scatter could not decide)
// plus the strip decomposition's pack / install / arrival kernels and read-back helpers.
#include <algorithm>

#include <cub/device/device_scan.cuh>

#include "common.h"
#include "device_math.cuh"

namespace ptoy {

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
#define PTOY_SCAN_THREADS 1024
#endif
#ifndef PTOY_SCAN_ITEMS
#define PTOY_SCAN_ITEMS 16
#endif
// Large tiles keep the look-back chain short (tiles / 32 rounds of one L2 round trip each): at
// 16M particles the histogram has 14M entries = 870 tiles of 16384 (1024 threads x 16: measured
// 0.057 ms against 0.068 ms for 512 x 32 and 0.089 ms for 256 x 16).
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
  auto exchanged = [&](int r) { return r != a.rank && (!a.neighbours_only || r == a.rank - 1 || r == a.rank + 1); };
  int off = 0, total = 0, mine = 0;
  for (int r = 0; r < a.world; ++r) {
    if (!exchanged(r)) continue;   // (its block was not received this iteration: stale)
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
  const int pid = reinterpret_cast<const int*>(blk + mig_off_id(a.in_cap))[k];
  a.pos[dst] = x;
  a.vel[dst] = reinterpret_cast<const float2*>(blk + mig_off_vel(a.in_cap))[k];
  a.id[dst] = pid;
  a.arr_vkey[off + k] = ((unsigned long long)src << 32) |
                        (unsigned)reinterpret_cast<const int*>(blk + mig_off_tag(a.in_cap))[k];
  if (a.bond_ell) {   // the migrant's bond row travels with it
    if (pid < a.id_cap) {
      const int4* row = reinterpret_cast<const int4*>(blk + mig_off_row(a.in_cap)) + 2 * (size_t)k;
      int4* out = reinterpret_cast<int4*>(a.bond_ell) + 2 * (size_t)pid;
      out[0] = row[0];
      out[1] = row[1];
    } else {
      atomicOr(&a.ctr->err_flags, 64);
    }
  }
  bool migrant;
  const int key = cell_key(a.g, cell_of_fast(a.g, x), migrant);  // the sender routed it here
  a.newcell[dst] = key;
  atomicAdd(a.cnt + key, 1);
}

// Copy the boundary sub-rows on each side into staging buffers (positions; optionally the
// matching slice of start[], rebased to 0, and the ids) for the neighbouring strips.
__global__ void __launch_bounds__(kThreads) ghost_pack_kernel(const __grid_constant__ GhostPackArgs a) {
  const int side = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  const int c0 = a.row_lo[side] * a.SJ;
  const int b0 = a.start[c0], b1 = a.start[c0 + a.nrows * a.SJ];
  if (t == 0 && b1 - b0 > a.ghost_cap) atomicOr(a.err, 4);
  if (t < b1 - b0 && t < a.ghost_cap) {
    a.out_pos[side][t] = a.X[b0 + t];
    if (a.out_id[side]) a.out_id[side][t] = a.id[b0 + t];
  }
  if (a.out_start[side] && t <= a.nrows * a.SJ) a.out_start[side][t] = a.start[c0 + t] - b0;
  if (t == 0) a.out_pos[side][a.ghost_cap] = make_float2(a.dmax ? __uint_as_float(*a.dmax) : 0.f, 0.f);
}

// Install the received boundary rows as ghost rows (see GhostInstallArgs).
__global__ void __launch_bounds__(kThreads) ghost_install_kernel(const __grid_constant__ GhostInstallArgs a) {
  const int side = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (a.in_pos[side] == nullptr) return;
  const int nk = a.nrows * a.SJ;                       // keys of the ghost rows on this side
  const int n = a.ctr->n_cur;
  const int g = min(a.in_start[side][nk], a.ghost_cap);   // ghosts on this side (the slice is rebased to 0)
  const int base = side == 0 ? -g : n;                  // slot of the first one
  if (t < g) {
    a.X[base + t] = a.in_pos[side][t];
    if (a.stage == 1 && a.in_id[side]) {
      const int pid = a.in_id[side][t];
      if (pid >= 0 && pid < a.id_cap) a.slot_of_id[pid] = base + t; else atomicOr(a.err, 64);
    }
  }
  if (a.stage == 1) {
    if (t <= nk) a.start[(side == 0 ? 0 : a.own_hi * a.SJ) + t] = base + a.in_start[side][t];
    if (t == 0) *a.count[side] = g;
  } else if (t == 0) {   // the neighbour's largest half-step drift widens this strip's search margin too
    atomicMax(&a.ctr->dmax, __float_as_uint(a.in_pos[side][a.ghost_cap].x));
  }
}

// Before the reorder: the ghosts leave the id -> slot map again (a slot outside [0, n) is a ghost's).
__global__ void __launch_bounds__(kThreads) ghost_unmark_kernel(const __grid_constant__ GhostUnmarkArgs a) {
  const int side = blockIdx.y;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (a.in_id[side] == nullptr || t >= *a.count[side]) return;
  const int pid = a.in_id[side][t];
  if (pid < 0 || pid >= a.id_cap) return;
  const int sl = a.slot_of_id[pid];
  if (sl != kSlotNone && (sl < 0 || sl >= a.ctr->n_cur)) a.slot_of_id[pid] = kSlotNone;
}

// One thread per listed portal block: the particles of a block this strip owns go into the
// zone table in (sub-row, slot) order; entries of other strips' blocks stay zero (the tables are
// summed over the strips afterwards).
__global__ void __launch_bounds__(kThreads) zone_pack_kernel(const __grid_constant__ PortalZoneArgs a) {
  const int l = blockIdx.x * kThreads + threadIdx.x;
  if (l >= a.n_entries) return;
  int r0, r1, c0, c1;
  block_extent(a.g, a.pblk[l], r0, r1, c0, c1);
  float2* out = a.X + a.zoff + (size_t)l * kZoneCap;
  int* oid = a.zid + (size_t)l * kZoneCap;
  int k = 0;
  for (int Rg = r0; Rg <= r1; ++Rg) {
    const int R = Rg - a.g.row0;
    if (R < a.g.own_lo || R >= a.g.own_hi) continue;
    const int beg = a.start[R * a.g.SJ + c0], end = a.start[R * a.g.SJ + c1 + 1];
    for (int q = beg; q < end; ++q, ++k)
      if (k < kZoneCap) { out[k] = a.X_in[q]; oid[k] = a.id[q]; }
  }
  if (k > kZoneCap) { atomicOr(a.err, 128); k = kZoneCap; }
  a.zcnt[l] = k;
}

// mark = 1: zone particles that are neither local nor ghosts enter the id -> slot map (a bond
// partner on the far side of a portal, on another strip); mark = 0: they leave it again.
__global__ void __launch_bounds__(kThreads) zone_mark_kernel(const __grid_constant__ PortalZoneArgs a, int mark) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  const int l = t / kZoneCap, k = t - l * kZoneCap;
  if (l >= a.n_entries || k >= a.zcnt[l]) return;
  const int pid = a.zid[t];
  if (pid < 0 || pid >= a.id_cap) return;
  const int slot = a.zoff + t;
  if (mark) { if (a.slot_of_id[pid] == kSlotNone) a.slot_of_id[pid] = slot; }
  else if (a.slot_of_id[pid] >= a.zoff) a.slot_of_id[pid] = kSlotNone;
}

// Wire format of the remote front end (ptoy_wasm.cpp:130-146 GetParticles): quantise the first
// n_max particles (engine order) to uint16 pixel pairs on the device.
__global__ void wire_kernel(const Counters* ctr, const float2* __restrict__ pos, int n_max, int width, int height,
                            unsigned* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ctr->n_cur || i >= n_max) return;
  const float2 p = pos[i];
  out[i] = wire_pixel(p.x, p.y, width, height);
}

// ---- render snapshot: the sub-cell order becomes the reference's block order on the device -----
// A reference block is two (the low band: up to four) runs of consecutive slots, one per sub-row.
__device__ __forceinline__ int block_rows(const GridDev& g, int blk, int& r0, int& r1, int& c0, int& c1) {
  block_extent(g, blk, r0, r1, c0, c1);          // global sub-rows / sub-columns of the block
  r0 -= g.row0; r1 -= g.row0;                     // local rows; only owned ones hold this strip's particles
  r0 = max(r0, g.own_lo); r1 = min(r1, g.own_hi - 1);
  return r1 - r0 + 1;
}
__global__ void __launch_bounds__(kThreads) block_counts_kernel(GridDev g, const int* __restrict__ start, int nblk,
                                                                int* __restrict__ count, int* max_count) {
  const int b = blockIdx.x * kThreads + threadIdx.x;
  int n = 0;
  if (b < nblk) {
    int r0, r1, c0, c1;
    block_rows(g, b, r0, r1, c0, c1);
    for (int R = r0; R <= r1; ++R) n += start[R * g.SJ + c1 + 1] - start[R * g.SJ + c0];
    count[b] = n;
  } else if (b == nblk) {
    count[b] = 0;
  }
  const int m = __reduce_max_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_count, m);
}
__global__ void __launch_bounds__(kThreads) block_gather_kernel(GridDev g, const int* __restrict__ start,
                                                                const int* __restrict__ block_off, int nblk,
                                                                const float2* __restrict__ pos, const float2* __restrict__ vel,
                                                                const int* __restrict__ id, float2* out_pos, float2* out_vel,
                                                                int* out_id, int* out_block) {
  const int b = blockIdx.x * kThreads + threadIdx.x;
  if (b >= nblk) return;
  int r0, r1, c0, c1;
  block_rows(g, b, r0, r1, c0, c1);
  int k = block_off[b];
  for (int R = r0; R <= r1; ++R)            // sub-row, then slot: the order a stable sort by block gives
    for (int q = start[R * g.SJ + c0]; q < start[R * g.SJ + c1 + 1]; ++q, ++k) {
      out_pos[k] = pos[q]; out_vel[k] = vel[q]; out_id[k] = id[q]; out_block[k] = b;
    }
}

__global__ void sum_i32_kernel(int* dst, const int* src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
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

// tile-staged kernels, one translation unit per (bonds, portals) combination (stage_tile.cu)
void launch_stage_tile_g0(int stage, bool extras, const StageArgs& a, int n_bound, cudaStream_t s);
void launch_stage_tile_g1(int stage, bool extras, const StageArgs& a, int n_bound, cudaStream_t s);
void launch_stage_tile_g2(int stage, bool extras, const StageArgs& a, int n_bound, cudaStream_t s);
void launch_stage_tile_g3(int stage, bool extras, const StageArgs& a, int n_bound, cudaStream_t s);

void launch_stage(int stage, const StageArgs& a, int n_bound, cudaStream_t s) {
  if (n_bound <= 0) return;
  const int feat = (a.has_bonds ? kFeatBonds : 0) | (a.n_sides > 0 ? kFeatPortals : 0) |
                   ((a.ph.force_enabled || a.ph.pick_enabled) ? kFeatExtras : 0);
  const bool extras = (feat & kFeatExtras) != 0;
  switch (feat & 3) {
    case 0: launch_stage_tile_g0(stage, extras, a, n_bound, s); break;
    case 1: launch_stage_tile_g1(stage, extras, a, n_bound, s); break;
    case 2: launch_stage_tile_g2(stage, extras, a, n_bound, s); break;
    default: launch_stage_tile_g3(stage, extras, a, n_bound, s); break;
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
void launch_ghost_install(const GhostInstallArgs& a, cudaStream_t s) {
  const size_t n = (size_t)max(a.ghost_cap, a.nrows * a.SJ + 1);
  ghost_install_kernel<<<dim3(grid_for(n), 2), kThreads, 0, s>>>(a);
}
void launch_ghost_unmark(const GhostUnmarkArgs& a, cudaStream_t s) {
  ghost_unmark_kernel<<<dim3(grid_for((size_t)a.ghost_cap), 2), kThreads, 0, s>>>(a);
}
void launch_zone_pack(const PortalZoneArgs& a, cudaStream_t s) {
  if (a.n_entries > 0) zone_pack_kernel<<<grid_for((size_t)a.n_entries), kThreads, 0, s>>>(a);
}
void launch_zone_mark(const PortalZoneArgs& a, int mark, cudaStream_t s) {
  if (a.n_entries > 0) zone_mark_kernel<<<grid_for((size_t)a.n_entries * kZoneCap), kThreads, 0, s>>>(a, mark);
}
void launch_wire(const Counters* ctr, const float2* pos, int n_max, int width, int height, unsigned* out, cudaStream_t s) {
  if (n_max > 0) wire_kernel<<<grid_for((size_t)n_max), kThreads, 0, s>>>(ctr, pos, n_max, width, height, out);
}
void launch_block_counts(const GridDev& g, const int* start, int nblk, int* count, int* max_count, cudaStream_t s) {
  block_counts_kernel<<<grid_for((size_t)nblk + 1), kThreads, 0, s>>>(g, start, nblk, count, max_count);
}
void launch_block_gather(const GridDev& g, const int* start, const int* block_off, int nblk, const float2* pos,
                         const float2* vel, const int* id, float2* out_pos, float2* out_vel, int* out_id, int* out_block,
                         cudaStream_t s) {
  block_gather_kernel<<<grid_for((size_t)nblk), kThreads, 0, s>>>(g, start, block_off, nblk, pos, vel, id, out_pos,
                                                                  out_vel, out_id, out_block);
}
// (library scan for this once-per-frame helper; the per-iteration scan of the hot path is scan_kernel above)
size_t scan_exclusive_i32_temp_bytes(int n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int*)nullptr, (int*)nullptr, n);
  return bytes;
}
void scan_exclusive_i32(void* temp, size_t temp_bytes, const int* in, int* out, int n, cudaStream_t s) {
  cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, n, s);
}
void launch_sum_i32(int* dst, const int* src, size_t n, cudaStream_t s) {
  if (n) sum_i32_kernel<<<grid_for(n), kThreads, 0, s>>>(dst, src, n);
}
void launch_arrivals(const ArrivalArgs& a, cudaStream_t s) {
  arrivals_kernel<<<grid_for((size_t)a.world * a.in_cap), kThreads, 0, s>>>(a);
}
void launch_ghost_pack(const GhostPackArgs& a, cudaStream_t s) {
  const size_t n = (size_t)max(a.ghost_cap, a.nrows * a.SJ + 1);
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

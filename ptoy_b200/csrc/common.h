// Shared host/device definitions of the ptoy_b200 engine.
//
// Data layout in HBM: particles live in structure-of-arrays
// buffers (float2 pos, float2 vel, int id) sorted by SUB-CELL key.  A sub-cell is
// half a reference block in each direction (blocks.h: block size 4r = 0.08, so a
// sub-cell is 2r = 0.04 = the pair cutoff); `start[key]` is the exclusive prefix
// sum of sub-cell populations.  Reference block (i,j) = sub-rows {2i+4, 2i+5}
// x sub-columns {2j+4, 2j+5} of the padded sub-grid (plus the low band {2,3} for
// i==0 / j==0, which holds the particles the reference's truncating FindBlock
// maps to row/column 0 from outside the grid, blocks.h:155-163).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ptoy {

constexpr int kMaxWalls = 16;
constexpr int kThreads = 256;
constexpr int kSlotNone = -0x7fffffff - 1;   // id -> slot map entry of an id that is not stored (left the grid / never added)

// Sub-cell grid.  S = floor(2*t) + 4 with t = fl(fl(x - A)/bs) exactly as
// blocks::FindBlock computes it (blocks.h:156-158).
struct GridDev {
  float Ax, Ay;    // blocks::domain_.A
  float bsx, bsy;  // blocks::block_size_
  int di, dj;      // blocks::dims_
  float fdi, fdj;  // (float)dims
  int SI, SJ;      // padded sub-grid dims: 2*d + 6
  int ncell;       // SI*SJ ; key == ncell is the trash bin (particle left the grid)
  float inv_hsx, inv_hsy;  // fl(2/bs): approximate sub-cell coordinate w~ = (x - A) * inv_hs
  float tol;       // |w~ - w| bound used by the division-free binning (kernels.cu cell_of_fast)
  float wmax;      // largest |w| for which tol holds
  // Strip decomposition: this engine stores sub-rows [row0, row0 + SI) of the
  // global padded sub-grid as local rows [0, SI); it OWNS local rows [own_lo, own_hi).  The
  // own_lo rows below and SI - own_hi rows above are GHOST rows: copies of the neighbour strips'
  // boundary rows, installed every stage right before slot 0 (negative slots) and right behind
  // the last owned particle, with start[] entries to match, so that a search window is one
  // contiguous slot range whether it lies in owned or ghost rows.  A single-GPU engine is the
  // case row0 = 0, own = [2, SI - 2): the (always empty) global padding rows.
  int row0;
  int own_lo, own_hi;
};

constexpr int kMaxRanks = 16;
// Where the strips begin, as global sub-rows: rank r owns rows [row_begin[r], row_begin[r+1]).
struct DistDev {
  int world, rank;
  int row_begin[kMaxRanks + 1];
  // migrant outbox: one fixed-size block per destination rank, laid out as
  //   int count (16 B header) | float2 pos[cap] | float2 vel[cap] | int id[cap] | int tag[cap] | int4 row[cap][2]
  // tag = old slot of the migrant on this rank: makes the receiver's order deterministic;
  // row = the migrant's bond row (bonds travel with their particle)
  char* out_base;          // nullptr: particles of other strips are simply dropped
  long long out_stride;
  int out_cap;
  const int* bond_ell;     // bond table by id (nullptr: the scene has no bonds)
  int* err;
};
__host__ __device__ inline long long mig_off_pos() { return 16; }
__host__ __device__ inline long long mig_off_vel(int cap) { return 16 + 8ll * cap; }
__host__ __device__ inline long long mig_off_id(int cap) { return 16 + 16ll * cap; }
__host__ __device__ inline long long mig_off_tag(int cap) { return 16 + 20ll * cap; }
__host__ __device__ inline long long mig_off_row(int cap) { return 16 + 24ll * cap; }
__host__ __device__ inline long long mig_stride(int cap) { return 16 + 56ll * cap; }

// Physics constants, computed on the host with the reference's own expression
// types (particles.cpp:7-24) so the bit patterns agree.
struct Phys {
  float dt;
  float c1, c2;          // fl(dt*0.5/m), fl(dt/m)        particles.cpp:116,145
  float gfx, gfy;        // gravity_*kMass                particles.cpp:850
  int gravity_enable;
  float diss;            // kDissipation*kMass            particles.cpp:880
  float RR, thr, r2c;    // pair: fl(R*R), r^2 threshold, first r^2 with zero force
  float r2c_hi;          // r2c grown by 8 ulp: bound for the fma prefilter of the pair loop
  float v2c;             // largest |v|^2 whose rounded sqrt is <= kVelocityLimit
  float bond_y;          // fl(1/bondR)
  float radius;          // kRadius
  float RRw;             // wall: fl(r*r)                 particles.cpp:574-578
  float wall_reach;      // conservative |p-Q| beyond which F12wall is exactly 0
  float RRe;             // portal edge: fl(R*R), R=3r    particles.cpp:740
  float vlimit;          // kVelocityLimit
  float bondR, bond_cutoff, sigma_bond;  // particles.cpp:728-729,777
  double bond_portal_reach;              // 2.*kPortalThickness + kCutoff   :808
  float pickR, sigma_pick;               // particles.cpp:752-753
  float portal_thickness;
  double two_pt;         // 2.*kPortalThickness (double)  :284,371,812
  int force_enabled, force_attractive;
  float fcx, fcy;
  float point_force, point_force_attractive;
  int pick_enabled, pick_id;
  float pkx, pky;
};

// Pixel coordinates of the remote front end's wire format (ptoy_wasm.cpp:130-198):
//   data[0] = (1 + p.x) * width * 0.5,  data[1] = (1 - p.y) * height * 0.5   -> uint16_t
// float sum, float product with the int width, product with the DOUBLE 0.5 (exact), truncation.
__host__ __device__ inline unsigned wire_pixel(float x, float y, int width, int height) {
#ifdef __CUDA_ARCH__
  const float fx = __fmul_rn(__fadd_rn(1.0f, x), (float)width), fy = __fmul_rn(__fsub_rn(1.0f, y), (float)height);
#else
  const float fx = (1.0f + x) * (float)width, fy = (1.0f - y) * (float)height;
#endif
  const unsigned qx = (unsigned)(int)((double)fx * 0.5) & 0xffffu, qy = (unsigned)(int)((double)fy * 0.5) & 0xffffu;
  return qx | (qy << 16);   // little endian: data[0] = x, data[1] = y
}

struct WallDev {
  float ax, ay, bx, by;          // line A,B (particles.h:35)
  float lox, loy, hix, hiy;      // bbox grown by wall_reach: outside it the force is 0
};

// One portal (pair p, side d) = entry 2p+d; r, n, rr as every user recomputes them
// (e.g. particles.cpp:293-296): r = end-begin, n = normalized(-r.y, r.x), rr = r.r
struct PortalSide {
  float ax, ay, bx, by;
  float rx, ry, nx, ny;
  float rr;
  float pad[3];
};

struct Counters {
  int n_cur;        // particles alive (valid slots [0,n_cur))
  int n_arrived;    // migrants appended behind them, waiting for the reorder
  int n_next;       // alive after the pending reorder
  int n_total;      // n_next + trash
  int removed_total;
  int max_sub_occupancy;
  int err_flags;
  int n_fix;        // cells filed for fixup_kernel by the current reorder
  unsigned dmax;    // float bits: largest per-component half-step drift measured by stage 1
  int lost[4];      // first bond whose partner could not be served (error bit 16): id, partner id, slot, n_cur
};

struct StageArgs {
  GridDev g;
  Phys ph;
  WallDev walls[kMaxWalls];
  int n_walls;
  float free_lox, free_loy, free_hix, free_hiy;  // strictly inside: no wall acts
  const Counters* ctr;
  const float2* pos;     // x0 (binning-time positions), sorted by sub-cell key
  const float2* vel;     // v0
  float2* pos_h;         // x_half
  float2* vel_h;         // v_half
  float2* pos_out;       // stage 2: x' (may alias pos)
  float2* vel_out;       // stage 2: v'
  const int* id;
  const int* cell;       // sort key (local sub-row * SJ + sub-column) of every sorted particle
  const int* start;      // [ncell+2] exclusive prefix of sub-cell populations
  DistDev dist;
  float margin;          // extra search margin in sub-cell units; stage 2: its upper bound
  // stage 2 of the tile kernels: margin = min(margin, margin_base + margin_scale * *dmax), where
  // *dmax (float bits) is the largest per-component half-step drift stage 1 measured
  float margin_base, margin_scale;
  unsigned* dmax;
  // frozen_ as a bitmask over ids
  const uint32_t* frozen_bits;
  int has_frozen;
  // bonds_ as CSR over ids: row = .first, columns ascending = std::set order
  const uint32_t* bond_ptr;
  const int* bond_col;
  const int* slot_of_id;
  int has_bonds;
  const int* bond_ell;         // bond table by id: 8 ints = first 6 partner ids (-1 pad), unused, degree
  int* bslot;                  // [cap][8] partner slots + degree: written by stage 1, read by stage 2
  int cap;                     // stride of the slot-major bslot array
  int id_cap;                  // entries of the id-indexed tables (slot_of_id, bond_ell, frozen_bits)
  int* err;                    // sticky error bits (Counters::err_flags)
  int* lost;                   // Counters::lost
  // portals
  int n_sides;           // 2 * pairs
  const PortalSide* sides;
  const int* pblk_ptr;   // [n_sides+1] into pblk
  const int* pblk;       // Portal::blocks, ascending
  const uint8_t* blk_flags;  // per reference block: bit0 in a portal list, bit1 near a portal end
  // strips: the particles of every listed portal block, gathered from the strips that own them
  // (PortalZoneArgs): entry l of pblk holds zcnt[l] particles at X[zoff + l * kZoneCap ...]
  const int* zcnt;       // nullptr: one strip, the blocks are read in place
  int zoff;
  int need_id;
  // outputs
  float2* force_out;     // optional, by slot
  int* newcell;          // stage 2: sort key after the update (old slot order); ncell = trash bin
  int* cnt;              // stage 2: sub-cell histogram (RED)
  int do_tail;           // stage 2: teleport + key + histogram fused
  int write_state;       // 0 = forces only (debug)
};

struct TailArgs {  // standalone teleport + key pass (resize iterations, add_particles)
  GridDev g;
  Phys ph;
  const Counters* ctr;
  float2* pos;
  float2* vel;
  int n_sides;
  const PortalSide* sides;
  const int* pblk_ptr;
  const int* pblk;
  const uint8_t* blk_flags;
  int do_teleport;
  DistDev dist;
  const int* id;
  int first;     // particles [first, n) are processed
  int* newcell;
  int* cnt;
};

struct ReorderArgs {
  const Counters* ctr;
  int ncell;
  int SJ;
  const int* newcell;
  const int* start;
  int* cnt;          // all-zero on entry (the scan consumed the histogram) and on exit
  int* tag;          // by new slot: old slot of a particle placed through the atomic counter
  int* fixlist;      // cells whose members were placed through the atomic counter
  int* n_fix;        // Counters::n_fix
  const float2* pos_in;
  const float2* vel_in;
  const int* id_in;
  float2* pos_out;
  float2* vel_out;
  int* id_out;
  int* cell_out;
  int* slot_of_id;   // optional
  // stable order across strips: locals sort as (rank, old slot), arrivals carry the (source
  // rank, old slot there) they had; equals the single-GPU stable order
  int rank;
  const unsigned long long* arr_vkey;   // [n_arrived]
};

struct ArrivalArgs {   // migrants received from every source rank -> appended behind the locals
  GridDev g;
  Counters* ctr;
  int world, in_cap;
  const char* in_base;      // [world] blocks in the outbox layout (DistDev)
  long long in_stride;
  float2* pos;
  float2* vel;
  int* id;
  int* newcell;
  int* cnt;
  unsigned long long* arr_vkey;
  int cap;
  int rank;
  int neighbours_only;      // only the blocks of rank - 1 / rank + 1 were exchanged this iteration
  int* bond_ell;            // receiver's bond table by id (nullptr: no bonds); rows of the arrivals go in
  int id_cap;
};

struct GhostPackArgs {   // boundary rows -> contiguous staging buffers for the neighbours
  const Counters* ctr;
  const float2* X;
  const int* start;
  int SJ;
  int nrows;          // ghost rows per side (GridDev::own_lo)
  int row_lo[2];      // first local row of the slice sent to the left (0) / right (1) neighbour
  int ghost_cap;
  float2* out_pos[2]; // [ghost_cap + 2]: the last element carries the sender's largest half-step drift
  int* out_start[2];  // nrows*SJ+1 entries, rebased to 0; nullptr = positions only (same layout as before)
  const int* id;      // with out_id: particle ids of the same rows (bonds across strips)
  int* out_id[2];
  const unsigned* dmax;
  int* err;
};

// The neighbours' boundary rows become this strip's ghost rows: positions go right before slot 0
// (left neighbour) / right behind the last owned particle (right neighbour), start[] of the ghost
// rows is rewritten to match, and (bonds) the ghost ids enter the id -> slot map.
struct GhostInstallArgs {
  Counters* ctr;
  int stage;                // 1: positions at x0 + start[] + ids; 2: positions at x_half + drift
  int SJ, nrows, own_hi, ghost_cap;
  const float2* in_pos[2];  // nullptr = no neighbour on that side
  const int* in_start[2];
  const int* in_id[2];      // nullptr: no bonds
  float2* X;                // pos (stage 1) / pos_h (stage 2), slot 0
  int* start;
  int* slot_of_id;
  int id_cap;
  int* count[2];            // ghosts installed per side (for ghost_unmark)
  int* err;
};
struct GhostUnmarkArgs {    // forget this iteration's ghosts in the id -> slot map (before the reorder)
  const Counters* ctr;
  int* slot_of_id;
  int id_cap;
  const int* in_id[2];
  const int* count[2];
  int ghost_cap;
};

// Cross-portal forces and through-portal bonds across strips (particles.cpp:318-381, 784-819):
// every strip copies the particles of the portal blocks it owns - in the order the single-GPU
// loop visits them: block-list entry, sub-row, slot - into a fixed-capacity table behind the
// ghost rows of the position array; the tables are summed over the strips (each entry has one
// owner, the others hold zeros), so every strip ends up with all portal-zone particles in the
// same order and the forces stay bit-identical to one GPU.
constexpr int kZoneCap = 24;   // particles per listed portal block
struct PortalZoneArgs {
  GridDev g;
  const Counters* ctr;
  const float2* X_in;     // pos (stage 1) / pos_h (stage 2), slot 0
  float2* X;              // the same array: the table goes to X[zoff ...]
  const int* id;
  const int* start;
  const int* pblk;
  int n_entries;          // pblk_ptr[n_sides]
  int zoff;
  int* zcnt;              // [n_entries]
  int* zid;               // [n_entries * kZoneCap]
  int* slot_of_id;        // mark / unmark (bonds): zone particles that are neither local nor ghosts
  int id_cap;
  int* err;
};

// launchers (kernels.cu)
void launch_zone_pack(const PortalZoneArgs& a, cudaStream_t s);
void launch_zone_mark(const PortalZoneArgs& a, int mark, cudaStream_t s);
void launch_sum_i32(int* dst, const int* src, size_t n, cudaStream_t s);
// render snapshot (SetParticleBuffer, particles.cpp:79-89): block-major compaction on the device
void launch_block_counts(const GridDev& g, const int* start, int nblk, int* count, int* max_count, cudaStream_t s);
void launch_block_gather(const GridDev& g, const int* start, const int* block_off, int nblk, const float2* pos,
                         const float2* vel, const int* id, float2* out_pos, float2* out_vel, int* out_id, int* out_block,
                         cudaStream_t s);
size_t scan_exclusive_i32_temp_bytes(int n);
void scan_exclusive_i32(void* temp, size_t temp_bytes, const int* in, int* out, int n, cudaStream_t s);
void launch_wire(const Counters* ctr, const float2* pos, int n_max, int width, int height, unsigned* out, cudaStream_t s);
void launch_stage(int stage, const StageArgs& a, int n_bound, cudaStream_t s);
void launch_tail(const TailArgs& a, int n_bound, cudaStream_t s);
void launch_scan(int* cnt, int* start, int n_items, int ncell, unsigned long long* tile_state,
                 int* ticket, unsigned epoch, Counters* ctr, cudaStream_t s);
void launch_scatter(const ReorderArgs& a, int n_bound, cudaStream_t s);
void launch_fixup(const ReorderArgs& a, int n_bound, cudaStream_t s);
void launch_finalize(Counters* ctr, int* ticket, cudaStream_t s);
void launch_arrivals(const ArrivalArgs& a, cudaStream_t s);
void launch_ghost_pack(const GhostPackArgs& a, cudaStream_t s);
void launch_ghost_install(const GhostInstallArgs& a, cudaStream_t s);
void launch_ghost_unmark(const GhostUnmarkArgs& a, cudaStream_t s);
void launch_find_blocks(const GridDev& g, const float2* pos, size_t n, long long* out, cudaStream_t s);
void launch_state_by_id(const GridDev& g, const Counters* ctr, const float2* pos, const float2* vel,
                        const int* id, int n_bound, float2* pos_by_id, float2* vel_by_id,
                        long long* block_by_id, cudaStream_t s);
void launch_scatter_by_id(const Counters* ctr, const float2* val, const int* id, int n_bound,
                          float2* out_by_id, cudaStream_t s);
void launch_block_of_slot(const GridDev& g, const Counters* ctr, const float2* pos, int n_bound,
                          int* block, cudaStream_t s);
void launch_slot_map(const Counters* ctr, const int* id, int n_bound, int* slot_of_id, cudaStream_t s);
void launch_fill_i32(int* p, int v, size_t n, cudaStream_t s);
int scan_tile_items();

}  // namespace ptoy

// Host side of the engine.  See engine.h.
#include "engine.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>

namespace ptoy {

#define CUDA_CHECK(expr)                                                                  \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess)                                                                \
      throw Error(-1, std::string(#expr) + ": " + cudaGetErrorString(e_));                \
  } while (0)

// kernels defined in kernels.cu used only here
void launch_max_id(const int* id, size_t n, int* out, cudaStream_t s);

// ---- constants, particles.cpp:7-24, with the reference's expression types -----------
static const float kRadius = 0.02f;
static const float kSigmaBond = 1e5f;
static const float kSigmaPick = 1e3f;
static const float kPointForce = 0.1f;
static const float kPointForceAttractive = 0.1f;
static const float kDissipation = 0.001f;
static const float kGravity = 10.f;
static const float kPortalThickness = 0.02f;
static const float kVelocityLimit = 10.f;
static const float kTimeStep = 0.0005f;
static inline float kMass() { return kRadius * kRadius * 100.f; }
static inline float kBlockSize() { return (float)(4. * (double)kRadius); }

// ---- Vect arithmetic on the host (geometry.h:18-90), no contraction (see build flags) --
static inline V2 vadd(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
static inline V2 vsub(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
static inline V2 vmul(V2 a, float k) { return V2{a.x * k, a.y * k}; }
static inline float vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline float vlen(V2 a) { return (float)std::sqrt((double)(a.x * a.x + a.y * a.y)); }
static inline float vdist(V2 a, V2 b) { return vlen(vsub(b, a)); }
static inline V2 vnormalized(V2 a) { float l = vlen(a); return V2{a.x / l, a.y / l}; }
static inline bool veq(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }
static inline float maxf_std(float a, float b) { return (a < b) ? b : a; }
static inline float minf_std(float a, float b) { return (b < a) ? b : a; }
static inline int f2i_trunc(float f) {
  if (!(f >= -2147483648.f && f < 2147483648.f)) return std::numeric_limits<int>::min();
  return (int)f;
}
static V2 seg_nearest(V2 A, V2 B, V2 p) {  // particles.h:37-46 / 67-77
  V2 ab = vsub(B, A);
  float lambda = vdot(ab, vsub(p, A)) / vdot(ab, ab);
  if (lambda > 0. && lambda < 1.) return vadd(A, vmul(ab, lambda));
  return (vdist(p, A) < vdist(p, B)) ? A : B;
}

// ---- DevBuf ----------------------------------------------------------------------------
template <class T>
void DevBuf<T>::release() {
  if (p) cudaFree(p - head);
  p = nullptr;
  n = 0;
  head = 0;
}
template <class T>
void DevBuf<T>::alloc(size_t count, size_t head_room) {
  if (count <= n && head_room <= head && p) return;
  release();
  if (count == 0) count = 1;
  T* raw = nullptr;
  CUDA_CHECK(cudaMalloc((void**)&raw, (count + head_room) * sizeof(T)));
  p = raw + head_room;
  n = count;
  head = head_room;
}
template <class T>
void DevBuf<T>::grow_keep(size_t count, cudaStream_t s, size_t head_room) {
  if (count <= n && head_room <= head && p) return;
  count = std::max(count, n);
  head_room = std::max(head_room, head);
  T* raw = nullptr;
  if (count == 0) count = 1;
  CUDA_CHECK(cudaMalloc((void**)&raw, (count + head_room) * sizeof(T)));
  T* np = raw + head_room;
  if (p && n) CUDA_CHECK(cudaMemcpyAsync(np - head, p - head, (n + head) * sizeof(T), cudaMemcpyDeviceToDevice, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (p) cudaFree(p - head);
  p = np;
  n = count;
  head = head_room;
}
template struct DevBuf<float2>;
template struct DevBuf<int>;
template struct DevBuf<unsigned long long>;
template struct DevBuf<Counters>;
template struct DevBuf<uint32_t>;
template struct DevBuf<PortalSide>;
template struct DevBuf<uint8_t>;  // also unsigned char

// ---- construction --------------------------------------------------------------------
Engine::Engine(int device) : device_(device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error(-1, std::string("no CUDA device: ptoy_b200 has no CPU fallback (") +
                        cudaGetErrorString(e) + ")");
  if (device < 0 || device >= count) throw Error(-2, "invalid CUDA device ordinal");
  CUDA_CHECK(cudaSetDevice(device));
  CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  CUDA_CHECK(cudaMallocHost((void**)&h_ctr_, sizeof(Counters)));
  CUDA_CHECK(cudaMallocHost((void**)&h_err_, sizeof(int)));
  *h_err_ = 0;
  CUDA_CHECK(cudaEventCreateWithFlags(&err_event_, cudaEventDisableTiming));
  ctr_.alloc(1);
  ticket_.alloc(1);
  CUDA_CHECK(cudaMemsetAsync(ctr_.p, 0, sizeof(Counters), stream_));
  CUDA_CHECK(cudaMemsetAsync(ticket_.p, 0, sizeof(int), stream_));
  // Particles::Particles() member initialisers (particles.cpp:26-33, 61-68)
  bs_ = V2{kBlockSize(), kBlockSize()};
  dt_ = kTimeStep;
  gravity_ = vmul(V2{0.f, -1.f}, kGravity);
  gravity_enable_ = true;
  renderer_ready_ = true;  // particles.h:210
  gA_ = gB_ = V2{0.f, 0.f};
  di_ = dj_ = 0;
  const V2 A{-1.f, -1.f}, B{1.f, 1.f};
  domA_ = A; domB_ = B;
  init_grid(A, B);
  rqA_ = A; rqB_ = B;
  reset_env_frame(A, B);
  t_ = 0.f;
}

Engine::~Engine() {
  cudaSetDevice(device_);
  if (owns_stream_ && stream_) cudaStreamSynchronize(stream_); else cudaDeviceSynchronize();
  for (auto& tl : timed_) { cudaEventDestroy(tl.a); cudaEventDestroy(tl.b); }
  for (auto ev : event_pool_) cudaEventDestroy(ev);
  if (h_ctr_) cudaFreeHost(h_ctr_);
  if (h_err_) cudaFreeHost(h_err_);
  if (err_event_) cudaEventDestroy(err_event_);
  delete transport_;
  if (stream_ && owns_stream_) cudaStreamDestroy(stream_);
}

// blocks::InitEmptyBlocks (blocks.h:256-276); particles are re-keyed by the caller
void Engine::init_grid(V2 A, V2 B) {
  gA_ = A; gB_ = B;
  const V2 size = vsub(B, A);
  di_ = f2i_trunc(size.x / bs_.x) + 1;
  dj_ = f2i_trunc(size.y / bs_.y) + 1;
  if (di_ <= 0 || dj_ <= 0) throw Error(-2, "degenerate block grid");
  // The reference's neighbour stencil wraps across short columns (blocks.h:268-275,
  // particles.cpp:885-889): with fewer than 3 blocks per column the wrapped cell is a real
  // neighbour and pairs would be double counted.  Not supported.
  if (dj_ < 3) throw Error(-2, "block grid needs at least 3 blocks per column (domain height >= 0.16)");
  alloc_cells();
  env_dirty_ = true;
}

// The loop of blocks::SetDomain (blocks.h:119-133): grow-only, by accumulation.
bool Engine::grown_grid(V2 pA, V2 pB, V2* nA, V2* nB) const {
  V2 A = gA_, B = gB_;
  while (pA.x < A.x) A.x -= bs_.x;
  while (pA.y < A.y) A.y -= bs_.y;
  while (pB.x > B.x) B.x += bs_.x;
  while (pB.y > B.y) B.y += bs_.y;
  *nA = A; *nB = B;
  return !(veq(A, gA_) && veq(B, gB_));
}

void Engine::alloc_cells() {
  const long long SIg = 2ll * di_ + 6, SJ = 2ll * dj_ + 6;
  // strip decomposition: global sub-rows owned by this rank; one rank owns all
  // rows between the global padding rows
  if (strip_.world == 1) { strip_.row_begin[0] = 2; strip_.row_begin[1] = (int)SIg - 2; }
  else {
    for (int r = 0; r <= strip_.world; ++r)
      strip_.row_begin[r] = r == 0 ? 2 : (r == strip_.world ? (int)SIg - 2 : 2 * strip_.col_begin[r] + 4);
    for (int r = 0; r < strip_.world; ++r)
      if (strip_.row_begin[r + 1] - strip_.row_begin[r] < 4) throw Error(-2, "a strip needs at least 2 block columns");
  }
  const int GR = strip_.world > 1 ? strip_.ghost_rows : 2;
  const int own_rows = strip_.row_begin[strip_.rank + 1] - strip_.row_begin[strip_.rank];
  for (int r = 0; r + 1 < strip_.world; ++r)
    if (strip_.row_begin[r + 1] - strip_.row_begin[r] < GR) throw Error(-2, "a strip is narrower than the ghost zone");
  const long long SI = own_rows + 2 * GR;  // + the ghost rows on either side
  const long long ncell = SI * SJ;
  if (ncell + 8 >= (1ll << 31)) throw Error(-4, "sub-cell grid exceeds 2^31 cells");
  if (SIg > 65535 || SJ > 65535) throw Error(-4, "block grid exceeds 32764 blocks per axis");
  grid_.Ax = gA_.x; grid_.Ay = gA_.y;
  grid_.bsx = bs_.x; grid_.bsy = bs_.y;
  grid_.di = di_; grid_.dj = dj_;
  grid_.fdi = (float)di_; grid_.fdj = (float)dj_;
  grid_.SI = (int)SI; grid_.SJ = (int)SJ;
  grid_.ncell = (int)ncell;
  grid_.row0 = strip_.row_begin[strip_.rank] - GR;
  grid_.own_lo = GR;
  grid_.own_hi = (int)SI - GR;
  grid_.inv_hsx = 2.0f / bs_.x;
  grid_.inv_hsy = 2.0f / bs_.y;
  grid_.wmax = (float)std::max(SIg, SJ) + 8.f;
  grid_.tol = grid_.wmax * std::ldexp(1.f, -21);  // >= 2 * |w| * 2^-22 for |w| < wmax
  n_items_ = (int)(((ncell + 2) + 3) / 4 * 4);
  cnt_.alloc((size_t)n_items_);
  start_.alloc((size_t)n_items_ + 4);
  CUDA_CHECK(cudaMemsetAsync(cnt_.p, 0, cnt_.n * sizeof(int), stream_));
  CUDA_CHECK(cudaMemsetAsync(start_.p, 0, start_.n * sizeof(int), stream_));
  const size_t tiles = ((size_t)n_items_ + scan_tile_items() - 1) / scan_tile_items();
  if (tiles > tile_state_.n) {
    tile_state_.alloc(tiles);
    CUDA_CHECK(cudaMemsetAsync(tile_state_.p, 0, tile_state_.n * sizeof(unsigned long long), stream_));
    epoch_ = 0;
  }
  blk_flags_.alloc((size_t)di_ * dj_);
  // search margins in sub-cell units
  const float W = (float)std::max(SIg, SJ);
  margin1_ = std::ldexp(1.f, -19) + W * std::ldexp(1.f, -21) + grid_.tol;
  const float drift = kVelocityLimit * dt_ * 0.5f;
  margin2_ = margin1_ + 2.f * drift / (0.5f * bs_.x) * 1.001f + W * std::ldexp(1.f, -21);
  // the same with the drift stage 1 measured instead of its bound (tile kernels)
  margin2_base_ = margin1_ + W * std::ldexp(1.f, -21);
  margin2_scale_ = 2.f / (0.5f * std::min(bs_.x, bs_.y)) * 1.001f;
  if (!(margin2_ < 0.5f)) throw Error(-2, "time step too large for the sub-cell search margin");
  if (strip_.world > 1) alloc_strip_buffers();
}

void Engine::ensure_capacity(size_t n) {
  // strips: the ghost rows live right before slot 0 and right behind the last owned particle
  const size_t gh = strip_.world > 1 ? (size_t)strip_.ghost_cap : 0;
  const size_t zone = strip_.world > 1 ? (size_t)zone_entries_ * kZoneCap : 0;   // portal zone table behind the ghosts
  if (n <= cap_ && pos_[0].head >= gh && pos_h_.head >= gh && pos_[0].n >= cap_ + gh + zone &&
      pos_[1].n >= cap_ + gh + zone && pos_h_.n >= cap_ + gh + zone)
    return;
  size_t nc = n <= cap_ ? cap_ : std::max<size_t>(n, cap_ + cap_ / 2);
  nc = (nc + 1023) / 1024 * 1024;
  for (int k = 0; k < 2; ++k) {
    pos_[k].grow_keep(nc + gh + zone, stream_, gh);
    vel_[k].grow_keep(nc, stream_);
    id_[k].grow_keep(nc, stream_);
  }
  pos_h_.alloc(nc + gh + zone, gh);
  vel_h_.alloc(nc);
  newcell_.alloc(nc);
  bslot_.alloc(nc * 8);
  for (int k = 0; k < 2; ++k) cell_[k].alloc(nc);  // rewritten by every reorder before use
  order_.alloc(nc);
  fixlist_.alloc(nc);
  force_dbg_.release();
  cap_ = nc;
}

void Engine::ensure_id_capacity(size_t n) {
  if (n <= id_cap_) return;
  if (n >= (1ull << 31)) throw Error(-2, "particle id too large");
  const size_t old = slot_of_id_.n && slot_of_id_.p ? id_cap_ : 0;
  size_t nc = std::max<size_t>(n, id_cap_ + id_cap_ / 2);
  slot_of_id_.grow_keep(nc + 1, stream_);
  launch_fill_i32(slot_of_id_.p + old, kSlotNone, nc + 1 - old, stream_);
  ++launches_;
  id_cap_ = nc;
  frozen_dirty_ = true;
  bonds_dirty_ = true;
}

// ---- scene ------------------------------------------------------------------------------
void Engine::reset_scene(V2 A, V2 B) {
  CUDA_CHECK(cudaSetDevice(device_));
  portals_.clear();
  portal_stage_ = 0;
  bonds_.clear(); bonds_dirty_ = true;
  frozen_.clear(); frozen_dirty_ = true;
  norend_buf_.clear();
  CUDA_CHECK(cudaMemsetAsync(ctr_.p, 0, sizeof(Counters), stream_));
  *h_err_ = 0;
  err_pending_ = false;
  n_bound_ = 0;
  removed_seen_ = 0;
  slot_map_valid_ = false;
  if (slot_of_id_.p) { launch_fill_i32(slot_of_id_.p, kSlotNone, slot_of_id_.n, stream_); ++launches_; }
  set_domain(A, B);
  rqA_ = A; rqB_ = B;
  reset_env_frame(A, B);
  t_ = 0.f;
  force_enabled_ = false;
  pick_enabled_ = false;
  remove_last_portal_ = false;
  renderer_ready_ = false;
}

// Particles::SetDomain (particles.h:108-111) -> blocks::SetDomain (blocks.h:119-141)
void Engine::set_domain(V2 A, V2 B) {
  CUDA_CHECK(cudaSetDevice(device_));
  domA_ = A; domB_ = B;
  V2 nA, nB;
  if (grown_grid(A, B, &nA, &nB)) {
    init_grid(nA, nB);
    if (n_bound_ > 0) {
      sync_env();
      rekey_and_reorder(0, false);  // AddParticles(old_data): re-bin everything
    }
  }
}

// particles.h:99-107 ResetEnvObjFrame: bottom, top, left, right
void Engine::reset_env_frame(V2 A, V2 B) {
  walls_ = {A.x, A.y, B.x, A.y, A.x, B.y, B.x, B.y, A.x, A.y, A.x, B.y, B.x, A.y, B.x, B.y};
  walls_are_frame_ = true;
  env_dirty_ = true;
}

void Engine::set_walls(const float* seg, size_t n) {
  if (n > (size_t)kMaxWalls) throw Error(-4, "too many walls (kMaxWalls = 16)");
  walls_.assign(seg, seg + 4 * n);
  walls_are_frame_ = false;
  env_dirty_ = true;
}

// Particles::UpdatePortalBlocks (particles.cpp:699-706) with Portal::IsClose
// (particles.h:78-80), blocks::GetCenter / GetCircumRadius (blocks.h:142-150).  Only
// blocks inside the segment's bounding box grown by 4 blocks can pass IsClose, so the
// scan is restricted to those (ascending block index, like the reference's loop).
void Engine::update_portal_blocks(PortalHost& p) const {
  p.blocks.clear();
  const float R = (float)((double)vlen(bs_) * 0.5);
  const double pad = 4.0 * (double)bs_.x;
  const double lox = std::min(p.begin.x, p.end.x) - pad, hix = std::max(p.begin.x, p.end.x) + pad;
  const double loy = std::min(p.begin.y, p.end.y) - pad, hiy = std::max(p.begin.y, p.end.y) + pad;
  auto clampi = [](double v, int lo, int hi) { return (int)std::max<double>(lo, std::min<double>(hi, v)); };
  const int i0 = clampi(std::floor((lox - gA_.x) / bs_.x), 0, di_ - 1);
  const int i1 = clampi(std::floor((hix - gA_.x) / bs_.x), 0, di_ - 1);
  const int j0 = clampi(std::floor((loy - gA_.y) / bs_.y), 0, dj_ - 1);
  const int j1 = clampi(std::floor((hiy - gA_.y) / bs_.y), 0, dj_ - 1);
  if (!(lox == lox && loy == loy)) return;  // NaN portal
  for (int mi = i0; mi <= i1; ++mi)
    for (int mj = j0; mj <= j1; ++mj) {
      const V2 off{(float)((0.5 + mi) * (double)bs_.x), (float)((0.5 + mj) * (double)bs_.y)};
      const V2 c = vadd(gA_, off);
      if (vdist(seg_nearest(p.begin, p.end, c), c) < R + 2 * kPortalThickness)
        p.blocks.push_back(mi * dj_ + mj);
    }
}

// largest float v2 with fl(sqrt(v2)) <= limit (sqrt is monotone): bisection on the bit pattern
static float find_v2c(float limit) {
  auto over = [&](float v2) { return (float)std::sqrt((double)v2) > limit; };
  uint32_t lo, hi;
  float a = limit * limit * 0.5f, b = limit * limit * 2.f;
  std::memcpy(&lo, &a, 4);
  std::memcpy(&hi, &b, 4);
  while (hi - lo > 1) {  // invariant: !over(lo), over(hi)
    uint32_t mid = lo + (hi - lo) / 2;
    float m;
    std::memcpy(&m, &mid, 4);
    if (over(m)) hi = mid; else lo = mid;
  }
  float r;
  std::memcpy(&r, &lo, 4);
  return r;
}

static float find_r2c(float thr, float RR) {
  // smallest float r2 >= thr for which F12's coefficient is clamped to zero, i.e.
  // fl(fl(1/r2) * RR) > 1 fails (see kernels.cu pair_force).  Monotone -> bisection on bits.
  auto nonzero = [&](float r2) { return ((float)(1. / (double)r2)) * RR > 1.0f; };
  uint32_t lo, hi;
  float one = 1.0f;
  std::memcpy(&lo, &thr, 4);
  std::memcpy(&hi, &one, 4);
  // invariant: nonzero(lo) true, nonzero(hi) false
  while (hi - lo > 1) {
    uint32_t mid = lo + (hi - lo) / 2;
    float m;
    std::memcpy(&m, &mid, 4);
    if (nonzero(m)) lo = mid; else hi = mid;
  }
  float r;
  std::memcpy(&r, &hi, 4);
  return r;
}

Phys Engine::make_phys() const {
  Phys p{};
  const float m = kMass();
  p.dt = dt_;
  p.c1 = (float)((double)dt_ * 0.5 / (double)m);
  p.c2 = (float)((double)dt_ / (double)m);
  const V2 gf = vmul(gravity_, m);
  p.gfx = gf.x; p.gfy = gf.y;
  p.gravity_enable = gravity_enable_;
  p.diss = kDissipation * m;
  const float R = (float)(2. * (double)kRadius);
  p.RR = R * R;
  p.thr = (float)(std::pow((double)kRadius, 2) * 1e-3);
  p.r2c = r2c_;
  p.r2c_hi = r2c_ * (1.0f + 8.0f * std::ldexp(1.f, -23));
  p.v2c = v2c_;
  p.radius = kRadius;
  p.RRw = kRadius * kRadius;
  p.wall_reach = wall_reach_;
  const float Re = 3.f * kRadius;
  p.RRe = Re * Re;
  p.vlimit = kVelocityLimit;
  p.bondR = R;
  p.bond_y = (float)(1. / (double)R);
  p.bond_cutoff = (float)(8. * (double)kRadius);
  p.sigma_bond = kSigmaBond;
  p.bond_portal_reach = 2. * (double)kPortalThickness + (double)p.bond_cutoff;
  p.pickR = (float)(0.1 * (double)kRadius);
  p.sigma_pick = kSigmaPick;
  p.portal_thickness = kPortalThickness;
  p.two_pt = 2. * (double)kPortalThickness;
  p.force_enabled = force_enabled_;
  p.force_attractive = force_attractive_;
  p.fcx = force_center_.x; p.fcy = force_center_.y;
  p.point_force = kPointForce;
  p.point_force_attractive = kPointForceAttractive;
  p.pick_enabled = pick_enabled_;
  p.pick_id = pick_id_;
  p.pkx = pick_pointer_.x; p.pky = pick_pointer_.y;
  return p;
}

// Upload everything derived from walls / portals / grid.
void Engine::sync_env() {
  if (!env_dirty_) return;
  if (r2c_ == 0.f) {
    const float R = (float)(2. * (double)kRadius);
    r2c_ = find_r2c((float)(std::pow((double)kRadius, 2) * 1e-3), R * R);
    v2c_ = find_v2c(kVelocityLimit);
  }
  // --- walls: force is exactly zero when |p - Q| >= r(1 + 2^-22); Q and the differences
  // carry rounding errors of a few ulp of the coordinates, so grow by 16 ulp of the
  // largest coordinate in play.
  float cmax = 1.f;
  for (float v : {gA_.x, gA_.y, gB_.x, gB_.y, domA_.x, domA_.y, domB_.x, domB_.y}) cmax = std::max(cmax, std::fabs(v));
  for (float v : walls_) cmax = std::max(cmax, std::fabs(v));
  const float ulp = std::ldexp(cmax, -23);
  wall_reach_ = kRadius * 1.0001f + 16.f * ulp;
  const size_t nw = walls_.size() / 4;
  for (size_t k = 0; k < nw; ++k) {
    WallDev& w = walls_dev_[k];
    w.ax = walls_[4 * k]; w.ay = walls_[4 * k + 1]; w.bx = walls_[4 * k + 2]; w.by = walls_[4 * k + 3];
    w.lox = std::min(w.ax, w.bx) - wall_reach_; w.hix = std::max(w.ax, w.bx) + wall_reach_;
    w.loy = std::min(w.ay, w.by) - wall_reach_; w.hiy = std::max(w.ay, w.by) + wall_reach_;
  }
  if (walls_are_frame_) {
    free_box_[0] = domA_.x + wall_reach_; free_box_[1] = domA_.y + wall_reach_;
    free_box_[2] = domB_.x - wall_reach_; free_box_[3] = domB_.y - wall_reach_;
  } else {
    free_box_[0] = free_box_[1] = 1.f; free_box_[2] = free_box_[3] = -1.f;  // empty
  }
  // --- portals
  const size_t nblk = (size_t)di_ * dj_;
  std::vector<uint8_t> flags;
  has_portal_tables_ = !portals_.empty();
  if (has_portal_tables_) {
    flags.assign(nblk, 0);
    std::vector<PortalSide> sides;
    std::vector<int> ptr{0}, blk;
    // conservative reach of F12_portal_edge around an endpoint, for a particle that may
    // have drifted half a step from where it was binned
    const double reach = 3.0 * kRadius * 1.001 + kVelocityLimit * dt_ * 0.5 * 1.01 + 16.0 * ulp;
    for (auto& pr : portals_) {
      PortalHost* two[2] = {&pr.first, &pr.second};
      for (int d = 0; d < 2; ++d) {
        PortalHost& p = *two[d];
        update_portal_blocks(p);  // UpdateEnvObj re-derives them (particles.cpp:720-724)
        PortalSide s{};
        s.ax = p.begin.x; s.ay = p.begin.y; s.bx = p.end.x; s.by = p.end.y;
        const V2 r = vsub(p.end, p.begin);
        const V2 n = vnormalized(V2{-r.y, r.x});
        s.rx = r.x; s.ry = r.y; s.nx = n.x; s.ny = n.y;
        s.rr = vdot(r, r);
        sides.push_back(s);
        for (int b : p.blocks) { blk.push_back(b); flags[b] |= 1; }
        ptr.push_back((int)blk.size());
        for (const V2& e : {p.begin, p.end}) {
          if (!(e.x == e.x && e.y == e.y)) continue;
          auto clampi = [](double v, int lo, int hi) { return (int)std::max<double>(lo, std::min<double>(hi, v)); };
          const int i0 = clampi(std::floor((e.x - reach - gA_.x) / bs_.x) - 1, 0, di_ - 1);
          const int i1 = clampi(std::floor((e.x + reach - gA_.x) / bs_.x) + 1, 0, di_ - 1);
          const int j0 = clampi(std::floor((e.y - reach - gA_.y) / bs_.y) - 1, 0, dj_ - 1);
          const int j1 = clampi(std::floor((e.y + reach - gA_.y) / bs_.y) + 1, 0, dj_ - 1);
          for (int mi = i0; mi <= i1; ++mi)
            for (int mj = j0; mj <= j1; ++mj) flags[(size_t)mi * dj_ + mj] |= 2;
        }
      }
    }
    if (strip_.world > 1) {
      // Strips: the portal-zone table also serves bond partners on the far side of a
      // portal.  Those need not lie in Portal::blocks, so the table is extended - behind the
      // regular entries, which keep their indices - by the blocks within bond reach of a portal.
      const float R = (float)((double)vlen(bs_) * 0.5);
      const double reach_b = 0.25;
      for (auto& pr : portals_)
        for (const PortalHost* p : {&pr.first, &pr.second}) {
          if (!(p->begin.x == p->begin.x && p->end.x == p->end.x)) continue;
          auto clampi = [](double v, int lo, int hi) { return (int)std::max<double>(lo, std::min<double>(hi, v)); };
          const double pad = reach_b + 2.0 * bs_.x;
          const int i0 = clampi(std::floor((std::min(p->begin.x, p->end.x) - pad - gA_.x) / bs_.x), 0, di_ - 1);
          const int i1 = clampi(std::floor((std::max(p->begin.x, p->end.x) + pad - gA_.x) / bs_.x), 0, di_ - 1);
          const int j0 = clampi(std::floor((std::min(p->begin.y, p->end.y) - pad - gA_.y) / bs_.y), 0, dj_ - 1);
          const int j1 = clampi(std::floor((std::max(p->begin.y, p->end.y) + pad - gA_.y) / bs_.y), 0, dj_ - 1);
          for (int mi = i0; mi <= i1; ++mi)
            for (int mj = j0; mj <= j1; ++mj) {
              const V2 c = vadd(gA_, V2{(float)((0.5 + mi) * (double)bs_.x), (float)((0.5 + mj) * (double)bs_.y)});
              const int b = mi * dj_ + mj;
              if (vdist(seg_nearest(p->begin, p->end, c), c) < R + reach_b &&
                  !std::binary_search(p->blocks.begin(), p->blocks.end(), b))
                blk.push_back(b);
            }
        }
    }
    sides_.alloc(sides.size());
    pblk_ptr_.alloc(ptr.size());
    pblk_.alloc(blk.size() + 1);
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    CUDA_CHECK(cudaMemcpy(sides_.p, sides.data(), sides.size() * sizeof(PortalSide), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(pblk_ptr_.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!blk.empty())
      CUDA_CHECK(cudaMemcpy(pblk_.p, blk.data(), blk.size() * sizeof(int), cudaMemcpyHostToDevice));
    blk_flags_.alloc(nblk);
    CUDA_CHECK(cudaMemcpy(blk_flags_.p, flags.data(), nblk, cudaMemcpyHostToDevice));
    zone_entries_ = (int)blk.size();
    if (strip_.world > 1) {   // room for the portal zone table
      zcnt_.alloc(blk.size() + 1);
      zid_.alloc(blk.size() * kZoneCap + 1);
      ensure_capacity(cap_);
    }
  } else {
    zone_entries_ = 0;
  }
  env_dirty_ = false;
}

// bonds_ -> CSR over ids (row = first, columns ascending = std::set iteration order)
void Engine::sync_bonds() {
  if (!bonds_dirty_) return;
  bonds_dirty_ = false;
  if (!has_bonds()) return;   // (strips: a rank without bonds of its own still keeps the table for arrivals)
  int maxid = 0;
  for (auto& b : bonds_) maxid = std::max(maxid, std::max(b.first, b.second));
  ensure_id_capacity((size_t)maxid + 1);
  bonds_dirty_ = false;
  std::vector<uint32_t> ptr(id_cap_ + 2, 0);
  std::vector<int> col(bonds_.size());
  for (auto& b : bonds_) ptr[(size_t)b.first + 1]++;
  for (size_t i = 0; i + 1 < ptr.size(); ++i) ptr[i + 1] += ptr[i];
  for (size_t e = 0; e < bonds_.size(); ++e) col[e] = bonds_[e].second;  // already row-major sorted
  // the same rows as a fixed-width table (first 6 partners + degree): two 16-byte loads per id
  std::vector<int> ell((id_cap_ + 1) * 8, -1);
  for (size_t id = 0; id <= id_cap_; ++id) {
    const uint32_t e0 = ptr[id], e1 = ptr[id + 1];
    for (uint32_t k = 0; k < e1 - e0 && k < 6; ++k) ell[id * 8 + k] = col[e0 + k];
    ell[id * 8 + 7] = (int)(e1 - e0);
  }
  bond_ell_.alloc(ell.size());
  bond_ptr_.alloc(ptr.size());
  bond_col_.alloc(col.size() + 1);
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  CUDA_CHECK(cudaMemcpy(bond_ptr_.p, ptr.data(), ptr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (!col.empty())
    CUDA_CHECK(cudaMemcpy(bond_col_.p, col.data(), col.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(bond_ell_.p, ell.data(), ell.size() * sizeof(int), cudaMemcpyHostToDevice));
}

void Engine::sync_frozen() {
  if (!frozen_dirty_) return;
  frozen_dirty_ = false;
  if (frozen_.empty()) return;
  ensure_id_capacity((size_t)frozen_.back() + 1);
  frozen_dirty_ = false;
  std::vector<uint32_t> bits((id_cap_ + 32) / 32, 0u);
  for (int id : frozen_) bits[(size_t)id >> 5] |= 1u << (id & 31);
  frozen_bits_.alloc(bits.size());
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  CUDA_CHECK(cudaMemcpy(frozen_bits_.p, bits.data(), bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
}

void Engine::ensure_slot_map() {
  if (slot_map_valid_ || id_cap_ == 0) return;
  launch_fill_i32(slot_of_id_.p, kSlotNone, id_cap_ + 1, stream_);
  launch_slot_map(ctr_.p, id_[cur_].p, n_bound_, slot_of_id_.p, stream_);
  launches_ += 2;
  slot_map_valid_ = true;
}

void Engine::set_bonds(const int32_t* pairs, size_t n) {
  bonds_.resize(n);
  for (size_t i = 0; i < n; ++i) {
    if (pairs[2 * i] < 0 || pairs[2 * i + 1] < 0) throw Error(-2, "negative particle id in bond");
    bonds_[i] = {pairs[2 * i], pairs[2 * i + 1]};
  }
  if (!std::is_sorted(bonds_.begin(), bonds_.end())) std::sort(bonds_.begin(), bonds_.end());
  bonds_.erase(std::unique(bonds_.begin(), bonds_.end()), bonds_.end());
  bonds_dirty_ = true;
  read_counters();
  removed_seen_ = h_ctr_->removed_total;
}

// particles.cpp:484-496: erase both directions if (a,b) is present, else insert both
void Engine::toggle_bond(int a, int b) {
  if (a < 0 || b < 0) throw Error(-2, "negative particle id in bond");
  prune_bonds();
  auto find = [&](int x, int y) { return std::lower_bound(bonds_.begin(), bonds_.end(), std::make_pair(x, y)); };
  auto it = find(a, b);
  if (it != bonds_.end() && *it == std::make_pair(a, b)) {
    bonds_.erase(it);
    auto jt = find(b, a);
    if (jt != bonds_.end() && *jt == std::make_pair(b, a)) bonds_.erase(jt);
  } else {
    bonds_.insert(it, {a, b});
    auto jt = find(b, a);
    if (jt == bonds_.end() || *jt != std::make_pair(b, a)) bonds_.insert(jt, {b, a});
  }
  bonds_dirty_ = true;
}

// CheckBonds (particles.cpp:205-215).  On the device a bond whose partner left the grid is
// skipped by the force kernel from that iteration on; the host list is pruned lazily.
void Engine::prune_bonds() {
  if (bonds_.empty()) return;
  // Decomposed: a slot < 0 may be a particle of another strip, and `removed_total` counts
  // migrants; whether a partner is gone is not a local question.  The device skips dead
  // partners either way; the host list keeps them (documented in include/ptoy_b200.h).
  if (strip_.world > 1) return;
  read_counters();
  if (h_ctr_->removed_total == removed_seen_) return;
  removed_seen_ = h_ctr_->removed_total;
  ensure_slot_map();
  std::vector<int> slot(id_cap_ + 1);
  CUDA_CHECK(cudaMemcpyAsync(slot.data(), slot_of_id_.p, slot.size() * sizeof(int), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  auto dead = [&](const std::pair<int, int>& b) {
    return (size_t)b.first >= id_cap_ || (size_t)b.second >= id_cap_ || slot[b.first] == kSlotNone ||
           slot[b.second] == kSlotNone;
  };
  bonds_.erase(std::remove_if(bonds_.begin(), bonds_.end(), dead), bonds_.end());
  bonds_dirty_ = true;
}

const std::vector<std::pair<int, int>>& Engine::bonds() {
  prune_bonds();
  return bonds_;
}

void Engine::set_frozen(const int32_t* ids, size_t n) {
  frozen_.assign(ids, ids + n);
  for (int id : frozen_) if (id < 0) throw Error(-2, "negative particle id in frozen set");
  std::sort(frozen_.begin(), frozen_.end());
  frozen_.erase(std::unique(frozen_.begin(), frozen_.end()), frozen_.end());
  frozen_dirty_ = true;
}

// PortalStart/PortalStop twice (particles.cpp:411-450)
void Engine::add_portal_pair(const float* s) {
  tool(3, 0, V2{s[0], s[1]}); tool(3, 2, V2{s[2], s[3]});
  tool(3, 0, V2{s[4], s[5]}); tool(3, 2, V2{s[6], s[7]});
}

void Engine::clear_portals() {
  portals_.clear();
  portal_stage_ = 0;
  env_dirty_ = true;
}

void Engine::read_counters() {
  CUDA_CHECK(cudaSetDevice(device_));
  CUDA_CHECK(cudaMemcpyAsync(h_ctr_, ctr_.p, sizeof(Counters), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  n_bound_ = h_ctr_->n_cur;
}

int Engine::error_flags() {
  read_counters();
  return h_ctr_->err_flags;
}

size_t Engine::num_particles() {
  read_counters();
  return (size_t)h_ctr_->n_cur;
}

void Engine::add_particles(const float* pos, const float* vel, const int32_t* id, size_t n, bool on_device) {
  CUDA_CHECK(cudaSetDevice(device_));
  if (n == 0) return;
  prune_bonds();  // a re-added id must not resurrect bonds CheckBonds erased
  sync_env();
  read_counters();
  const size_t n_old = (size_t)h_ctr_->n_cur;
  if (n_old + n >= (1ull << 31) - 1024) throw Error(-4, "more than 2^31 particles");
  ensure_capacity(n_old + n + (size_t)arrivals_bound());  // strips: room for migrants
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CUDA_CHECK(cudaMemcpyAsync(pos_[cur_].p + n_old, pos, n * sizeof(float2), kind, stream_));
  CUDA_CHECK(cudaMemcpyAsync(vel_[cur_].p + n_old, vel, n * sizeof(float2), kind, stream_));
  CUDA_CHECK(cudaMemcpyAsync(id_[cur_].p + n_old, id, n * sizeof(int), kind, stream_));
  int maxid = -1, minid = 0;
  if (on_device) {
    int* d_tmp = order_.p;  // scratch: two ints
    const int init[2] = {-1, std::numeric_limits<int>::max()};
    CUDA_CHECK(cudaMemcpyAsync(d_tmp, init, sizeof(init), cudaMemcpyHostToDevice, stream_));
    launch_max_id(id_[cur_].p + n_old, n, d_tmp, stream_);
    ++launches_;
    int out[2];
    CUDA_CHECK(cudaMemcpyAsync(out, d_tmp, sizeof(out), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    maxid = out[0]; minid = out[1];
  } else {
    for (size_t i = 0; i < n; ++i) { maxid = std::max(maxid, id[i]); minid = std::min(minid, id[i]); }
  }
  if (minid < 0) throw Error(-2, "negative particle id");  // blocks.h:111 assert
  ensure_id_capacity((size_t)maxid + 1);
  const int n_new = (int)(n_old + n);
  CUDA_CHECK(cudaMemcpyAsync(&ctr_.p->n_cur, &n_new, sizeof(int), cudaMemcpyHostToDevice, stream_));
  n_bound_ = n_new;
  rekey_and_reorder(0, false);
}

void Engine::upload_state(const float* pos, const float* vel, const int32_t* id, size_t n) {
  CUDA_CHECK(cudaSetDevice(device_));
  prune_bonds();  // CheckBonds against the state being replaced, before it disappears
  const int zero = 0;
  CUDA_CHECK(cudaMemcpyAsync(&ctr_.p->n_cur, &zero, sizeof(int), cudaMemcpyHostToDevice, stream_));
  n_bound_ = 0;
  slot_map_valid_ = false;
  if (slot_of_id_.p && has_bonds()) { launch_fill_i32(slot_of_id_.p, kSlotNone, id_cap_ + 1, stream_); ++launches_; }
  add_particles(pos, vel, id, n, false);
}

// Pipelined forms: everything is enqueued on this engine's stream and nothing waits, so
// several handles (one stream each) overlap upload, stepping and download of independent
// batches.  n_old = 0 and the id bound come from the caller instead of a device read / host scan.
void Engine::upload_state_async(const float* pos, const float* vel, const int32_t* id, size_t n, int32_t max_id) {
  CUDA_CHECK(cudaSetDevice(device_));
  if (max_id < 0 && n) throw Error(-2, "upload_state_async needs the largest id of the batch");
  if (n >= (1ull << 31) - 1024) throw Error(-4, "more than 2^31 particles");
  sync_env();
  slot_map_valid_ = false;
  if (n) {
    ensure_capacity(n + (size_t)arrivals_bound());
    ensure_id_capacity((size_t)max_id + 1);
  }
  if (slot_of_id_.p && has_bonds()) { launch_fill_i32(slot_of_id_.p, kSlotNone, id_cap_ + 1, stream_); ++launches_; }
  launch_fill_i32(&ctr_.p->n_cur, (int)n, 1, stream_);
  ++launches_;
  n_bound_ = (int)n;
  if (!n) return;
  CUDA_CHECK(cudaMemcpyAsync(pos_[cur_].p, pos, n * sizeof(float2), cudaMemcpyHostToDevice, stream_));
  CUDA_CHECK(cudaMemcpyAsync(vel_[cur_].p, vel, n * sizeof(float2), cudaMemcpyHostToDevice, stream_));
  CUDA_CHECK(cudaMemcpyAsync(id_[cur_].p, id, n * sizeof(int), cudaMemcpyHostToDevice, stream_));
  rekey_and_reorder(0, false);
}

void Engine::download_state_async(float* pos, float* vel, int32_t* id, size_t cap) {
  CUDA_CHECK(cudaSetDevice(device_));
  const size_t n = std::min(cap, (size_t)std::max(n_bound_, 0));
  if (n) {
    if (pos) CUDA_CHECK(cudaMemcpyAsync(pos, pos_[cur_].p, n * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
    if (vel) CUDA_CHECK(cudaMemcpyAsync(vel, vel_[cur_].p, n * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
    if (id) CUDA_CHECK(cudaMemcpyAsync(id, id_[cur_].p, n * sizeof(int), cudaMemcpyDeviceToHost, stream_));
  }
  CUDA_CHECK(cudaMemcpyAsync(h_ctr_, ctr_.p, sizeof(Counters), cudaMemcpyDeviceToHost, stream_));
}

size_t Engine::download_state(float* pos, float* vel, int32_t* id, size_t cap) {
  read_counters();
  const size_t n = std::min(cap, (size_t)h_ctr_->n_cur);
  if (n) {
    if (pos) CUDA_CHECK(cudaMemcpyAsync(pos, pos_[cur_].p, n * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
    if (vel) CUDA_CHECK(cudaMemcpyAsync(vel, vel_[cur_].p, n * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
    if (id) CUDA_CHECK(cudaMemcpyAsync(id, id_[cur_].p, n * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
  }
  return (size_t)h_ctr_->n_cur;
}

// ---- one iteration -------------------------------------------------------------------------
StageArgs Engine::make_stage_args(int stage) {
  StageArgs a{};
  a.g = grid_;
  a.ph = make_phys();
  const size_t nw = walls_.size() / 4;
  for (size_t k = 0; k < nw; ++k) a.walls[k] = walls_dev_[k];
  a.n_walls = (int)nw;
  a.free_lox = free_box_[0]; a.free_loy = free_box_[1]; a.free_hix = free_box_[2]; a.free_hiy = free_box_[3];
  a.ctr = ctr_.p;
  a.pos = pos_[cur_].p; a.vel = vel_[cur_].p;
  a.pos_h = pos_h_.p; a.vel_h = vel_h_.p;
  a.pos_out = pos_[cur_].p; a.vel_out = vel_[cur_].p;
  a.id = id_[cur_].p;
  a.start = start_.p;
  a.dist = make_dist(true);
  a.margin = (stage == 1) ? margin1_ : margin2_;
  a.margin_base = margin2_base_;
  a.margin_scale = margin2_scale_;
  a.dmax = &ctr_.p->dmax;
  a.has_frozen = !frozen_.empty();
  a.frozen_bits = frozen_bits_.p;
  a.has_bonds = has_bonds();
  a.bond_ptr = bond_ptr_.p; a.bond_col = bond_col_.p; a.slot_of_id = slot_of_id_.p;
  a.bslot = bslot_.p; a.bond_ell = bond_ell_.p;
  a.cap = (int)cap_;
  a.id_cap = (int)id_cap_;
  a.err = &ctr_.p->err_flags;
  a.lost = ctr_.p->lost;
  a.n_sides = has_portal_tables_ ? (int)portals_.size() * 2 : 0;
  a.sides = sides_.p; a.pblk_ptr = pblk_ptr_.p; a.pblk = pblk_.p;
  a.blk_flags = has_portal_tables_ ? blk_flags_.p : nullptr;
  a.zcnt = has_zones() ? zcnt_.p : nullptr;
  a.zoff = (int)zone_off();
  a.need_id = a.has_frozen || a.has_bonds || pick_enabled_ || strip_.world > 1;
  a.force_out = nullptr;
  a.newcell = newcell_.p; a.cnt = cnt_.p;
  a.cell = cell_[cur_].p;
  a.do_tail = 1;
  a.write_state = 1;
  return a;
}

void Engine::begin_timed(int kind) {
  ++launches_;
  if (!timing_) return;
  TimedLaunch tl;
  tl.kind = kind;
  for (cudaEvent_t* e : {&tl.a, &tl.b}) {
    if (event_pool_.empty()) CUDA_CHECK(cudaEventCreate(e));
    else { *e = event_pool_.back(); event_pool_.pop_back(); }
  }
  CUDA_CHECK(cudaEventRecord(tl.a, stream_));
  timed_.push_back(tl);
}
void Engine::end_timed() {
  if (!timing_) return;
  CUDA_CHECK(cudaEventRecord(timed_.back().b, stream_));
}
void Engine::drain_timed() {
  if (timed_.empty()) return;
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  for (auto& tl : timed_) {
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, tl.a, tl.b));
    kind_ms_[tl.kind] += ms;
    kind_launches_[tl.kind]++;
    event_pool_.push_back(tl.a);
    event_pool_.push_back(tl.b);
  }
  timed_.clear();
}
void Engine::enable_kernel_timing(bool e) {
  drain_timed();
  timing_ = e;
  for (int k = 0; k < 6; ++k) { kind_ms_[k] = 0; kind_launches_[k] = 0; }
}
void Engine::kernel_times(double* ms6, int64_t* launches6) {
  drain_timed();
  for (int k = 0; k < 6; ++k) { ms6[k] = kind_ms_[k]; launches6[k] = kind_launches_[k]; }
}

void Engine::reorder() {
  ++epoch_;
  begin_timed(2);
  launch_scan(cnt_.p, start_.p, n_items_, grid_.ncell, tile_state_.p, ticket_.p, epoch_, ctr_.p, stream_);
  end_timed();
  ReorderArgs r{};
  r.ctr = ctr_.p;
  r.ncell = grid_.ncell;
  r.newcell = newcell_.p;
  r.SJ = grid_.SJ;
  r.cell_out = cell_[cur_ ^ 1].p;
  r.start = start_.p;
  r.cnt = cnt_.p;
  r.tag = order_.p;
  r.fixlist = fixlist_.p;
  r.n_fix = &ctr_.p->n_fix;
  r.pos_in = pos_[cur_].p; r.vel_in = vel_[cur_].p; r.id_in = id_[cur_].p;
  r.pos_out = pos_[cur_ ^ 1].p; r.vel_out = vel_[cur_ ^ 1].p; r.id_out = id_[cur_ ^ 1].p;
  r.rank = strip_.rank;
  r.arr_vkey = strip_.arr_vkey.p;
  const bool keep_map = has_bonds();
  r.slot_of_id = keep_map ? slot_of_id_.p : nullptr;
  if (!keep_map) slot_map_valid_ = false;
  begin_timed(3);
  launch_scatter(r, n_bound_ + arrivals_bound(), stream_);
  end_timed();
  begin_timed(4);
  launch_fixup(r, n_bound_ + arrivals_bound(), stream_);
  end_timed();
  begin_timed(5);
  launch_finalize(ctr_.p, ticket_.p, stream_);
  end_timed();
  cur_ ^= 1;
}

void Engine::rekey_and_reorder(int first, bool teleport) {
  TailArgs a{};
  a.g = grid_;
  a.ph = make_phys();
  a.ctr = ctr_.p;
  a.pos = pos_[cur_].p; a.vel = vel_[cur_].p;
  a.n_sides = has_portal_tables_ ? (int)portals_.size() * 2 : 0;
  a.sides = sides_.p; a.pblk_ptr = pblk_ptr_.p; a.pblk = pblk_.p;
  a.blk_flags = has_portal_tables_ ? blk_flags_.p : nullptr;
  a.do_teleport = teleport && a.n_sides > 0;
  a.dist = make_dist(false);   // (re)binning never routes: particles of other strips are dropped
  a.id = id_[cur_].p;
  a.first = first;
  a.newcell = newcell_.p; a.cnt = cnt_.p;
  begin_timed(5);
  launch_tail(a, n_bound_, stream_);
  end_timed();
  reorder();
}

// One pass of the while loop at particles.cpp:96-201, as phases, so that the exchanges of the
// strip decomposition (dist.cu) can run between them.
void Engine::phase_prepare() {
  CUDA_CHECK(cudaSetDevice(device_));
  sync_env();
  sync_bonds();
  sync_frozen();
  if (has_bonds()) ensure_slot_map();
  t_half_ = (float)((double)t_ + 0.5 * (double)dt_);                      // :126
  // The serial tail's resize decision (:157-175) depends on host state only, so it is
  // taken up front: if the block grid grows, portals are detected after the re-binning.
  rs_nA_ = domA_; rs_nB_ = domB_;
  const float lim = 0.01f;
  auto limit = [&](float& cur, float target) { cur = minf_std(cur + lim, maxf_std(cur - lim, target)); };
  limit(rs_nA_.x, rqA_.x); limit(rs_nA_.y, rqA_.y); limit(rs_nB_.x, rqB_.x); limit(rs_nB_.y, rqB_.y);
  const bool gate = f2i_trunc(t_half_ / dt_) % (int)(0.01 / (double)dt_) == 0;  // :170
  resize_now_ = gate && !(veq(rs_nA_, domA_) && veq(rs_nB_, domB_));
  grid_changes_ = resize_now_ && grown_grid(rs_nA_, rs_nB_, &rs_gA_, &rs_gB_);
  if (grid_changes_ && strip_.world > 1) throw Error(-3, "the block grid cannot grow while strip-decomposed");
}

void Engine::phase_stage1() {
  if (n_bound_ <= 0) return;
  StageArgs a1 = make_stage_args(1);
  begin_timed(0);
  launch_stage(1, a1, n_bound_, stream_);
  end_timed();
}

void Engine::phase_stage2() {
  if (strip_.world > 1) {  // empty the migrant outbox headers (one strided memset)
    CUDA_CHECK(cudaMemset2DAsync(strip_.outbox.p, (size_t)mig_stride(strip_.out_cap), 0, 16, strip_.world, stream_));
  }
  if (n_bound_ <= 0) return;
  StageArgs a2 = make_stage_args(2);
  a2.do_tail = !grid_changes_;
  begin_timed(1);
  launch_stage(2, a2, n_bound_, stream_);
  end_timed();
}

void Engine::phase_reorder() {
  if (resize_now_) {
    domA_ = rs_nA_; domB_ = rs_nB_;                                       // SetDomain(new_domain)
    if (grid_changes_) init_grid(rs_gA_, rs_gB_);
    reset_env_frame(rs_nA_, rs_nB_);                                      // :173
    sync_env();
  }
  if (strip_.world > 1) {  // migrants from the other strips join before the sort
    ArrivalArgs ar{};
    ar.g = grid_;
    ar.ctr = ctr_.p;
    ar.world = strip_.world;
    ar.in_cap = strip_.out_cap;
    ar.in_base = reinterpret_cast<const char*>(strip_.inbox.p);
    ar.in_stride = mig_stride(strip_.out_cap);
    ar.pos = pos_[cur_].p; ar.vel = vel_[cur_].p; ar.id = id_[cur_].p;
    ar.newcell = newcell_.p; ar.cnt = cnt_.p;
    ar.arr_vkey = strip_.arr_vkey.p;
    ar.cap = (int)cap_;
    ar.rank = strip_.rank;
    ar.neighbours_only = migrants_neighbours_only();
    ar.bond_ell = has_bonds() ? bond_ell_.p : nullptr;
    ar.id_cap = (int)id_cap_;
    begin_timed(5);
    launch_arrivals(ar, stream_);
    end_timed();
  }
  if (n_bound_ > 0 || strip_.world > 1) {
    if (grid_changes_) rekey_and_reorder(0, true);                        // :177-179 on the new grid
    else reorder();                                                       // :179
  }
  if (strip_.world > 1) n_bound_ = (int)std::min<size_t>(cap_, (size_t)n_bound_ + arrivals_bound());
}

void Engine::phase_end() {
  // CheckBonds (:180): dead partners are already skipped on the device (prune_bonds()).
  if (remove_last_portal_) {                                              // :182-187
    if (!portals_.empty()) { portals_.pop_back(); env_dirty_ = true; }
    remove_last_portal_ = false;
  }
  if (renderer_ready_) {                                                  // :190-196
    take_snapshot();
    compute_no_rendering();
    renderer_ready_ = false;
  }
  t_ = (float)((double)t_half_ + 0.5 * (double)dt_);                      // :199
}

void Engine::iterate() {
  phase_prepare();
  if (transport_) {
    phase_pack_halo(1); transport_->halo(*this, 1); phase_install_halo(1);
    if (has_zones()) { phase_pack_zones(1); transport_->zones(*this, 1); phase_mark_zones(1); }
  }
  phase_stage1();
  if (transport_) {
    phase_pack_halo(2); transport_->halo(*this, 2); phase_install_halo(2);
    if (has_zones()) { phase_pack_zones(2); transport_->zones(*this, 2); }
  }
  phase_stage2();
  if (transport_) { phase_unmark_halo(); transport_->migrants(*this); }
  phase_reorder();
  phase_end();
}

// Sticky device error bits (include/ptoy_b200.h): the flags word is copied to pinned host memory
// behind every batch of iterations; whoever looks next (without waiting, unless asked to) throws.
void Engine::check_device_errors(bool wait) {
  if (err_pending_) {
    if (wait) CUDA_CHECK(cudaEventSynchronize(err_event_));
    else if (cudaEventQuery(err_event_) != cudaSuccess) return;
    err_pending_ = false;
  }
  const int f = *h_err_;
  if (!f) return;
  const int code = (f & (16 | 64)) ? -5 : -4;
  std::string where;
  if (f & 16) {
    read_counters();
    where = " [first lost bond: particle id " + std::to_string(h_ctr_->lost[0]) + " (slot " + std::to_string(h_ctr_->lost[2]) +
            " of " + std::to_string(h_ctr_->lost[3]) + ") -> partner id " + std::to_string(h_ctr_->lost[1]) + ", rank " +
            std::to_string(strip_.rank) + "]";
  }
  throw Error(code, "device error flags " + std::to_string(f) + where +
                        " (see ptoy_error_flags in include/ptoy_b200.h): the simulation state is no longer valid");
}
void Engine::step_to(float time_target) {
  CUDA_CHECK(cudaSetDevice(device_));
  check_device_errors(false);
  while (t_ < time_target) iterate();                                    // :96 (quit = false)
  CUDA_CHECK(cudaMemcpyAsync(h_err_, &ctr_.p->err_flags, sizeof(int), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaEventRecord(err_event_, stream_));
  err_pending_ = true;
}
void Engine::step_n(int n) {
  CUDA_CHECK(cudaSetDevice(device_));
  check_device_errors(false);
  for (int i = 0; i < n; ++i) iterate();
  CUDA_CHECK(cudaMemcpyAsync(h_err_, &ctr_.p->err_flags, sizeof(int), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaEventRecord(err_event_, stream_));
  err_pending_ = true;
}
void Engine::synchronize() {
  CUDA_CHECK(cudaSetDevice(device_));
  // (a fresh copy of the flags: iterations driven by a strip group do not pass through step_n)
  CUDA_CHECK(cudaMemcpyAsync(h_err_, &ctr_.p->err_flags, sizeof(int), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  CUDA_CHECK(cudaGetLastError());
  err_pending_ = false;
  check_device_errors(true);
}

// ---- read-back ------------------------------------------------------------------------------
void Engine::domain(float* o) const { o[0] = domA_.x; o[1] = domA_.y; o[2] = domB_.x; o[3] = domB_.y; }
void Engine::geometry(float* o, int* dims) const {
  domain(o);
  o[4] = gA_.x; o[5] = gA_.y; o[6] = gB_.x; o[7] = gB_.y;
  dims[0] = di_; dims[1] = dj_;
}

// SetParticleBuffer (particles.cpp:79-89): device state -> host, flattened block-major.  The
// block-major order is produced ON THE DEVICE (counts per reference block from start[], exclusive
// scan, one gather into the idle halves of the ping-pong buffers), so the host only receives the
// finished arrays: no host sort, no FindBlock.
void Engine::take_snapshot() {
  read_counters();
  const size_t n = (size_t)h_ctr_->n_cur;
  const size_t nblk = (size_t)di_ * dj_;
  snap_pos_.resize(n); snap_vel_.resize(n); snap_id_.resize(n); snap_block_.resize(n);
  snap_block_start_.assign(nblk + 1, 0);
  size_t mx = 0;
  if (n) {
    blk_cnt_.alloc(nblk + 2);
    blk_off_.alloc(nblk + 2);
    const size_t tb = scan_exclusive_i32_temp_bytes((int)nblk + 1);
    scan_tmp_.alloc(tb + 16);
    int* d_max = blk_cnt_.p + nblk + 1;
    CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof(int), stream_));
    launch_block_counts(grid_, start_.p, (int)nblk, blk_cnt_.p, d_max, stream_);
    scan_exclusive_i32(scan_tmp_.p, tb, blk_cnt_.p, blk_off_.p, (int)nblk + 1, stream_);
    float2* o_pos = pos_[cur_ ^ 1].p;       // (free between iterations)
    float2* o_vel = vel_[cur_ ^ 1].p;
    int* o_id = id_[cur_ ^ 1].p;
    int* o_blk = order_.p;
    launch_block_gather(grid_, start_.p, blk_off_.p, (int)nblk, pos_[cur_].p, vel_[cur_].p, id_[cur_].p, o_pos, o_vel,
                        o_id, o_blk, stream_);
    launches_ += 3;
    std::vector<int> off(nblk + 1);
    int h_max = 0;
    CUDA_CHECK(cudaMemcpyAsync(snap_pos_.data(), o_pos, n * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaMemcpyAsync(snap_vel_.data(), o_vel, n * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaMemcpyAsync(snap_id_.data(), o_id, n * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaMemcpyAsync(snap_block_.data(), o_blk, n * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaMemcpyAsync(off.data(), blk_off_.p, (nblk + 1) * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    for (size_t b = 0; b <= nblk; ++b) snap_block_start_[b] = (size_t)off[b];   // block b: entries [start[b], start[b+1])
    mx = (size_t)h_max;
  }
  snap_num_particles_ = n;
  snap_num_per_cell_ = mx;
  snap_di_ = di_; snap_dj_ = dj_; snap_gA_ = gA_;
}

size_t Engine::snapshot(float* pos, float* vel, int32_t* id, int64_t* block, size_t cap) {
  const size_t n = std::min(cap, snap_id_.size());
  if (pos) std::memcpy(pos, snap_pos_.data(), n * sizeof(V2));
  if (vel) std::memcpy(vel, snap_vel_.data(), n * sizeof(V2));
  if (id) std::memcpy(id, snap_id_.data(), n * sizeof(int));
  if (block) for (size_t i = 0; i < n; ++i) block[i] = snap_block_[i];
  return snap_id_.size();
}

// The extra RHS_bonds() of the snapshot branch (particles.cpp:192) only matters for
// no_rendering_: bonds that act through a portal (:805-817).  Host-side, on the snapshot.
void Engine::compute_no_rendering() {
  norend_buf_.clear();
  if (portals_.empty() || bonds_.empty()) return;
  const auto& bl = bonds();
  std::vector<int> where(id_cap_ + 1, -1);
  for (size_t k = 0; k < snap_id_.size(); ++k) where[snap_id_[k]] = (int)k;
  const float kCutoff = (float)(8. * (double)kRadius);
  for (auto& bd : bl) {
    const int ia = where[bd.first], ib = where[bd.second];
    if (ia < 0 || ib < 0) continue;
    const V2 p_pos = snap_pos_[ia], q_pos = snap_pos_[ib];
    if (vdist(p_pos, q_pos) < kCutoff) continue;
    for (auto& pr : portals_)
      for (int d = 0; d <= 1; ++d) {
        const PortalHost& portal = d ? pr.second : pr.first;
        const PortalHost& other = d ? pr.first : pr.second;
        const V2 a = portal.begin, r = vsub(portal.end, portal.begin);
        const V2 n = vnormalized(V2{-r.y, r.x});
        const V2 oa = other.begin, orr = vsub(other.end, other.begin);
        const V2 on = vnormalized(V2{-orr.y, orr.x});
        const float p_lambda = vdot(r, vsub(p_pos, a)) / vdot(r, r);
        const float p_offset = vdot(vsub(p_pos, a), n);
        const float q_lambda = vdot(orr, vsub(q_pos, oa)) / vdot(orr, orr);
        const float q_offset = vdot(vsub(q_pos, oa), on);
        if (p_lambda > 0. && p_lambda < 1. && q_lambda > 0. && q_lambda < 1. &&
            p_offset * q_offset < 0. &&
            (double)std::fabs(p_offset - q_offset) < 2. * (double)kPortalThickness + (double)kCutoff) {
          norend_buf_.push_back({bd.first, bd.second});
          norend_buf_.push_back({bd.second, bd.first});
        }
      }
  }
  std::sort(norend_buf_.begin(), norend_buf_.end());
  norend_buf_.erase(std::unique(norend_buf_.begin(), norend_buf_.end()), norend_buf_.end());
}

void Engine::state_by_id(size_t id_cap, float* pos, float* vel, int64_t* block) {
  CUDA_CHECK(cudaSetDevice(device_));
  read_counters();
  const size_t m = std::max(id_cap, id_cap_) + 1;
  DevBuf<float2> dp, dv;
  DevBuf<int> dummy;
  long long* db = nullptr;
  dp.alloc(m); dv.alloc(m);
  CUDA_CHECK(cudaMalloc((void**)&db, m * sizeof(long long)));
  CUDA_CHECK(cudaMemsetAsync(dp.p, 0xff, m * sizeof(float2), stream_));  // NaN payload
  CUDA_CHECK(cudaMemsetAsync(dv.p, 0xff, m * sizeof(float2), stream_));
  CUDA_CHECK(cudaMemsetAsync(db, 0xff, m * sizeof(long long), stream_));  // -1
  launch_state_by_id(grid_, ctr_.p, pos_[cur_].p, vel_[cur_].p, id_[cur_].p, n_bound_, dp.p, dv.p, db, stream_);
  ++launches_;
  if (pos) CUDA_CHECK(cudaMemcpyAsync(pos, dp.p, id_cap * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
  if (vel) CUDA_CHECK(cudaMemcpyAsync(vel, dv.p, id_cap * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
  if (block) CUDA_CHECK(cudaMemcpyAsync(block, db, id_cap * sizeof(long long), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  cudaFree(db);
}

void Engine::eval_forces(int stage, size_t id_cap, float* force) {
  CUDA_CHECK(cudaSetDevice(device_));
  if (stage != 1 && stage != 2) throw Error(-2, "stage must be 1 or 2");
  sync_env(); sync_bonds(); sync_frozen();
  read_counters();
  if (has_bonds()) ensure_slot_map();
  force_dbg_.alloc(std::max<size_t>(cap_, 1));
  const size_t m = std::max(id_cap, id_cap_) + 1;
  DevBuf<float2> by_id;
  by_id.alloc(m);
  CUDA_CHECK(cudaMemsetAsync(by_id.p, 0xff, m * sizeof(float2), stream_));
  if (n_bound_ > 0) {
    StageArgs a1 = make_stage_args(1);
    if (stage == 1) { a1.write_state = 0; a1.force_out = force_dbg_.p; }
    launch_stage(1, a1, n_bound_, stream_);
    ++launches_;
    if (stage == 2) {
      StageArgs a2 = make_stage_args(2);
      a2.write_state = 0; a2.force_out = force_dbg_.p;
      launch_stage(2, a2, n_bound_, stream_);
      ++launches_;
    }
    launch_scatter_by_id(ctr_.p, force_dbg_.p, id_[cur_].p, n_bound_, by_id.p, stream_);
    ++launches_;
  }
  CUDA_CHECK(cudaMemcpyAsync(force, by_id.p, id_cap * sizeof(float2), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void Engine::find_blocks(const float* pos, size_t n, int64_t* block) {
  CUDA_CHECK(cudaSetDevice(device_));
  if (!n) return;
  DevBuf<float2> dp;
  dp.alloc(n);
  long long* db = nullptr;
  CUDA_CHECK(cudaMalloc((void**)&db, n * sizeof(long long)));
  CUDA_CHECK(cudaMemcpyAsync(dp.p, pos, n * sizeof(float2), cudaMemcpyHostToDevice, stream_));
  launch_find_blocks(grid_, dp.p, n, db, stream_);
  ++launches_;
  CUDA_CHECK(cudaMemcpyAsync(block, db, n * sizeof(long long), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  cudaFree(db);
}

// ---- wire format for remote front ends (ptoy_wasm.cpp:130-198) ---------------------------------
// GetParticles: quantised on the device from the live state, culled to max_size entries; only the
// 4 bytes per particle cross PCIe instead of 24.  (Engine order, not block order: the consumer
// draws points, ptoy_inc.js:69-126; the cap of 10 000 entries keeps whichever come first.)
int Engine::wire_particles(uint16_t* data, int max_size, int width, int height) {
  CUDA_CHECK(cudaSetDevice(device_));
  if (max_size < 2 || !data) return 0;
  read_counters();
  const int n = std::min(h_ctr_->n_cur, max_size / 2);
  if (n <= 0) return 0;
  wire_.alloc((size_t)n);
  launch_wire(ctr_.p, pos_[cur_].p, n, width, height, wire_.p, stream_);
  ++launches_;
  CUDA_CHECK(cudaMemcpyAsync(data, wire_.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  return 2 * n;
}

// GetPortals / GetBonds / GetFrozen (:147-198) from the host state and the render snapshot, in the
// order UpdateScene (:47-116) builds the lists.
int Engine::wire_scene(int what, uint16_t* data, int max_size, int width, int height) {
  int i = 0;
  auto append = [&](V2 p) {
    const unsigned q = wire_pixel(p.x, p.y, width, height);
    data[i] = (uint16_t)(q & 0xffffu); data[i + 1] = (uint16_t)(q >> 16);
    i += 2;
  };
  if (what == 1) {          // portals, then the pair being drawn (:61-88)
    std::vector<std::array<V2, 4>> list;
    for (auto& pr : portals_) list.push_back({pr.first.begin, pr.first.end, pr.second.begin, pr.second.end});
    const V2 none{-1.f, -1.f};
    if (portal_stage_ == 0) {
      if (portal_mouse_moving_) list.push_back({portal_begin_, portal_current_, none, none});
    } else {
      list.push_back({portal_prev_first_, portal_prev_second_, none, none});
      if (portal_mouse_moving_) { list.back()[2] = portal_begin_; list.back()[3] = portal_current_; }
    }
    for (auto& q : list) {
      if (i + 8 > max_size) break;
      for (const V2& p : q) append(p);
    }
    return i;
  }
  std::vector<int> where(id_cap_ + 1, -1);
  for (size_t k = 0; k < snap_id_.size(); ++k) where[snap_id_[k]] = (int)k;
  if (what == 2) {          // bonds that are drawn (:90-104)
    for (auto& b : bonds()) {
      if (i + 4 > max_size) break;
      if (std::binary_search(norend_buf_.begin(), norend_buf_.end(), b)) continue;
      const int ia = where[b.first], ib = where[b.second];
      if (ia < 0 || ib < 0) continue;
      append(snap_pos_[ia]); append(snap_pos_[ib]);
    }
    return i;
  }
  if (what == 3) {          // frozen particles (:106-115)
    for (int id : frozen_) {
      if (i + 2 > max_size) break;
      if ((size_t)id > id_cap_ || where[id] < 0) continue;
      append(snap_pos_[where[id]]);
    }
    return i;
  }
  throw Error(-2, "wire_scene: what must be 1 (portals), 2 (bonds) or 3 (frozen)");
}

// ---- the constructor's scene (particles.cpp:34-76) ---------------------------------------------
void Engine::load_default_scene() {
  reset_scene(V2{-1.f, -1.f}, V2{1.f, 1.f});
  gravity_ = vmul(V2{0.f, -1.f}, kGravity);
  gravity_enable_ = true;
  const size_t rows = 25, columns = 25;
  const float width = (float)((double)columns * 2. * (double)kRadius);
  const float height = (float)((double)rows * std::sqrt(3.) * (double)kRadius);
  const V2 boxA{(float)(-0.5 * (double)width), (float)-1.};
  const V2 boxB{(float)(0.5 * (double)width), (float)(-1. + (double)height)};
  std::vector<float> pos, vel(2 * rows * columns, 0.f);
  std::vector<int> ids;
  for (size_t j = 0; j < rows; ++j)
    for (size_t i = 0; i < columns; ++i) {
      pos.push_back((float)((double)boxA.x + (double)kRadius * (2. * (double)i + 1. + (double)(j % 2))));
      pos.push_back((float)((double)boxA.y + (double)kRadius * (std::sqrt(3.) * (double)j + 1.)));
      ids.push_back((int)ids.size());
    }
  add_particles(pos.data(), vel.data(), ids.data(), ids.size(), false);
  take_snapshot();
  const float dx = kPortalThickness;
  const float s[8] = {boxA.x - dx, boxA.y, boxA.x - dx, (float)((double)boxB.y + 0.2),
                      boxB.x + dx, boxA.y, boxB.x + dx, (float)((double)boxB.y + 0.2)};
  add_portal_pair(s);
  renderer_ready_ = true;
}

// ---- UI tools (particles.cpp:228-262, 411-560) on the render snapshot -----------------------
void Engine::portal_tool_state(int* sm, float* xy) const {
  sm[0] = portal_stage_; sm[1] = portal_mouse_moving_;
  xy[0] = portal_begin_.x; xy[1] = portal_begin_.y; xy[2] = portal_current_.x; xy[3] = portal_current_.y;
  xy[4] = portal_prev_first_.x; xy[5] = portal_prev_first_.y; xy[6] = portal_prev_second_.x; xy[7] = portal_prev_second_.y;
}

// The UI tools look particles up by position (particles.cpp:228-248, 452-501, 516-553): the
// reference walks the whole snapshot; here the snapshot's block structure answers the same
// question from the few blocks around the pointer.  Same candidates in the same (block-major)
// order, the same float distance, hence the same answer - including which of several matches wins.
//
// visit(k) for every snapshot entry k of the blocks within `rings` blocks of the one holding
// `point`, in ascending snapshot order.
template <class F>
void Engine::snapshot_near(V2 point, int rings, F&& visit) const {
  if (snap_block_start_.empty()) return;
  auto clampi = [](double v, int lo, int hi) { return (int)std::max<double>(lo, std::min<double>(hi, v)); };
  const int bi = clampi(std::floor(((double)point.x - snap_gA_.x) / bs_.x), 0, snap_di_ - 1);
  const int bj = clampi(std::floor(((double)point.y - snap_gA_.y) / bs_.y), 0, snap_dj_ - 1);
  for (int mi = std::max(0, bi - rings); mi <= std::min(snap_di_ - 1, bi + rings); ++mi)
    for (int mj = std::max(0, bj - rings); mj <= std::min(snap_dj_ - 1, bj + rings); ++mj) {
      const size_t b = (size_t)mi * snap_dj_ + mj;
      for (size_t k = snap_block_start_[b]; k < snap_block_start_[b + 1]; ++k) visit(k);
    }
}

void Engine::tool(int tool, int phase, V2 point) {
  const size_t ns = snap_id_.size();
  switch (tool * 3 + phase) {
    case 0: {  // BondsStart :452-466 (last match in block order wins); kRadius < one block: 3x3 blocks
      bonds_enabled_ = true;
      int id = -1;
      snapshot_near(point, 1, [&](size_t k) { if (vdist(snap_pos_[k], point) < kRadius) id = snap_id_[k]; });
      bonds_prev_id_ = id;
      break;
    }
    case 1: {  // BondsMove :468-501
      if (!bonds_enabled_) return;
      int id = -1;
      snapshot_near(point, 1, [&](size_t k) {
        if (vdist(snap_pos_[k], point) < kRadius && snap_id_[k] != bonds_prev_id_) id = snap_id_[k];
      });
      if (bonds_prev_id_ != -1 && id != -1) toggle_bond(bonds_prev_id_, id);
      if (id != -1) bonds_prev_id_ = id;
      break;
    }
    case 2: if (bonds_enabled_) bonds_enabled_ = false; break;                 // :503-508
    case 3: {  // PickStart :228-248 (first strict minimum over ALL particles)
      bool have = false;
      size_t best = 0;
      float min_dist = 0.f;
      // rings of blocks around the pointer, until nothing outside them can be nearer: a particle
      // of a block more than k blocks away is at least k block sizes from a pointer inside the grid
      const bool inside = point.x >= snap_gA_.x && point.y >= snap_gA_.y &&
                          point.x < snap_gA_.x + snap_di_ * bs_.x && point.y < snap_gA_.y + snap_dj_ * bs_.y;
      const float bsmin = std::min(bs_.x, bs_.y);
      bool done = false;
      if (inside && ns > 64)
        for (int k = 1; k <= std::max(snap_di_, snap_dj_) && !done; k = k < 4 ? k + 1 : 2 * k) {
          have = false;
          snapshot_near(point, k, [&](size_t q) {
            const float d = vdist(snap_pos_[q], point);
            if (!have || d < min_dist) { have = true; best = q; min_dist = d; }
          });
          done = have && min_dist < (float)k * bsmin - 1e-4f;
        }
      if (!done) {   // (pointer outside the grid, a handful of particles, or nothing within reach)
        have = false;
        for (size_t k = 0; k < ns; ++k)
          if (!have || vdist(snap_pos_[k], point) < min_dist) {
            have = true; best = k; min_dist = vdist(snap_pos_[k], point);
          }
      }
      if (!have) return;
      pick_id_ = snap_id_[best];
      pick_pointer_ = point;
      pick_enabled_ = true;
      break;
    }
    case 4: if (pick_enabled_) pick_pointer_ = point; break;                   // :250-255
    case 5: if (pick_enabled_) pick_enabled_ = false; break;                   // :257-262
    case 6: freeze_last_id_ = -1; freeze_enabled_ = true;                      // :510-514
      // fallthrough
    case 7: {  // FreezeMove :516-553
      if (!freeze_enabled_) return;
      const float kFreezeRadius = kRadius * 3;
      float min_dist = kFreezeRadius;
      size_t best = 0;
      snapshot_near(point, 1, [&](size_t k) {   // 3 kRadius < one block: the 3x3 blocks around the pointer
        const float dist = vdist(snap_pos_[k], point);
        if (dist < min_dist) { min_dist = dist; best = k; }
      });
      if (min_dist < kFreezeRadius) {
        const int id = snap_id_[best];
        if (id != freeze_last_id_) {
          auto it = std::lower_bound(frozen_.begin(), frozen_.end(), id);
          if (it != frozen_.end() && *it == id) frozen_.erase(it); else frozen_.insert(it, id);
          frozen_dirty_ = true;
          freeze_last_id_ = id;
        }
      }
      break;
    }
    case 8: if (freeze_enabled_) freeze_enabled_ = false; break;               // :555-560
    case 9:  // PortalStart :411-416
      portal_enabled_ = true; portal_begin_ = point; portal_current_ = point; portal_mouse_moving_ = true;
      break;
    case 10: if (portal_enabled_) portal_current_ = point; break;              // :418-423
    case 11: {  // PortalStop :425-450
      if (!portal_enabled_) return;
      portal_enabled_ = false;
      portal_mouse_moving_ = false;
      if (portal_stage_ == 0) {
        portal_prev_first_ = portal_begin_;
        portal_prev_second_ = point;
        portal_stage_ = 1;
      } else {
        PortalHost p0, p1;
        p0.begin = portal_prev_first_; p0.end = portal_prev_second_;
        p1.begin = portal_begin_;
        p1.end = vadd(p1.begin, vmul(vnormalized(vsub(point, p1.begin)), vdist(p0.begin, p0.end)));
        update_portal_blocks(p0);
        update_portal_blocks(p1);
        portals_.push_back({p0, p1});
        portal_stage_ = 0;
        env_dirty_ = true;
      }
      break;
    }
    default: throw Error(-2, "unknown tool / phase");
  }
}

}  // namespace ptoy

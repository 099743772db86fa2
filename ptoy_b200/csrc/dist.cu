// Strip decomposition across GPUs (SURVEY.md §8e — new work, the reference is
// single-process).
//
// The 2-D domain is cut into strips over the SLOW block index (x): a strip is a contiguous range
// of block columns = of sub-rows = of sorted particles.  Each rank steps its own strip with the
// same kernels; what crosses a strip boundary per iteration is
//   * before stage 1: the neighbour's two boundary sub-rows at x0 (positions + the matching
//     slice of its start[]), before stage 2 the same rows at x_half  -> "ghost slices", plain
//     contiguous copies because a strip's boundary rows are contiguous in its sorted arrays;
//   * after stage 2: particles whose new sub-row belongs to another strip (any strip after a
//     portal teleport) -> fixed-size per-destination outbox blocks, appended behind the
//     receiver's particles and merged by its reorder.  Migrants carry (source rank, old slot), so
//     the receiver's stable sort reproduces the single-GPU order: N strips == 1 GPU bit for bit.
// No host synchronisation inside an iteration: all message sizes are fixed capacities.
//
// Transports: LocalGroup (several strips in one process on one device: the test vehicle, also
// how the logic was developed without N GPUs) and NcclTransport (one process per GPU).
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound at run time (see NcclApi)

#include <algorithm>
#include <cstring>

#include "engine.h"

namespace ptoy {

// NCCL is bound with dlopen/dlsym instead of a link-time dependency: a process that also runs
// PyTorch must end up with ONE libnccl.so.2 (PyTorch's bundled one, which is newer than the
// system copy and which libtorch_cuda requires); dlopen by soname returns whichever is already
// loaded, and single-GPU users never load NCCL at all.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  static NcclApi& get() {
    static NcclApi api = load();
    return api;
  }
  static NcclApi load() {
    NcclApi a;
    void* h = nullptr;
    if (const char* p = std::getenv("PTOY_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw Error(-5, std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* n) {
      void* s = dlsym(h, n);
      if (!s) throw Error(-5, std::string("libnccl.so.2 lacks ") + n);
      return s;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    return a;
  }
};

#define CUDA_CHECK(expr)                                                                  \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess)                                                                \
      throw Error(-1, std::string(#expr) + ": " + cudaGetErrorString(e_));                \
  } while (0)
#define NCCL_CHECK(expr)                                                                  \
  do {                                                                                    \
    ncclResult_t r_ = (expr);                                                             \
    if (r_ != ncclSuccess)                                                                \
      throw Error(-5, std::string(#expr) + ": " + NcclApi::get().GetErrorString(r_));                \
  } while (0)

void Engine::configure_strips(int rank, int world, const int* col_begin, int ghost_rows, size_t id_capacity, bool bonds) {
  CUDA_CHECK(cudaSetDevice(device_));
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) throw Error(-2, "bad rank / world");
  if (ghost_rows <= 0) ghost_rows = 2;
  if (ghost_rows < 2 || ghost_rows > 64) throw Error(-2, "ghost_rows must be in [2, 64]");
  read_counters();
  if (h_ctr_->n_cur != 0) throw Error(-3, "configure_strips must precede add_particles");
  strip_.world = world;
  strip_.rank = rank;
  strip_.ghost_rows = world > 1 ? ghost_rows : 2;
  strip_.bonds = world > 1 && bonds;
  for (int r = 0; r <= world; ++r) strip_.col_begin[r] = col_begin ? col_begin[r] : 0;
  if (world > 1) {
    if (!col_begin || col_begin[0] != 0 || col_begin[world] != di_) throw Error(-2, "col_begin must run from 0 to dims.i");
    for (int r = 0; r < world; ++r)
      if (col_begin[r + 1] - col_begin[r] < 2) throw Error(-2, "a strip needs at least 2 block columns");
  }
  alloc_cells();
  // every id-indexed table must cover the ids of ALL strips: migrants and ghosts bring ids this
  // rank never loaded itself
  if (id_capacity) ensure_id_capacity(id_capacity);
  if (strip_.bonds) bonds_dirty_ = true;
  env_dirty_ = true;
}

void Engine::alloc_strip_buffers() {
  StripState& s = strip_;
  // the boundary sub-rows hold about ghost_rows*SJ*1.2 particles at dense packing; allow 8 per sub-cell
  s.ghost_cap = std::max(4096, s.ghost_rows * grid_.SJ * 8);
  s.out_cap = 1 << 14;
  if (const char* v = std::getenv("PTOY_MIGRANT_CAP")) s.out_cap = std::max(16, atoi(v));
  s.nstart = (size_t)s.ghost_rows * grid_.SJ + 1;
  auto up16 = [](size_t b) { return (b + 15) / 16 * 16; };
  s.halo_pos_bytes = up16((size_t)(s.ghost_cap + 2) * sizeof(float2));
  const size_t start_bytes = up16(s.nstart * sizeof(int));
  s.halo_bytes = s.halo_pos_bytes + start_bytes + up16((size_t)s.ghost_cap * sizeof(int));
  for (int sd = 0; sd < 2; ++sd) {
    s.send_raw[sd].alloc(s.halo_bytes);
    s.recv_raw[sd].alloc(s.halo_bytes);
    CUDA_CHECK(cudaMemsetAsync(s.send_raw[sd].p, 0, s.halo_bytes, stream_));
    CUDA_CHECK(cudaMemsetAsync(s.recv_raw[sd].p, 0, s.halo_bytes, stream_));
    for (auto pr : {std::make_pair(&s.send[sd], s.send_raw[sd].p), std::make_pair(&s.ghost[sd], s.recv_raw[sd].p)}) {
      pr.first->pos = reinterpret_cast<float2*>(pr.second);
      pr.first->start = reinterpret_cast<int*>(pr.second + s.halo_pos_bytes);
      pr.first->id = reinterpret_cast<int*>(pr.second + s.halo_pos_bytes + start_bytes);
    }
  }
  s.ghost_n.alloc(2);
  CUDA_CHECK(cudaMemsetAsync(s.ghost_n.p, 0, 2 * sizeof(int), stream_));
  const size_t bytes = (size_t)mig_stride(s.out_cap) * s.world;
  s.outbox.alloc(bytes);
  s.inbox.alloc(bytes);
  CUDA_CHECK(cudaMemsetAsync(s.outbox.p, 0, bytes, stream_));
  CUDA_CHECK(cudaMemsetAsync(s.inbox.p, 0, bytes, stream_));
  s.arr_vkey.alloc((size_t)s.world * s.out_cap);
}

DistDev Engine::make_dist(bool with_outbox) const {
  DistDev d{};
  d.world = strip_.world;
  d.rank = strip_.rank;
  for (int r = 0; r <= strip_.world; ++r) d.row_begin[r] = strip_.row_begin[r];
  d.out_cap = strip_.out_cap;
  d.out_stride = mig_stride(strip_.out_cap);
  d.out_base = (with_outbox && strip_.world > 1) ? reinterpret_cast<char*>(strip_.outbox.p) : nullptr;
  d.bond_ell = (strip_.world > 1 && has_bonds()) ? bond_ell_.p : nullptr;
  d.err = &ctr_.p->err_flags;
  return d;
}

void Engine::set_stream(cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (owns_stream_) cudaStreamDestroy(stream_);
  stream_ = s;
  owns_stream_ = false;
}

// Boundary rows -> staging buffers.  stage 1: positions at x0 + start[] slice (+ ids); 2: x_half
// + this strip's largest half-step drift.
void Engine::phase_pack_halo(int stage) {
  if (strip_.world == 1) return;
  GhostPackArgs g{};
  g.ctr = ctr_.p;
  g.X = stage == 1 ? pos_[cur_].p : pos_h_.p;
  g.start = start_.p;
  g.SJ = grid_.SJ;
  g.nrows = strip_.ghost_rows;
  g.row_lo[0] = grid_.own_lo;
  g.row_lo[1] = grid_.own_hi - strip_.ghost_rows;
  g.ghost_cap = strip_.ghost_cap;
  for (int sd = 0; sd < 2; ++sd) {
    g.out_pos[sd] = strip_.send[sd].pos;
    g.out_start[sd] = stage == 1 ? strip_.send[sd].start : nullptr;
    g.out_id[sd] = (stage == 1 && has_bonds()) ? strip_.send[sd].id : nullptr;
  }
  g.id = id_[cur_].p;
  g.dmax = stage == 2 ? &ctr_.p->dmax : nullptr;
  g.err = &ctr_.p->err_flags;
  if (cap_ == 0) return;
  begin_timed(5);
  launch_ghost_pack(g, stream_);
  end_timed();
}

// The received boundary rows become ghost rows: positions right before slot 0 / right behind the
// last owned particle, start[] of the ghost rows to match, and (bonds) the ghost ids in the
// id -> slot map, so that a bond partner in a ghost row is found like a local one.
void Engine::phase_install_halo(int stage) {
  if (strip_.world == 1 || cap_ == 0) return;
  StripState& s = strip_;
  GhostInstallArgs g{};
  g.ctr = ctr_.p;
  g.stage = stage;
  g.SJ = grid_.SJ;
  g.nrows = s.ghost_rows;
  g.own_hi = grid_.own_hi;
  g.ghost_cap = s.ghost_cap;
  for (int sd = 0; sd < 2; ++sd) {
    const bool have = sd == 0 ? s.rank > 0 : s.rank + 1 < s.world;
    g.in_pos[sd] = have ? s.ghost[sd].pos : nullptr;
    g.in_start[sd] = s.ghost[sd].start;
    g.in_id[sd] = (have && has_bonds()) ? s.ghost[sd].id : nullptr;
    g.count[sd] = s.ghost_n.p + sd;
  }
  g.X = stage == 1 ? pos_[cur_].p : pos_h_.p;
  g.start = start_.p;
  g.slot_of_id = slot_of_id_.p;
  g.id_cap = (int)id_cap_;
  g.err = &ctr_.p->err_flags;
  begin_timed(5);
  launch_ghost_install(g, stream_);
  end_timed();
}

// ---- portal zones: cross-portal forces and through-portal bonds across strips ------------------
static PortalZoneArgs zone_args(Engine& e, const GridDev& g, Counters* ctr, float2* X, const int* id, const int* start,
                                const int* pblk, int n_entries, size_t zoff, int* zcnt, int* zid, int* slot_of_id,
                                size_t id_cap) {
  PortalZoneArgs z{};
  z.g = g;
  z.ctr = ctr;
  z.X_in = X; z.X = X;
  z.id = id; z.start = start; z.pblk = pblk;
  z.n_entries = n_entries;
  z.zoff = (int)zoff;
  z.zcnt = zcnt; z.zid = zid;
  z.slot_of_id = slot_of_id;
  z.id_cap = (int)id_cap;
  z.err = &ctr->err_flags;
  (void)e;
  return z;
}

void Engine::zone_buffers(int stage, int** pos_i32, size_t* n_pos, int** cnt, size_t* n_cnt, int** ids, size_t* n_ids) {
  float2* X = stage == 1 ? pos_[cur_].p : pos_h_.p;
  *pos_i32 = reinterpret_cast<int*>(X + zone_off());
  *n_pos = (size_t)zone_entries_ * kZoneCap * 2;
  *cnt = zcnt_.p; *n_cnt = (size_t)zone_entries_;
  *ids = zid_.p; *n_ids = (size_t)zone_entries_ * kZoneCap;
}

void Engine::phase_pack_zones(int stage) {
  if (!has_zones()) return;
  int *zp, *zc, *zi;
  size_t np, nc, ni;
  zone_buffers(stage, &zp, &np, &zc, &nc, &zi, &ni);
  CUDA_CHECK(cudaMemsetAsync(zp, 0, np * sizeof(int), stream_));
  CUDA_CHECK(cudaMemsetAsync(zc, 0, nc * sizeof(int), stream_));
  CUDA_CHECK(cudaMemsetAsync(zi, 0, ni * sizeof(int), stream_));
  PortalZoneArgs z = zone_args(*this, grid_, ctr_.p, stage == 1 ? pos_[cur_].p : pos_h_.p, id_[cur_].p, start_.p, pblk_.p,
                               zone_entries_, zone_off(), zcnt_.p, zid_.p, slot_of_id_.p, id_cap_);
  begin_timed(5);
  launch_zone_pack(z, stream_);
  end_timed();
}

void Engine::phase_mark_zones(int stage) {
  if (!has_zones() || !has_bonds() || stage != 1) return;
  PortalZoneArgs z = zone_args(*this, grid_, ctr_.p, pos_[cur_].p, id_[cur_].p, start_.p, pblk_.p, zone_entries_,
                               zone_off(), zcnt_.p, zid_.p, slot_of_id_.p, id_cap_);
  begin_timed(5);
  launch_zone_mark(z, 1, stream_);
  end_timed();
}

void Engine::phase_unmark_halo() {
  if (has_zones() && has_bonds() && cap_ != 0) {
    PortalZoneArgs z = zone_args(*this, grid_, ctr_.p, pos_[cur_].p, id_[cur_].p, start_.p, pblk_.p, zone_entries_,
                                 zone_off(), zcnt_.p, zid_.p, slot_of_id_.p, id_cap_);
    begin_timed(5);
    launch_zone_mark(z, 0, stream_);
    end_timed();
  }
  if (strip_.world == 1 || !has_bonds() || cap_ == 0) return;
  StripState& s = strip_;
  GhostUnmarkArgs g{};
  g.ctr = ctr_.p;
  g.slot_of_id = slot_of_id_.p;
  g.id_cap = (int)id_cap_;
  g.ghost_cap = s.ghost_cap;
  for (int sd = 0; sd < 2; ++sd) {
    const bool have = sd == 0 ? s.rank > 0 : s.rank + 1 < s.world;
    g.in_id[sd] = have ? s.ghost[sd].id : nullptr;
    g.count[sd] = s.ghost_n.p + sd;
  }
  begin_timed(5);
  launch_ghost_unmark(g, stream_);
  end_timed();
}

// ---------------------------------------------------------------------------------------------
// In-process group of strips on one device (shared stream => phases are naturally ordered).
// ---------------------------------------------------------------------------------------------
struct LocalGroup : Transport {
  std::vector<Engine*> e;
  cudaStream_t stream = nullptr;

  explicit LocalGroup(const std::vector<Engine*>& engines) : e(engines) {
    if (e.empty()) throw Error(-2, "empty group");
    CUDA_CHECK(cudaSetDevice(e[0]->device()));
    CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (size_t r = 0; r < e.size(); ++r) {
      if (e[r]->strip().world != (int)e.size() || e[r]->strip().rank != (int)r)
        throw Error(-2, "group engines must be configured as ranks 0..n-1 of an n-strip decomposition");
      if (e[r]->device() != e[0]->device()) throw Error(-2, "a local group lives on one device");
      e[r]->set_stream(stream);
      e[r]->set_transport(nullptr);  // the group drives the phases itself
      e[r]->set_local_group(true);   // (its migrant exchange is always all-to-all)
    }
  }
  ~LocalGroup() override {
    if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
  }
  void copy(void* dst, const void* src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream));
  }
  void halo_all(int stage) {
    const int n = (int)e.size();
    for (int r = 0; r < n; ++r) {
      StripState& me = e[r]->strip_mut();
      const size_t bytes = stage == 1 ? me.halo_bytes : me.halo_pos_bytes;
      if (r > 0) copy(e[r - 1]->strip_mut().recv_raw[1].p, me.send_raw[0].p, bytes);   // my first rows -> left neighbour's right ghost
      if (r + 1 < n) copy(e[r + 1]->strip_mut().recv_raw[0].p, me.send_raw[1].p, bytes);   // my last rows -> right neighbour's left ghost
    }
  }
  void migrants_all() {
    const int n = (int)e.size();
    for (int s = 0; s < n; ++s)
      for (int d = 0; d < n; ++d) {
        if (s == d) continue;
        const size_t st = (size_t)mig_stride(e[s]->strip().out_cap);
        copy(e[d]->strip_mut().inbox.p + s * st, e[s]->strip_mut().outbox.p + d * st, st);
      }
  }
  void halo(Engine&, int) override {}
  void migrants(Engine&) override {}
  void zones(Engine&, int) override {}
  void zones_all(int stage) {   // sum of the tables of all strips (each entry has one owner), given to everybody
    if (!e[0]->has_zones()) return;
    int *p0, *c0, *i0;
    size_t np, nc, ni;
    e[0]->zone_buffers(stage, &p0, &np, &c0, &nc, &i0, &ni);
    for (size_t r = 1; r < e.size(); ++r) {
      int *p, *c, *i;
      e[r]->zone_buffers(stage, &p, &np, &c, &nc, &i, &ni);
      launch_sum_i32(p0, p, np, stream);
      launch_sum_i32(c0, c, nc, stream);
      launch_sum_i32(i0, i, ni, stream);
    }
    for (size_t r = 1; r < e.size(); ++r) {
      int *p, *c, *i;
      e[r]->zone_buffers(stage, &p, &np, &c, &nc, &i, &ni);
      copy(p, p0, np * sizeof(int));
      copy(c, c0, nc * sizeof(int));
      copy(i, i0, ni * sizeof(int));
    }
  }

  void step_n(int iters) {
    for (int it = 0; it < iters; ++it) {
      for (auto* x : e) x->phase_prepare();
      for (auto* x : e) x->phase_pack_halo(1);
      halo_all(1);
      for (auto* x : e) x->phase_install_halo(1);
      for (auto* x : e) x->phase_pack_zones(1);
      zones_all(1);
      for (auto* x : e) x->phase_mark_zones(1);
      for (auto* x : e) x->phase_stage1();
      for (auto* x : e) x->phase_pack_halo(2);
      halo_all(2);
      for (auto* x : e) x->phase_install_halo(2);
      for (auto* x : e) x->phase_pack_zones(2);
      zones_all(2);
      for (auto* x : e) x->phase_stage2();
      for (auto* x : e) x->phase_unmark_halo();
      migrants_all();
      for (auto* x : e) x->phase_reorder();
      for (auto* x : e) x->phase_end();
    }
  }
};

// ---------------------------------------------------------------------------------------------
// One process per GPU: NCCL point-to-point over NVLink / NVSwitch on the engine's stream.
// ---------------------------------------------------------------------------------------------
struct NcclTransport : Transport {
  ncclComm_t comm = nullptr;
  explicit NcclTransport(Engine& e, const ncclUniqueId& id) {
    CUDA_CHECK(cudaSetDevice(e.device()));
    NCCL_CHECK(NcclApi::get().CommInitRank(&comm, e.strip().world, id, e.strip().rank));
    // Establish the neighbour connections now, with one tiny symmetric exchange, instead of
    // lazily inside the first iteration: connection setup is a rendezvous between the two ranks,
    // and doing it here keeps it out of the stepping pipeline (and makes a transport problem show
    // up at connect time, where the caller can time out, rather than as a stalled step).
    StripState& s = e.strip_mut();
    DevBuf<int> probe;
    probe.alloc(2 * (size_t)s.world);
    CUDA_CHECK(cudaMemsetAsync(probe.p, 0, 2 * s.world * sizeof(int), e.stream()));
    NCCL_CHECK(NcclApi::get().GroupStart());
    for (int r = 0; r < s.world; ++r) {
      if (r != s.rank - 1 && r != s.rank + 1) continue;  // the neighbours: all a portal-free scene ever talks to
      NCCL_CHECK(NcclApi::get().Send(probe.p + r, sizeof(int), ncclChar, r, comm, e.stream()));
      NCCL_CHECK(NcclApi::get().Recv(probe.p + s.world + r, sizeof(int), ncclChar, r, comm, e.stream()));
    }
    NCCL_CHECK(NcclApi::get().GroupEnd());
    CUDA_CHECK(cudaStreamSynchronize(e.stream()));
  }
  ~NcclTransport() override {
    if (comm) NcclApi::get().CommDestroy(comm);
  }
  void halo(Engine& e, int stage) override {
    StripState& s = e.strip_mut();
    // ONE message per neighbour: positions | start[] slice | ids (stage 2: the positions only)
    const size_t bytes = stage == 1 ? s.halo_bytes : s.halo_pos_bytes;
    NCCL_CHECK(NcclApi::get().GroupStart());
    for (int sd = 0; sd < 2; ++sd) {
      const int peer = sd == 0 ? s.rank - 1 : s.rank + 1;
      if (peer < 0 || peer >= s.world) continue;
      // my boundary rows on side sd become the peer's ghost on its opposite side; symmetric recv
      NCCL_CHECK(NcclApi::get().Send(s.send_raw[sd].p, bytes, ncclChar, peer, comm, e.stream()));
      NCCL_CHECK(NcclApi::get().Recv(s.recv_raw[sd].p, bytes, ncclChar, peer, comm, e.stream()));
    }
    NCCL_CHECK(NcclApi::get().GroupEnd());
  }
  void zones(Engine& e, int stage) override {
    // every entry of the zone table is written by exactly one strip (the owner of the block) and
    // zero elsewhere, so the integer sum over the strips is that strip's value, bit for bit
    int *p, *c, *i;
    size_t np, nc, ni;
    e.zone_buffers(stage, &p, &np, &c, &nc, &i, &ni);
    NCCL_CHECK(NcclApi::get().GroupStart());
    NCCL_CHECK(NcclApi::get().AllReduce(p, p, np, ncclInt32, ncclSum, comm, e.stream()));
    NCCL_CHECK(NcclApi::get().AllReduce(c, c, nc, ncclInt32, ncclSum, comm, e.stream()));
    NCCL_CHECK(NcclApi::get().AllReduce(i, i, ni, ncclInt32, ncclSum, comm, e.stream()));
    NCCL_CHECK(NcclApi::get().GroupEnd());
  }
  void migrants(Engine& e) override {
    StripState& s = e.strip_mut();
    const size_t st = (size_t)mig_stride(s.out_cap);
    const bool all_to_all = !e.migrants_neighbours_only();   // (see engine.h)
    NCCL_CHECK(NcclApi::get().GroupStart());
    for (int r = 0; r < s.world; ++r) {
      if (r == s.rank) continue;
      if (!all_to_all && r != s.rank - 1 && r != s.rank + 1) continue;
      NCCL_CHECK(NcclApi::get().Send(s.outbox.p + r * st, st, ncclChar, r, comm, e.stream()));
      NCCL_CHECK(NcclApi::get().Recv(s.inbox.p + r * st, st, ncclChar, r, comm, e.stream()));
    }
    NCCL_CHECK(NcclApi::get().GroupEnd());
  }
};

void Engine::delete_transport() {
  delete transport_;
  transport_ = nullptr;
}
void Engine::connect_nccl(const void* uid128) {
  if (strip_.world < 2) throw Error(-3, "configure_strips first");
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(&id, uid128, sizeof(id));
  set_transport(new NcclTransport(*this, id));
}
void nccl_unique_id(void* out128) {
  ncclUniqueId id;
  NCCL_CHECK(NcclApi::get().GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
}
LocalGroup* make_local_group(const std::vector<Engine*>& engines) { return new LocalGroup(engines); }
void local_group_step(LocalGroup* g, int n) { g->step_n(n); }
void local_group_destroy(LocalGroup* g) { delete g; }

}  // namespace ptoy

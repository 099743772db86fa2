// Drop-in replacement for the reference's src/particles.cpp.
//
// Compile THIS file instead of src/particles.cpp, against the reference's UNCHANGED
// src/particles.h / blocks.h / geometry.h, and link libptoy_b200.so.  Every
// member of `class Particles` (particles.h:60-228) is implemented by forwarding to the C ABI
// in include/ptoy_b200.h; main.cpp, view_gl.cpp, control.cpp, game.h and ptoy_wasm.cpp compile
// and behave as before.
//
// What stays on the host, in the reference's own containers, is exactly what its getters hand
// out by reference: the render snapshot (`blocks_buffer_`, `particle_buffer_`), `bonds_`,
// `frozen_`, `portals_`, `no_rendering_buffer_` and the public portal-drawing fields.  They are
// refreshed from the engine when the reference would refresh them: the snapshot on the first
// sub-step of a step() call that follows SetRendererReadyForNext(true) (particles.cpp:190-196),
// the sets after the tool handlers that mutate them.
//
// There is no CPU physics in this file and no fallback: if the engine cannot be created the
// constructor throws.

#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

// The helper functions of this file are implementation details of `Particles` that the
// unchanged header cannot declare as members; they reach its private section this way (the
// class layout is unaffected; the standard headers are already included above).
#define private public
#include "particles.h"
#undef private

#include "ptoy_b200.h"

const Scal kRadius = 0.02;           // particles.cpp:7  (declared extern in particles.h:12)
const Scal kPortalThickness = 0.02;  // particles.cpp:20 (declared extern in particles.h:13)
static const Scal kBlockSize = 4. * kRadius;  // particles.cpp:18

namespace {

std::mutex g_mu;
std::unordered_map<const Particles*, ptoy_engine*> g_engines;

ptoy_engine* engine_of(const Particles* p) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_engines.find(p);
  if (it == g_engines.end()) throw std::runtime_error("ptoy_b200: Particles object has no engine");
  return it->second;
}

void check(int rc, const char* what) {
  if (rc != PTOY_OK)
    throw std::runtime_error(std::string("ptoy_b200: ") + what + ": " + ptoy_last_error());
}

struct CoutMute {  // blocks::SetDomain reports growth on stdout (blocks.h:138); the engine already did
  std::streambuf* old;
  std::ostringstream sink;
  CoutMute() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~CoutMute() { std::cout.rdbuf(old); }
};

}  // namespace

// ---- construction ---------------------------------------------------------------------------

Particles::Particles()
    : domain(RectVect(Vect(-1, -1), Vect(1, 1)))
    , Blocks(domain, Vect(kBlockSize, kBlockSize))
    , blocks_buffer_(Blocks) {
  force_enabled = false;
  force_center = Vect(0, 0);
  remove_last_portal_ = false;
  ptoy_engine* e = nullptr;
  check(ptoy_create(&e, 0), "ptoy_create");
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_engines[this] = e;
  }
  check(ptoy_load_default_scene(e), "ptoy_load_default_scene");  // particles.cpp:34-76
  check(ptoy_time(e, &t), "ptoy_time");
  check(ptoy_dt(e, &dt), "ptoy_dt");
  float g[2];
  check(ptoy_get_gravity_vect(e, g), "gravity");
  gravity_ = Vect(g[0], g[1]);
  resize_queue_ = domain;
  SetParticleBuffer();
  // portals_, public portal fields
  PortalMove(Vect(0, 0));  // no-op on the engine (no portal being drawn); refreshes the mirrors
}

Particles::~Particles() {
  ptoy_engine* e = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_engines.find(this);
    if (it != g_engines.end()) {
      e = it->second;
      g_engines.erase(it);
    }
  }
  if (e) ptoy_destroy(e);
}

// ---- mirrors ----------------------------------------------------------------------------------

namespace {

void pull_portals(Particles* self, ptoy_engine* e) {
  size_t n = 0;
  check(ptoy_num_portal_pairs(e, &n), "portals");
  std::vector<float> seg(8 * n + 8);
  if (n) check(ptoy_get_portals(e, seg.data()), "portals");
  self->portals_.resize(n);
  for (size_t p = 0; p < n; ++p)
    for (int d = 0; d < 2; ++d) {
      auto& po = self->portals_[p][d];
      const float* s = &seg[(2 * p + d) * 4];
      po.begin = Vect(s[0], s[1]);
      po.end = Vect(s[2], s[3]);
      size_t nb = 0;
      check(ptoy_get_portal_blocks(e, p, d, nullptr, 0, &nb), "portal blocks");
      std::vector<int64_t> b(nb + 1);
      check(ptoy_get_portal_blocks(e, p, d, b.data(), nb, &nb), "portal blocks");
      po.blocks.assign(b.begin(), b.begin() + nb);
    }
  int sm[2];
  float xy[8];
  check(ptoy_get_portal_tool_state(e, sm, xy), "portal tool state");
  self->portal_stage_ = sm[0];
  self->portal_mouse_moving_ = sm[1] != 0;
  self->portal_begin_ = Vect(xy[0], xy[1]);
  self->portal_current_ = Vect(xy[2], xy[3]);
  self->portal_prev_ = {Vect(xy[4], xy[5]), Vect(xy[6], xy[7])};
}

void pull_pairs(int (*fn)(ptoy_engine*, int32_t*, size_t, size_t*), ptoy_engine* e,
                std::set<std::pair<int, int>>& out) {
  size_t n = 0;
  check(fn(e, nullptr, 0, &n), "pairs");
  std::vector<int32_t> buf(2 * n + 2);
  check(fn(e, buf.data(), n, &n), "pairs");
  out.clear();
  for (size_t i = 0; i < n; ++i) out.emplace_hint(out.end(), buf[2 * i], buf[2 * i + 1]);
}

}  // namespace

// SetGravity / SetGravityVect / SetForceAttractive / PushResize / RemoveLastPortal / SetDomain are
// inline in particles.h and only touch host members; the engine picks those members up right
// before it runs, so the header stays untouched.
static void sync_inline_state(Particles* self, ptoy_engine* e) {
  check(ptoy_set_gravity(e, self->gravity_enable_), "gravity");
  check(ptoy_set_gravity_vect(e, self->gravity_.x, self->gravity_.y), "gravity");
  check(ptoy_set_point_force(e, self->force_enabled, self->force_attractive_, self->force_center.x,
                             self->force_center.y), "point force");
  check(ptoy_push_resize(e, self->resize_queue_.A.x, self->resize_queue_.A.y, self->resize_queue_.B.x,
                         self->resize_queue_.B.y), "resize");
  if (self->remove_last_portal_) check(ptoy_remove_last_portal(e), "remove_last_portal");
  float dom[4];
  check(ptoy_get_domain(e, dom), "domain");
  if (!(self->domain == RectVect(Vect(dom[0], dom[1]), Vect(dom[2], dom[3]))))  // SetDomain() was called
    check(ptoy_set_domain(e, self->domain.A.x, self->domain.A.y, self->domain.B.x, self->domain.B.y), "set_domain");
}

// Engine render snapshot -> Blocks / blocks_buffer_ / particle_buffer_ (particles.cpp:79-89).
// The engine hands the particles over block-major, with the block of every particle, so the
// blocks are filled directly - no FindBlock, no SortParticles on the host.
static void rebuild_snapshot(Particles* self, ptoy_engine* e) {
  size_t n = 0;
  check(ptoy_get_snapshot(e, nullptr, nullptr, nullptr, nullptr, 0, &n), "snapshot");
  ArrayVect p(n), v(n);
  std::vector<int> id(n);
  std::vector<int64_t> blk(n);
  if (n) check(ptoy_get_snapshot(e, &p[0].x, &v[0].x, id.data(), blk.data(), n, &n), "snapshot");
  {
    CoutMute m;
    // keep the host grid congruent with the engine's (same grow-only loop, blocks.h:119-141)
    float dom[4];
    check(ptoy_get_domain(e, dom), "domain");
    self->domain = RectVect(Vect(dom[0], dom[1]), Vect(dom[2], dom[3]));
    blocks& B = self->Blocks;
    auto& d = B.GetData();
    d.clear();
    d.resize(B.GetNumBlocks());
    B.SetDomain(self->domain);
    const size_t nblk = B.GetNumBlocks();
    size_t k = 0, most = 0;
    while (k < n) {                       // one run of the block-major arrays per block
      const size_t b = (size_t)blk[k];
      size_t k1 = k;
      while (k1 < n && (size_t)blk[k1] == b) ++k1;
      if (b < nblk) {
        const size_t cnt = k1 - k;
        d.position[b].assign(p.begin() + k, p.begin() + k1);
        d.velocity[b].assign(v.begin() + k, v.begin() + k1);
        d.id[b].assign(id.begin() + k, id.begin() + k1);
        d.position_tmp[b].assign(cnt, GetNan<Vect>());   // as BlockData::AddParticle leaves them (blocks.h:102-108)
        d.velocity_tmp[b].assign(cnt, GetNan<Vect>());
        d.force[b].assign(cnt, GetNan<Vect>());
        for (size_t q = k; q < k1; ++q) {
          const size_t pid = (size_t)id[q];
          if (B.block_by_id_.size() <= pid) B.block_by_id_.resize(pid + 1, {blocks::kBlockNone, 0});
          B.block_by_id_[pid] = {b, q - k};
        }
        most = std::max(most, cnt);
      }
      k = k1;
    }
    B.num_particles_ = n;
    B.num_per_cell_ = most;
  }
  self->blocks_buffer_ = self->Blocks;
  std::vector<particle> res;
  res.reserve(n);
  for (size_t q = 0; q < n; ++q) res.emplace_back(p[q], v[q]);   // block order = the order handed over
  self->particle_buffer_ = res;
}

void Particles::CheckBonds() {  // particles.cpp:205-215: the engine prunes; refresh bonds_
  pull_pairs(ptoy_get_bonds, engine_of(this), bonds_);
}

void Particles::SetParticleBuffer() {  // particles.cpp:79-89
  ptoy_engine* e = engine_of(this);
  check(ptoy_set_particle_buffer(e), "ptoy_set_particle_buffer");
  rebuild_snapshot(this, e);
}

// ---- stepping ------------------------------------------------------------------------------------

void Particles::step(Scal time_target, bool quit) {  // particles.cpp:93-203
  if (quit) return;
  ptoy_engine* e = engine_of(this);
  sync_inline_state(this, e);
  const RectVect domain_before = domain;
  const bool snap = renderer_ready_for_next_;
  check(ptoy_set_renderer_ready(e, snap), "renderer_ready");
  float before = 0, after = 0;
  check(ptoy_time(e, &before), "time");
  check(ptoy_step(e, time_target), "ptoy_step");
  check(ptoy_time(e, &after), "time");
  t = after;
  if (after == before) return;  // t >= time_target: the loop body never ran
  if (snap) {  // the snapshot was taken on the first sub-step (:190-196)
    rebuild_snapshot(this, e);
    pull_pairs(ptoy_get_bonds, e, bonds_);
    pull_pairs(ptoy_get_no_rendering, e, no_rendering_buffer_);
    no_rendering_ = no_rendering_buffer_;
    renderer_ready_for_next_ = false;
  } else {
    float dom[4];
    check(ptoy_get_domain(e, dom), "domain");
    domain = RectVect(Vect(dom[0], dom[1]), Vect(dom[2], dom[3]));
  }
  if (remove_last_portal_ || domain != domain_before) {
    // :182-187 happened inside the engine / the frame moved: UpdateEnvObj re-derived Portal::blocks
    remove_last_portal_ = false;
    pull_portals(this, e);
  }
}

// ---- mutators ---------------------------------------------------------------------------------------

void Particles::SetForce(Vect center, bool enabled) {  // particles.cpp:217-220
  force_center = center;
  force_enabled = enabled;
}
void Particles::SetForce(Vect center) { force_center = center; }  // :221-223
void Particles::SetForce(bool enabled) { force_enabled = enabled; }  // :224-226

void Particles::AddEnvObj(env_object* env) {  // particles.cpp:90-92 (takes ownership)
  ENVOBJ.push_back(std::unique_ptr<env_object>(env));
}

void Particles::UpdateEnvObj() {  // particles.cpp:708-725: walls -> device
  std::vector<float> seg;
  for (auto& o : ENVOBJ) {
    const line* l = dynamic_cast<const line*>(o.get());
    if (!l) throw std::runtime_error("ptoy_b200: only `line` environment objects run on the device");
    // `line` (particles.h:34-58) keeps its end points private and offers no accessor, but the
    // header's own inline ResetEnvObjFrame hands `line` objects to AddEnvObj and the device
    // needs their geometry: read them through the object layout (vptr, A, B), which the
    // unchanged header fixes.
    struct LineLayout { void* vptr; Scal ax, ay, bx, by; };
    static_assert(sizeof(line) == sizeof(LineLayout), "unexpected layout of class line");
    const LineLayout* m = reinterpret_cast<const LineLayout*>(l);
    seg.insert(seg.end(), {m->ax, m->ay, m->bx, m->by});
  }
  ptoy_engine* e = engine_of(this);
  check(ptoy_set_walls(e, seg.data(), ENVOBJ.size()), "ptoy_set_walls");
  pull_portals(this, e);  // UpdateEnvObj re-derives Portal::blocks (:720-724)
}

void Particles::UpdatePortalBlocks(Portal& portal) {  // particles.cpp:699-706
  portal.blocks.clear();
  for (size_t ib = 0; ib < Blocks.GetNumBlocks(); ++ib)
    if (portal.IsClose(Blocks.GetCenter(ib), Blocks.GetCircumRadius())) portal.blocks.push_back(ib);
}

// DetectPortals / ApplyPortals / ApplyPortalsForces are phases of one iteration of step()
// (particles.cpp:102-107, 177-178).  On the device they are fused into the stage kernels and
// cannot be run in isolation; no caller outside step() uses them.
void Particles::DetectPortals() {}
void Particles::ApplyPortals() {}
void Particles::ApplyPortalsForces() {}

// particles.cpp:264-286, restated (pure function of its arguments)
void Particles::MoveToPortal(Vect& position, Vect& velocity, const Portal& src, const Portal& dest) {
  const Vect sr = src.end - src.begin, dr = dest.end - dest.begin;
  const Vect sn = Vect(-sr.y, sr.x).GetNormalized(), dn = Vect(-dr.y, dr.x).GetNormalized();
  const Scal lp = sr.dot(position - src.begin) / sr.dot(sr);
  const Scal op = sn.dot(position - src.begin);
  const Scal lv = sr.dot(velocity) / sr.dot(sr);
  const Scal ov = sn.dot(velocity);
  const Scal sign = (op > 0. ? 1. : -1.);
  position = dest.begin + dr * lp + dn * (op - sign * 2. * kPortalThickness);
  velocity = dr * lv + dn * ov;
}

// ---- UI tools (particles.cpp:228-262, 411-560): forwarded, mirrors refreshed ---------------------------

namespace {
void tool(Particles* self, int which, int phase, Vect p) {
  ptoy_engine* e = engine_of(self);
  sync_inline_state(self, e);
  check(ptoy_tool(e, which, phase, p.x, p.y), "ptoy_tool");
}
}  // namespace

void Particles::PickStart(Vect p) { tool(this, PTOY_TOOL_PICK, PTOY_PHASE_START, p); pick_enabled_ = true; pick_pointer_ = p; }
void Particles::PickMove(Vect p) { tool(this, PTOY_TOOL_PICK, PTOY_PHASE_MOVE, p); if (pick_enabled_) pick_pointer_ = p; }
void Particles::PickStop(Vect p) { tool(this, PTOY_TOOL_PICK, PTOY_PHASE_STOP, p); pick_enabled_ = false; }

void Particles::BondsStart(Vect p) { tool(this, PTOY_TOOL_BONDS, PTOY_PHASE_START, p); bonds_enabled_ = true; }
void Particles::BondsMove(Vect p) {
  tool(this, PTOY_TOOL_BONDS, PTOY_PHASE_MOVE, p);
  pull_pairs(ptoy_get_bonds, engine_of(this), bonds_);
}
void Particles::BondsStop(Vect p) { tool(this, PTOY_TOOL_BONDS, PTOY_PHASE_STOP, p); bonds_enabled_ = false; }

namespace {
void pull_frozen(Particles* self) {
  ptoy_engine* e = engine_of(self);
  size_t n = 0;
  check(ptoy_get_frozen(e, nullptr, 0, &n), "frozen");
  std::vector<int32_t> ids(n + 1);
  check(ptoy_get_frozen(e, ids.data(), n, &n), "frozen");
  std::set<int> next(ids.begin(), ids.begin() + n);
  // the reference reports each toggle on stdout (particles.cpp:545,548)
  for (int id : next)
    if (!self->frozen_.count(id)) std::cout << "Freeze particle id=" << id << std::endl;
  for (int id : self->frozen_)
    if (!next.count(id)) std::cout << "Unfreeze particle id=" << id << std::endl;
  self->frozen_ = next;
}
}  // namespace

void Particles::FreezeStart(Vect p) { tool(this, PTOY_TOOL_FREEZE, PTOY_PHASE_START, p); freeze_enabled_ = true; pull_frozen(this); }
void Particles::FreezeMove(Vect p) { tool(this, PTOY_TOOL_FREEZE, PTOY_PHASE_MOVE, p); pull_frozen(this); }
void Particles::FreezeStop(Vect p) { tool(this, PTOY_TOOL_FREEZE, PTOY_PHASE_STOP, p); freeze_enabled_ = false; }

void Particles::PortalStart(Vect p) { tool(this, PTOY_TOOL_PORTAL, PTOY_PHASE_START, p); portal_enabled_ = true; pull_portals(this, engine_of(this)); }
void Particles::PortalMove(Vect p) { tool(this, PTOY_TOOL_PORTAL, PTOY_PHASE_MOVE, p); pull_portals(this, engine_of(this)); }
void Particles::PortalStop(Vect p) { tool(this, PTOY_TOOL_PORTAL, PTOY_PHASE_STOP, p); portal_enabled_ = false; pull_portals(this, engine_of(this)); }


// TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load what this builds.
//
// Thin C wrapper around the UNMODIFIED reference physics (compiled where it
// lies: /root/reference/src/particles.cpp + geometry.cpp, see oracle/Makefile).
// It drives `class Particles` (particles.h:60-228) through its own methods and
// reaches private state with the access trick of SURVEY.md Appendix C
// (`#define private public` in THIS translation unit only; the class layout is
// unchanged so it links against the normally compiled particles.cpp).
//
// Nothing here re-implements physics: every ref_* entry point forwards to the
// reference.  Output goes to oracle/_ref/ (git-ignored, travels with gpurun).

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <set>
#include <sstream>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define private public
#include "particles.h"
#undef private

namespace {
struct Ref {
  Particles ps;
};
// blocks::SetDomain prints to stdout when the grid grows (blocks.h:138);
// silence it so pytest/bench output stays clean.
struct CoutMute {
  std::streambuf* old;
  std::ostringstream sink;
  CoutMute() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~CoutMute() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// Default constructor scene (particles.cpp:26-77): 625 particles, one portal pair.
void* ref_create_default() {
  CoutMute m;
  Ref* r = new Ref();
  r->ps.renderer_ready_for_next_ = false;
  return r;
}

void ref_destroy(void* h) { delete static_cast<Ref*>(h); }

// Empty the scene and install a new domain the reference way (Appendix C):
// grid grown by blocks::SetDomain (blocks.h:119-141), frame walls by
// ResetEnvObjFrame (particles.h:99-107).  Never assigns `Blocks` (dangling parent).
void ref_reset_scene(void* h, float ax, float ay, float bx, float by) {
  CoutMute m;
  Particles& ps = static_cast<Ref*>(h)->ps;
  ps.portals_.clear();
  ps.portal_stage_ = 0;
  ps.bonds_.clear();
  ps.frozen_.clear();
  ps.no_rendering_.clear();
  ps.particle_to_move_.clear();
  auto& d = ps.Blocks.GetData();
  d.clear();
  d.resize(ps.Blocks.GetNumBlocks());
  ps.Blocks.block_by_id_.clear();
  ps.Blocks.num_particles_ = 0;
  ps.Blocks.num_per_cell_ = 0;
  RectVect dom(Vect(ax, ay), Vect(bx, by));
  ps.SetDomain(dom);
  ps.resize_queue_ = dom;
  ps.ResetEnvObjFrame(dom);
  ps.t = 0;
  ps.force_enabled = false;
  ps.pick_enabled_ = false;
  ps.remove_last_portal_ = false;
  ps.renderer_ready_for_next_ = false;
}

// blocks::AddParticles(position, velocity, id)  (blocks.h:179-193)
void ref_add_particles(
    void* h, const float* pos, const float* vel, const int* id, size_t n) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  ArrayVect p(n), v(n);
  std::vector<int> ids(id, id + n);
  for (size_t i = 0; i < n; ++i) {
    p[i] = Vect(pos[2 * i], pos[2 * i + 1]);
    v[i] = Vect(vel[2 * i], vel[2 * i + 1]);
  }
  ps.Blocks.AddParticles(p, v, ids);
}

// Directed pairs inserted as given (the UI inserts both directions,
// particles.cpp:493-494; the caller does the same).
void ref_set_bonds(void* h, const int* pairs, size_t n) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  ps.bonds_.clear();
  // sorted input → hinted insertion is O(1) per pair
  for (size_t i = 0; i < n; ++i) {
    ps.bonds_.emplace_hint(ps.bonds_.end(), pairs[2 * i], pairs[2 * i + 1]);
  }
}

size_t ref_num_bonds(void* h) { return static_cast<Ref*>(h)->ps.bonds_.size(); }

size_t ref_get_bonds(void* h, int* pairs, size_t cap) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  size_t k = 0;
  for (auto b : ps.bonds_) {
    if (k >= cap) break;
    pairs[2 * k] = b.first;
    pairs[2 * k + 1] = b.second;
    ++k;
  }
  return ps.bonds_.size();
}

size_t ref_get_no_rendering(void* h, int* pairs, size_t cap) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  size_t k = 0;
  for (auto b : ps.no_rendering_) {
    if (k >= cap) break;
    pairs[2 * k] = b.first;
    pairs[2 * k + 1] = b.second;
    ++k;
  }
  return ps.no_rendering_.size();
}

void ref_set_frozen(void* h, const int* ids, size_t n) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  ps.frozen_.clear();
  ps.frozen_.insert(ids, ids + n);
}

size_t ref_get_frozen(void* h, int* ids, size_t cap) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  size_t k = 0;
  for (int id : ps.frozen_) {
    if (k >= cap) break;
    ids[k++] = id;
  }
  return ps.frozen_.size();
}

// A portal pair created the way the UI does: PortalStart/PortalStop twice
// (particles.cpp:411-450); the second portal is length-normalised (:441-443).
void ref_add_portal_pair(void* h, const float* s /*b0 e0 b1 e1, xy each*/) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  ps.PortalStart(Vect(s[0], s[1]));
  ps.PortalStop(Vect(s[2], s[3]));
  ps.PortalStart(Vect(s[4], s[5]));
  ps.PortalStop(Vect(s[6], s[7]));
}

void ref_clear_portals(void* h) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  ps.portals_.clear();
  ps.portal_stage_ = 0;
}

size_t ref_num_portal_pairs(void* h) {
  return static_cast<Ref*>(h)->ps.portals_.size();
}

// out: [pairs][2][4] = begin.xy end.xy
void ref_get_portals(void* h, float* out) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  size_t k = 0;
  for (auto& pair : ps.portals_)
    for (int d = 0; d < 2; ++d) {
      out[k++] = pair[d].begin.x;
      out[k++] = pair[d].begin.y;
      out[k++] = pair[d].end.x;
      out[k++] = pair[d].end.y;
    }
}

size_t ref_get_portal_blocks(
    void* h, size_t pair, int side, int64_t* out, size_t cap) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  const auto& b = ps.portals_[pair][side].blocks;
  for (size_t i = 0; i < b.size() && i < cap; ++i) out[i] = (int64_t)b[i];
  return b.size();
}

void ref_remove_last_portal(void* h) {
  static_cast<Ref*>(h)->ps.RemoveLastPortal();
}

void ref_set_gravity(void* h, int enabled, float gx, float gy) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  ps.SetGravity(enabled != 0);
  ps.SetGravityVect(Vect(gx, gy));
}

void ref_get_gravity(void* h, int* enabled, float* g) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  *enabled = ps.GetGravity();
  g[0] = ps.GetGravityVect().x;
  g[1] = ps.GetGravityVect().y;
}

void ref_set_point_force(
    void* h, int enabled, int attractive, float cx, float cy) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  ps.SetForce(Vect(cx, cy), enabled != 0);
  ps.SetForceAttractive(attractive != 0);
}

void ref_set_pick(void* h, int enabled, int id, float px, float py) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  ps.pick_enabled_ = enabled != 0;
  ps.pick_particle_id_ = id;
  ps.pick_pointer_ = Vect(px, py);
}

void ref_push_resize(void* h, float ax, float ay, float bx, float by) {
  static_cast<Ref*>(h)->ps.PushResize(RectVect(Vect(ax, ay), Vect(bx, by)));
}

void ref_set_renderer_ready(void* h, int v) {
  static_cast<Ref*>(h)->ps.SetRendererReadyForNext(v != 0);
}

// Particles::step(time_target, quit)  (particles.cpp:93)
void ref_step_to(void* h, float time_target) {
  CoutMute m;
  static_cast<Ref*>(h)->ps.step(time_target, false);
}

// Exactly n iterations of the while loop at particles.cpp:96, one per call
// (step(GetTime() + 0.25f*dt, false)).
void ref_step_n(void* h, int n, int snapshot_each) {
  CoutMute m;
  Particles& ps = static_cast<Ref*>(h)->ps;
  for (int i = 0; i < n; ++i) {
    ps.renderer_ready_for_next_ = snapshot_each != 0;
    ps.step(ps.GetTime() + 0.25f * ps.dt, false);
  }
}

// Stage-1 force evaluation exactly as the first half of an iteration
// (particles.cpp:97-107), without the update.  Mutates force[] (and zeroes the
// velocity of frozen particles, as ApplyFrozen does).
void ref_eval_forces(void* h) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  for (size_t i = 0; i < ps.Blocks.GetNumBlocks(); ++i) ps.calc_forces(i);
  ps.RHS_bonds();
  ps.ApplyPortalsForces();
  ps.ApplyFrozen();
}

float ref_time(void* h) { return static_cast<Ref*>(h)->ps.GetTime(); }
float ref_dt(void* h) { return static_cast<Ref*>(h)->ps.dt; }
size_t ref_num_steps(void* h) { return static_cast<Ref*>(h)->ps.GetNumSteps(); }
size_t ref_num_particles(void* h) {
  return static_cast<Ref*>(h)->ps.Blocks.GetNumParticles();
}
size_t ref_num_per_cell(void* h) {
  return static_cast<Ref*>(h)->ps.Blocks.GetNumPerCell();
}
size_t ref_num_blocks(void* h) {
  return static_cast<Ref*>(h)->ps.Blocks.GetNumBlocks();
}
size_t ref_id_capacity(void* h) {
  return static_cast<Ref*>(h)->ps.Blocks.block_by_id_.size();
}

// out[0..3] = Particles::domain, out[4..7] = blocks::domain_ (grid), dims i,j
void ref_get_geometry(void* h, float* dom8, int* dims2) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  dom8[0] = ps.domain.A.x;
  dom8[1] = ps.domain.A.y;
  dom8[2] = ps.domain.B.x;
  dom8[3] = ps.domain.B.y;
  dom8[4] = ps.Blocks.domain_.A.x;
  dom8[5] = ps.Blocks.domain_.A.y;
  dom8[6] = ps.Blocks.domain_.B.x;
  dom8[7] = ps.Blocks.domain_.B.y;
  dims2[0] = ps.Blocks.dims_.i;
  dims2[1] = ps.Blocks.dims_.j;
}

void ref_get_neighbor_offsets(void* h, int* out9) {
  auto o = static_cast<Ref*>(h)->ps.Blocks.GetNeighborOffsets();
  for (int i = 0; i < 9; ++i) out9[i] = o[i];
}

// Id-indexed state (SURVEY.md Appendix C "Read back by id").  Arrays have
// `cap` entries; ids that are absent get block = -1 and NaN payload.
void ref_get_state_by_id(
    void* h, size_t cap, float* pos, float* vel, float* force,
    int64_t* block, int64_t* slot) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  const auto& bbi = ps.Blocks.block_by_id_;
  const auto& d = ps.Blocks.GetData();
  const float nan = std::nanf("");
  for (size_t id = 0; id < cap; ++id) {
    bool ok = id < bbi.size() && bbi[id].first != blocks::kBlockNone;
    if (ok) {
      const size_t b = bbi[id].first, s = bbi[id].second;
      if (pos) { pos[2 * id] = d.position[b][s].x; pos[2 * id + 1] = d.position[b][s].y; }
      if (vel) { vel[2 * id] = d.velocity[b][s].x; vel[2 * id + 1] = d.velocity[b][s].y; }
      if (force) { force[2 * id] = d.force[b][s].x; force[2 * id + 1] = d.force[b][s].y; }
      if (block) block[id] = (int64_t)b;
      if (slot) slot[id] = (int64_t)s;
    } else {
      if (pos) pos[2 * id] = pos[2 * id + 1] = nan;
      if (vel) vel[2 * id] = vel[2 * id + 1] = nan;
      if (force) force[2 * id] = force[2 * id + 1] = nan;
      if (block) block[id] = -1;
      if (slot) slot[id] = -1;
    }
  }
}

// Per-block id lists in the reference's own (history-dependent) order:
// ids[block_start[b] .. block_start[b+1]).  block_start has nblocks+1 entries.
size_t ref_get_block_lists(void* h, int* ids, size_t cap, int64_t* block_start) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  const auto& d = ps.Blocks.GetData();
  size_t k = 0;
  for (size_t b = 0; b < ps.Blocks.GetNumBlocks(); ++b) {
    block_start[b] = (int64_t)k;
    for (size_t s = 0; s < d.id[b].size(); ++s) {
      if (k < cap) ids[k] = d.id[b][s];
      ++k;
    }
  }
  block_start[ps.Blocks.GetNumBlocks()] = (int64_t)k;
  return k;
}

// Render snapshot (particles.cpp:79-89): GetParticles() flattened p,v.
size_t ref_get_particle_buffer(void* h, float* pv, size_t cap) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  const auto& buf = ps.GetParticles();
  for (size_t i = 0; i < buf.size() && i < cap; ++i) {
    pv[4 * i] = buf[i].p.x;
    pv[4 * i + 1] = buf[i].p.y;
    pv[4 * i + 2] = buf[i].v.x;
    pv[4 * i + 3] = buf[i].v.y;
  }
  return buf.size();
}

// UI tool handlers, forwarded verbatim (control.cpp:125-180 is their caller).
void ref_tool(void* h, int tool, int phase, float x, float y) {
  CoutMute m;
  Particles& ps = static_cast<Ref*>(h)->ps;
  const Vect p(x, y);
  switch (tool * 3 + phase) {
    case 0: ps.BondsStart(p); break;
    case 1: ps.BondsMove(p); break;
    case 2: ps.BondsStop(p); break;
    case 3: ps.PickStart(p); break;
    case 4: ps.PickMove(p); break;
    case 5: ps.PickStop(p); break;
    case 6: ps.FreezeStart(p); break;
    case 7: ps.FreezeMove(p); break;
    case 8: ps.FreezeStop(p); break;
    case 9: ps.PortalStart(p); break;
    case 10: ps.PortalMove(p); break;
    case 11: ps.PortalStop(p); break;
    default: break;
  }
}

void ref_set_particle_buffer(void* h) {
  static_cast<Ref*>(h)->ps.SetParticleBuffer();
}

// Free functions of the force laws, for known-answer tests.
void ref_F12(const float* p, const float* q, float* out) {
  Vect f = F12(Vect(p[0], p[1]), Vect(q[0], q[1]));
  out[0] = f.x;
  out[1] = f.y;
}
void ref_F12wall(const float* p, const float* q, float* out) {
  Vect f = F12wall(Vect(p[0], p[1]), Vect(q[0], q[1]));
  out[0] = f.x;
  out[1] = f.y;
}
int64_t ref_find_block(void* h, float x, float y) {
  size_t b = static_cast<Ref*>(h)->ps.Blocks.FindBlock(Vect(x, y));
  return b == blocks::kBlockNone ? -1 : (int64_t)b;
}
void ref_move_to_portal(void* h, size_t pair, int side, float* pos, float* vel) {
  Particles& ps = static_cast<Ref*>(h)->ps;
  Vect p(pos[0], pos[1]), v(vel[0], vel[1]);
  ps.MoveToPortal(p, v, ps.portals_[pair][side], ps.portals_[pair][1 - side]);
  pos[0] = p.x; pos[1] = p.y; vel[0] = v.x; vel[1] = v.y;
}
}  // extern "C"

// File-scope force laws that particles.h does not declare.
Vect F12_bond(Vect p1, Vect p2);
Vect F12_portal_edge(Vect p1, Vect p2);
Vect F12_pick(Vect p1, Vect p2);
extern "C" {
void ref_F12_bond(const float* p, const float* q, float* out) {
  Vect f = F12_bond(Vect(p[0], p[1]), Vect(q[0], q[1]));
  out[0] = f.x;
  out[1] = f.y;
}
void ref_F12_portal_edge(const float* p, const float* q, float* out) {
  Vect f = F12_portal_edge(Vect(p[0], p[1]), Vect(q[0], q[1]));
  out[0] = f.x;
  out[1] = f.y;
}
void ref_F12_pick(const float* p, const float* q, float* out) {
  Vect f = F12_pick(Vect(p[0], p[1]), Vect(q[0], q[1]));
  out[0] = f.x;
  out[1] = f.y;
}
int ref_openmp_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}

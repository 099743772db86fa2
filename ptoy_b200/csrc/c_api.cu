// extern "C" boundary: include/ptoy_b200.h.  Every entry point catches C++ exceptions and
// turns them into an error code + message (the reference has no error codes, SURVEY.md §8b;
// its asserts become PTOY_ERR_ARG here).
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ptoy_b200.h"
#include "engine.h"

namespace ptoy {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
}  // namespace ptoy

using ptoy::Engine;
using ptoy::V2;

struct ptoy_engine {
  Engine* e;
};

#define PTOY_TRY(h, body)                                   \
  if (!(h) || !(h)->e) {                                    \
    ptoy::set_last_error("null engine handle");             \
    return PTOY_ERR_ARG;                                    \
  }                                                         \
  try {                                                     \
    Engine& E = *(h)->e;                                    \
    (void)E;                                                \
    body;                                                   \
    return PTOY_OK;                                         \
  } catch (const ptoy::Error& err) {                        \
    ptoy::set_last_error(err.what());                       \
    return err.code;                                        \
  } catch (const std::exception& err) {                     \
    ptoy::set_last_error(err.what());                       \
    return PTOY_ERR_STATE;                                  \
  }

namespace ptoy {
struct LocalGroup;
void nccl_unique_id(void* out128);
LocalGroup* make_local_group(const std::vector<Engine*>& engines);
void local_group_step(LocalGroup* g, int n);
void local_group_destroy(LocalGroup* g);
}  // namespace ptoy
struct ptoy_group {
  ptoy::LocalGroup* g;
};

extern "C" {

const char* ptoy_last_error(void) { return ptoy::g_last_error.c_str(); }
int ptoy_abi_version(void) { return 2; }

int ptoy_create(ptoy_engine** out, int device) {
  if (!out) { ptoy::set_last_error("null out pointer"); return PTOY_ERR_ARG; }
  *out = nullptr;
  try {
    ptoy_engine* h = new ptoy_engine{nullptr};
    h->e = new Engine(device);
    *out = h;
    return PTOY_OK;
  } catch (const ptoy::Error& err) {
    ptoy::set_last_error(err.what());
    return err.code;
  } catch (const std::exception& err) {
    ptoy::set_last_error(err.what());
    return PTOY_ERR_STATE;
  }
}
void ptoy_destroy(ptoy_engine* h) {
  if (!h) return;
  delete h->e;
  delete h;
}
int ptoy_load_default_scene(ptoy_engine* h) { PTOY_TRY(h, E.load_default_scene()) }
int ptoy_reset_scene(ptoy_engine* h, float ax, float ay, float bx, float by) {
  PTOY_TRY(h, E.reset_scene(V2{ax, ay}, V2{bx, by}))
}
int ptoy_add_particles(ptoy_engine* h, const float* pos, const float* vel, const int32_t* id, size_t n) {
  PTOY_TRY(h, E.add_particles(pos, vel, id, n, false))
}
int ptoy_add_particles_device(ptoy_engine* h, const float* pos, const float* vel, const int32_t* id, size_t n) {
  PTOY_TRY(h, E.add_particles(pos, vel, id, n, true))
}
int ptoy_upload_state(ptoy_engine* h, const float* pos, const float* vel, const int32_t* id, size_t n) {
  PTOY_TRY(h, E.upload_state(pos, vel, id, n))
}
int ptoy_upload_state_async(ptoy_engine* h, const float* pos, const float* vel, const int32_t* id, size_t n,
                            int32_t max_id) {
  PTOY_TRY(h, E.upload_state_async(pos, vel, id, n, max_id))
}
int ptoy_download_state_async(ptoy_engine* h, float* pos, float* vel, int32_t* id, size_t cap) {
  PTOY_TRY(h, E.download_state_async(pos, vel, id, cap))
}
int ptoy_downloaded_count(ptoy_engine* h, size_t* n) {
  PTOY_TRY(h, *n = E.downloaded_count())
}
int ptoy_download_state(ptoy_engine* h, float* pos, float* vel, int32_t* id, size_t cap, size_t* n) {
  PTOY_TRY(h, *n = E.download_state(pos, vel, id, cap))
}
int ptoy_step(ptoy_engine* h, float time_target) { PTOY_TRY(h, E.step_to(time_target)) }
int ptoy_step_n(ptoy_engine* h, int n) { PTOY_TRY(h, E.step_n(n)) }
int ptoy_synchronize(ptoy_engine* h) { PTOY_TRY(h, E.synchronize()) }
int ptoy_set_renderer_ready(ptoy_engine* h, int ready) { PTOY_TRY(h, E.set_renderer_ready(ready != 0)) }

int ptoy_set_gravity(ptoy_engine* h, int enabled) { PTOY_TRY(h, E.set_gravity(enabled != 0)) }
int ptoy_get_gravity(ptoy_engine* h, int* enabled) { PTOY_TRY(h, *enabled = E.gravity()) }
int ptoy_set_gravity_vect(ptoy_engine* h, float gx, float gy) { PTOY_TRY(h, E.set_gravity_vect(V2{gx, gy})) }
int ptoy_get_gravity_vect(ptoy_engine* h, float* g) {
  PTOY_TRY(h, { V2 v = E.gravity_vect(); g[0] = v.x; g[1] = v.y; })
}
int ptoy_set_point_force(ptoy_engine* h, int enabled, int attractive, float cx, float cy) {
  PTOY_TRY(h, E.set_point_force(enabled != 0, attractive != 0, V2{cx, cy}))
}
int ptoy_set_pick(ptoy_engine* h, int enabled, int32_t id, float px, float py) {
  PTOY_TRY(h, E.set_pick(enabled != 0, id, V2{px, py}))
}
int ptoy_push_resize(ptoy_engine* h, float ax, float ay, float bx, float by) {
  PTOY_TRY(h, E.push_resize(V2{ax, ay}, V2{bx, by}))
}
int ptoy_set_domain(ptoy_engine* h, float ax, float ay, float bx, float by) {
  PTOY_TRY(h, E.set_domain(V2{ax, ay}, V2{bx, by}))
}
int ptoy_set_walls(ptoy_engine* h, const float* seg, size_t n) { PTOY_TRY(h, E.set_walls(seg, n)) }
int ptoy_reset_env_frame(ptoy_engine* h, float ax, float ay, float bx, float by) {
  PTOY_TRY(h, E.reset_env_frame(V2{ax, ay}, V2{bx, by}))
}

int ptoy_set_bonds(ptoy_engine* h, const int32_t* pairs, size_t n) { PTOY_TRY(h, E.set_bonds(pairs, n)) }
int ptoy_toggle_bond(ptoy_engine* h, int32_t a, int32_t b) { PTOY_TRY(h, E.toggle_bond(a, b)) }
static size_t copy_pairs(const std::vector<std::pair<int, int>>& v, int32_t* out, size_t cap) {
  const size_t m = v.size() < cap ? v.size() : cap;
  for (size_t i = 0; i < m; ++i) { out[2 * i] = v[i].first; out[2 * i + 1] = v[i].second; }
  return v.size();
}
int ptoy_get_bonds(ptoy_engine* h, int32_t* pairs, size_t cap, size_t* n) {
  PTOY_TRY(h, *n = copy_pairs(E.bonds(), pairs, cap))
}
int ptoy_get_no_rendering(ptoy_engine* h, int32_t* pairs, size_t cap, size_t* n) {
  PTOY_TRY(h, *n = copy_pairs(E.no_rendering(), pairs, cap))
}
int ptoy_set_frozen(ptoy_engine* h, const int32_t* ids, size_t n) { PTOY_TRY(h, E.set_frozen(ids, n)) }
int ptoy_get_frozen(ptoy_engine* h, int32_t* ids, size_t cap, size_t* n) {
  PTOY_TRY(h, {
    const auto& f = E.frozen();
    const size_t m = f.size() < cap ? f.size() : cap;
    if (m) std::memcpy(ids, f.data(), m * sizeof(int));
    *n = f.size();
  })
}
int ptoy_add_portal_pair(ptoy_engine* h, const float* s) { PTOY_TRY(h, E.add_portal_pair(s)) }
int ptoy_clear_portals(ptoy_engine* h) { PTOY_TRY(h, E.clear_portals()) }
int ptoy_remove_last_portal(ptoy_engine* h) { PTOY_TRY(h, E.remove_last_portal()) }
int ptoy_num_portal_pairs(ptoy_engine* h, size_t* n) { PTOY_TRY(h, *n = E.portals().size()) }
int ptoy_get_portals(ptoy_engine* h, float* out) {
  PTOY_TRY(h, {
    size_t k = 0;
    for (const auto& pr : E.portals())
      for (const ptoy::PortalHost* p : {&pr.first, &pr.second}) {
        out[k++] = p->begin.x; out[k++] = p->begin.y; out[k++] = p->end.x; out[k++] = p->end.y;
      }
  })
}
int ptoy_get_portal_blocks(ptoy_engine* h, size_t pair, int side, int64_t* out, size_t cap, size_t* n) {
  PTOY_TRY(h, {
    if (pair >= E.portals().size() || side < 0 || side > 1) throw ptoy::Error(PTOY_ERR_ARG, "no such portal");
    const auto& b = side ? E.portals()[pair].second.blocks : E.portals()[pair].first.blocks;
    for (size_t i = 0; i < b.size() && i < cap; ++i) out[i] = b[i];
    *n = b.size();
  })
}
int ptoy_tool(ptoy_engine* h, int tool, int phase, float x, float y) {
  PTOY_TRY(h, {
    if (tool < 0 || tool > 3 || phase < 0 || phase > 2) throw ptoy::Error(PTOY_ERR_ARG, "unknown tool / phase");
    E.tool(tool, phase, V2{x, y});
  })
}
int ptoy_get_portal_tool_state(ptoy_engine* h, int* sm, float* xy8) { PTOY_TRY(h, E.portal_tool_state(sm, xy8)) }

int ptoy_time(ptoy_engine* h, float* t) { PTOY_TRY(h, *t = E.time()) }
int ptoy_dt(ptoy_engine* h, float* dt) { PTOY_TRY(h, *dt = E.dt()) }
int ptoy_num_steps(ptoy_engine* h, size_t* n) { PTOY_TRY(h, *n = E.num_steps()) }
int ptoy_get_domain(ptoy_engine* h, float* o) { PTOY_TRY(h, E.domain(o)) }
int ptoy_get_geometry(ptoy_engine* h, float* o8, int* d2) { PTOY_TRY(h, E.geometry(o8, d2)) }
int ptoy_num_particles(ptoy_engine* h, size_t* n) { PTOY_TRY(h, *n = E.num_particles()) }
int ptoy_snapshot_counts(ptoy_engine* h, size_t* np, size_t* npc) { PTOY_TRY(h, E.snapshot_counts(np, npc)) }
int ptoy_set_particle_buffer(ptoy_engine* h) { PTOY_TRY(h, E.take_snapshot()) }
int ptoy_get_snapshot(ptoy_engine* h, float* pos, float* vel, int32_t* id, int64_t* block, size_t cap, size_t* n) {
  PTOY_TRY(h, *n = E.snapshot(pos, vel, id, block, cap))
}
int ptoy_get_state_by_id(ptoy_engine* h, size_t id_cap, float* pos, float* vel, int64_t* block) {
  PTOY_TRY(h, E.state_by_id(id_cap, pos, vel, block))
}
int ptoy_eval_forces(ptoy_engine* h, int stage, size_t id_cap, float* force) {
  PTOY_TRY(h, E.eval_forces(stage, id_cap, force))
}
int ptoy_id_capacity(ptoy_engine* h, size_t* n) { PTOY_TRY(h, *n = E.id_capacity()) }
int ptoy_find_blocks(ptoy_engine* h, const float* pos, size_t n, int64_t* block) {
  PTOY_TRY(h, E.find_blocks(pos, n, block))
}
int ptoy_wire_particles(ptoy_engine* h, uint16_t* data, int max_size, int width, int height, int* n) {
  PTOY_TRY(h, *n = E.wire_particles(data, max_size, width, height))
}
int ptoy_wire_scene(ptoy_engine* h, int what, uint16_t* data, int max_size, int width, int height, int* n) {
  PTOY_TRY(h, *n = E.wire_scene(what, data, max_size, width, height))
}
// ---- strip decomposition -----------------------------------------------------------------

int ptoy_configure_strips(ptoy_engine* h, int rank, int world, const int32_t* col_begin, int ghost_rows,
                          size_t id_capacity, int bonds) {
  PTOY_TRY(h, E.configure_strips(rank, world, col_begin, ghost_rows, id_capacity, bonds != 0))
}
int ptoy_nccl_unique_id(void* out128) {
  try {
    ptoy::nccl_unique_id(out128);
    return PTOY_OK;
  } catch (const ptoy::Error& err) {
    ptoy::set_last_error(err.what());
    return err.code;
  }
}
int ptoy_nccl_connect(ptoy_engine* h, const void* uid128) { PTOY_TRY(h, E.connect_nccl(uid128)) }
int ptoy_group_create(ptoy_group** out, ptoy_engine** engines, int n) {
  if (!out || !engines || n < 1) { ptoy::set_last_error("bad group arguments"); return PTOY_ERR_ARG; }
  try {
    std::vector<Engine*> v;
    for (int i = 0; i < n; ++i) v.push_back(engines[i]->e);
    *out = new ptoy_group{ptoy::make_local_group(v)};
    return PTOY_OK;
  } catch (const ptoy::Error& err) {
    ptoy::set_last_error(err.what());
    return err.code;
  }
}
int ptoy_group_step_n(ptoy_group* g, int n) {
  if (!g) { ptoy::set_last_error("null group"); return PTOY_ERR_ARG; }
  try {
    ptoy::local_group_step(g->g, n);
    return PTOY_OK;
  } catch (const ptoy::Error& err) {
    ptoy::set_last_error(err.what());
    return err.code;
  } catch (const std::exception& err) {
    ptoy::set_last_error(err.what());
    return PTOY_ERR_STATE;
  }
}
void ptoy_group_destroy(ptoy_group* g) {
  if (!g) return;
  ptoy::local_group_destroy(g->g);
  delete g;
}
int ptoy_error_flags(ptoy_engine* h, int* flags) { PTOY_TRY(h, *flags = E.error_flags()) }

int ptoy_enable_kernel_timing(ptoy_engine* h, int enable) { PTOY_TRY(h, E.enable_kernel_timing(enable != 0)) }
int ptoy_get_kernel_times(ptoy_engine* h, double* ms6, int64_t* l6) { PTOY_TRY(h, E.kernel_times(ms6, l6)) }
int ptoy_launch_count(ptoy_engine* h, int64_t* n) { PTOY_TRY(h, *n = E.launch_count()) }
int ptoy_stream(ptoy_engine* h, void** s) { PTOY_TRY(h, *s = (void*)E.stream()) }

}  // extern "C"

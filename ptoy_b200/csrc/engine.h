// Host side of the ptoy_b200 engine: the state `class Particles` keeps on the host
// (particles.h:184-228) and the sequencing of one iteration of Particles::step
// (particles.cpp:96-201) as kernel launches on one CUDA stream.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "common.h"

namespace ptoy {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct V2 { float x, y; };

struct PortalHost {
  V2 begin, end;
  std::vector<int> blocks;  // Portal::blocks (particles.h:66)
};

template <class T>
struct DevBuf {
  T* p = nullptr;      // element 0; `head` more elements are addressable before it (ghost rows)
  size_t n = 0;
  size_t head = 0;
  void alloc(size_t count, size_t head_room = 0);          // discards contents
  void grow_keep(size_t count, cudaStream_t s, size_t head_room = 0);  // preserves [-head, n)
  void release();
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// Strip decomposition state: which block columns this rank owns, the staging
// buffers of the halo exchange and the migrant in/outboxes.
struct StripState {
  int world = 1, rank = 0;
  int col_begin[kMaxRanks + 1] = {0};  // first block column (blocks::dims_.i index) of every rank
  int row_begin[kMaxRanks + 1] = {0};  // the same as global sub-rows
  int ghost_rows = 2;                  // boundary sub-rows mirrored on either side (>= 2; wider for long bonds)
  bool bonds = false;                  // the decomposed scene has bonds (agreed across ranks at configure time)
  int ghost_cap = 0, out_cap = 0;
  // One message per neighbour and stage: positions (+ the sender's drift in the last element) |
  // start[] slice | ids, contiguous in `send_raw` / `recv_raw`; stage 2 sends the positions only.
  struct HaloView { float2* pos = nullptr; int* start = nullptr; int* id = nullptr; };
  DevBuf<unsigned char> send_raw[2], recv_raw[2];
  HaloView send[2], ghost[2];          // my boundary rows for the left / right neighbour; theirs as received
  size_t halo_pos_bytes = 0, halo_bytes = 0;   // bytes of the positions part / of the whole message
  size_t nstart = 0;
  DevBuf<int> ghost_n;                 // [2] ghosts installed per side this iteration
  DevBuf<unsigned char> outbox, inbox; // world blocks each (common.h DistDev layout)
  DevBuf<unsigned long long> arr_vkey;
};

class Engine {
 public:
  explicit Engine(int device);
  ~Engine();

  // scene
  void reset_scene(V2 A, V2 B);
  void load_default_scene();
  void add_particles(const float* pos, const float* vel, const int32_t* id, size_t n, bool on_device);
  void upload_state(const float* pos, const float* vel, const int32_t* id, size_t n);
  size_t download_state(float* pos, float* vel, int32_t* id, size_t cap);
  void upload_state_async(const float* pos, const float* vel, const int32_t* id, size_t n, int32_t max_id);
  void download_state_async(float* pos, float* vel, int32_t* id, size_t cap);
  size_t downloaded_count() const { return (size_t)h_ctr_->n_cur; }

  // stepping
  void step_to(float time_target);
  void step_n(int n);
  void synchronize();
  void set_renderer_ready(bool v) { renderer_ready_ = v; }

  // mutators
  void set_gravity(bool e) { gravity_enable_ = e; }
  bool gravity() const { return gravity_enable_; }
  void set_gravity_vect(V2 g) { gravity_ = g; }
  V2 gravity_vect() const { return gravity_; }
  void set_point_force(bool enabled, bool attractive, V2 c) { force_enabled_ = enabled; force_attractive_ = attractive; force_center_ = c; }
  void set_pick(bool enabled, int id, V2 p) { pick_enabled_ = enabled; pick_id_ = id; pick_pointer_ = p; }
  void push_resize(V2 A, V2 B) { rqA_ = A; rqB_ = B; }
  void set_domain(V2 A, V2 B);
  void set_walls(const float* seg, size_t n);
  void reset_env_frame(V2 A, V2 B);
  void set_bonds(const int32_t* pairs, size_t n);
  void toggle_bond(int a, int b);
  const std::vector<std::pair<int, int>>& bonds();  // pruned (CheckBonds)
  const std::vector<std::pair<int, int>>& no_rendering() const { return norend_buf_; }
  void set_frozen(const int32_t* ids, size_t n);
  const std::vector<int>& frozen() const { return frozen_; }
  void add_portal_pair(const float* s);
  void clear_portals();
  void remove_last_portal() { remove_last_portal_ = true; }
  const std::vector<std::pair<PortalHost, PortalHost>>& portals() const { return portals_; }
  void tool(int tool, int phase, V2 p);
  void portal_tool_state(int* stage_moving, float* xy8) const;

  // read-back
  float time() const { return t_; }
  float dt() const { return dt_; }
  size_t num_steps() const { return static_cast<size_t>(t_ / dt_); }  // particles.h:144-146
  void domain(float* out4) const;
  void geometry(float* out8, int* dims2) const;
  size_t num_particles();
  void take_snapshot();  // SetParticleBuffer
  size_t snapshot(float* pos, float* vel, int32_t* id, int64_t* block, size_t cap);
  void snapshot_counts(size_t* np, size_t* npc) const { *np = snap_num_particles_; *npc = snap_num_per_cell_; }
  void state_by_id(size_t id_cap, float* pos, float* vel, int64_t* block);
  void eval_forces(int stage, size_t id_cap, float* force);
  size_t id_capacity() const { return id_cap_; }
  // wire format of the remote front end (ptoy_wasm.cpp:130-198): uint16 pixel coordinates
  int wire_particles(uint16_t* data, int max_size, int width, int height);   // device quantisation, engine order
  int wire_scene(int what, uint16_t* data, int max_size, int width, int height);  // 1 portals, 2 bonds, 3 frozen (host, snapshot)
  void find_blocks(const float* pos, size_t n, int64_t* block);

  // strip decomposition (dist.cu).  Phases of one iteration, so that a transport (in-process
  // group of strips, or NCCL between processes) can run the exchanges between them.
  void configure_strips(int rank, int world, const int* col_begin, int ghost_rows, size_t id_capacity, bool bonds);
  const StripState& strip() const { return strip_; }
  StripState& strip_mut() { return strip_; }
  void phase_prepare();        // host bookkeeping: dirty state -> device, resize decision
  void phase_pack_halo(int stage);   // boundary rows -> send buffers (stage 1: x0 + start slice, 2: x_half)
  void phase_install_halo(int stage);  // received boundary rows -> ghost rows (positions, start[], id -> slot map)
  void phase_unmark_halo();    // ghosts leave the id -> slot map again (before the reorder)
  void phase_pack_zones(int stage);    // portal-zone particles of the blocks this strip owns -> zone table
  void phase_mark_zones(int stage);    // after the tables were summed over the strips
  bool has_zones() const { return strip_.world > 1 && has_portal_tables_ && zone_entries_ > 0; }
  // the three arrays of the zone table for `stage` (device pointers, as int32 counts) for the transports
  void zone_buffers(int stage, int** pos_i32, size_t* n_pos, int** cnt, size_t* n_cnt, int** ids, size_t* n_ids);
  void phase_stage1();
  bool has_bonds() const { return !bonds_.empty() || (strip_.world > 1 && strip_.bonds); }
  bool has_portals() const { return !portals_.empty(); }
  void phase_stage2();
  void phase_reorder();        // arrivals + reorder
  void phase_end();            // snapshot, time
  void set_transport(struct Transport* t) { delete_transport(); transport_ = t; }
  void delete_transport();
  void connect_nccl(const void* uid128);
  void set_stream(cudaStream_t s);
  int device() const { return device_; }
  int error_flags();
  void check_device_errors(bool wait);   // throws PTOY_ERR_CAPACITY / PTOY_ERR_DIST if a sticky error bit is set

  // measurement
  void enable_kernel_timing(bool e);
  void kernel_times(double* ms6, int64_t* launches6);
  int64_t launch_count() const { return launches_; }
  cudaStream_t stream() const { return stream_; }

 private:
  // ---- host state mirroring Particles --------------------------------------------
  V2 domA_, domB_;     // Particles::domain
  V2 rqA_, rqB_;       // resize_queue_
  V2 gA_, gB_, bs_;    // blocks::domain_, block_size_
  int di_ = 0, dj_ = 0;
  float t_ = 0.f, dt_ = 0.f;
  V2 gravity_;
  bool gravity_enable_ = true;
  std::vector<float> walls_;  // 4 floats per line
  bool walls_are_frame_ = false;
  V2 force_center_{0.f, 0.f};
  bool force_enabled_ = false, force_attractive_ = false;
  bool pick_enabled_ = false;
  int pick_id_ = 0;
  V2 pick_pointer_{0.f, 0.f};
  std::vector<std::pair<int, int>> bonds_;  // sorted directed pairs = std::set order
  bool bonds_dirty_ = false;                // device CSR out of date
  int removed_seen_ = 0;                    // removed_total at the last prune
  std::vector<int> frozen_;
  bool frozen_dirty_ = false;
  std::vector<std::pair<PortalHost, PortalHost>> portals_;
  bool env_dirty_ = true;
  bool remove_last_portal_ = false;
  bool renderer_ready_ = false;
  std::vector<std::pair<int, int>> norend_buf_;
  // UI tool state (particles.h:199-218)
  int bonds_prev_id_ = -1;
  bool bonds_enabled_ = false, freeze_enabled_ = false, portal_enabled_ = false;
  int freeze_last_id_ = -1;
  int portal_stage_ = 0;
  bool portal_mouse_moving_ = false;
  V2 portal_begin_{0, 0}, portal_current_{0, 0}, portal_prev_first_{0, 0}, portal_prev_second_{0, 0};
  // render snapshot, block-major
  std::vector<V2> snap_pos_, snap_vel_;
  std::vector<int> snap_id_, snap_block_;
  size_t snap_num_particles_ = 0, snap_num_per_cell_ = 0;
  std::vector<size_t> snap_block_start_;   // block-major prefix of the snapshot, on the grid it was taken on
  int snap_di_ = 0, snap_dj_ = 0;
  V2 snap_gA_{0, 0};
  template <class F> void snapshot_near(V2 point, int rings, F&& visit) const;

  // ---- device state -------------------------------------------------------------------
  int device_;
  cudaStream_t stream_ = nullptr;
  size_t cap_ = 0;      // particle capacity
  int n_bound_ = 0;     // host upper bound of the device particle count
  size_t id_cap_ = 0;
  int cur_ = 0;         // which of the ping-pong buffers holds the sorted state
  DevBuf<float2> pos_[2], vel_[2], pos_h_, vel_h_, force_dbg_;
  DevBuf<int> id_[2], order_, fixlist_;   // order_: reorder tags / scratch
  DevBuf<int> cell_[2], newcell_;         // sort keys: of the sorted particles / after the update
  DevBuf<int> bslot_;            // [cap][8] bond partner slots + degree: stage 1 -> stage 2
  DevBuf<int> bond_ell_;         // [id_cap+1][8] bond table by id
  DevBuf<int> cnt_, start_;
  int n_items_ = 0;     // padded scan length
  DevBuf<unsigned long long> tile_state_;
  DevBuf<int> ticket_;
  unsigned epoch_ = 0;
  DevBuf<Counters> ctr_;
  Counters* h_ctr_ = nullptr;  // pinned
  int* h_err_ = nullptr;       // pinned: err_flags as of the last completed iteration (async copy)
  cudaEvent_t err_event_ = nullptr;
  bool err_pending_ = false;
  DevBuf<int> slot_of_id_;
  bool slot_map_valid_ = false;
  DevBuf<uint32_t> frozen_bits_, bond_ptr_;
  DevBuf<int> bond_col_;
  DevBuf<PortalSide> sides_;
  DevBuf<int> pblk_ptr_, pblk_;
  DevBuf<uint32_t> wire_;         // quantised pixel pairs
  DevBuf<int> blk_cnt_, blk_off_; // render snapshot: particles per reference block, their exclusive prefix
  DevBuf<unsigned char> scan_tmp_;
  DevBuf<int> zcnt_, zid_;        // portal zone table (strips): particles per listed block, their ids
  int zone_entries_ = 0;          // listed portal blocks (pblk_ptr[n_sides])
  size_t zone_off() const { return cap_ + (size_t)strip_.ghost_cap; }   // slot of the table's first particle
  DevBuf<uint8_t> blk_flags_;
  bool has_portal_tables_ = false;
  GridDev grid_{};
  float wall_reach_ = 0.f;
  WallDev walls_dev_[kMaxWalls];
  float free_box_[4];
  float r2c_ = 0.f, v2c_ = 0.f;
  float margin1_ = 0.f, margin2_ = 0.f, margin2_base_ = 0.f, margin2_scale_ = 0.f;

  // timing
  bool timing_ = false;
  struct TimedLaunch { int kind; cudaEvent_t a, b; };
  std::vector<TimedLaunch> timed_;
  std::vector<cudaEvent_t> event_pool_;
  double kind_ms_[6] = {0, 0, 0, 0, 0, 0};
  int64_t kind_launches_[6] = {0, 0, 0, 0, 0, 0};
  int64_t launches_ = 0;

  StripState strip_;
  struct Transport* transport_ = nullptr;
  bool owns_stream_ = true;
  float t_half_ = 0.f;
  bool resize_now_ = false, grid_changes_ = false;
  V2 rs_nA_{0, 0}, rs_nB_{0, 0}, rs_gA_{0, 0}, rs_gB_{0, 0};
  void alloc_strip_buffers();
  DistDev make_dist(bool with_outbox) const;
  int arrivals_bound() const { return strip_.world > 1 ? strip_.world * strip_.out_cap : 0; }
 public:
  // Without portals a particle moves less than one sub-row per iteration, so it can only enter an
  // adjacent strip: migrants are exchanged with the two neighbours only.  With portals a teleport can
  // land on any strip: all-to-all.  (Every rank holds the same portal list, so all ends agree.)
  bool migrants_neighbours_only() const { return !local_group_ && portals_.empty(); }
  void set_local_group(bool v) { local_group_ = v; }
 private:
  bool local_group_ = false;

  // ---- helpers --------------------------------------------------------------------------
  void init_grid(V2 A, V2 B);                 // blocks::InitEmptyBlocks
  bool grown_grid(V2 pA, V2 pB, V2* nA, V2* nB) const;  // blocks::SetDomain's loop
  void alloc_cells();
  void ensure_capacity(size_t n);
  void ensure_id_capacity(size_t n);
  void update_portal_blocks(PortalHost& p) const;  // Particles::UpdatePortalBlocks
  void sync_env();      // walls, portal tables, block flags -> device
  void sync_bonds();    // CSR -> device
  void sync_frozen();
  void ensure_slot_map();
  Phys make_phys() const;
  StageArgs make_stage_args(int stage);
  void rekey_and_reorder(int first, bool teleport);  // tail + scan + fill + gather
  void reorder();                                    // scan + fill + gather + finalize
  void iterate();                                    // particles.cpp:96-201
  void read_counters();
  void prune_bonds();
  void compute_no_rendering();
  struct Scope;
  void begin_timed(int kind);
  void end_timed();
  void drain_timed();
};

// Exchanges between the strips of one iteration.  halo(stage): send_pos (+ send_start when
// stage == 1) of every rank -> ghost_pos / ghost_pos_h (+ ghost_start) of its neighbours;
// migrants(): outbox block d of rank s -> inbox block s of rank d.
struct Transport {
  virtual ~Transport() {}
  virtual void halo(Engine& e, int stage) = 0;
  virtual void migrants(Engine& e) = 0;
  virtual void zones(Engine& e, int stage) = 0;   // sum the portal zone tables over the strips
};

void set_last_error(const std::string& m);

}  // namespace ptoy

/* ptoy_b200 — C ABI of the B200-native particle-stepping engine.
 *
 * Drop-in boundary for ptoy's physics path: every entry point below replaces one
 * public member of the reference `class Particles`
 * (reference: src/particles.h:60-228, implemented in src/particles.cpp) or one
 * operation of its cell grid `class blocks` (src/blocks.h:13-277).  The reference
 * has no FFI for this path — the boundary there is the C++ class itself
 *  — so the replacement `particles.cpp` façade
 * (ptoy_b200/host/particles_b200.cpp) forwards each member to
 * the function named here, and the Python host mirror (ptoy_b200/particles.py)
 * binds the same symbols with ctypes.
 *
 * Conventions: plain C types, host pointers unless the name ends in `_device`;
 * positions/velocities are interleaved xy float pairs (layout of `Vect`,
 * geometry.h:18-90); every function returns PTOY_OK (0) or a negative error code
 * and records a message retrievable with ptoy_last_error().  All calls on one
 * handle must come from one host thread (the reference's callers are
 * single-threaded).  There is NO CPU fallback: without a CUDA
 * device ptoy_create fails with PTOY_ERR_CUDA.
 */
#ifndef PTOY_B200_H_
#define PTOY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptoy_engine ptoy_engine;

enum {
  PTOY_OK = 0,
  PTOY_ERR_CUDA = -1,      /* CUDA runtime error / no device */
  PTOY_ERR_ARG = -2,       /* invalid argument */
  PTOY_ERR_STATE = -3,     /* call not valid in the current state */
  PTOY_ERR_CAPACITY = -4,  /* a device-side capacity was exceeded */
  PTOY_ERR_DIST = -5       /* multi-GPU transport error */
};

/* Tools of the reference UI controller (control.cpp:125-180 → particles.cpp). */
enum { PTOY_TOOL_BONDS = 0, PTOY_TOOL_PICK = 1, PTOY_TOOL_FREEZE = 2, PTOY_TOOL_PORTAL = 3 };
enum { PTOY_PHASE_START = 0, PTOY_PHASE_MOVE = 1, PTOY_PHASE_STOP = 2 };

const char* ptoy_last_error(void);
/* ABI version of this header (bumped on incompatible change). */
int ptoy_abi_version(void);

/* ---- construction --------------------------------------------------------------- */

/* Particles::Particles() (particles.cpp:26-77) WITHOUT the default scene: domain
 * [-1,1]^2, block size 4r, dt = 0.0005, gravity (0,-10) enabled, four frame walls,
 * no particles, no portals.  `device` = CUDA ordinal. */
int ptoy_create(ptoy_engine** out, int device);
/* The constructor's scene (particles.cpp:34-76): 25x25 hex pile + one portal pair. */
int ptoy_load_default_scene(ptoy_engine* h);
void ptoy_destroy(ptoy_engine* h);

/* Empty the scene and install a new domain: grid grown the reference way
 * (blocks::SetDomain, blocks.h:119-141), frame walls rebuilt
 * (ResetEnvObjFrame, particles.h:99-107), t = 0, bonds/frozen/portals cleared. */
int ptoy_reset_scene(ptoy_engine* h, float ax, float ay, float bx, float by);

/* blocks::AddParticles(position, velocity, id) (blocks.h:179-193): particles whose
 * position maps outside the block grid are skipped; within a cell existing
 * particles stay first and new ones follow in input order.  ids must be >= 0. */
int ptoy_add_particles(ptoy_engine* h, const float* pos_xy, const float* vel_xy,
                       const int32_t* id, size_t n);
/* Same, inputs already resident in device memory (used for device-side timing). */
int ptoy_add_particles_device(ptoy_engine* h, const float* d_pos_xy, const float* d_vel_xy,
                              const int32_t* d_id, size_t n);

/* Replace the particle state (positions, velocities, ids) in one call while keeping bonds,
 * frozen set, portals, walls and time: the bulk form of emptying the grid and calling
 * blocks::AddParticles (SURVEY.md Appendix C "load a custom scene").  Host buffers; pinned
 * memory makes the copy asynchronous. */
int ptoy_upload_state(ptoy_engine* h, const float* pos_xy, const float* vel_xy,
                      const int32_t* id, size_t n);
/* Current device state in engine order (sub-cell sorted) copied straight into host buffers
 * (cap entries each, any pointer may be NULL); *n = live particle count.  Synchronises. */
int ptoy_download_state(ptoy_engine* h, float* pos_xy, float* vel_xy, int32_t* id,
                        size_t cap, size_t* n);

/* Pipelined forms of the two calls above, for streaming independent batches through several
 * handles (every handle owns a CUDA stream): the copies and the re-binning are only ENQUEUED.
 * The host buffers must be pinned and stay untouched until ptoy_synchronize(h) returns.
 * max_id = largest id in the batch (the synchronous form scans for it).  The async upload does
 * not run CheckBonds (particles.cpp:205-215) against the state it replaces.
 * ptoy_download_state_async copies min(cap, upper bound of the live count) entries;
 * ptoy_downloaded_count returns the exact live count once ptoy_synchronize has returned. */
int ptoy_upload_state_async(ptoy_engine* h, const float* pos_xy, const float* vel_xy,
                            const int32_t* id, size_t n, int32_t max_id);
int ptoy_download_state_async(ptoy_engine* h, float* pos_xy, float* vel_xy, int32_t* id, size_t cap);
int ptoy_downloaded_count(ptoy_engine* h, size_t* n);

/* ---- stepping --------------------------------------------------------------------- */

/* Particles::step(time_target, quit=false) (particles.cpp:93-203): iterate
 * `while (t < time_target)`; each iteration = two force evaluations + re-binning. */
int ptoy_step(ptoy_engine* h, float time_target);
/* Exactly n iterations of that loop (what step(GetTime()+0.25f*dt) does, n times).
 * Does not synchronise the device. */
int ptoy_step_n(ptoy_engine* h, int n);
/* Block until all queued device work of this engine has finished. */
int ptoy_synchronize(ptoy_engine* h);
/* Particles::SetRendererReadyForNext (particles.h:179-181): the next iteration ends
 * with a render snapshot (particles.cpp:190-196). */
int ptoy_set_renderer_ready(ptoy_engine* h, int ready);

/* ---- mutators used between steps ------------------------------------------------------ */

int ptoy_set_gravity(ptoy_engine* h, int enabled);                    /* particles.h:155 SetGravity */
int ptoy_get_gravity(ptoy_engine* h, int* enabled);                   /* particles.h:152 GetGravity */
int ptoy_set_gravity_vect(ptoy_engine* h, float gx, float gy);        /* particles.h:161 SetGravityVect */
int ptoy_get_gravity_vect(ptoy_engine* h, float* g_xy);               /* particles.h:158 GetGravityVect */
/* SetForce(center,enabled) / SetForce(center) / SetForce(enabled) / SetForceAttractive
 * (particles.cpp:217-226, particles.h:123-125): pass the full state. */
int ptoy_set_point_force(ptoy_engine* h, int enabled, int attractive, float cx, float cy);
/* pick_enabled_/pick_particle_id_/pick_pointer_ (particles.cpp:245-247, 874-876) */
int ptoy_set_pick(ptoy_engine* h, int enabled, int32_t id, float px, float py);
int ptoy_push_resize(ptoy_engine* h, float ax, float ay, float bx, float by); /* particles.h:112 PushResize */
int ptoy_set_domain(ptoy_engine* h, float ax, float ay, float bx, float by);  /* particles.h:108 SetDomain */
/* ClearEnvObj + AddEnvObj(new line(A,B)) x n + UpdateEnvObj (particles.h:93-107). */
int ptoy_set_walls(ptoy_engine* h, const float* seg_axaybxby, size_t n);
int ptoy_reset_env_frame(ptoy_engine* h, float ax, float ay, float bx, float by); /* particles.h:99 */

/* bonds_ (particles.h:206): replace the whole set of DIRECTED pairs (the UI stores
 * both directions, particles.cpp:493-494).  Any order; duplicates are merged. */
int ptoy_set_bonds(ptoy_engine* h, const int32_t* pairs_ab, size_t n_pairs);
/* Toggle both directions of a bond as BondsMove does (particles.cpp:484-496). */
int ptoy_toggle_bond(ptoy_engine* h, int32_t a, int32_t b);
/* GetBonds() after CheckBonds (particles.cpp:205-215): sorted directed pairs.
 * Writes up to cap pairs, always returns the full count in *n. */
int ptoy_get_bonds(ptoy_engine* h, int32_t* pairs_ab, size_t cap, size_t* n);
/* GetNoRendering() (particles.h:225-227): bonds acting through a portal at the
 * last render snapshot. */
int ptoy_get_no_rendering(ptoy_engine* h, int32_t* pairs_ab, size_t cap, size_t* n);

int ptoy_set_frozen(ptoy_engine* h, const int32_t* ids, size_t n);    /* frozen_ (particles.h:207) */
int ptoy_get_frozen(ptoy_engine* h, int32_t* ids, size_t cap, size_t* n); /* particles.h:118 GetFrozen */

/* One portal pair given as begin0 end0 begin1 end1 (xy each), created exactly as
 * PortalStart/PortalStop x2 do (particles.cpp:411-450; second portal length-normalised). */
int ptoy_add_portal_pair(ptoy_engine* h, const float* b0e0b1e1_xy);
int ptoy_clear_portals(ptoy_engine* h);
int ptoy_remove_last_portal(ptoy_engine* h);                           /* particles.h:90 RemoveLastPortal */
int ptoy_num_portal_pairs(ptoy_engine* h, size_t* n);
/* GetPortals() (particles.h:82-84): out[pairs][2][4] = begin.xy end.xy */
int ptoy_get_portals(ptoy_engine* h, float* out);
/* Portal::blocks (particles.h:66, particles.cpp:699-706) of (pair, side). */
int ptoy_get_portal_blocks(ptoy_engine* h, size_t pair, int side, int64_t* out, size_t cap, size_t* n);

/* UI tool handlers Bonds/Pick/Freeze/Portal x Start/Move/Stop
 * (particles.cpp:228-262, 411-560); they act on the last render snapshot. */
int ptoy_tool(ptoy_engine* h, int tool, int phase, float x, float y);
/* public portal-drawing state (particles.h:214-218):
 * out = {portal_stage_, portal_mouse_moving_}, xy6 = begin, current, prev.first, prev.second */
int ptoy_get_portal_tool_state(ptoy_engine* h, int* stage_moving2, float* xy8);

/* ---- read-back ------------------------------------------------------------------------ */

int ptoy_time(ptoy_engine* h, float* t);                 /* particles.h:141 GetTime */
int ptoy_dt(ptoy_engine* h, float* dt);
int ptoy_num_steps(ptoy_engine* h, size_t* n);           /* particles.h:144 GetNumSteps */
int ptoy_get_domain(ptoy_engine* h, float* axaybxby);    /* particles.h:147 GetDomain */
/* out8 = Particles::domain, blocks::domain_ ; dims2 = dims_.i, dims_.j (blocks.h:247-251) */
int ptoy_get_geometry(ptoy_engine* h, float* out8, int* dims2);
/* Live particle count on the device (blocks::GetNumParticles, blocks.h:206). Synchronises. */
int ptoy_num_particles(ptoy_engine* h, size_t* n);
/* GetNumParticles()/GetNumPerCell() of the render snapshot (particles.h:167-172). */
int ptoy_snapshot_counts(ptoy_engine* h, size_t* num_particles, size_t* num_per_cell);

/* SetParticleBuffer (particles.cpp:79-89) on demand: take a render snapshot now. */
int ptoy_set_particle_buffer(ptoy_engine* h);
/* The render snapshot in reference block order (block ascending): what
 * GetParticles()/GetBlockData()/GetBlockById() expose (particles.h:164-178).
 * Arrays hold up to cap particles; *n = snapshot size.  Any pointer may be NULL. */
int ptoy_get_snapshot(ptoy_engine* h, float* pos_xy, float* vel_xy, int32_t* id,
                      int64_t* block, size_t cap, size_t* n);

/* Current device state indexed by particle id (test / parity hook; the analogue of
 * reading data.position[b][s] through block_by_id_).
 * Arrays have id_cap entries; absent ids get block = -1 and NaN payload.
 * Any pointer may be NULL.  Synchronises. */
int ptoy_get_state_by_id(ptoy_engine* h, size_t id_cap, float* pos_xy, float* vel_xy,
                         int64_t* block);
/* Evaluate the forces of `stage` (1: at the current state; 2: at the midpoint of an
 * iteration started from the current state) WITHOUT advancing the state, indexed by
 * particle id — the value of data.force after particles.cpp:97-107 / :128-138. */
int ptoy_eval_forces(ptoy_engine* h, int stage, size_t id_cap, float* force_xy);
/* id capacity = 1 + largest id ever added (size of block_by_id_, blocks.h:253). */
int ptoy_id_capacity(ptoy_engine* h, size_t* n);
/* blocks::FindBlock (blocks.h:155-163) evaluated on the device for n points; -1 = none. */
int ptoy_find_blocks(ptoy_engine* h, const float* pos_xy, size_t n, int64_t* block);

/* ---- wire format of the remote (WASM / browser) front end ------------------------------------
 *
 * The exports of ptoy_wasm.cpp:130-198 hand the browser uint16 pixel coordinates,
 *   data[0] = (1 + p.x) * width * 0.5,   data[1] = (1 - p.y) * height * 0.5,
 * `max_size` = capacity of `data` in uint16 entries (ptoy_inc.js:9-27 uses 10 000); *n = entries
 * written.  ptoy_wire_particles replaces GetParticles (:130-146): the positions are quantised and
 * culled ON THE DEVICE from the live state, so 4 bytes per drawn particle cross PCIe; the order
 * is the engine's, not block order (the consumer draws points).  ptoy_wire_scene replaces
 * GetPortals (what = 1, 8 entries per pair, the pair being drawn included, :147-166),
 * GetBonds (2, 4 entries per drawn bond, :167-182) and GetFrozen (3, :183-198), built from the
 * host state and the last render snapshot in the order UpdateScene (:47-116) builds its lists. */
int ptoy_wire_particles(ptoy_engine* h, uint16_t* data, int max_size, int width, int height, int* n);
int ptoy_wire_scene(ptoy_engine* h, int what, uint16_t* data, int max_size, int width, int height, int* n);

/* ---- strip decomposition across GPUs (new work: the reference is single-process) ---
 *
 * The domain is cut into strips over the slow block index (blocks.h:160: block = i*dims.j + j):
 * rank r owns block columns [col_begin[r], col_begin[r+1]).  Per iteration the strips exchange
 * their two boundary sub-rows before each force evaluation and the particles that crossed a
 * strip boundary (or teleported) after the update.  N strips give bit-identical results to one
 * GPU.  Configure BEFORE adding particles; ptoy_add_particles then keeps only the particles of
 * the rank's own strip. */
int ptoy_configure_strips(ptoy_engine* h, int rank, int world, const int32_t* col_begin /* world+1 */,
                          int ghost_rows, size_t id_capacity, int bonds);
/*   ghost_rows   boundary sub-rows (0.04 each) mirrored on either side of a strip; 0 = the default 2,
 *                which covers the pair force.  A bond partner must lie within the ghost zone of the
 *                strip that holds the particle: scenes whose bonds stretch across a strip boundary
 *                want 4 to 6 (a partner beyond it raises error bit 16).
 *   id_capacity  1 + the largest particle id of the WHOLE scene (all strips): migrants and ghosts
 *                carry ids this rank never loaded, and every id-indexed table must cover them.
 *                0 = grow with the ids this rank sees (single-strip use, tests that load every
 *                particle into every rank).
 *   bonds        non-zero if ANY rank will set bonds: the halo then also carries particle ids and a
 *                migrant takes its bond row (first 6 partners) along.  Must agree across ranks.
 * While decomposed: ptoy_set_bonds / ptoy_set_frozen take the entries of this rank's own
 * particles (more is allowed); bonds should be set before stepping (a later edit re-uploads the
 * rank's host list and forgets rows that arrived with migrants); ptoy_get_bonds returns the
 * rank's list without the CheckBonds pruning (whether a partner is gone is not a local question;
 * the device skips dead partners either way); the block grid cannot grow. */
/* One process per GPU: rank 0 creates the 128-byte id, every rank connects with it (collective). */
int ptoy_nccl_unique_id(void* out128);
int ptoy_nccl_connect(ptoy_engine* h, const void* uid128);
/* Several strips in ONE process on ONE device (tests / development): the group steps them in
 * lock step with device-to-device copies in place of NCCL.  Destroy the group before its engines. */
typedef struct ptoy_group ptoy_group;
int ptoy_group_create(ptoy_group** out, ptoy_engine** engines, int n);
int ptoy_group_step_n(ptoy_group* g, int n);
void ptoy_group_destroy(ptoy_group* g);
/* Sticky device-side error bits:
 *    2  arriving migrants exceed the particle capacity        (PTOY_ERR_CAPACITY)
 *    4  a boundary slice exceeds the ghost capacity           (PTOY_ERR_CAPACITY)
 *    8  a migrant outbox overflowed (PTOY_MIGRANT_CAP)        (PTOY_ERR_CAPACITY)
 *   16  a bond partner is neither local nor in a ghost row    (PTOY_ERR_DIST)
 *   32  a migrant has more than 6 bonds: the rest of its row does not travel   (PTOY_ERR_CAPACITY)
 *   64  an id beyond id_capacity arrived from another strip   (PTOY_ERR_DIST)
 * Once a bit is set the simulation state is no longer trustworthy: ptoy_synchronize returns the
 * error shown, and so do ptoy_step / ptoy_step_n as soon as the flag has reached the host (they
 * do not wait for the device). */
int ptoy_error_flags(ptoy_engine* h, int* flags);

/* ---- measurement hooks -------------------------------------------------------------------- */

/* Device time in ms of the kernels of the most recent ptoy_step_n(.., n), summed per
 * kernel kind over its launches, measured with CUDA events on the engine's stream.
 * kinds: 0 stage1, 1 stage2, 2 scan, 3 scatter, 4 fixup, 5 other.  Enable first. */
int ptoy_enable_kernel_timing(ptoy_engine* h, int enable);
int ptoy_get_kernel_times(ptoy_engine* h, double* ms6, int64_t* launches6);
/* Number of kernel launches issued by this engine since creation. */
int ptoy_launch_count(ptoy_engine* h, int64_t* n);
/* cudaStream_t of the engine (as void*), so callers can record events on it. */
int ptoy_stream(ptoy_engine* h, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* PTOY_B200_H_ */

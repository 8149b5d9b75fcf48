/* sphb200.h — C ABI of libsphb200.so: the per-step SPH hot path of
 * f1nalspace/nbodysimulation_experiment's Demo 4 solver, rebuilt for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  Every entry point replaces one member of the reference's
 * plugin interface `class BaseSimulation` (NBodySimulation/base.h:8-39) as implemented by
 * `Demo4::ParticleSimulation` (NBodySimulation/demo4.h:137-226, demo4.cpp); the citation next
 * to each declaration names the member it stands in for.  INTEGRATION.md shows the
 * `BaseSimulation` subclass a maintainer of the reference would add on top of these calls.
 *
 * Conventions
 *   - every function returns 0 (SPH_OK) or a negative SphStatus; sph_last_error() has the text.
 *   - plain pointers and sizes only; all pointers are HOST pointers unless named `dev_*`.
 *   - one caller thread per handle (the reference calls everything from its main thread).
 *   - sph_step() enqueues one Update() on the simulation's CUDA stream and returns; every
 *     sph_read_* / sph_get_stats / sph_sync call waits for the work it needs.
 *   - there is NO CPU fallback: without a CUDA device sph_create() fails with SPH_ERR_CUDA.
 */
#ifndef SPHB200_H
#define SPHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB200_ABI_VERSION 1

typedef struct SphSim *SphHandle;

typedef enum SphStatus {
	SPH_OK = 0,
	SPH_ERR_INVALID = -1,  /* bad argument / bad handle */
	SPH_ERR_CAPACITY = -2, /* particle, body, emitter, halo or cell capacity exceeded (the reference asserts: demo4.cpp:45,87,156,199) */
	SPH_ERR_CUDA = -3,     /* CUDA runtime error, or no device */
	SPH_ERR_COMM = -4,     /* multi-GPU exchange failed */
	SPH_ERR_STATE = -5     /* call not valid in the current state */
} SphStatus;

/* fp_mode: how the three pair loops (density, displacement, viscosity) round.
 *   EXACT: one IEEE-754 rounding per reference operation (no FMA contraction, correctly rounded
 *          sqrt and 1/x), i.e. the arithmetic of the reference's scalar SSE2 build — bit-for-bit
 *          equal to the CPU oracle's gather mode.
 *   FAST : FMA contraction and rsqrt in the pair loops; integrate, predict, cell keys, collisions
 *          and the velocity update stay exact in both modes. */
enum { SPH_FP_EXACT = 0, SPH_FP_FAST = 1 };

/* solver: how the two in-place pair sweeps of the reference (viscosity, demo4.cpp:223-237, and
 * pressure displacement, demo4.cpp:239-255) are parallelised.
 *   COLORED_GS: the reference's own in-place half-weight pair updates, swept race-free over nine
 *               cell colours (cx mod 3, cy mod 3), ascending id inside a cell.  Keeps the
 *               Gauss-Seidel self-damping the reference relies on; the default.
 *   GATHER    : every pair term evaluated from the pass's input state and summed per particle
 *               (Jacobi); cheaper to exchange across GPUs but unstable on stiff scenes unless
 *               under-relaxed (SphConfig.relaxation). */
enum { SPH_SOLVER_COLORED_GS = 0, SPH_SOLVER_GATHER = 1 };

enum {
	SPH_FLAG_PHASE_TIMING = 1u << 0, /* bracket every phase with CUDA events and fill SphStats.time_* (sph.h:131-141) */
	SPH_FLAG_NO_GRAPHS = 1u << 1,    /* launch every kernel of a step individually instead of replaying a CUDA graph */
	SPH_FLAG_SWEEP_TEAM = 1u << 2,   /* coloured sweeps: always one thread block per cell (default: chosen by particle count) */
	SPH_FLAG_SWEEP_WARP = 1u << 3,   /* coloured sweeps: nine launches, one warp per cell */
	SPH_FLAG_SWEEP_FLOW = 1u << 4,   /* coloured sweeps: one launch for all colours, persistent warps + per-cell dependency flags
	                                    (default for >= 131072 particles); all three kernels give identical bits */
	SPH_FLAG_EXCHANGE_NCCL = 1u << 5 /* strips: ship the halo records as ncclSend/ncclRecv messages instead of storing them straight
	                                    into the neighbour's memory (the fallback sph_comm_init picks by itself when CUDA IPC or peer
	                                    access is not available) */
};

/* Runtime replacement for the compile-time world of sph.h:18-72. */
typedef struct SphConfig {
	uint32_t struct_size;    /* sizeof(SphConfig), for ABI checks */
	float domain_width;      /* kSPHBoundaryWidth  (sph.h:19); the domain is centred on the origin */
	float domain_height;     /* kSPHBoundaryHeight (sph.h:20) */
	float cell_size;         /* kSPHGridCellSize   (sph.h:60); grid = (int)(W/cell) x (int)(H/cell), sph.h:61-62 */
	uint64_t max_particles;  /* kSPHMaxParticleCount (sph.h:70); per rank when world_size > 1 */
	int32_t device;          /* CUDA device ordinal */
	int32_t fp_mode;         /* SPH_FP_EXACT | SPH_FP_FAST */
	uint32_t flags;          /* SPH_FLAG_* */
	float relaxation;        /* GATHER only: omega of the displacement pass, x += omega*dx (1 = plain Jacobi) */
	int32_t solver;          /* SPH_SOLVER_COLORED_GS | SPH_SOLVER_GATHER */
	uint32_t sweep_capacity; /* COLORED_GS: candidates of one 3x3 block staged in shared memory (0 = 512) */
	/* y-strip decomposition; world_size = 1 for a single GPU */
	int32_t rank;
	int32_t world_size;
	uint64_t halo_capacity;  /* particles per direction per step the exchange buffers hold (0 = max_particles/4) */
	int32_t halo_rows;       /* ghost rows kept on each side of the strip (0 = 7) */
	int32_t reserved0;
} SphConfig;

/* SPHParameters, sph.h:77-87, same field order. */
typedef struct SphParams {
	float kernel_height, cell_size, particle_spacing, inv_kernel_height, rest_density, stiffness,
	    near_stiffness, linear_viscosity, quadratic_viscosity;
} SphParams;

/* SPHStatistics, sph.h:125-141.  Neighbour counts are the per-step min/max candidate-list length
 * (demo4.cpp:369-376); cell counts are the min over occupied cells / max occupancy seen since
 * sph_reset_stats (the reference tracks them per insert/remove event, demo4.cpp:51-53,73-75). */
typedef struct SphStats {
	uint64_t min_particle_neighbor_count, max_particle_neighbor_count;
	uint64_t min_cell_particle_count, max_cell_particle_count;
	float time_emitters, time_integration, time_viscosity_forces, time_predict, time_update_grid,
	    time_neighbor_search, time_density_and_pressure, time_delta_positions, time_collisions;
	uint64_t steps;          /* Update() calls since creation */
	uint64_t pair_candidates; /* sum over particles of the candidate-list length in the last step */
} SphStats;

/* ---- lifecycle ------------------------------------------------------------------------ */
/* Demo4::ParticleSimulation::ParticleSimulation(), demo4.cpp:14-26 */
int sph_config_default(SphConfig *cfg); /* the reference's world: 10 x 5.625, cell 0.3, 10 000 particles */
int sph_create(const SphConfig *cfg, SphHandle *out);
/* ~ParticleSimulation(), demo4.cpp:28-35 (BaseSimulation has no virtual dtor: call this explicitly) */
int sph_destroy(SphHandle h);
int sph_abi_version(void);
int sph_last_error(SphHandle h, char *buf, size_t n); /* h may be NULL: error of the last failed sph_create */

/* ---- parameters ----------------------------------------------------------------------- */
int sph_set_params(SphHandle h, const SphParams *p);   /* SetParams, demo4.h:223 (+ copy-ctor rule sph.h:100-110) */
int sph_get_params(SphHandle h, SphParams *out);       /* GetParams, demo4.h:217 */
int sph_set_gravity(SphHandle h, float gx, float gy);  /* SetGravity, demo4.h:213 */
int sph_add_external_force(SphHandle h, float fx, float fy); /* AddExternalForces, demo4.h:189 */
int sph_clear_external_force(SphHandle h);             /* ClearExternalForce, demo4.h:192 */
int sph_set_relaxation(SphHandle h, float omega);
int sph_grid_dims(SphHandle h, int32_t *gx, int32_t *gy); /* kSPHGridCountX/Y, sph.h:61-62 */

/* ---- bodies (demo4.cpp:78-123) -------------------------------------------------------- */
int sph_clear_bodies(SphHandle h);                                        /* ClearBodies */
int sph_add_plane(SphHandle h, float nx, float ny, float distance);       /* AddPlane */
int sph_add_circle(SphHandle h, float x, float y, float radius);          /* AddCircle */
int sph_add_segment(SphHandle h, float ax, float ay, float bx, float by); /* AddLineSegment */
int sph_add_polygon(SphHandle h, size_t vertex_count, const float *xy);   /* AddPolygon (<= 8 vertices, sph.h:161) */
int sph_body_count(SphHandle h, size_t *out);

/* ---- particles, volumes, emitters (demo4.cpp:125-181, 257-284) ------------------------ */
int sph_clear_particles(SphHandle h); /* ClearParticles */
int sph_clear_emitters(SphHandle h);  /* ClearEmitters */
/* AddParticle in bulk: n positions and initial accelerations ("force", consumed by the first
 * integrate, demo4.cpp:146,306-308); acc_xy may be NULL (zero).  Indices are creation order. */
/* On a y-strip (world_size > 1) every rank is given the WHOLE list, in the same order, and keeps the particles whose
 * cell row it owns; creation indices are the same on every rank.  sph_add_volume, the emitters (every rank runs the
 * same emitter clocks and libc rand() sequence; sph_load_scenario seeds it) and sph_load_scenario go through here. */
int sph_add_particles(SphHandle h, size_t n, const float *pos_xy, const float *acc_xy, uint64_t *first_index);
/* AddVolume, demo4.cpp:169-181: row-major block with the libc rand() jitter of vecmath.h:317-322 */
int sph_add_volume(SphHandle h, float cx, float cy, float fx, float fy, int count_x, int count_y, float spacing);
/* Large synthetic scenes: same lattice, jitter from a counter-based hash of (seed, index) generated
 * on the device, ids first_id + row-major index.  With world_size > 1 every rank calls this with the
 * same arguments and keeps the particles of its own strip. */
int sph_add_volume_hashed(SphHandle h, float cx, float cy, float fx, float fy, int64_t count_x, int64_t count_y,
                          float spacing, uint64_t seed);
/* AddEmitter, demo4.cpp:155-167; the emitter clock and rand() stay on the host (demo4.cpp:257-284) */
int sph_add_emitter(SphHandle h, float px, float py, float dx, float dy, float radius, float speed, float rate, float duration);
/* The reference's 8 built-in scenes (SPHScenarios[], sph.h:315-437) set up by the call sequence of
 * DemoApplication::LoadScenario (app.cpp:477-534).  seed >= 0 calls srand(seed) first; seed < 0 keeps
 * libc's rand() state (the reference never seeds: glibc starts at seed 1). */
int sph_scenario_count(void);
const char *sph_scenario_name(int idx);
int sph_load_scenario(SphHandle h, int idx, int seed);
int sph_particle_count(SphHandle h, uint64_t *out); /* GetParticleCount; global count when world_size > 1 */
int sph_local_particle_count(SphHandle h, uint64_t *out); /* particles this rank owns */

/* ---- the hot path --------------------------------------------------------------------- */
/* Update(deltaTime), demo4.cpp:286-451: emitters, integrate, viscosity (previous step's grid),
 * predict, grid rebuild, density/pressure, displacement, collisions, velocity. */
int sph_step(SphHandle h, float dt);
int sph_sync(SphHandle h);
/* one phase at a time, for per-pass parity against the oracle from an injected state */
enum {
	SPH_PASS_INTEGRATE = 1, SPH_PASS_VISCOSITY = 2, SPH_PASS_PREDICT = 3, SPH_PASS_GRID = 4,
	SPH_PASS_DENSITY = 5, SPH_PASS_DELTA = 6, SPH_PASS_COLLIDE = 7, SPH_PASS_VELOCITY = 8
};
int sph_run_pass(SphHandle h, int pass, float dt);

/* ---- statistics (sph.h:125-150) ------------------------------------------------------- */
int sph_reset_stats(SphHandle h);              /* ResetStats */
int sph_get_stats(SphHandle h, SphStats *out); /* GetStats */

/* ---- readback / injection ------------------------------------------------------------- */
/* Record layout = Demo4::ParticleData (demo4.h:81-99): cur, prev, acc, vel, density, nearDensity,
 * pressure, nearPressure = 12 floats.  Record i is the i-th created particle, written at
 * dst + i*stride (stride >= 48).  With world_size > 1 only this rank's particles are written. */
int sph_read_particles(SphHandle h, void *dst, size_t stride);
/* inject cur/prev/acc/vel (+ rho,P fields) and re-file the grid; on a y-strip every rank is given the whole state
 * (sph_particle_count rows, creation order) and keeps the rows of its window, ghost rows included */
int sph_write_particles(SphHandle h, const void *src, size_t stride);
/* Render()'s particle section, demo4.cpp:520-531: positions at pos_stride (>= 8) and the colours of
 * SPHGetParticleColor (sph.h:683-695) at color_stride (>= 16), both in creation order.  The state is
 * snapshotted on the device and copied on a second stream: with pinned buffers (sph_host_alloc) the
 * call returns at once and the copy overlaps the next sph_step; the data is valid after
 * sph_wait_render (or sph_sync).  The reference's contract - pointers valid until the frame is
 * drawn (render.h:342-353) - is met by calling sph_wait_render before drawing. */
int sph_render_particles(SphHandle h, void *positions, size_t pos_stride, void *colors, size_t color_stride);
int sph_wait_render(SphHandle h);
int sph_read_cell_counts(SphHandle h, uint32_t *out);        /* Cell::count per cell, row-major (demo4.h:118-121) */
int sph_read_cell_of_particle(SphHandle h, int32_t *out_xy); /* ParticleIndex::cellIndex, creation order (demo4.h:112) */
/* the grid as the GPU holds it: particle ids in cell-sorted order (n entries) and the exclusive
 * prefix cell_start (cells+1 entries); the neighbour candidates of a particle are the id ranges of
 * its 3x3 block, dy outer, dx inner (demo4.cpp:183-206) */
int sph_read_sorted_ids(SphHandle h, uint32_t *ids);
int sph_read_cell_start(SphHandle h, uint32_t *cell_start);

/* pinned host memory for the readback path */
int sph_host_alloc(void **out, size_t bytes);
int sph_host_free(void *p);
/* the cudaStream_t all work of this handle is enqueued on (for event timing by the caller) */
int sph_get_stream(SphHandle h, void **stream_out);
/* wall-clock-free timing helpers on that stream: elapsed ms between two marks */
int sph_mark(SphHandle h, int slot);                         /* slot 0..7 */
int sph_elapsed_ms(SphHandle h, int from_slot, int to_slot, float *ms); /* waits for to_slot */
/* average device time per phase (ms) over the steps since the last sph_reset_stats; needs
 * SPH_FLAG_PHASE_TIMING.  Order: integrate, viscosity, predict+key, scan, reorder, density, delta,
 * collide+velocity, exchange. */
#define SPH_NUM_PHASES 9
int sph_get_phase_ms(SphHandle h, float out[SPH_NUM_PHASES], uint64_t *steps);

/* ---- multi-GPU plumbing (one process per GPU) --------------------------- */
/* NCCL bootstrap: rank 0 fills a 128-byte id, the host broadcasts it by any channel (the Python
 * host uses torch.distributed), every rank then calls sph_comm_init before adding particles. */
int sph_comm_unique_id(uint8_t id128[128]);
int sph_comm_init(SphHandle h, const uint8_t id128[128]);
/* sph_comm_init also maps the two neighbours' halo mailboxes into this process (CUDA IPC): from then on
 * predict_key_kernel stores a strip's migrating and halo particles straight into the neighbour's HBM over NVLink, a
 * device-side sequence number tells the neighbour when they are complete, and a step is one CUDA graph with no host
 * or NCCL call inside.  All NCCL connections are opened here, so the first sph_step costs what every step costs
 * (the reference's contract: every Update is one frame, app.cpp:228-236).
 *
 * The same for strips that all live in THIS process - one host thread driving several GPUs like the reference's
 * single-threaded app loop would, or several strips on one GPU (tests): create the handles with rank = 0..n-1,
 * world_size = n, wire them with sph_comm_init_local and step them together with sph_step_group (sph_step refuses
 * such a handle: the group call enqueues every strip's sends before any strip's wait). */
int sph_comm_init_local(SphHandle *handles, int32_t n);
int sph_step_group(SphHandle *handles, int32_t n, float dt);
/* rows [row_begin, row_end) of the grid this rank owns (even split of occupied rows at init) */
int sph_set_strip(SphHandle h, int32_t row_begin, int32_t row_end);
int sph_get_strip(SphHandle h, int32_t *row_begin, int32_t *row_end);
/* Periodic re-balancing (off by default): every `every_steps` steps the ranks all-reduce the particles
 * per grid row, plan the same new split (prefix over the histogram, each boundary moving at most `max_shift_rows`
 * rows, 0 = 2) and the rows that change hands travel with that step's neighbour exchange.  Call before adding
 * particles: the cell arrays are re-sized for the whole grid so that a strip can move without reallocating.
 * Results stay bit-identical to one GPU.  every_steps = 0 switches it off. */
int sph_set_rebalance(SphHandle h, int32_t every_steps, int32_t max_shift_rows);
/* The planner on its own (host arithmetic, no device): row_counts[grid_y], old_bounds[world + 1] (= 0, B_1, ...,
 * grid_y) -> new_bounds[world + 1].  Exposed for tests and for hosts that want to inspect the split. */
int sph_plan_strip_bounds(const uint32_t *row_counts, int32_t grid_y, const int32_t *old_bounds, int32_t world, int32_t halo_rows,
                          int32_t max_shift_rows, int32_t *new_bounds);
/* The particles this rank owns, compacted in arbitrary order: creation ids, ParticleData records
 * (as sph_read_particles), and/or Render()'s positions and colours.  Any output pointer may be NULL;
 * *count receives the number of owned particles.  Works for world_size = 1 too. */
int sph_read_owned(SphHandle h, uint32_t *ids, void *records, size_t record_stride, void *positions, size_t pos_stride,
                   void *colors, size_t color_stride, uint64_t *count);
/* The same readback for the renderer, overlapped like sph_render_particles (demo4.cpp:520-531 on a strip): the
 * snapshot is taken on the simulation's stream, the copies run on a second stream while the next sph_step
 * executes, and the host arrays are complete after sph_wait_render_owned, which also returns the count.  One frame
 * in flight; alternate two sets of (pinned) host arrays to overlap a frame's copy with the next step. */
int sph_render_owned(SphHandle h, uint32_t *ids, void *positions, size_t pos_stride, void *colors, size_t color_stride);
int sph_wait_render_owned(SphHandle h, uint64_t *count);

#ifdef __cplusplus
}
#endif
#endif /* SPHB200_H */

/* TEST INFRASTRUCTURE ONLY.  CPU restatement ("Oracle B") of the per-step SPH hot path of
 * f1nalspace/nbodysimulation_experiment's Demo 4 solver, with a runtime-sized domain.
 *
 * Who may use this: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
 * `--impl reference` legs — as the CHECKER or the timed CPU baseline, never as the product.
 * Nothing under nbodysimulation_experiment_b200/ links, loads or calls it.
 *
 * Parity pin: mode ORACLE_MODE_GS_INDEX, single-threaded, at the reference's constants is
 * bit-for-bit equal to the reference's own demo4.cpp compiled headless (oracle/_ref/libsphref.so,
 * see tests/test_oracle_vs_reference.py and the fixtures in tests/golden/ made by
 * tools/make_golden.py).  The reference itself ships no tests or golden vectors.
 *
 * Reference lines each function follows are cited in sph_oracle.c.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SphOracle SphOracle;

enum {
	ORACLE_MODE_GS_INDEX = 0, /* the reference's semantics: in-place half-weight pair updates in
	                             particle-index order over neighbour lists in the reference's own
	                             cell-storage order (demo4.cpp:223-255) */
	ORACLE_MODE_JACOBI = 1,   /* the deterministic per-particle gather the GPU implements: every
	                             pair term evaluated from the pass's input state; candidates walked
	                             in (dy, dx, ascending particle id) order */
	ORACLE_MODE_COLORED = 2,  /* the reference's in-place half-weight pair updates swept race-free:
	                             nine cell colours (cx mod 3, cy mod 3), ascending id inside a cell,
	                             32-lane evaluation of each particle's loop (see sph_oracle.c) */
	ORACLE_MODE_HYBRID = 3    /* viscosity as in JACOBI (gather), pressure displacement as in COLORED */
};

enum { ORACLE_BODY_PLANE = 1, ORACLE_BODY_CIRCLE = 2, ORACLE_BODY_SEGMENT = 3, ORACLE_BODY_POLYGON = 4 };

/* domain W x H centred on the origin; grid = (int)(W/cell) x (int)(H/cell) cells (sph.h:60-65) */
SphOracle *oracle_create(float width, float height, float cell_size);
void oracle_destroy(SphOracle *o);

void oracle_set_mode(SphOracle *o, int mode);
void oracle_set_threads(SphOracle *o, int threads); /* 1 = caller thread only (SetMultiThreading(false)) */
int oracle_get_threads(const SphOracle *o);
void oracle_set_relaxation(SphOracle *o, float omega); /* JACOBI only: x += omega * dx */

void oracle_reset_stats(SphOracle *o);
void oracle_clear_bodies(SphOracle *o);
void oracle_clear_particles(SphOracle *o);
void oracle_clear_emitters(SphOracle *o);
void oracle_set_params(SphOracle *o, const float p9[9]); /* SPHParameters, sph.h:77-87 order */
void oracle_get_params(const SphOracle *o, float p9[9]);
void oracle_set_gravity(SphOracle *o, float gx, float gy);
void oracle_get_gravity(const SphOracle *o, float g2[2]);
void oracle_add_external_force(SphOracle *o, float fx, float fy);
void oracle_clear_external_force(SphOracle *o);

void oracle_add_plane(SphOracle *o, float nx, float ny, float distance);
void oracle_add_circle(SphOracle *o, float x, float y, float radius);
void oracle_add_segment(SphOracle *o, float ax, float ay, float bx, float by);
void oracle_add_polygon(SphOracle *o, int n, const float *xy);
int oracle_body_count(const SphOracle *o);
void oracle_get_body(const SphOracle *o, int idx, int32_t *type, int32_t *nverts, float out16[16]);

uint64_t oracle_add_particle(SphOracle *o, float x, float y, float ax, float ay);
uint64_t oracle_add_particles(SphOracle *o, uint64_t n, const float *pos_xy, const float *acc_xy);
void oracle_add_volume(SphOracle *o, float cx, float cy, float fx, float fy, int nx, int ny, float spacing);
void oracle_add_emitter(SphOracle *o, float px, float py, float dx, float dy, float radius, float speed, float rate, float duration);

/* the reference's 8 built-in scenes (sph.h:315-437) through the LoadScenario call order
 * (app.cpp:477-534); seed < 0 keeps libc's rand() state */
int oracle_scenario_count(void);
const char *oracle_scenario_name(int idx);
void oracle_load_scenario(SphOracle *o, int idx, int seed);

void oracle_step(SphOracle *o, float dt);
double oracle_step_timed(SphOracle *o, float dt, int steps); /* wall seconds inside the steps */

/* single passes, for per-pass parity from an injected state */
void oracle_pass_update_grid(SphOracle *o);
void oracle_pass_neighbor_search(SphOracle *o);
void oracle_pass_density(SphOracle *o);
void oracle_pass_viscosity(SphOracle *o, float dt);
void oracle_pass_delta(SphOracle *o, float dt);
void oracle_pass_collide(SphOracle *o);

uint64_t oracle_particle_count(const SphOracle *o);
void oracle_grid_dims(const SphOracle *o, int32_t out2[2]);
void oracle_get_particles(const SphOracle *o, float *out12);   /* ParticleData order, 12 floats each */
void oracle_set_particles(SphOracle *o, const float *in12);    /* then re-files the grid */
void oracle_get_cell_of_particle(const SphOracle *o, int32_t *out_xy);
void oracle_get_cell_counts(const SphOracle *o, uint32_t *out);
uint32_t oracle_get_cell_members(const SphOracle *o, int cell, uint32_t *out);
void oracle_get_neighbor_counts(const SphOracle *o, uint32_t *out);
uint32_t oracle_get_neighbors(const SphOracle *o, uint64_t i, uint32_t *out);
void oracle_get_stats(const SphOracle *o, uint64_t counters4[4], float times9[9]);
void oracle_get_colors(const SphOracle *o, float *out4);

/* collision solvers of sph.h:514-681 on one point */
void oracle_solve_plane(float pxy[2], float nx, float ny, float d);
void oracle_solve_circle(float pxy[2], float cx, float cy, float r);
void oracle_solve_segment(float pxy[2], float ax, float ay, float bx, float by);
void oracle_solve_polygon(float pxy[2], int n, const float *xy);
void oracle_cell_index(const SphOracle *o, float x, float y, int32_t out2[2]);

#ifdef __cplusplus
}
#endif
#endif

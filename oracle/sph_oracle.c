/* TEST INFRASTRUCTURE ONLY — see sph_oracle.h for who may use this and how it is pinned.
 *
 * CPU restatement of Demo4::ParticleSimulation (reference: NBodySimulation/demo4.cpp, sph.h,
 * vecmath.h, threading.h, app.cpp:477-534).  Citations are file:line of /root/reference/NBodySimulation.
 * Every float expression keeps the reference's operation order; build with -ffp-contract=off.
 */
#define _GNU_SOURCE
#include "sph_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- constants (sph.h:18-72) ------------------------------------------------------------- */
static const float kParticleRadius = 0.05f;                       /* sph.h:35 */
#define kKernelHeight (6.0f * kParticleRadius)                    /* sph.h:36 */
#define kCollisionRadius kParticleRadius                          /* sph.h:38 */
static const float kRestDensity = 20.0f;                          /* sph.h:40 */
static const float kStiffness = 0.6f;                             /* sph.h:41 */
static const float kLinearViscosity = 0.5f;                       /* sph.h:43 */
static const float kQuadraticViscosity = 0.3f;                    /* sph.h:44 */
static const float kVolumeDistributionScale = 0.01f;              /* sph.h:51 */
#define kCollisionMargin (0.005f * 2.0f)                          /* sph.h:54 */
#define kMaxNeighborStat 1000u                                    /* sph.h:69 (stats seed, demo4.cpp:369) */
#define kMaxCellStat 500u                                         /* sph.h:68 (stats seed, sph.h:143-147) */
#define kMaxPolyVerts 8                                           /* sph.h:161 */
#define kMaxEmitters 8                                            /* sph.h:72 */

typedef struct { float x, y; } V2;

/* vecmath.h:233-260 */
static inline V2 v2(float x, float y) { V2 r = { x, y }; return r; }
static inline V2 v2_add(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 v2_sub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 v2_scale(V2 a, float s) { return v2(a.x * s, a.y * s); }
static inline V2 v2_neg(V2 a) { return v2(-a.x, -a.y); }
static inline float v2_dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }           /* :262 */
static inline float v2_len(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }          /* :267 */
static inline V2 v2_normalize(V2 a) {                                              /* :272-280 */
	float l = v2_len(a);
	if (l == 0) l = 1;
	float inv = 1.0f / l;
	return v2_scale(a, inv);
}
static inline float v2_distsq_quirk(V2 a, V2 b) {                                  /* :292-296 (sic) */
	float f = (b.x - a.x) * (b.y - a.y);
	return f * f;
}
static inline float lerpf(float a, float t, float b) { return (1.0f - t) * a + t * b; } /* :228 */

/* ParticleData, demo4.h:81-99 (48 bytes, same field order) */
typedef struct {
	V2 cur, prev, acc, vel;
	float rho, rhoNear, P, PNear;
} Particle;

typedef struct { int32_t cx, cy; uint32_t inCell; } ParticleIdx;   /* demo4.h:111-116 minus the list */
typedef struct { uint32_t *idx; uint32_t count, cap; } Cell;       /* demo4.h:118-121, growable */
typedef struct { int32_t type, nverts; float f[16]; } Body;        /* demo4.h:29-79 flattened */
typedef struct {                                                   /* demo4.h:123-133 */
	V2 position, direction;
	float radius, speed, rate, duration, elapsed, totalElapsed;
	int32_t active;
} Emitter;

/* ---- thread pool with the reference's task split (threading.h:111-129) -------------------- */
typedef void (*RangeFn)(SphOracle *o, int64_t start, int64_t end_inclusive, float dt);

typedef struct {
	pthread_t *threads;
	int nthreads;
	pthread_mutex_t mu;
	pthread_cond_t cvWork, cvDone;
	/* current batch */
	RangeFn fn;
	SphOracle *owner;
	float dt;
	int64_t itemCount, chunk;
	int64_t nextTask, taskCount, pending;
	int stop;
} Pool;

struct SphOracle {
	float params[9]; /* kernelHeight, cellSize, particleSpacing, invKernelHeight, restDensity,
	                    stiffness, nearStiffness, linearViscosity, quadraticViscosity */
	V2 gravity, extForce;
	float width, height, halfW, halfH, cellSize;
	int32_t gridX, gridY;
	int mode;
	float omega;

	Particle *p;
	ParticleIdx *pi;
	uint64_t n, cap;

	Cell *cells;
	/* neighbour lists as built by the last NeighborSearch (CSR); particles added afterwards
	 * have no list and are in nobody's list (demo4.cpp:148) */
	uint64_t *nbrOff; /* nbrN + 1 */
	uint32_t *nbr;
	uint64_t nbrN, nbrCap, nbrOffCap;

	Body *bodies;
	int nbodies, bodyCap;
	Emitter emitters[kMaxEmitters];
	int nemitters;

	uint64_t statMinNbr, statMaxNbr, statMinCell, statMaxCell;
	float times[9];

	/* jacobi scratch */
	V2 *scratch;
	uint64_t scratchCap;

	Pool pool;
	int threads;
};

static double now_ms(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int pool_take(Pool *pl, int64_t *s, int64_t *e) {
	if (pl->nextTask >= pl->taskCount) return 0;
	int64_t k = pl->nextTask++;
	*s = k * pl->chunk;
	int64_t last = *s + pl->chunk - 1;
	*e = last < pl->itemCount - 1 ? last : pl->itemCount - 1;
	return 1;
}

static void *pool_worker(void *arg) {
	Pool *pl = (Pool *)arg;
	pthread_mutex_lock(&pl->mu);
	for (;;) {
		int64_t s, e;
		while (!pl->stop && !pool_take(pl, &s, &e)) pthread_cond_wait(&pl->cvWork, &pl->mu);
		if (pl->stop) break;
		pthread_mutex_unlock(&pl->mu);
		pl->fn(pl->owner, s, e, pl->dt);
		pthread_mutex_lock(&pl->mu);
		if (--pl->pending == 0) pthread_cond_signal(&pl->cvDone);
	}
	pthread_mutex_unlock(&pl->mu);
	return NULL;
}

static void pool_start(Pool *pl, int nthreads) {
	memset(pl, 0, sizeof(*pl));
	pthread_mutex_init(&pl->mu, NULL);
	pthread_cond_init(&pl->cvWork, NULL);
	pthread_cond_init(&pl->cvDone, NULL);
	pl->nthreads = nthreads;
	pl->threads = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
	for (int i = 0; i < nthreads; ++i) pthread_create(&pl->threads[i], NULL, pool_worker, pl);
}

static void pool_stop(Pool *pl) {
	if (!pl->threads) return;
	pthread_mutex_lock(&pl->mu);
	pl->stop = 1;
	pthread_cond_broadcast(&pl->cvWork);
	pthread_mutex_unlock(&pl->mu);
	for (int i = 0; i < pl->nthreads; ++i) pthread_join(pl->threads[i], NULL);
	free(pl->threads);
	pthread_mutex_destroy(&pl->mu);
	pthread_cond_destroy(&pl->cvWork);
	pthread_cond_destroy(&pl->cvDone);
	memset(pl, 0, sizeof(*pl));
}

/* CreateTasks + WaitUntilDone: chunk = max(1, N / threads), inclusive ranges, the caller only waits */
static void run_ranges(SphOracle *o, RangeFn fn, float dt) {
	int64_t N = (int64_t)o->n;
	if (N == 0) return;
	if (o->threads <= 1) {
		fn(o, 0, N - 1, dt); /* demo4.cpp:323 etc. */
		return;
	}
	Pool *pl = &o->pool;
	pthread_mutex_lock(&pl->mu);
	pl->fn = fn;
	pl->owner = o;
	pl->dt = dt;
	pl->itemCount = N;
	pl->chunk = N / pl->nthreads > 1 ? N / pl->nthreads : 1;
	pl->taskCount = (N + pl->chunk - 1) / pl->chunk;
	pl->nextTask = 0;
	pl->pending = pl->taskCount;
	pthread_cond_broadcast(&pl->cvWork);
	while (pl->pending > 0) pthread_cond_wait(&pl->cvDone, &pl->mu);
	pl->taskCount = 0;
	pl->nextTask = 0;
	pthread_mutex_unlock(&pl->mu);
}

/* ---- grid (sph.h:439-463, demo4.cpp:37-76) ------------------------------------------------ */
static inline void cell_index(const SphOracle *o, V2 pos, int32_t *cx, int32_t *cy) {
	int x = (int)((pos.x + o->halfW) / o->cellSize); /* sph.h:451 */
	int y = (int)((pos.y + o->halfH) / o->cellSize); /* sph.h:452 */
	if (x < 0) x = 0;                                /* sph.h:459-460 */
	if (x > o->gridX - 1) x = o->gridX - 1;
	if (y < 0) y = 0;
	if (y > o->gridY - 1) y = o->gridY - 1;
	*cx = x;
	*cy = y;
}

static inline int in_grid(const SphOracle *o, int x, int y) { /* sph.h:439-442 */
	return x >= 0 && x < o->gridX && y >= 0 && y < o->gridY;
}

static void grid_insert(SphOracle *o, uint64_t i) { /* demo4.cpp:37-54 */
	int32_t cx, cy;
	cell_index(o, o->p[i].cur, &cx, &cy);
	Cell *c = &o->cells[(size_t)cy * o->gridX + cx];
	if (c->count == c->cap) {
		c->cap = c->cap ? c->cap * 2 : 8;
		c->idx = (uint32_t *)realloc(c->idx, c->cap * sizeof(uint32_t));
	}
	uint32_t at = c->count++;
	c->idx[at] = (uint32_t)i;
	o->pi[i].cx = cx;
	o->pi[i].cy = cy;
	o->pi[i].inCell = at;
	if (c->count < o->statMinCell) o->statMinCell = c->count;
	if (c->count > o->statMaxCell) o->statMaxCell = c->count;
}

static void grid_remove(SphOracle *o, uint64_t i) { /* demo4.cpp:56-76: swap with last, then shrink */
	Cell *c = &o->cells[(size_t)o->pi[i].cy * o->gridX + o->pi[i].cx];
	uint32_t at = o->pi[i].inCell, last = c->count - 1;
	if (at != last) {
		c->idx[at] = c->idx[last];
		c->idx[last] = (uint32_t)i;
		o->pi[c->idx[at]].inCell = at;
	}
	--c->count;
	if (c->count < o->statMinCell) o->statMinCell = c->count;
	if (c->count > o->statMaxCell) o->statMaxCell = c->count;
}

/* "Update grid" (demo4.cpp:342-356).  JACOBI mode re-files every particle from scratch in index
 * order instead, so each cell lists its members by ascending id — the canonical order the GPU's
 * per-cell id ranking produces. */
void oracle_pass_update_grid(SphOracle *o) {
	if (o->mode != ORACLE_MODE_GS_INDEX) {
		size_t ncell = (size_t)o->gridX * o->gridY;
		for (size_t c = 0; c < ncell; ++c) o->cells[c].count = 0;
		for (uint64_t i = 0; i < o->n; ++i) grid_insert(o, i);
		return;
	}
	for (uint64_t i = 0; i < o->n; ++i) {
		int32_t cx, cy;
		cell_index(o, o->p[i].cur, &cx, &cy);
		if (cx != o->pi[i].cx || cy != o->pi[i].cy) {
			grid_remove(o, i);
			grid_insert(o, i);
		}
	}
}

/* NeighborSearch (demo4.cpp:183-206): concatenate the 3x3 block's member lists, dy outer, dx inner,
 * no distance test, self included.  Two sweeps (count, fill) because the lists live in CSR. */
void oracle_pass_neighbor_search(SphOracle *o) {
	uint64_t N = o->n;
	if (N + 1 > o->nbrOffCap) {
		o->nbrOffCap = N + 1;
		o->nbrOff = (uint64_t *)realloc(o->nbrOff, o->nbrOffCap * sizeof(uint64_t));
	}
	uint64_t total = 0;
	for (uint64_t i = 0; i < N; ++i) {
		o->nbrOff[i] = total;
		for (int dy = -1; dy <= 1; ++dy)
			for (int dx = -1; dx <= 1; ++dx) {
				int x = o->pi[i].cx + dx, y = o->pi[i].cy + dy;
				if (in_grid(o, x, y)) total += o->cells[(size_t)y * o->gridX + x].count;
			}
	}
	o->nbrOff[N] = total;
	if (total > o->nbrCap) {
		o->nbrCap = total + total / 4;
		free(o->nbr);
		o->nbr = (uint32_t *)malloc(o->nbrCap * sizeof(uint32_t));
	}
	for (uint64_t i = 0; i < N; ++i) {
		uint32_t *dst = o->nbr + o->nbrOff[i];
		for (int dy = -1; dy <= 1; ++dy)
			for (int dx = -1; dx <= 1; ++dx) {
				int x = o->pi[i].cx + dx, y = o->pi[i].cy + dy;
				if (!in_grid(o, x, y)) continue;
				const Cell *c = &o->cells[(size_t)y * o->gridX + x];
				memcpy(dst, c->idx, c->count * sizeof(uint32_t));
				dst += c->count;
			}
	}
	o->nbrN = N;
	/* demo4.cpp:369-376 */
	o->statMinNbr = kMaxNeighborStat;
	o->statMaxNbr = 0;
	for (uint64_t i = 0; i < N; ++i) {
		uint64_t cnt = o->nbrOff[i + 1] - o->nbrOff[i];
		if (cnt < o->statMinNbr) o->statMinNbr = cnt;
		if (cnt > o->statMaxNbr) o->statMaxNbr = cnt;
	}
}

static inline uint64_t nbr_begin(const SphOracle *o, uint64_t i) { return i < o->nbrN ? o->nbrOff[i] : 0; }
static inline uint64_t nbr_end(const SphOracle *o, uint64_t i) { return i < o->nbrN ? o->nbrOff[i + 1] : 0; }

/* ---- pair kernels (sph.h:465-512) -------------------------------------------------------- */
static void density_range(SphOracle *o, int64_t s, int64_t e, float dt) { /* demo4.cpp:208-221 */
	(void)dt;
	const float h = o->params[0], invH = o->params[3];
	for (int64_t i = s; i <= e; ++i) {
		Particle *a = &o->p[i];
		float rho = 0, rhoNear = 0;
		for (uint64_t k = nbr_begin(o, (uint64_t)i); k < nbr_end(o, (uint64_t)i); ++k) {
			V2 rij = v2_sub(o->p[o->nbr[k]].cur, a->cur);     /* sph.h:466 */
			float r2 = v2_dot(rij, rij);
			if (r2 < (h * h)) {                               /* sph.h:470 */
				float r = sqrtf(r2);
				float term = 1.0f - r * invH;                 /* sph.h:472 */
				rho += (term * term);
				rhoNear += (term * term * term);
			}
		}
		a->rho = rho;
		a->rhoNear = rhoNear;
		a->P = o->params[5] * (rho - o->params[4]);           /* sph.h:479 */
		a->PNear = o->params[6] * rhoNear;                    /* sph.h:480 */
	}
}

/* reference semantics: in place on both particles of the pair (demo4.cpp:223-237) */
static void viscosity_range_gs(SphOracle *o, int64_t s, int64_t e, float dt) {
	const float h = o->params[0], invH = o->params[3], sigma = o->params[7], beta = o->params[8];
	for (int64_t i = s; i <= e; ++i) {
		Particle *a = &o->p[i];
		for (uint64_t k = nbr_begin(o, (uint64_t)i); k < nbr_end(o, (uint64_t)i); ++k) {
			Particle *b = &o->p[o->nbr[k]];
			V2 force = v2(0, 0);
			V2 rij = v2_sub(b->cur, a->cur);                  /* sph.h:500 */
			float r2 = v2_dot(rij, rij);
			if (r2 < (h * h)) {
				float r = sqrtf(r2);
				float q = r * invH;
				V2 nrm = v2_normalize(rij);
				float u = v2_dot(v2_sub(a->vel, b->vel), nrm);
				if (u > 0.0f) {
					float f = (1.0f - q) * (sigma * u + beta * (u * u)); /* sph.h:508 */
					force = v2(f * nrm.x, f * nrm.y);
				}
			}
			V2 half = v2_scale(v2_scale(force, 0.5f), dt);    /* demo4.cpp:233: force * 0.5f * deltaTime */
			a->vel = v2_sub(a->vel, half);
			b->vel = v2_add(half, b->vel);                    /* operator+= is b + a (vecmath.h:249) */
		}
	}
}

/* gather form: v_i' = v_i - dt * sum_j F_ij, every F from the pass's input velocities.  Each
 * unordered pair is visited from both ends in the reference with half weight and F_ji = -F_ij
 * bit-for-bit, so the two halves fold into one full-weight term. */
static void viscosity_range_jacobi(SphOracle *o, int64_t s, int64_t e, float dt) {
	const float h = o->params[0], invH = o->params[3], sigma = o->params[7], beta = o->params[8];
	for (int64_t i = s; i <= e; ++i) {
		const Particle *a = &o->p[i];
		V2 vnew = a->vel;
		for (uint64_t k = nbr_begin(o, (uint64_t)i); k < nbr_end(o, (uint64_t)i); ++k) {
			const Particle *b = &o->p[o->nbr[k]];
			V2 rij = v2_sub(b->cur, a->cur);
			float r2 = v2_dot(rij, rij);
			if (r2 < (h * h)) {
				float r = sqrtf(r2);
				float q = r * invH;
				V2 nrm = v2_normalize(rij);
				float u = v2_dot(v2_sub(a->vel, b->vel), nrm);
				if (u > 0.0f) {
					float f = (1.0f - q) * (sigma * u + beta * (u * u));
					float fdt = f * dt;
					vnew.x = vnew.x - fdt * nrm.x;
					vnew.y = vnew.y - fdt * nrm.y;
				}
			}
		}
		o->scratch[i] = vnew;
	}
}

/* reference semantics (demo4.cpp:239-255) */
static void delta_range_gs(SphOracle *o, int64_t s, int64_t e, float dt) {
	const float h = o->params[0], invH = o->params[3];
	for (int64_t i = s; i <= e; ++i) {
		Particle *a = &o->p[i];
		V2 dx = v2(0, 0);
		for (uint64_t k = nbr_begin(o, (uint64_t)i); k < nbr_end(o, (uint64_t)i); ++k) {
			Particle *b = &o->p[o->nbr[k]];
			V2 delta = v2(0, 0);
			V2 rij = v2_sub(b->cur, a->cur);                  /* sph.h:486 */
			float r2 = v2_dot(rij, rij);
			if (r2 < (h * h)) {
				float r = sqrtf(r2);
				V2 nrm = v2_normalize(rij);
				float term = 1.0f - r * invH;
				float d = (dt * dt) * (a->P * term + a->PNear * (term * term)); /* sph.h:492 */
				delta = v2(d * nrm.x, d * nrm.y);
			}
			V2 half = v2_scale(delta, 0.5f);
			b->cur = v2_add(half, b->cur);                    /* demo4.cpp:250 */
			dx = v2_sub(dx, half);                            /* demo4.cpp:251 */
		}
		a->cur = v2_add(dx, a->cur);                          /* demo4.cpp:253 */
	}
}

/* gather form: dx_i = -(dt^2/2) * sum_j [(P_i+P_j) t + (Pn_i+Pn_j) t^2] n_ij, all from the
 * pass's input positions; x_i' = x_i + omega * dx_i */
static void delta_range_jacobi(SphOracle *o, int64_t s, int64_t e, float dt) {
	const float h = o->params[0], invH = o->params[3];
	const float halfDt2 = (dt * dt) * 0.5f;
	for (int64_t i = s; i <= e; ++i) {
		const Particle *a = &o->p[i];
		float dxx = 0.0f, dxy = 0.0f;
		for (uint64_t k = nbr_begin(o, (uint64_t)i); k < nbr_end(o, (uint64_t)i); ++k) {
			const Particle *b = &o->p[o->nbr[k]];
			V2 rij = v2_sub(b->cur, a->cur);
			float r2 = v2_dot(rij, rij);
			if (r2 < (h * h)) {
				float r = sqrtf(r2);
				V2 nrm = v2_normalize(rij);
				float term = 1.0f - r * invH;
				float w = halfDt2 * ((a->P + b->P) * term + (a->PNear + b->PNear) * (term * term));
				dxx = dxx - w * nrm.x;
				dxy = dxy - w * nrm.y;
			}
		}
		o->scratch[i] = v2(a->cur.x + o->omega * dxx, a->cur.y + o->omega * dxy);
	}
}

/* ---- 9-colour cell Gauss-Seidel (ORACLE_MODE_COLORED) ---------------------------------------
 * The reference's in-place sweeps (demo4.cpp:223-255) visit particles in index order, which a GPU
 * cannot do in parallel, and its own multithreaded mode visits them in a racy order.  Cells whose
 * (cx mod 3, cy mod 3) agree have disjoint 3x3 footprints, so sweeping the nine colours one after
 * another, cells of a colour in any order, particles of a cell by ascending id, is a legitimate
 * race-free in-place sweep.  Inside one particle's loop the pair terms are evaluated by 32 "lanes"
 * (the n-th candidate within h belongs to lane n mod 32) from the particle's state at loop entry; partners are
 * updated immediately, the particle's own change is summed per lane, combined by a 5-stage
 * butterfly and applied at the end - exactly what the CUDA kernel does, so results are bit-equal. */
static float butterfly_sum(float part[32]) {
	for (int o = 16; o; o >>= 1) {
		float next[32];
		for (int l = 0; l < 32; ++l) next[l] = part[l] + part[l ^ o];
		memcpy(part, next, sizeof(next));
	}
	return part[0];
}

static inline int color_of(const SphOracle *o, uint64_t i) { return (o->pi[i].cy % 3) * 3 + (o->pi[i].cx % 3); }

static void delta_colored(SphOracle *o, float dt) {
	const float h = o->params[0], invH = o->params[3];
	for (int color = 0; color < 9; ++color)
		for (uint64_t i = 0; i < o->nbrN; ++i) {
			if (color_of(o, i) != color) continue;
			Particle *a = &o->p[i];
			const V2 xi = a->cur;
			float px[32], py[32];
			memset(px, 0, sizeof(px));
			memset(py, 0, sizeof(py));
			uint64_t b0 = nbr_begin(o, i), e0 = nbr_end(o, i);
			uint32_t inRange = 0; /* the n-th candidate within h is evaluated by lane n mod 32 */
			for (uint64_t k = b0; k < e0; ++k) {
				Particle *b = &o->p[o->nbr[k]];
				V2 rij = v2_sub(b->cur, xi);
				float r2 = v2_dot(rij, rij);
				if (r2 < (h * h)) {
					int lane = (int)(inRange++ & 31u);
					float r = sqrtf(r2);
					V2 nrm = v2_normalize(rij);
					float term = 1.0f - r * invH;
					float d = (dt * dt) * (a->P * term + a->PNear * (term * term)); /* sph.h:492 */
					V2 half = v2(d * nrm.x * 0.5f, d * nrm.y * 0.5f);
					b->cur = v2_add(half, b->cur);  /* demo4.cpp:250 */
					px[lane] = px[lane] - half.x;   /* demo4.cpp:251 */
					py[lane] = py[lane] - half.y;
				}
			}
			float dx = butterfly_sum(px), dy = butterfly_sum(py);
			a->cur = v2_add(v2(dx, dy), a->cur);    /* demo4.cpp:253 */
		}
}

static void viscosity_colored(SphOracle *o, float dt) {
	const float h = o->params[0], invH = o->params[3], sigma = o->params[7], beta = o->params[8];
	for (int color = 0; color < 9; ++color)
		for (uint64_t i = 0; i < o->nbrN; ++i) {
			if (color_of(o, i) != color) continue;
			Particle *a = &o->p[i];
			const V2 xi = a->cur, vi = a->vel;
			float px[32], py[32];
			memset(px, 0, sizeof(px));
			memset(py, 0, sizeof(py));
			uint64_t b0 = nbr_begin(o, i), e0 = nbr_end(o, i);
			uint32_t inRange = 0;
			for (uint64_t k = b0; k < e0; ++k) {
				Particle *b = &o->p[o->nbr[k]];
				V2 rij = v2_sub(b->cur, xi);
				float r2 = v2_dot(rij, rij);
				if (r2 < (h * h)) {
					int lane = (int)(inRange++ & 31u);
					float r = sqrtf(r2);
					float q = r * invH;
					V2 nrm = v2_normalize(rij);
					float u = v2_dot(v2_sub(vi, b->vel), nrm);
					if (u > 0.0f) {
						float f = (1.0f - q) * (sigma * u + beta * (u * u)); /* sph.h:508 */
						V2 half = v2((f * nrm.x * 0.5f) * dt, (f * nrm.y * 0.5f) * dt); /* demo4.cpp:233 */
						b->vel = v2_add(half, b->vel);
						px[lane] = px[lane] - half.x;
						py[lane] = py[lane] - half.y;
					}
				}
			}
			float dx = butterfly_sum(px), dy = butterfly_sum(py);
			a->vel = v2_add(v2(dx, dy), a->vel);
		}
}

static void ensure_scratch(SphOracle *o) {
	if (o->scratchCap < o->n) {
		o->scratchCap = o->n;
		o->scratch = (V2 *)realloc(o->scratch, o->scratchCap * sizeof(V2));
	}
}

void oracle_pass_density(SphOracle *o) { run_ranges(o, density_range, 0.0f); }

void oracle_pass_viscosity(SphOracle *o, float dt) {
	if (o->mode == ORACLE_MODE_COLORED) {
		viscosity_colored(o, dt);
	} else if (o->mode == ORACLE_MODE_JACOBI || o->mode == ORACLE_MODE_HYBRID) {
		ensure_scratch(o);
		run_ranges(o, viscosity_range_jacobi, dt);
		for (uint64_t i = 0; i < o->n; ++i) o->p[i].vel = o->scratch[i];
	} else {
		run_ranges(o, viscosity_range_gs, dt);
	}
}

void oracle_pass_delta(SphOracle *o, float dt) {
	if (o->mode == ORACLE_MODE_COLORED || o->mode == ORACLE_MODE_HYBRID) {
		delta_colored(o, dt);
	} else if (o->mode == ORACLE_MODE_JACOBI) {
		ensure_scratch(o);
		run_ranges(o, delta_range_jacobi, dt);
		for (uint64_t i = 0; i < o->n; ++i) o->p[i].cur = o->scratch[i];
	} else {
		run_ranges(o, delta_range_gs, dt);
	}
}

/* ---- collisions (sph.h:514-681) ---------------------------------------------------------- */
static V2 solve_plane(V2 pos, V2 normal, float distance) { /* sph.h:514-524 */
	V2 p = v2_scale(normal, distance);
	V2 delta = v2_sub(pos, p);
	float proj = v2_dot(delta, normal);
	if (proj <= kCollisionRadius) {
		float penetration = kCollisionRadius - proj;
		pos = v2_add(v2_scale(normal, penetration), pos);
	}
	return pos;
}

static V2 solve_circle(V2 pos, V2 c, float radius) { /* sph.h:526-542 */
	float both = radius + kCollisionRadius;
	V2 d = v2_sub(pos, c);
	float d2 = v2_dot(d, d);
	if (d2 <= both * both) {
		if (fabsf(d2) > 0) { /* a particle exactly at the centre is left where it is */
			float dist = sqrtf(d2);
			V2 normal = v2_scale(d, 1.0f / dist);
			float penetration = both - dist;
			pos = v2_add(v2_scale(normal, penetration), pos);
		}
	}
	return pos;
}

static V2 solve_segment(V2 pos, V2 a, V2 b) { /* sph.h:544-598 */
	float both = kCollisionMargin + kCollisionRadius;
	V2 e = v2_sub(b, a);
	float u = v2_dot(e, v2_sub(b, pos));
	float v = v2_dot(e, v2_sub(pos, a));
	V2 closest, normal;
	if (v <= 0.0f) { /* region A */
		closest = a;
		V2 d = v2_sub(pos, closest);
		if (v2_dot(d, d) > both * both) return pos;
		normal = v2_normalize(v2_sub(pos, closest));
	} else if (u <= 0.0f) { /* region B */
		closest = b;
		V2 d = v2_sub(pos, closest);
		if (v2_dot(d, d) > both * both) return pos;
		normal = v2_normalize(v2_sub(pos, closest));
	} else { /* region AB */
		float den = v2_dot(e, e);
		closest = v2_scale(v2_add(v2_scale(a, u), v2_scale(b, v)), 1.0f / den);
		V2 d = v2_sub(pos, closest);
		if (v2_dot(d, d) > both * both) return pos;
		normal = v2(-e.y, e.x);
		if (v2_dot(normal, v2_sub(pos, a)) < 0.0f) normal = v2_neg(normal);
		normal = v2_normalize(normal);
	}
	V2 dp = v2_sub(pos, closest);
	float distance = v2_dot(normal, dp);
	float penetration = both - distance;
	return v2_add(v2_scale(normal, penetration), pos);
}

static int find_mtv_circle_polygon(V2 c, int n, const V2 *verts, V2 *mtv) { /* sph.h:600-672 */
	int edge = 0;
	V2 normal = v2(0, 0);
	float separation = -FLT_MAX;
	float radius = kCollisionMargin + kCollisionRadius;
	for (int i = 0; i < n; ++i) {
		V2 a = verts[i], b = verts[(i + 1) % n];
		V2 eab = v2_sub(b, a);
		V2 nn = v2_normalize(v2(1.0f * eab.y, -1.0f * eab.x)); /* Vec2Cross(b - a, 1.0f), vecmath.h:299 */
		float s = v2_dot(nn, v2_sub(c, a));
		if (s > radius) return 0;
		if (s > separation) {
			normal = nn;
			separation = s;
			edge = i;
		}
	}
	V2 v1 = verts[edge], v2_ = verts[(edge + 1) % n];
	if (separation < FLT_EPSILON) { /* centre inside */
		float penetration = radius - separation;
		*mtv = v2_scale(normal, penetration);
		return 1;
	}
	float u1 = v2_dot(v2_sub(c, v1), v2_sub(v2_, v1));
	float u2 = v2_dot(v2_sub(c, v2_), v2_sub(v1, v2_));
	if (u1 <= 0.0f) {
		if (v2_distsq_quirk(c, v1) > radius * radius) return 0; /* sph.h:637 */
		V2 d = v2_sub(c, v1);
		normal = v2_normalize(d);
		float penetration = radius - v2_dot(normal, d);
		*mtv = v2_scale(normal, penetration);
		return 1;
	} else if (u2 <= 0.0f) {
		if (v2_distsq_quirk(c, v2_) > radius * radius) return 0; /* sph.h:648 */
		V2 d = v2_sub(c, v2_);
		normal = v2_normalize(d);
		float penetration = radius - v2_dot(normal, d);
		*mtv = v2_scale(normal, penetration);
		return 1;
	} else {
		V2 fc = v2(lerpf(v1.x, 0.5f, v2_.x), lerpf(v1.y, 0.5f, v2_.y));
		V2 d = v2_sub(c, fc);
		float s = v2_dot(d, normal);
		if (s > radius) return 0;
		float penetration = radius - s;
		*mtv = v2_scale(normal, penetration);
		return 1;
	}
}

static V2 solve_polygon(V2 pos, int n, const V2 *verts) { /* sph.h:674-681 */
	V2 mtv = v2(0, 0);
	if (find_mtv_circle_polygon(pos, n, verts, &mtv)) pos = v2_add(mtv, pos);
	return pos;
}

static V2 collide_all(const SphOracle *o, V2 pos) { /* demo4.cpp:414-441, bodies in insertion order */
	for (int b = 0; b < o->nbodies; ++b) {
		const Body *bd = &o->bodies[b];
		switch (bd->type) {
			case ORACLE_BODY_PLANE: pos = solve_plane(pos, v2(bd->f[0], bd->f[1]), bd->f[2]); break;
			case ORACLE_BODY_CIRCLE: pos = solve_circle(pos, v2(bd->f[0], bd->f[1]), bd->f[2]); break;
			case ORACLE_BODY_SEGMENT: pos = solve_segment(pos, v2(bd->f[0], bd->f[1]), v2(bd->f[2], bd->f[3])); break;
			case ORACLE_BODY_POLYGON: pos = solve_polygon(pos, bd->nverts, (const V2 *)bd->f); break;
			default: break;
		}
	}
	return pos;
}

void oracle_pass_collide(SphOracle *o) {
	for (uint64_t i = 0; i < o->n; ++i) o->p[i].cur = collide_all(o, o->p[i].cur);
}

/* ---- particles, volumes, emitters (demo4.cpp:142-181, 257-284) --------------------------- */
static void ensure_particles(SphOracle *o, uint64_t need) {
	if (need <= o->cap) return;
	uint64_t cap = o->cap ? o->cap : 1024;
	while (cap < need) cap *= 2;
	o->p = (Particle *)realloc(o->p, cap * sizeof(Particle));
	o->pi = (ParticleIdx *)realloc(o->pi, cap * sizeof(ParticleIdx));
	o->cap = cap;
}

uint64_t oracle_add_particle(SphOracle *o, float x, float y, float ax, float ay) { /* demo4.cpp:142-153 */
	ensure_particles(o, o->n + 1);
	uint64_t i = o->n++;
	Particle *p = &o->p[i];
	memset(p, 0, sizeof(*p));
	p->cur = p->prev = v2(x, y);
	p->acc = v2(ax, ay);
	memset(&o->pi[i], 0, sizeof(o->pi[i]));
	grid_insert(o, i);
	return i;
}

uint64_t oracle_add_particles(SphOracle *o, uint64_t n, const float *pos, const float *acc) {
	uint64_t first = o->n;
	for (uint64_t k = 0; k < n; ++k)
		oracle_add_particle(o, pos[2 * k], pos[2 * k + 1], acc ? acc[2 * k] : 0.0f, acc ? acc[2 * k + 1] : 0.0f);
	return first;
}

static V2 random_direction(void) { /* vecmath.h:317-322 */
	float d = rand() / (float)RAND_MAX;
	float angle = d * ((float)M_PI * 2.0f);
	return v2(cosf(angle), sinf(angle));
}

void oracle_add_volume(SphOracle *o, float cx, float cy, float fx, float fy, int nx, int ny, float spacing) {
	/* demo4.cpp:169-181 */
	V2 offset = v2_scale(v2(nx * spacing, ny * spacing), 0.5f);
	V2 base = v2_sub(v2(cx, cy), offset);
	for (int yi = 0; yi < ny; ++yi)
		for (int xi = 0; xi < nx; ++xi) {
			V2 p = v2_scale(v2((float)xi, (float)yi), spacing);
			p = v2_add(v2(spacing * 0.5f, spacing * 0.5f), p);
			p = v2_add(base, p);
			V2 jitter = v2_scale(v2_scale(random_direction(), kKernelHeight), kVolumeDistributionScale);
			p = v2_add(jitter, p);
			oracle_add_particle(o, p.x, p.y, fx, fy);
		}
}

void oracle_add_emitter(SphOracle *o, float px, float py, float dx, float dy, float radius, float speed, float rate, float duration) {
	if (o->nemitters >= kMaxEmitters) return;
	Emitter *e = &o->emitters[o->nemitters++]; /* demo4.cpp:155-167 */
	e->position = v2(px, py);
	e->direction = v2(dx, dy);
	e->radius = radius;
	e->speed = speed;
	e->rate = rate;
	e->duration = duration;
	e->elapsed = 0;
	e->totalElapsed = 0;
	e->active = 1;
}

static void update_emitter(SphOracle *o, Emitter *em, float dt) { /* demo4.cpp:257-284 */
	const float spacing = o->params[2];
	const float invDt = 1.0f / dt;
	if (!em->active) return;
	const float rate = 1.0f / em->rate;
	em->elapsed += dt;
	em->totalElapsed += dt;
	if (em->elapsed >= rate) {
		em->elapsed = 0;
		V2 acc = v2_scale(v2_scale(em->direction, em->speed), invDt);
		V2 dir = v2(-1.0f * em->direction.y, 1.0f * em->direction.x); /* Vec2Cross(1.0f, dir), vecmath.h:304 */
		int count = (int)floor(em->radius / spacing);
		V2 offset = v2_scale(v2_scale(v2_scale(dir, (float)count), spacing), 0.5f);
		V2 base = v2_sub(em->position, offset);
		for (int k = 0; k < count; ++k) {
			V2 p = v2_scale(v2_scale(dir, (float)k), spacing);
			p = v2_add(v2_scale(v2_scale(dir, spacing), 0.5f), p);
			p = v2_add(base, p);
			V2 jitter = v2_scale(v2_scale(random_direction(), kKernelHeight), kVolumeDistributionScale);
			p = v2_add(jitter, p);
			oracle_add_particle(o, p.x, p.y, acc.x, acc.y);
		}
	}
	if (em->totalElapsed >= em->duration) em->active = 0;
}

/* ---- Update (demo4.cpp:286-451) ------------------------------------------------------------ */
void oracle_step(SphOracle *o, float dt) {
	const float invDt = 1.0f / dt;
	double t0 = now_ms(), t1;

	for (int e = 0; e < o->nemitters; ++e) update_emitter(o, &o->emitters[e], dt); /* :291-299 */
	t1 = now_ms(); o->times[0] = (float)(t1 - t0); t0 = t1;

	V2 force = v2_add(o->gravity, o->extForce);
	for (uint64_t i = 0; i < o->n; ++i) { /* :304-309 */
		Particle *p = &o->p[i];
		p->acc = v2_add(force, p->acc);
		p->vel = v2_add(v2_scale(p->acc, dt), p->vel);
		p->acc = v2(0, 0);
	}
	t1 = now_ms(); o->times[1] = (float)(t1 - t0); t0 = t1;

	oracle_pass_viscosity(o, dt); /* :315-327, on the previous step's lists */
	t1 = now_ms(); o->times[2] = (float)(t1 - t0); t0 = t1;

	for (uint64_t i = 0; i < o->n; ++i) { /* :332-336 */
		Particle *p = &o->p[i];
		p->prev = p->cur;
		p->cur = v2_add(v2_scale(p->vel, dt), p->cur);
	}
	t1 = now_ms(); o->times[3] = (float)(t1 - t0); t0 = t1;

	oracle_pass_update_grid(o); /* :342-356 */
	t1 = now_ms(); o->times[4] = (float)(t1 - t0); t0 = t1;

	oracle_pass_neighbor_search(o); /* :359-379 */
	t1 = now_ms(); o->times[5] = (float)(t1 - t0); t0 = t1;

	oracle_pass_density(o); /* :382-394 */
	t1 = now_ms(); o->times[6] = (float)(t1 - t0); t0 = t1;

	oracle_pass_delta(o, dt); /* :397-409 */
	t1 = now_ms(); o->times[7] = (float)(t1 - t0); t0 = t1;

	oracle_pass_collide(o); /* :412-444 */
	t1 = now_ms(); o->times[8] = (float)(t1 - t0);

	for (uint64_t i = 0; i < o->n; ++i) { /* :447-450 */
		Particle *p = &o->p[i];
		p->vel = v2_scale(v2_sub(p->cur, p->prev), invDt);
	}
}

double oracle_step_timed(SphOracle *o, float dt, int steps) {
	double t0 = now_ms();
	for (int s = 0; s < steps; ++s) oracle_step(o, dt);
	return (now_ms() - t0) * 1e-3;
}

/* ---- lifecycle / configuration ----------------------------------------------------------- */
static void default_params(float p[9]) { /* SPHParameters(), sph.h:88-98 */
	p[0] = kKernelHeight;
	p[1] = kKernelHeight;          /* kSPHGridCellSize, sph.h:60 */
	p[2] = kKernelHeight * 0.5f;   /* kSPHParticleSpacing, sph.h:37 */
	p[3] = 1.0f / p[0];
	p[4] = kRestDensity;
	p[5] = kStiffness;
	p[6] = kStiffness * 10.0f;
	p[7] = kLinearViscosity;
	p[8] = kQuadraticViscosity;
}

SphOracle *oracle_create(float width, float height, float cell) {
	SphOracle *o = (SphOracle *)calloc(1, sizeof(SphOracle));
	o->width = width;
	o->height = height;
	o->halfW = width * 0.5f;  /* sph.h:21 */
	o->halfH = height * 0.5f; /* sph.h:22 */
	o->cellSize = cell;
	o->gridX = (int)(width / cell);  /* sph.h:61 */
	o->gridY = (int)(height / cell); /* sph.h:62 */
	o->cells = (Cell *)calloc((size_t)o->gridX * o->gridY, sizeof(Cell));
	default_params(o->params);
	o->omega = 1.0f;
	o->threads = 1;
	oracle_reset_stats(o);
	return o;
}

void oracle_destroy(SphOracle *o) {
	if (!o) return;
	pool_stop(&o->pool);
	size_t ncell = (size_t)o->gridX * o->gridY;
	for (size_t c = 0; c < ncell; ++c) free(o->cells[c].idx);
	free(o->cells);
	free(o->p);
	free(o->pi);
	free(o->nbrOff);
	free(o->nbr);
	free(o->bodies);
	free(o->scratch);
	free(o);
}

void oracle_set_mode(SphOracle *o, int mode) { o->mode = mode; }
void oracle_set_relaxation(SphOracle *o, float omega) { o->omega = omega; }
int oracle_get_threads(const SphOracle *o) { return o->threads; }
void oracle_set_threads(SphOracle *o, int threads) {
	if (threads < 1) threads = 1;
	if (threads == o->threads) return;
	pool_stop(&o->pool);
	o->threads = threads;
	if (threads > 1) pool_start(&o->pool, threads);
}

void oracle_reset_stats(SphOracle *o) { /* SPHStatistics(), sph.h:143-149 */
	o->statMinNbr = kMaxCellStat;
	o->statMaxNbr = 0;
	o->statMinCell = kMaxCellStat;
	o->statMaxCell = 0;
	memset(o->times, 0, sizeof(o->times));
}
void oracle_clear_bodies(SphOracle *o) { o->nbodies = 0; }
void oracle_clear_particles(SphOracle *o) { /* demo4.cpp:125-132 */
	size_t ncell = (size_t)o->gridX * o->gridY;
	for (size_t c = 0; c < ncell; ++c) o->cells[c].count = 0;
	o->n = 0;
	o->nbrN = 0;
}
void oracle_clear_emitters(SphOracle *o) { o->nemitters = 0; }
void oracle_set_params(SphOracle *o, const float p9[9]) { /* copy-ctor, sph.h:100-110 */
	memcpy(o->params, p9, sizeof(o->params));
	o->params[3] = 1.0f / o->params[0];
}
void oracle_get_params(const SphOracle *o, float p9[9]) { memcpy(p9, o->params, sizeof(o->params)); }
void oracle_set_gravity(SphOracle *o, float gx, float gy) { o->gravity = v2(gx, gy); }
void oracle_get_gravity(const SphOracle *o, float g2[2]) { g2[0] = o->gravity.x; g2[1] = o->gravity.y; }
void oracle_add_external_force(SphOracle *o, float fx, float fy) { o->extForce = v2_add(v2(fx, fy), o->extForce); }
void oracle_clear_external_force(SphOracle *o) { o->extForce = v2(0, 0); }

static Body *new_body(SphOracle *o, int type) {
	if (o->nbodies == o->bodyCap) {
		o->bodyCap = o->bodyCap ? o->bodyCap * 2 : 16;
		o->bodies = (Body *)realloc(o->bodies, (size_t)o->bodyCap * sizeof(Body));
	}
	Body *b = &o->bodies[o->nbodies++];
	memset(b, 0, sizeof(*b));
	b->type = type;
	return b;
}
void oracle_add_plane(SphOracle *o, float nx, float ny, float d) { Body *b = new_body(o, ORACLE_BODY_PLANE); b->f[0] = nx; b->f[1] = ny; b->f[2] = d; }
void oracle_add_circle(SphOracle *o, float x, float y, float r) { Body *b = new_body(o, ORACLE_BODY_CIRCLE); b->f[0] = x; b->f[1] = y; b->f[2] = r; }
void oracle_add_segment(SphOracle *o, float ax, float ay, float bx, float by) { Body *b = new_body(o, ORACLE_BODY_SEGMENT); b->f[0] = ax; b->f[1] = ay; b->f[2] = bx; b->f[3] = by; }
void oracle_add_polygon(SphOracle *o, int n, const float *xy) {
	if (n > kMaxPolyVerts) n = kMaxPolyVerts;
	Body *b = new_body(o, ORACLE_BODY_POLYGON);
	b->nverts = n;
	memcpy(b->f, xy, (size_t)n * 2 * sizeof(float));
}
int oracle_body_count(const SphOracle *o) { return o->nbodies; }
void oracle_get_body(const SphOracle *o, int idx, int32_t *type, int32_t *nverts, float out16[16]) {
	*type = o->bodies[idx].type;
	*nverts = o->bodies[idx].nverts;
	memcpy(out16, o->bodies[idx].f, 16 * sizeof(float));
}

/* ---- the built-in scenes (sph.h:307-437) and LoadScenario (app.cpp:477-534) -------------- */
#define BW 10.0f                       /* kSPHBoundaryWidth, sph.h:19 */
#define BH (BW / (16.0f / 9.0f))       /* kSPHBoundaryHeight, sph.h:20 */
#define BHW (BW * 0.5f)
#define BHH (BH * 0.5f)
#define DEG2RAD ((float)M_PI / 180.0f) /* vecmath.h:9 */

typedef struct { int type; float px, py, rot, a, b, c, d; } SceneBody; /* plane: a,b = normal; circle: a = radius; box: a,b = half extents */
typedef struct { float px, py, w, h, fx, fy; } SceneVolume;
typedef struct { float px, py, dx, dy, radius, speed, rate, duration; } SceneEmitter;
typedef struct {
	const char *name;
	float gx, gy;
	int nvol; SceneVolume vol[2];
	int nemit; SceneEmitter emit[1];
	int nbody; SceneBody body[8];
	float spacing, nearStiffness;
} Scene;

enum { SB_PLANE = 1, SB_CIRCLE = 2, SB_BOX = 3 };
#define WALLS4 \
	{ SB_PLANE, 0, -BHH, 0, 0, 1, 0, 0 }, { SB_PLANE, 0, BHH, 0, 0, -1, 0, 0 }, \
	{ SB_PLANE, -BHW, 0, 0, 1, 0, 0, 0 }, { SB_PLANE, BHW, 0, 0, -1, 0, 0, 0 }

#define DAM_WALL_W (BW * 0.05f)
#define DAM_WALL_H (BH * 0.85f)
#define DAM_VOL_W (BW * 0.25f)
#define DAM_VOL_H (BH * 0.95f)
#define BLOB_W (BW * 0.5f)
#define BLOB_H (BH * 0.5f)

static int scene_table(Scene *out) {
	/* the 8-arg SPHParameters ctor (sph.h:112-122) ignores its kernelHeight and restDensity
	 * arguments, so only spacing and nearStiffness vary between scenes */
	const Scene scenes[] = {
		{ "Dambreak", 0, -10, 1, { { -BHW + DAM_VOL_W * 0.5f, 0, DAM_VOL_W, DAM_VOL_H, 0, 0 } }, 0, { { 0 } }, 5,
		  { WALLS4, { SB_BOX, -BHW + DAM_VOL_W + DAM_WALL_W * 0.5f + kCollisionRadius, BH * 0.1f, 0.0f, DAM_WALL_W * 0.5f, DAM_WALL_H * 0.5f, 0, 0 } },
		  kKernelHeight / 6.0f, kStiffness * 10.0f },
		{ "Dambreak x 2", 0, -10, 2,
		  { { -BHW + DAM_VOL_W * 0.5f, 0, DAM_VOL_W, DAM_VOL_H, 0, 0 }, { BHW - DAM_VOL_W * 0.5f, 0, DAM_VOL_W, DAM_VOL_H, 0, 0 } },
		  0, { { 0 } }, 4, { WALLS4 }, kKernelHeight / 3.0f, kStiffness * 20.0f },
		{ "Blob", 0, 0, 1, { { 0, 0, BLOB_W, BLOB_H, 0, 0 } }, 0, { { 0 } }, 4, { WALLS4 }, kKernelHeight / 3.0f, kStiffness * 10.0f },
		{ "Blob x 2", 0, 0, 2,
		  { { -BLOB_H * 0.75f, 0, BLOB_H * 0.75f, BLOB_H * 0.75f, 10, 0 }, { BLOB_H * 0.75f, 0, BLOB_H * 0.75f, BLOB_H * 0.75f, -10, 0 } },
		  0, { { 0 } }, 4, { WALLS4 }, kKernelHeight / 3.0f, kStiffness * 10.0f },
		{ "Liquid", 0, -2, 0, { { 0 } }, 1, { { -3.5f, 0.0f, 1, 0, kKernelHeight * 3, 2.5f, 15.0f, 30.0f } }, 4, { WALLS4 },
		  kKernelHeight / 4.0f, kStiffness * 10.0f },
		{ "Glass", 0, -10, 0, { { 0 } }, 1, { { -1.5f, 2.0f, 1, 0, kKernelHeight * 3, 2.5f, 15.0f, 25.0f } }, 7,
		  { WALLS4, { SB_BOX, 0.0f, -2.0f, 0.0f, 1.0f, 0.2f, 0, 0 }, { SB_BOX, -1.0f, -0.5f, 0.0f, 0.2f, 1.5f, 0, 0 }, { SB_BOX, 1.0f, -0.5f, 0.0f, 0.2f, 1.5f, 0, 0 } },
		  kKernelHeight / 4.0f, kStiffness * 6.0f },
		{ "Fontain", 0, -10, 0, { { 0 } }, 1, { { 0, -BHH + 1.0f, 0, 1, kKernelHeight * 4, 8.0f, 15.0f, 25.0f } }, 4, { WALLS4 },
		  kKernelHeight / 4.0f, kStiffness * 2.0f },
		{ "Fun", 0, -10, 0, { { 0 } }, 1, { { -4, 2, 1, 0, kKernelHeight * 4, 3.5f, 15.0f, 20.0f } }, 7,
		  { { SB_PLANE, 0, -BHH, 0, 0, 1, 0, 0 }, { SB_PLANE, -BHW, 0, 0, 1, 0, 0, 0 }, { SB_PLANE, BHW, 0, 0, -1, 0, 0, 0 },
		    { SB_BOX, -1.5f, 1.0f, DEG2RAD * -2.5f, 3.5f, 0.1f, 0, 0 }, { SB_BOX, 1.5f, -0.25f, DEG2RAD * 2.5f, 3.5f, 0.1f, 0, 0 },
		    { SB_CIRCLE, -4.0f, -1.5f, 0, 0.5f, 0, 0, 0 }, { SB_BOX, 0, -BHH + 0.5f, 0, 0.3f, 1.0f, 0, 0 } },
		  kKernelHeight / 4.0f, kStiffness * 6.0f },
	};
	int count = (int)(sizeof(scenes) / sizeof(scenes[0]));
	if (out) memcpy(out, scenes, sizeof(scenes));
	return count;
}

int oracle_scenario_count(void) { return scene_table(NULL); }
const char *oracle_scenario_name(int idx) {
	static Scene table[8];
	scene_table(table);
	return table[idx].name;
}

static V2 mat2_mul(V2 col1, V2 col2, V2 v) { /* Vec2MultMat2, vecmath.h:287-290 */
	return v2(col1.x * v.x + col2.x * v.y, col1.y * v.x + col2.y * v.y);
}

void oracle_load_scenario(SphOracle *o, int idx, int seed) {
	Scene table[8];
	scene_table(table);
	const Scene *sc = &table[idx];
	if (seed >= 0) srand((unsigned)seed);
	oracle_reset_stats(o);
	oracle_clear_bodies(o);
	oracle_clear_particles(o);
	oracle_clear_emitters(o);
	oracle_set_gravity(o, sc->gx, sc->gy);
	float p[9];
	default_params(p);
	p[2] = sc->spacing;
	p[6] = sc->nearStiffness;
	oracle_set_params(o, p);
	for (int b = 0; b < sc->nbody; ++b) {
		const SceneBody *sb = &sc->body[b];
		V2 pos = v2(sb->px, sb->py);
		if (sb->type == SB_PLANE) { /* CreatePlane + app.cpp:491-495 */
			V2 normal = v2(sb->a, sb->b);
			oracle_add_plane(o, normal.x, normal.y, v2_dot(normal, pos));
		} else if (sb->type == SB_CIRCLE) {
			oracle_add_circle(o, pos.x, pos.y, sb->a);
		} else { /* CreateBox (sph.h:204-216) + app.cpp:507-515 */
			float s = sinf(sb->rot), c = cosf(sb->rot);
			V2 col1 = v2(c, s), col2 = v2(-s, c); /* Mat2FromAngle, vecmath.h:358-365 */
			V2 local[4] = { v2(sb->a, sb->b), v2(-sb->a, sb->b), v2(-sb->a, -sb->b), v2(sb->a, -sb->b) };
			float xy[8];
			for (int v = 0; v < 4; ++v) {
				V2 w = v2_add(mat2_mul(col1, col2, local[v]), pos);
				xy[2 * v] = w.x;
				xy[2 * v + 1] = w.y;
			}
			oracle_add_polygon(o, 4, xy);
		}
	}
	const float spacing = o->params[2];
	for (int v = 0; v < sc->nvol; ++v) { /* app.cpp:519-527 (zero-sized extra volumes add nothing) */
		int numX = (int)floor((sc->vol[v].w / spacing));
		int numY = (int)floor((sc->vol[v].h / spacing));
		oracle_add_volume(o, sc->vol[v].px, sc->vol[v].py, sc->vol[v].fx, sc->vol[v].fy, numX, numY, spacing);
	}
	for (int e = 0; e < sc->nemit; ++e) {
		const SceneEmitter *se = &sc->emit[e];
		oracle_add_emitter(o, se->px, se->py, se->dx, se->dy, se->radius, se->speed, se->rate, se->duration);
	}
}

/* ---- state access ------------------------------------------------------------------------- */
uint64_t oracle_particle_count(const SphOracle *o) { return o->n; }
void oracle_grid_dims(const SphOracle *o, int32_t out2[2]) { out2[0] = o->gridX; out2[1] = o->gridY; }
void oracle_get_particles(const SphOracle *o, float *out12) { memcpy(out12, o->p, o->n * sizeof(Particle)); }
void oracle_set_particles(SphOracle *o, const float *in12) {
	memcpy(o->p, in12, o->n * sizeof(Particle));
	oracle_pass_update_grid(o);
}
void oracle_get_cell_of_particle(const SphOracle *o, int32_t *out) {
	for (uint64_t i = 0; i < o->n; ++i) { out[2 * i] = o->pi[i].cx; out[2 * i + 1] = o->pi[i].cy; }
}
void oracle_get_cell_counts(const SphOracle *o, uint32_t *out) {
	size_t ncell = (size_t)o->gridX * o->gridY;
	for (size_t c = 0; c < ncell; ++c) out[c] = o->cells[c].count;
}
uint32_t oracle_get_cell_members(const SphOracle *o, int cell, uint32_t *out) {
	memcpy(out, o->cells[cell].idx, o->cells[cell].count * sizeof(uint32_t));
	return o->cells[cell].count;
}
void oracle_get_neighbor_counts(const SphOracle *o, uint32_t *out) {
	for (uint64_t i = 0; i < o->n; ++i) out[i] = (uint32_t)(nbr_end(o, i) - nbr_begin(o, i));
}
uint32_t oracle_get_neighbors(const SphOracle *o, uint64_t i, uint32_t *out) {
	uint64_t b = nbr_begin(o, i), e = nbr_end(o, i);
	memcpy(out, o->nbr + b, (e - b) * sizeof(uint32_t));
	return (uint32_t)(e - b);
}
void oracle_get_stats(const SphOracle *o, uint64_t c[4], float t[9]) {
	c[0] = o->statMinNbr; c[1] = o->statMaxNbr; c[2] = o->statMinCell; c[3] = o->statMaxCell;
	memcpy(t, o->times, sizeof(o->times));
}
void oracle_get_colors(const SphOracle *o, float *out4) { /* sph.h:683-695 */
	for (uint64_t i = 0; i < o->n; ++i) {
		const Particle *p = &o->p[i];
		float r = p->P / (-10.0f);
		float g = p->rho / o->params[4];
		float b = v2_len(p->vel) / 10.0f;
		out4[4 * i] = fmaxf(fminf(r, 1.0f), 0.0f);
		out4[4 * i + 1] = fmaxf(fminf(g, 1.0f), 0.0f);
		out4[4 * i + 2] = fmaxf(fminf(b, 1.0f), 0.0f);
		out4[4 * i + 3] = 1.0f;
	}
}

void oracle_solve_plane(float p[2], float nx, float ny, float d) { V2 r = solve_plane(v2(p[0], p[1]), v2(nx, ny), d); p[0] = r.x; p[1] = r.y; }
void oracle_solve_circle(float p[2], float cx, float cy, float rad) { V2 r = solve_circle(v2(p[0], p[1]), v2(cx, cy), rad); p[0] = r.x; p[1] = r.y; }
void oracle_solve_segment(float p[2], float ax, float ay, float bx, float by) { V2 r = solve_segment(v2(p[0], p[1]), v2(ax, ay), v2(bx, by)); p[0] = r.x; p[1] = r.y; }
void oracle_solve_polygon(float p[2], int n, const float *xy) { V2 r = solve_polygon(v2(p[0], p[1]), n, (const V2 *)xy); p[0] = r.x; p[1] = r.y; }
void oracle_cell_index(const SphOracle *o, float x, float y, int32_t out2[2]) { cell_index(o, v2(x, y), &out2[0], &out2[1]); }

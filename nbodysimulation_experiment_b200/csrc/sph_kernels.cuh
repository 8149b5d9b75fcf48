// CUDA kernels of the SPH step (sm_100a).  One step is, in the reference's phase order
// (demo4.cpp:286-451):
//
//   integrate -> viscosity gather (previous step's grid) -> predict + cell key + histogram
//   -> exclusive scan -> id scatter -> canonical reorder -> density gather -> displacement gather
//   -> collide + velocity
//
// Data layout: cell-sorted SoA.  `pos/prev/vel` are float2, `cellOf` is the packed (cy<<16|cx) cell
// of each sorted particle, `id` its creation index, `cellStart` the exclusive prefix over the
// row-major cells (key = iy*gridX+ix, sph.h:444-448), so the 3 cells (cx-1..cx+1) of one grid row
// are ONE contiguous slab of the sorted arrays and a particle's candidate list (demo4.cpp:183-206)
// is three slabs.  Within a cell particles are ordered by id, which makes every floating-point
// sum order canonical: runs are bit-reproducible and independent of atomics or of how many GPUs
// share the domain.
//
// All per-particle kernels are warp-uniform grid-stride loops over device-side counts
// (Counters), so no phase needs a host round trip even when the count changes (emitters,
// migration between strips).
#pragma once
#include "sph_math.cuh"

namespace sphb200 {

struct Counters {
	uint32_t n;        // particles in the arrays at step start ([0,nSorted) sorted, the rest appended since)
	uint32_t nSorted;  // particles covered by cellStart / cellOf (the previous step's grid)
	uint32_t nIn;      // n + particles received from neighbour strips this step
	uint32_t nOut;     // particles kept by this step's grid build (= cellStart[nCells])
	uint32_t sendDown, sendUp; // halo/migration records packed for the lower / upper strip
	uint32_t lost;     // particles that left the local rows with nowhere to go
	uint32_t overflow; // capacity overflow flags
	uint32_t minNbr, maxNbr, minCell, maxCell;
	unsigned long long pairCandidates;
};

#define SPH_KEY_NONE 0xFFFFFFFFu
#define SPH_THREADS 256

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// warp-uniform grid-stride: every lane of a warp runs the same number of trips
#define SPH_WARP_LOOP(i, n)                                                                        \
	for (uint32_t i##_base = (blockIdx.x * blockDim.x + (threadIdx.x & ~31u)), i = i##_base + lane_id(); \
	     i##_base < (n); i##_base += gridDim.x * blockDim.x, i = i##_base + lane_id())

__device__ __forceinline__ uint32_t warp_min(uint32_t v) {
	for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ uint32_t warp_max(uint32_t v) {
	for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// ---- phase 1: integrate forces (demo4.cpp:302-312) --------------------------------------
// a += g + fext; v += a*dt; a = 0.  Only particles added since the last step carry a non-zero
// acceleration (demo4.cpp:146), so `acc` is read from index accFrom on and zeroed.
__global__ void __launch_bounds__(SPH_THREADS) integrate_kernel(const Counters *__restrict__ ctr, float2 *__restrict__ vel,
                                                               float2 *__restrict__ acc, uint32_t accFrom, float2 force, float dt) {
	const uint32_t n = ctr->n;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		float2 a = make_float2(0.0f, 0.0f);
		if (i >= accFrom) {
			a = acc[i];
			acc[i] = make_float2(0.0f, 0.0f);
		}
		a.x = __fadd_rn(force.x, a.x); // operator+= evaluates b + a (vecmath.h:249-252)
		a.y = __fadd_rn(force.y, a.y);
		float2 v = vel[i];
		v.x = __fadd_rn(__fmul_rn(a.x, dt), v.x);
		v.y = __fadd_rn(__fmul_rn(a.y, dt), v.y);
		vel[i] = v;
	}
}

// ---- the 3x3 candidate walk shared by the gather kernels ---------------------------------
// Calls body(j) for every candidate j of a particle in cell (cx,cy), in the reference's order:
// dy outer, dx inner (demo4.cpp:188-189), ascending id inside a cell.  Returns the list length.
template <class F>
__device__ __forceinline__ uint32_t for_each_candidate(const GridDesc &g, const uint32_t *__restrict__ cellStart, int cx, int cy, F body) {
	uint32_t total = 0;
	const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.gx - 1);
#pragma unroll 1
	for (int y = cy - 1; y <= cy + 1; ++y) {
		if (y < g.rowLo || y >= g.rowHi) continue; // SPHIsPositionInGrid, sph.h:439-442 (and strip edge)
		const uint32_t base = (uint32_t)(y - g.rowLo) * (uint32_t)g.gx;
		const uint32_t lo = cellStart[base + x0], hi = cellStart[base + x1 + 1];
		total += hi - lo;
		for (uint32_t j = lo; j < hi; ++j) body(j);
	}
	return total;
}

// ---- phase 2: viscosity impulses on the previous step's lists (demo4.cpp:223-237, 315-327) ---
template <class M>
__global__ void __launch_bounds__(SPH_THREADS) viscosity_kernel(GridDesc g, PairParams k, const Counters *__restrict__ ctr,
                                                                const float2 *__restrict__ pos, const float2 *__restrict__ vel,
                                                                const uint32_t *__restrict__ cellOf, const uint32_t *__restrict__ cellStart,
                                                                float2 *__restrict__ velOut) {
	const uint32_t n = ctr->n, nSorted = ctr->nSorted;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		float2 vi = vel[i];
		float vx = vi.x, vy = vi.y;
		if (i < nSorted) { // particles created since the last grid build have no list (demo4.cpp:148)
			const float2 pi = pos[i];
			const uint32_t c = cellOf[i];
			for_each_candidate(g, cellStart, (int)(c & 0xffffu), (int)(c >> 16), [&](uint32_t j) {
				viscosity_pair<M>(k, pi, vi, __ldg(&pos[j]), __ldg(&vel[j]), vx, vy);
			});
		}
		velOut[i] = make_float2(vx, vy);
	}
}

// ---- phases 3+4a: predict, cell key, histogram (demo4.cpp:330-356) -------------------------
// prev = x; x += v*dt; cell = SPHComputeCellIndex(x); rank = cellCount[cell]++ (warp-aggregated).
// With doPredict = 0 only the grid part runs (sph_run_pass(GRID) / state injection).
__global__ void __launch_bounds__(SPH_THREADS) predict_key_kernel(GridDesc g, Counters *__restrict__ ctr, float2 *__restrict__ pos,
                                                                 float2 *__restrict__ prev, const float2 *__restrict__ vel,
                                                                 uint32_t *__restrict__ cellNew, uint32_t *__restrict__ rank,
                                                                 uint32_t *__restrict__ cellCount, float dt, int doPredict) {
	const uint32_t n = ctr->n;
	SPH_WARP_LOOP(i, n) {
		const bool in = i < n;
		uint32_t key = 0xFFFFFF00u | lane_id(); // unique per lane: never matches
		uint32_t packed = SPH_KEY_NONE;
		if (in) {
			float2 p = pos[i];
			if (doPredict) {
				const float2 v = vel[i];
				prev[i] = p;
				p.x = __fadd_rn(__fmul_rn(v.x, dt), p.x);
				p.y = __fadd_rn(__fmul_rn(v.y, dt), p.y);
				pos[i] = p;
			}
			int cx, cy;
			cell_of(g, p, cx, cy);
			if (cy >= g.rowLo && cy < g.rowHi) {
				key = (uint32_t)(cy - g.rowLo) * (uint32_t)g.gx + (uint32_t)cx;
				packed = pack_cell(cx, cy);
			} else {
				atomicAdd(&ctr->lost, 1u);
			}
		}
		// one atomic per distinct cell per warp
		const uint32_t peers = __match_any_sync(0xffffffffu, key);
		const int leader = __ffs(peers) - 1;
		uint32_t base = 0;
		if (packed != SPH_KEY_NONE && (int)lane_id() == leader) base = atomicAdd(&cellCount[key], (uint32_t)__popc(peers));
		base = __shfl_sync(0xffffffffu, base, leader);
		if (in) {
			cellNew[i] = packed;
			rank[i] = base + (uint32_t)__popc(peers & ((1u << lane_id()) - 1u));
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) ctr->nIn = n;
}

// ---- phase 4b: exclusive scan of the cell histogram ----------------------------------------
// Three small kernels (tile-local scan, scan of tile sums, add back); cells*12 B of traffic.
#define SPH_SCAN_ITEMS 16
#define SPH_SCAN_TILE (SPH_THREADS * SPH_SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
	__shared__ uint32_t warpSums[SPH_THREADS / 32];
	const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
	uint32_t inc = v;
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= (uint32_t)o) inc += t;
	}
	if (lane == 31) warpSums[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t s = lane < SPH_THREADS / 32 ? warpSums[lane] : 0;
		for (int o = 1; o < SPH_THREADS / 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
			if (lane >= (uint32_t)o) s += t;
		}
		if (lane < SPH_THREADS / 32) warpSums[lane] = s;
	}
	__syncthreads();
	const uint32_t before = w ? warpSums[w - 1] : 0;
	*total = warpSums[SPH_THREADS / 32 - 1];
	__syncthreads();
	return before + inc - v;
}

__global__ void __launch_bounds__(SPH_THREADS) scan_tiles_kernel(const uint32_t *__restrict__ cellCount, uint32_t *__restrict__ cellStart,
                                                                uint32_t *__restrict__ tileSums, uint32_t nCells) {
	const uint32_t first = blockIdx.x * SPH_SCAN_TILE + threadIdx.x * SPH_SCAN_ITEMS;
	uint32_t v[SPH_SCAN_ITEMS], sum = 0;
#pragma unroll
	for (int k = 0; k < SPH_SCAN_ITEMS; ++k) {
		v[k] = (first + k < nCells) ? cellCount[first + k] : 0u;
		sum += v[k];
	}
	uint32_t total;
	uint32_t run = block_exclusive_scan(sum, &total);
#pragma unroll
	for (int k = 0; k < SPH_SCAN_ITEMS; ++k) {
		if (first + k < nCells) cellStart[first + k] = run;
		run += v[k];
	}
	if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums in place, total -> cellStart[nCells] and ctr->nOut
__global__ void __launch_bounds__(SPH_THREADS) scan_sums_kernel(uint32_t *__restrict__ tileSums, uint32_t nTiles, uint32_t *__restrict__ cellStart,
                                                               uint32_t nCells, Counters *__restrict__ ctr) {
	uint32_t carry = 0;
	for (uint32_t base = 0; base < nTiles; base += SPH_THREADS) {
		const uint32_t idx = base + threadIdx.x;
		const uint32_t v = idx < nTiles ? tileSums[idx] : 0u;
		uint32_t total;
		const uint32_t ex = block_exclusive_scan(v, &total);
		if (idx < nTiles) tileSums[idx] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0) {
		cellStart[nCells] = carry;
		ctr->nOut = carry;
	}
}

__global__ void __launch_bounds__(SPH_THREADS) scan_add_kernel(uint32_t *__restrict__ cellStart, const uint32_t *__restrict__ tileSums, uint32_t nCells) {
	const uint32_t off = tileSums[blockIdx.x];
	const uint32_t first = blockIdx.x * SPH_SCAN_TILE;
	for (uint32_t k = threadIdx.x; k < SPH_SCAN_TILE; k += SPH_THREADS)
		if (first + k < nCells) cellStart[first + k] += off;
}

// ---- phase 4c: drop each id at (cell start + arrival rank) ----------------------------------
__global__ void __launch_bounds__(SPH_THREADS) scatter_ids_kernel(GridDesc g, const Counters *__restrict__ ctr, const uint32_t *__restrict__ cellNew,
                                                                 const uint32_t *__restrict__ rank, const uint32_t *__restrict__ id,
                                                                 const uint32_t *__restrict__ cellStart, uint32_t *__restrict__ slotId) {
	const uint32_t n = ctr->nIn;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		const uint32_t c = cellNew[i];
		if (c == SPH_KEY_NONE) continue;
		const uint32_t key = ((c >> 16) - (uint32_t)g.rowLo) * (uint32_t)g.gx + (c & 0xffffu);
		slotId[cellStart[key] + rank[i]] = id[i];
	}
}

// ---- phase 4d: canonical reorder --------------------------------------------------------------
// Arrival ranks come from atomics, so they are not reproducible.  Each particle re-ranks itself
// inside its cell by id (count the smaller ids among the cell's members) and moves its payload
// there: the result is the stable (cell, id) order whatever the atomics did.
__global__ void __launch_bounds__(SPH_THREADS) reorder_kernel(GridDesc g, Counters *__restrict__ ctr, const uint32_t *__restrict__ cellNew,
                                                             const uint32_t *__restrict__ id, const uint32_t *__restrict__ cellStart,
                                                             const uint32_t *__restrict__ slotId, const float2 *__restrict__ pos,
                                                             const float2 *__restrict__ prev, float2 *__restrict__ posOut,
                                                             float2 *__restrict__ prevOut, uint32_t *__restrict__ idOut,
                                                             uint32_t *__restrict__ cellOut) {
	const uint32_t n = ctr->nIn;
	uint32_t occMin = 0xffffffffu, occMax = 0;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		const uint32_t c = cellNew[i];
		if (c == SPH_KEY_NONE) continue;
		const uint32_t key = ((c >> 16) - (uint32_t)g.rowLo) * (uint32_t)g.gx + (c & 0xffffu);
		const uint32_t lo = cellStart[key], hi = cellStart[key + 1];
		const uint32_t me = id[i];
		uint32_t r = 0;
		for (uint32_t s = lo; s < hi; ++s) r += (slotId[s] < me) ? 1u : 0u;
		const uint32_t dst = lo + r;
		posOut[dst] = pos[i];
		prevOut[dst] = prev[i];
		idOut[dst] = me;
		cellOut[dst] = c;
		occMin = min(occMin, hi - lo);
		occMax = max(occMax, hi - lo);
	}
	occMin = warp_min(occMin);
	occMax = warp_max(occMax);
	if (lane_id() == 0 && occMax) {
		atomicMin(&ctr->minCell, occMin);
		atomicMax(&ctr->maxCell, occMax);
	}
}

// ---- phase 6: density and pressure (demo4.cpp:208-221) ----------------------------------------
template <class M>
__global__ void __launch_bounds__(SPH_THREADS) density_kernel(GridDesc g, PairParams k, Counters *__restrict__ ctr, const float2 *__restrict__ pos,
                                                             const uint32_t *__restrict__ cellOf, const uint32_t *__restrict__ cellStart,
                                                             float2 *__restrict__ dens, float2 *__restrict__ press) {
	const uint32_t n = ctr->nOut;
	uint32_t cMin = 0xffffffffu, cMax = 0, cSum = 0;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		const float2 pi = pos[i];
		const uint32_t c = cellOf[i];
		float rho = 0.0f, rhoNear = 0.0f;
		const uint32_t cand = for_each_candidate(g, cellStart, (int)(c & 0xffffu), (int)(c >> 16),
		                                         [&](uint32_t j) { density_pair<M>(k, pi, __ldg(&pos[j]), rho, rhoNear); });
		dens[i] = make_float2(rho, rhoNear);
		// SPHComputePressure, sph.h:478-481
		press[i] = make_float2(__fmul_rn(k.stiffness, __fsub_rn(rho, k.restDensity)), __fmul_rn(k.nearStiffness, rhoNear));
		cMin = min(cMin, cand);
		cMax = max(cMax, cand);
		cSum += cand;
	}
	// neighbour statistics of demo4.cpp:369-376
	cMin = warp_min(cMin);
	cMax = warp_max(cMax);
	cSum = warp_sum(cSum);
	if (lane_id() == 0 && cMax) {
		atomicMin(&ctr->minNbr, cMin);
		atomicMax(&ctr->maxNbr, cMax);
		atomicAdd(&ctr->pairCandidates, (unsigned long long)cSum);
	}
}

// ---- phase 7: pressure displacement, gather form of demo4.cpp:239-255 ---------------------------
template <class M>
__global__ void __launch_bounds__(SPH_THREADS) delta_kernel(GridDesc g, PairParams k, const Counters *__restrict__ ctr, const float2 *__restrict__ pos,
                                                           const float2 *__restrict__ press, const uint32_t *__restrict__ cellOf,
                                                           const uint32_t *__restrict__ cellStart, float2 *__restrict__ posOut) {
	const uint32_t n = ctr->nOut;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		const float2 pi = pos[i], ppi = press[i];
		const uint32_t c = cellOf[i];
		float dx = 0.0f, dy = 0.0f;
		for_each_candidate(g, cellStart, (int)(c & 0xffffu), (int)(c >> 16),
		                   [&](uint32_t j) { delta_pair<M>(k, pi, ppi, __ldg(&pos[j]), __ldg(&press[j]), dx, dy); });
		posOut[i] = make_float2(__fadd_rn(pi.x, __fmul_rn(k.omega, dx)), __fadd_rn(pi.y, __fmul_rn(k.omega, dy)));
	}
}

// ---- phases 8+9: body collisions and velocity (demo4.cpp:412-450) ---------------------------------
// Also closes the step: publishes n = nSorted = nOut for the next one.
__global__ void __launch_bounds__(SPH_THREADS) collide_velocity_kernel(Counters *__restrict__ ctr, float2 *__restrict__ pos, const float2 *__restrict__ prev,
                                                                      float2 *__restrict__ vel, const DevBody *__restrict__ bodies, int nbodies,
                                                                      float invDt, int doCollide, int doVelocity, int commit) {
	const uint32_t n = commit ? ctr->nOut : ctr->n;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		float2 p = pos[i];
		if (doCollide) {
			p = col::all(p, bodies, nbodies);
			pos[i] = p;
		}
		if (doVelocity) {
			const float2 q = prev[i];
			vel[i] = make_float2(__fmul_rn(__fsub_rn(p.x, q.x), invDt), __fmul_rn(__fsub_rn(p.y, q.y), invDt));
		}
	}
}

__global__ void commit_kernel(Counters *ctr) {
	ctr->n = ctr->nOut;
	ctr->nSorted = ctr->nOut;
}

// per-step reset of the statistics that are per-step in the reference (demo4.cpp:369-370)
__global__ void begin_step_kernel(Counters *ctr) {
	ctr->minNbr = 0xffffffffu;
	ctr->maxNbr = 0;
	ctr->pairCandidates = 0ull;
	ctr->sendDown = 0;
	ctr->sendUp = 0;
}

// ---- readback / injection -----------------------------------------------------------------------
// Demo4::ParticleData record (demo4.h:81-99), written at the particle's creation index.
struct ParticleRecord {
	float2 cur, prev, acc, vel;
	float rho, rhoNear, P, PNear;
};

__global__ void __launch_bounds__(SPH_THREADS) gather_records_kernel(const Counters *__restrict__ ctr, const uint32_t *__restrict__ id,
                                                                    const float2 *__restrict__ pos, const float2 *__restrict__ prev,
                                                                    const float2 *__restrict__ vel, const float2 *__restrict__ acc,
                                                                    const float2 *__restrict__ dens, const float2 *__restrict__ press,
                                                                    ParticleRecord *__restrict__ out, uint32_t idBase, uint32_t idCount) {
	const uint32_t n = ctr->n;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		const uint32_t k = id[i] - idBase;
		if (k >= idCount) continue;
		ParticleRecord r;
		r.cur = pos[i];
		r.prev = prev[i];
		r.acc = acc[i];
		r.vel = vel[i];
		const float2 d = dens[i], pr = press[i];
		r.rho = d.x;
		r.rhoNear = d.y;
		r.P = pr.x;
		r.PNear = pr.y;
		out[k] = r;
	}
}

__global__ void __launch_bounds__(SPH_THREADS) scatter_records_kernel(uint32_t n, const ParticleRecord *__restrict__ in, uint32_t *__restrict__ id,
                                                                     float2 *__restrict__ pos, float2 *__restrict__ prev, float2 *__restrict__ vel,
                                                                     float2 *__restrict__ acc, float2 *__restrict__ dens, float2 *__restrict__ press) {
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		const ParticleRecord r = in[i];
		id[i] = i;
		pos[i] = r.cur;
		prev[i] = r.prev;
		vel[i] = r.vel;
		acc[i] = r.acc;
		dens[i] = make_float2(r.rho, r.rhoNear);
		press[i] = make_float2(r.P, r.PNear);
	}
}

// Render()'s particle section (demo4.cpp:520-531): position + SPHGetParticleColor (sph.h:683-695)
struct RenderVertex {
	float2 pos;
	float4 color;
};
__global__ void __launch_bounds__(SPH_THREADS) render_kernel(const Counters *__restrict__ ctr, const uint32_t *__restrict__ id, const float2 *__restrict__ pos,
                                                            const float2 *__restrict__ vel, const float2 *__restrict__ dens,
                                                            const float2 *__restrict__ press, float restDensity, float2 *__restrict__ outPos,
                                                            float4 *__restrict__ outColor, uint32_t idBase, uint32_t idCount) {
	const uint32_t n = ctr->n;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		const uint32_t k = id[i] - idBase;
		if (k >= idCount) continue;
		const float2 v = vel[i];
		const float r = __fdiv_rn(press[i].x, -10.0f);
		const float gcol = __fdiv_rn(dens[i].x, restDensity);
		const float b = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))), 10.0f);
		outPos[k] = pos[i];
		outColor[k] = make_float4(fmaxf(fminf(r, 1.0f), 0.0f), fmaxf(fminf(gcol, 1.0f), 0.0f), fmaxf(fminf(b, 1.0f), 0.0f), 1.0f);
	}
}

__global__ void __launch_bounds__(SPH_THREADS) cell_of_particle_kernel(const Counters *__restrict__ ctr, const uint32_t *__restrict__ id,
                                                                      const uint32_t *__restrict__ cellOf, int2 *__restrict__ out, uint32_t idBase,
                                                                      uint32_t idCount) {
	const uint32_t n = ctr->nSorted;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		const uint32_t k = id[i] - idBase;
		if (k >= idCount) continue;
		const uint32_t c = cellOf[i];
		out[k] = make_int2((int)(c & 0xffffu), (int)(c >> 16));
	}
}

// AddVolume's lattice (demo4.cpp:169-181) with a counter-hash jitter instead of libc rand(), for
// scenes too large to build on the host.  Appends the particles whose cell row is in [ownLo,ownHi).
__device__ __forceinline__ uint32_t hash32(uint64_t x) { // splitmix64 finaliser
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return (uint32_t)((x ^ (x >> 31)) >> 32);
}

__global__ void __launch_bounds__(SPH_THREADS) volume_hashed_kernel(GridDesc g, Counters *__restrict__ ctr, uint32_t capacity, float2 *__restrict__ pos,
                                                                   float2 *__restrict__ prev, float2 *__restrict__ vel, float2 *__restrict__ acc,
                                                                   float2 *__restrict__ dens, float2 *__restrict__ press, uint32_t *__restrict__ id,
                                                                   float baseX, float baseY, float2 force, long long countX, long long rowFirst,
                                                                   long long rowCount, float spacing, float jitterScale, uint64_t seed,
                                                                   uint32_t firstId) {
	const unsigned long long total = (unsigned long long)countX * (unsigned long long)rowCount;
	for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < total;
	     t += (unsigned long long)gridDim.x * blockDim.x) {
		const long long yi = rowFirst + (long long)(t / (unsigned long long)countX), xi = (long long)(t % (unsigned long long)countX);
		const unsigned long long lin = (unsigned long long)yi * (unsigned long long)countX + (unsigned long long)xi;
		float2 p = make_float2(__fmul_rn((float)xi, spacing), __fmul_rn((float)yi, spacing));
		p.x = __fadd_rn(__fadd_rn(__fmul_rn(spacing, 0.5f), p.x), baseX);
		p.y = __fadd_rn(__fadd_rn(__fmul_rn(spacing, 0.5f), p.y), baseY);
		const float u = (float)(hash32(seed ^ (lin * 0xD1B54A32D192ED03ull)) >> 8) * (1.0f / 16777216.0f);
		float s, c;
		sincosf(u * 6.28318530717958647692f, &s, &c);
		p.x = __fadd_rn(__fmul_rn(c, jitterScale), p.x);
		p.y = __fadd_rn(__fmul_rn(s, jitterScale), p.y);
		int cx, cy;
		cell_of(g, p, cx, cy);
		if (cy < g.ownLo || cy >= g.ownHi) continue;
		const uint32_t slot = atomicAdd(&ctr->n, 1u);
		if (slot >= capacity) {
			atomicOr(&ctr->overflow, 1u);
			continue;
		}
		pos[slot] = p;
		prev[slot] = p;
		vel[slot] = make_float2(0.0f, 0.0f);
		acc[slot] = force;
		dens[slot] = make_float2(0.0f, 0.0f);
		press[slot] = make_float2(0.0f, 0.0f);
		id[slot] = firstId + (uint32_t)lin;
	}
}

__global__ void clamp_count_kernel(Counters *ctr, uint32_t capacity) {
	if (ctr->n > capacity) ctr->n = capacity;
}

} // namespace sphb200

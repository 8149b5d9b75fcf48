// CUDA kernels of the SPH step (sm_100a).  One step is, in the reference's phase order
// (demo4.cpp:286-451):
//
//   integrate -> viscosity (previous step's grid) -> predict + cell key + histogram [-> strip exchange]
//   -> exclusive scan (+ colour lists) -> id scatter -> canonical reorder -> density gather
//   -> pressure displacement -> collide + velocity
//
// Viscosity and displacement are 9-colour in-place Gauss-Seidel sweeps by default (color_sweep_kernel)
// or plain gathers (viscosity_kernel / delta_kernel) with SPH_SOLVER_GATHER.
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
	uint32_t xseq;     // strip exchanges published so far (the mailbox protocol's sequence number, see publish_halo_kernel)
	uint32_t reserved1;
	uint32_t lost;     // particles that left the local rows with nowhere to go
	uint32_t overflow; // capacity overflow flags: 1 particles, 2 halo buffer, 4 sweep queue, 8 neighbour strip did not publish in time
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
// Block 0 also opens the step: it resets the statistics that are per step in the reference
// (demo4.cpp:369-370); nothing else in this kernel touches them.
// The three streaming kernels (integrate, predict + key, collide + velocity) move 16-36 B per particle and do almost
// no arithmetic: with one particle per thread there were ~2.4 MB of loads in flight per GPU where HBM needs ~6 MB
// (6.4 TB/s x ~1 us), and they ran at 1.5-2 TB/s.  A thread now takes particle PAIRS (one 16-byte access per float2
// array) and keeps two pairs in flight.
struct Pair2 { // particles 2q and 2q+1 of a float2 array as one 16-byte access (the arrays are 256-byte aligned)
	float2 a, b;
};
__device__ __forceinline__ Pair2 load_pair(const float2 *arr, uint32_t q, bool both) {
	Pair2 r;
	if (both) {
		const float4 v = *reinterpret_cast<const float4 *>(arr + 2u * q);
		r.a = make_float2(v.x, v.y);
		r.b = make_float2(v.z, v.w);
	} else {
		r.a = arr[2u * q];
		r.b = make_float2(0.0f, 0.0f);
	}
	return r;
}
__device__ __forceinline__ void store_pair(float2 *arr, uint32_t q, bool both, Pair2 v) {
	if (both) *reinterpret_cast<float4 *>(arr + 2u * q) = make_float4(v.a.x, v.a.y, v.b.x, v.b.y);
	else arr[2u * q] = v.a;
}

__device__ __forceinline__ float2 integrate_one(float2 v, float2 a, float2 force, float dt) {
	a.x = __fadd_rn(force.x, a.x); // operator+= evaluates b + a (vecmath.h:249-252)
	a.y = __fadd_rn(force.y, a.y);
	v.x = __fadd_rn(__fmul_rn(a.x, dt), v.x);
	v.y = __fadd_rn(__fmul_rn(a.y, dt), v.y);
	return v;
}

__global__ void __launch_bounds__(SPH_THREADS) integrate_kernel(Counters *__restrict__ ctr, float2 *__restrict__ vel,
                                                               float2 *__restrict__ acc, uint32_t accFrom, float2 force, float dt) {
	const uint32_t n = ctr->n;
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		ctr->minNbr = 0xffffffffu;
		ctr->maxNbr = 0;
		ctr->pairCandidates = 0ull;
	}
	const uint32_t nPairs = (n + 1u) >> 1, T = gridDim.x * blockDim.x;
	for (uint32_t q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < nPairs; q0 += 2u * T) {
		const uint32_t q1 = q0 + T;
		const bool has1 = q1 < nPairs;
		const bool both0 = 2u * q0 + 1u < n, both1 = has1 && 2u * q1 + 1u < n;
		Pair2 v0 = load_pair(vel, q0, both0), v1 = {};
		if (has1) v1 = load_pair(vel, q1, both1);
		Pair2 a0 = {}, a1 = {}; // only particles added since the last step carry an acceleration (demo4.cpp:146)
		if (2u * q0 + 1u >= accFrom) {
			if (2u * q0 >= accFrom) { a0.a = acc[2u * q0]; acc[2u * q0] = make_float2(0.0f, 0.0f); }
			if (both0) { a0.b = acc[2u * q0 + 1u]; acc[2u * q0 + 1u] = make_float2(0.0f, 0.0f); }
		}
		if (has1 && 2u * q1 + 1u >= accFrom) {
			if (2u * q1 >= accFrom) { a1.a = acc[2u * q1]; acc[2u * q1] = make_float2(0.0f, 0.0f); }
			if (both1) { a1.b = acc[2u * q1 + 1u]; acc[2u * q1 + 1u] = make_float2(0.0f, 0.0f); }
		}
		v0.a = integrate_one(v0.a, a0.a, force, dt);
		v0.b = integrate_one(v0.b, a0.b, force, dt);
		store_pair(vel, q0, both0, v0);
		if (has1) {
			v1.a = integrate_one(v1.a, a1.a, force, dt);
			v1.b = integrate_one(v1.b, a1.b, force, dt);
			store_pair(vel, q1, both1, v1);
		}
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
#pragma unroll 4
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
// Multi-GPU (y-strips): a rank advances every particle it holds, but only the ones
// it OWNED on the previous grid are authoritative.  After predict each of those is (a) kept if its
// new row is inside the local window [rowLo,rowHi) = owned rows +- halo, (b) copied to the lower /
// upper neighbour if the row is inside that neighbour's window.  Stale ghost copies are dropped;
// the neighbour sends fresh ones every step.  One exchange per step carries migration and halo.
struct HaloRecord { // what a neighbour needs to file a particle into this step's grid
	float2 pos, prev;
	uint32_t id, pad;
};
// A mailbox: header + records.  Every rank holds four of them (from the lower / upper neighbour, two exchange
// parities); the SENDER fills them: predict_key_kernel stores the records straight into the neighbour's mailbox (peer
// memory over NVLink, mapped with CUDA IPC or - several strips in one process - plain device pointers), and
// publish_halo_kernel then writes the count and, with release semantics at system scope, the exchange's sequence
// number.  The receiver polls `seq` (wait_halo_kernel) and files the records (unpack_kernel).  No host in the loop, no
// fixed message size, and the whole step stays one CUDA graph.  Two parities because a sender may run one exchange
// ahead: it writes box e+1 while the receiver still reads box e; it cannot reach e+2 before it has seen the receiver's
// own e+1, which the receiver publishes after it has unpacked e (stream order).
struct HaloBuffer {
	uint32_t count, seq, pad[6]; // 32-byte header
	// HaloRecord records[capacity] follow
};
__device__ __forceinline__ HaloRecord *halo_records(HaloBuffer *b) { return reinterpret_cast<HaloRecord *>(b + 1); }
__device__ __forceinline__ const HaloRecord *halo_records(const HaloBuffer *b) { return reinterpret_cast<const HaloRecord *>(b + 1); }

struct StripDesc {
	int32_t rank, world, halo;  // halo rows on each side
	uint32_t haloCap;           // records per HaloBuffer
	// where the records for the lower / upper neighbour go, by exchange parity: the neighbour's mailbox (peer memory)
	// or, with the NCCL transport, a local send buffer of the same layout
	HaloBuffer *outDown[2], *outUp[2];
	uint32_t *sendCount;        // [0] records packed for the lower neighbour in this exchange, [1] for the upper, [2] peak since the last look
	// Rows this rank owned when the PREVIOUS grid was built: what it is authoritative for.  Equal to GridDesc's
	// ownLo/ownHi except in the one step that re-balances the strips: there the particles are kept / sent by the
	// new rows (GridDesc) while authority still follows the old ones (sph_set_rebalance).
	int32_t authLo, authHi;
};

// shared tail of predict_key / unpack: claim a rank inside the cell, one atomic per distinct cell per warp
__device__ __forceinline__ uint32_t claim_cell_rank(uint32_t key, bool valid, uint32_t *__restrict__ cellCount) {
	const uint32_t peers = __match_any_sync(0xffffffffu, key);
	const int leader = __ffs(peers) - 1;
	uint32_t base = 0;
	if (valid && (int)lane_id() == leader) base = atomicAdd(&cellCount[key], (uint32_t)__popc(peers));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + (uint32_t)__popc(peers & ((1u << lane_id()) - 1u));
}

// one particle of predict_key_kernel: predict (demo4.cpp:330-339), then decide where it is filed / sent
// (toDown / toUp: a copy goes to that neighbour's mailbox; every lane of the warp returns, the caller ships warp-wide)
__device__ __forceinline__ void predict_file_one(const GridDesc &g, const StripDesc &sd, Counters *__restrict__ ctr, bool in, uint32_t i, uint32_t nSorted,
                                                 float2 &p, float2 &q, float2 v, const uint32_t *__restrict__ cellOld, float dt,
                                                 int doPredict, uint32_t &key, uint32_t &packed, bool &toDown, bool &toUp) {
	key = 0xFFFFFF00u | lane_id(); // unique per lane: never matches
	packed = SPH_KEY_NONE;
	toDown = toUp = false;
	if (!in) return;
	if (doPredict) {
		q = p; // prevPosition = curPosition
		p.x = __fadd_rn(__fmul_rn(v.x, dt), p.x);
		p.y = __fadd_rn(__fmul_rn(v.y, dt), p.y);
	}
	bool authoritative = true;
	if (sd.world > 1 && i < nSorted) { // ghosts of the previous grid are not ours to file
		const int rowOld = (int)(cellOld[i] >> 16);
		authoritative = rowOld >= sd.authLo && rowOld < sd.authHi;
	}
	if (!authoritative) return;
	int cx, cy;
	cell_of(g, p, cx, cy);
	bool placed = false;
	if (cy >= g.rowLo && cy < g.rowHi) {
		key = (uint32_t)(cy - g.rowLo) * (uint32_t)g.gx + (uint32_t)cx;
		packed = pack_cell(cx, cy);
		placed = true;
	}
	if (sd.world > 1) {
		toDown = sd.rank > 0 && cy < g.ownLo + sd.halo;
		toUp = sd.rank + 1 < sd.world && cy >= g.ownHi - sd.halo;
		placed = placed || toDown || toUp;
	}
	if (!placed) atomicAdd(&ctr->lost, 1u);
}

// Warp-wide: the lanes with `want` claim consecutive slots of the exchange's out-box (one atomic per warp on the LOCAL
// counter) and store their records there - for a peer mailbox these are posted writes over NVLink, 24 contiguous
// bytes per lane.  Slots past the capacity are dropped; publish_halo_kernel reports them.
__device__ __forceinline__ void ship_records(bool want, HaloBuffer *box, uint32_t *counter, uint32_t cap, float2 p, float2 q, uint32_t id) {
	const uint32_t mask = __ballot_sync(0xffffffffu, want);
	if (!mask) return;
	const int leader = __ffs(mask) - 1;
	uint32_t base = 0;
	if ((int)lane_id() == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
	base = __shfl_sync(0xffffffffu, base, leader);
	if (want) {
		const uint32_t at = base + (uint32_t)__popc(mask & ((1u << lane_id()) - 1u));
		if (at < cap) {
			HaloRecord rec;
			rec.pos = p;
			rec.prev = q;
			rec.id = id;
			rec.pad = 0;
			halo_records(box)[at] = rec;
		}
	}
}

// A thread takes the particle pair (2q, 2q+1): one 16-byte access per float2 array (see integrate_kernel).
__global__ void __launch_bounds__(SPH_THREADS) predict_key_kernel(GridDesc g, StripDesc sd, Counters *__restrict__ ctr, float2 *__restrict__ pos,
                                                                 float2 *__restrict__ prev, const float2 *__restrict__ vel,
                                                                 const uint32_t *__restrict__ id, const uint32_t *__restrict__ cellOld,
                                                                 uint32_t *__restrict__ cellNew, uint32_t *__restrict__ rank,
                                                                 uint32_t *__restrict__ cellCount, float dt, int doPredict) {
	const uint32_t n = ctr->n, nSorted = ctr->nSorted;
	const uint32_t nPairs = (n + 1u) >> 1;
	// this exchange's number is xseq + 1 (publish_halo_kernel, which runs after this kernel, makes it current)
	const uint32_t par = sd.world > 1 ? ((ctr->xseq + 1u) & 1u) : 0u;
	HaloBuffer *const boxDown = sd.outDown[par], *const boxUp = sd.outUp[par];
	SPH_WARP_LOOP(qi, nPairs) {
		const bool in0 = 2u * qi < n, in1 = 2u * qi + 1u < n; // in1 implies in0
		Pair2 p = {}, q = {}, v = {};
		if (in0) {
			p = load_pair(pos, qi, in1);
			q = load_pair(prev, qi, in1);
			if (doPredict) v = load_pair(vel, qi, in1);
		}
		uint32_t key0, packed0, key1, packed1;
		bool down0, up0, down1, up1;
		predict_file_one(g, sd, ctr, in0, 2u * qi, nSorted, p.a, q.a, v.a, cellOld, dt, doPredict, key0, packed0, down0, up0);
		predict_file_one(g, sd, ctr, in1, 2u * qi + 1u, nSorted, p.b, q.b, v.b, cellOld, dt, doPredict, key1, packed1, down1, up1);
		if (in0 && doPredict) {
			store_pair(prev, qi, in1, q);
			store_pair(pos, qi, in1, p);
		}
		if (sd.world > 1 && __any_sync(0xffffffffu, down0 || up0 || down1 || up1)) {
			const uint32_t id0 = (down0 || up0) ? id[2u * qi] : 0u, id1 = (down1 || up1) ? id[2u * qi + 1u] : 0u;
			ship_records(down0, boxDown, sd.sendCount + 0, sd.haloCap, p.a, q.a, id0);
			ship_records(down1, boxDown, sd.sendCount + 0, sd.haloCap, p.b, q.b, id1);
			ship_records(up0, boxUp, sd.sendCount + 1, sd.haloCap, p.a, q.a, id0);
			ship_records(up1, boxUp, sd.sendCount + 1, sd.haloCap, p.b, q.b, id1);
		}
		const uint32_t r0 = claim_cell_rank(key0, packed0 != SPH_KEY_NONE, cellCount);
		const uint32_t r1 = claim_cell_rank(key1, packed1 != SPH_KEY_NONE, cellCount);
		if (in1) {
			*reinterpret_cast<uint2 *>(cellNew + 2u * qi) = make_uint2(packed0, packed1);
			*reinterpret_cast<uint2 *>(rank + 2u * qi) = make_uint2(r0, r1);
		} else if (in0) {
			cellNew[2u * qi] = packed0;
			rank[2u * qi] = r0;
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) ctr->nIn = n;
}

// ---- the strip exchange: publish, wait, unpack ------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
	uint32_t v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

// Runs after predict_key_kernel (all records of this exchange are stored): writes the counts into the out-boxes and
// then, released at system scope, the exchange number - the receiver's cue.  Resets the local slot counters for the
// next exchange and makes the exchange current (ctr->xseq).  One thread.
__global__ void publish_halo_kernel(StripDesc sd, Counters *__restrict__ ctr) {
	const uint32_t e = ctr->xseq + 1u, par = e & 1u;
	HaloBuffer *box[2] = { sd.outDown[par], sd.outUp[par] };
	uint32_t most = 0;
#pragma unroll
	for (int d = 0; d < 2; ++d) {
		const uint32_t c = sd.sendCount[d];
		sd.sendCount[d] = 0u;
		if (!box[d]) continue;
		if (c > sd.haloCap) atomicOr(&ctr->overflow, 2u);
		most = max(most, c);
		box[d]->count = min(c, sd.haloCap);
	}
	if (most > sd.sendCount[2]) sd.sendCount[2] = most;
	__threadfence_system(); // the records (previous kernel) and the counts before the sequence numbers
#pragma unroll
	for (int d = 0; d < 2; ++d)
		if (box[d]) st_release_sys(&box[d]->seq, e);
	ctr->xseq = e;
}

// Lane 0 waits for the lower neighbour's mailbox of the current exchange, lane 1 for the upper one.  A neighbour that
// does not publish within `timeoutNs` (it died, or its host stopped stepping) must not hang this GPU: the wait gives
// up, raises overflow bit 8 - sph_get_stats turns it into SPH_ERR_COMM - and later waits return at once.
__global__ void wait_halo_kernel(Counters *__restrict__ ctr, const HaloBuffer *down0, const HaloBuffer *down1, const HaloBuffer *up0, const HaloBuffer *up1,
                                 unsigned long long timeoutNs) {
	const uint32_t e = ctr->xseq, par = e & 1u;
	const HaloBuffer *box = threadIdx.x == 0 ? (par ? down1 : down0) : (threadIdx.x == 1 ? (par ? up1 : up0) : nullptr);
	if (!box || (ctr->overflow & 8u)) return;
	const unsigned long long t0 = global_timer_ns();
	// sequence numbers only grow; (int) difference so that a wrap after 2^32 exchanges stays harmless
	while ((int32_t)(ld_acquire_sys(&box->seq) - e) < 0) {
		__nanosleep(200);
		if (global_timer_ns() - t0 > timeoutNs) {
			atomicOr(&ctr->overflow, 8u);
			break;
		}
	}
}

// append the records received from one neighbour behind the local particles and file them; the mailbox is the one of
// the current exchange's parity, `before` (same parity) is the neighbour unpacked ahead of this one
__global__ void __launch_bounds__(SPH_THREADS) unpack_kernel(GridDesc g, Counters *__restrict__ ctr, const HaloBuffer *in0, const HaloBuffer *in1, uint32_t haloCap,
                                                            uint32_t capacity, const HaloBuffer *before0, const HaloBuffer *before1,
                                                            float2 *__restrict__ pos, float2 *__restrict__ prev, uint32_t *__restrict__ id,
                                                            uint32_t *__restrict__ cellNew, uint32_t *__restrict__ rank, uint32_t *__restrict__ cellCount) {
	const uint32_t e = ctr->xseq, par = e & 1u;
	const HaloBuffer *in = par ? in1 : in0, *before = par ? before1 : before0;
	if ((int32_t)(in->seq - e) < 0) return; // the wait timed out: nothing arrived (already reported)
	// records land at [n + (count of the buffer unpacked before this one), ...)
	if (blockIdx.x == 0 && threadIdx.x == 0 && in->count > haloCap) atomicOr(&ctr->overflow, 2u); // the sender packed more than a message ships
	const uint32_t count = min(in->count, haloCap);
	const uint32_t first = ctr->n + ((before && (int32_t)(before->seq - e) >= 0) ? min(before->count, haloCap) : 0u);
	SPH_WARP_LOOP(k, count) {
		const bool ok = k < count && first + k < capacity;
		uint32_t key = 0xFFFFFF00u | lane_id();
		uint32_t packed = SPH_KEY_NONE;
		if (ok) {
			const HaloRecord rec = halo_records(in)[k];
			int cx, cy;
			cell_of(g, rec.pos, cx, cy);
			if (cy >= g.rowLo && cy < g.rowHi) {
				key = (uint32_t)(cy - g.rowLo) * (uint32_t)g.gx + (uint32_t)cx;
				packed = pack_cell(cx, cy);
			}
			pos[first + k] = rec.pos;
			prev[first + k] = rec.prev;
			id[first + k] = rec.id;
		} else if (k < count) {
			atomicOr(&ctr->overflow, 1u);
		}
		const uint32_t r = claim_cell_rank(key, packed != SPH_KEY_NONE, cellCount);
		if (ok) {
			cellNew[first + k] = packed;
			rank[first + k] = r;
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		const uint32_t end = min(first + count, capacity);
		atomicMax(&ctr->nIn, end);
	}
}

// Particles per grid row of the rows this rank owns (read off the previous grid's prefix), written at the row's GLOBAL
// index, and the rank's first row at out[gy + rank]: summed over the ranks (all-reduce) every rank sees the whole
// histogram and all boundaries, and plans the same new split (plan_strip_bounds).
__global__ void row_counts_kernel(GridDesc g, const uint32_t *__restrict__ cellStart, uint32_t *__restrict__ out, int rank) {
	const int r = g.ownLo + (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (r < g.ownHi) {
		const uint32_t local = (uint32_t)(r - g.rowLo);
		out[r] = cellStart[(local + 1u) * (uint32_t)g.gx] - cellStart[local * (uint32_t)g.gx];
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) out[g.gy + rank] = (uint32_t)g.ownLo;
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

// 16 consecutive cells of one thread as four 16-byte accesses (the cell arrays are 256-byte aligned and a tile starts at
// a multiple of 4096 cells); cells past the end read as empty
__device__ __forceinline__ void load_cells16(const uint32_t *__restrict__ src, uint32_t first, uint32_t nCells, uint32_t v[SPH_SCAN_ITEMS]) {
	if (first + SPH_SCAN_ITEMS <= nCells) {
#pragma unroll
		for (int q = 0; q < SPH_SCAN_ITEMS / 4; ++q) {
			const uint4 x = *reinterpret_cast<const uint4 *>(src + first + 4 * q);
			v[4 * q] = x.x;
			v[4 * q + 1] = x.y;
			v[4 * q + 2] = x.z;
			v[4 * q + 3] = x.w;
		}
	} else {
#pragma unroll
		for (int k = 0; k < SPH_SCAN_ITEMS; ++k) v[k] = (first + k < nCells) ? src[first + k] : 0u;
	}
}

// Two launches (was three, one of them a single block): the tiles' sums, then every tile adds up the sums of the tiles
// before it by itself (at most a few thousand words, 256 threads) and scans its cells behind that offset.
__global__ void __launch_bounds__(SPH_THREADS) scan_tile_sums_kernel(const uint32_t *__restrict__ cellCount, uint32_t *__restrict__ tileSums, uint32_t nCells) {
	uint32_t v[SPH_SCAN_ITEMS], sum = 0;
	load_cells16(cellCount, blockIdx.x * SPH_SCAN_TILE + threadIdx.x * SPH_SCAN_ITEMS, nCells, v);
#pragma unroll
	for (int k = 0; k < SPH_SCAN_ITEMS; ++k) sum += v[k];
	__shared__ uint32_t warpSums[SPH_THREADS / 32];
	sum = warp_sum(sum);
	if (lane_id() == 0) warpSums[threadIdx.x >> 5] = sum;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t total = 0;
#pragma unroll
		for (int w = 0; w < SPH_THREADS / 32; ++w) total += warpSums[w];
		tileSums[blockIdx.x] = total;
	}
}

__global__ void __launch_bounds__(SPH_THREADS) scan_apply_kernel(const uint32_t *__restrict__ cellCount, const uint32_t *__restrict__ tileSums,
                                                                uint32_t *__restrict__ cellStart, uint32_t nCells, Counters *__restrict__ ctr) {
	// sum of the tiles before this one
	__shared__ uint32_t warpPart[SPH_THREADS / 32];
	uint32_t before = 0;
	for (uint32_t j = threadIdx.x; j < blockIdx.x; j += SPH_THREADS) before += tileSums[j];
	before = warp_sum(before);
	if (lane_id() == 0) warpPart[threadIdx.x >> 5] = before;
	const uint32_t first = blockIdx.x * SPH_SCAN_TILE + threadIdx.x * SPH_SCAN_ITEMS;
	uint32_t v[SPH_SCAN_ITEMS], sum = 0;
	load_cells16(cellCount, first, nCells, v);
#pragma unroll
	for (int k = 0; k < SPH_SCAN_ITEMS; ++k) sum += v[k];
	__syncthreads();
	uint32_t offset = 0;
#pragma unroll
	for (int w = 0; w < SPH_THREADS / 32; ++w) offset += warpPart[w];
	uint32_t total;
	uint32_t run = offset + block_exclusive_scan(sum, &total);
	if (first + SPH_SCAN_ITEMS <= nCells) {
#pragma unroll
		for (int q = 0; q < SPH_SCAN_ITEMS / 4; ++q) {
			uint4 x;
			x.x = run;
			run += v[4 * q];
			x.y = run;
			run += v[4 * q + 1];
			x.z = run;
			run += v[4 * q + 2];
			x.w = run;
			run += v[4 * q + 3];
			*reinterpret_cast<uint4 *>(cellStart + first + 4 * q) = x;
		}
	} else {
#pragma unroll
		for (int k = 0; k < SPH_SCAN_ITEMS; ++k) {
			if (first + k < nCells) cellStart[first + k] = run;
			run += v[k];
		}
	}
	if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { // the last tile knows the total
		cellStart[nCells] = offset + total;
		ctr->nOut = offset + total;
	}
}

// ---- colour lists in row-major order ---------------------------------------------------------------
// The one-launch sweep (color_sweep_flow_kernel) hands cells out in list order and a cell waits for its
// lower-colour neighbours; if a colour's list walks the grid front to back, the cells a newcomer depends on
// were handed out a whole colour earlier and nobody ever waits.  Lists appended with atomics are shuffled over
// a quarter of the domain, so these three kernels build them sorted instead: one warp per (grid row, cx mod 3)
// counts its occupied cells, one block turns the counts into list offsets (rows of equal colour in ascending
// order), and the same warps write the cells behind their offset.  Deterministic as a side effect.
//
// LIGHT and HEAVY cells.  A cell's sweep is a serial chain over its particles, and the nine colours are a serial
// chain over the cells of a neighbourhood: where a column of fluid compresses (cells with 40-80 particles and 400
// candidates next to the usual 9 and 81) one warp per cell leaves the GPU waiting for nine heavy cells in a row.
// The lists therefore come in two classes per colour: LIGHT cells (the 3x3 block fits one warp's staging area and its
// work m x T is below `workHeavy`) are swept one warp per cell, HEAVY cells by a whole thread block per cell
// (sweep_cell_team).  Same arithmetic, same bits (tests compare every kernel with the oracle).  The light list of a
// colour grows from the front of its region of `colorList`, the heavy list from the back; colorCount[0..8] are the
// light counts, colorCount[9..17] the heavy ones.  The class is decided once, in color_rows_count_kernel, which
// leaves it in bit 31 of the cell's histogram word (cellCount is not read again after the scan).
#define SPH_ROWLIST_WARPS 6 // two rows (x three column classes) per block
#define SPH_CELL_HEAVY 0x80000000u

struct SweepClass {
	uint32_t capLight;  // a light cell's padded candidate count fits this (the per-warp staging capacity of the sweeps)
	uint32_t workHeavy; // m x T from which a cell is heavy whatever its size
};

// Also resets the done flags of the one-launch sweep for this grid: 0 = occupied, not swept in any pass yet;
// SPH_FLOW_EMPTY = nothing to wait for, ever.  (flow[0..3] are the ticket counters, the flags start at flow + 4.)
#define SPH_FLOW_EMPTY 0xFFFFFFFFu
#define SPH_FLOW_FLAGS 4u
__global__ void __launch_bounds__(SPH_ROWLIST_WARPS * 32) color_rows_count_kernel(GridDesc g, SweepClass cls, uint32_t *__restrict__ cellCount,
                                                                                const uint32_t *__restrict__ cellStart, uint32_t *__restrict__ rowColor,
                                                                                uint32_t *__restrict__ colorCount, uint32_t *__restrict__ flow) {
	const uint32_t wid = blockIdx.x * SPH_ROWLIST_WARPS + (threadIdx.x >> 5), lane = lane_id();
	const uint32_t nRows = (uint32_t)(g.rowHi - g.rowLo);
	if (blockIdx.x == 0 && threadIdx.x < 18) colorCount[threadIdx.x] = 0u; // color_rows_fill_kernel writes the totals of the colours that have rows
	if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 36) flow[threadIdx.x - 32] = 0u; // ticket counters (light, heavy) of the next two passes over this grid
	if (wid >= nRows * 3u) return;
	const uint32_t row = wid / 3u, a = wid - row * 3u;
	uint32_t *cc = cellCount + (size_t)row * (uint32_t)g.gx;
	uint32_t *flags = flow + SPH_FLOW_FLAGS + (size_t)row * (uint32_t)g.gx;
	uint32_t nLight = 0, nHeavy = 0;
	for (uint32_t c0 = a + 3u * lane; c0 < (uint32_t)g.gx; c0 += 4u * 96u) { // four loads in flight per round trip
		uint32_t v[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) v[u] = (c0 + 96u * (uint32_t)u < (uint32_t)g.gx) ? cc[c0 + 96u * (uint32_t)u] : 0u;
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const uint32_t cx = c0 + 96u * (uint32_t)u;
			if (cx < (uint32_t)g.gx) {
				flags[cx] = v[u] ? 0u : SPH_FLOW_EMPTY;
				if (v[u]) { // candidates of the 3x3 block, read off the prefix (only occupied cells pay for this)
					const uint32_t x0 = cx ? cx - 1u : 0u, x1 = min(cx + 1u, (uint32_t)g.gx - 1u) + 1u;
					uint32_t T = 0;
#pragma unroll
					for (int dr = -1; dr <= 1; ++dr) {
						const int y = (int)row + dr;
						if (y >= 0 && y < (int)nRows) T += cellStart[(uint32_t)y * (uint32_t)g.gx + x1] - cellStart[(uint32_t)y * (uint32_t)g.gx + x0];
					}
					const bool heavy = ((T + 31u) & ~31u) > cls.capLight || (unsigned long long)v[u] * T >= cls.workHeavy;
					if (heavy) cc[cx] = v[u] | SPH_CELL_HEAVY;
					nHeavy += heavy ? 1u : 0u;
					nLight += heavy ? 0u : 1u;
				}
			}
		}
	}
	nLight = warp_sum(nLight);
	nHeavy = warp_sum(nHeavy);
	if (lane == 0) {
		rowColor[wid] = nLight;
		rowColor[nRows * 3u + wid] = nHeavy;
	}
}

__global__ void __launch_bounds__(SPH_ROWLIST_WARPS * 32) color_rows_fill_kernel(GridDesc g, const uint32_t *__restrict__ cellCount, const uint32_t *__restrict__ rowColor,
                                                                               uint32_t *__restrict__ colorList, uint32_t listStride,
                                                                               uint32_t *__restrict__ colorCount) {
	const uint32_t wid = blockIdx.x * SPH_ROWLIST_WARPS + (threadIdx.x >> 5), lane = lane_id();
	const uint32_t nRows = (uint32_t)(g.rowHi - g.rowLo);
	if (wid >= nRows * 3u) return;
	const uint32_t row = wid / 3u, a = wid - row * 3u;
	const uint32_t k = ((row + (uint32_t)g.rowLo) % 3u) * 3u + a;
	const uint32_t *cc = cellCount + (size_t)row * (uint32_t)g.gx;
	uint32_t *list = colorList + (size_t)k * listStride;
	// Where this row's cells start in its colour's lists: the counts of the rows of the same colour below it (rows
	// row-3, row-6, ...), summed by the warp itself - a few hundred words instead of a single-block scan launch.  The
	// colour's topmost row also knows the totals.
	uint32_t at = 0, atHeavy = 0;
	for (int r = (int)row - 3 * (int)(lane + 1u); r >= 0; r -= 96) {
		at += rowColor[(uint32_t)r * 3u + a];
		atHeavy += rowColor[nRows * 3u + (uint32_t)r * 3u + a];
	}
	at = warp_sum(at);
	atHeavy = warp_sum(atHeavy);
	if (row + 3u >= nRows && lane == 0) {
		colorCount[k] = at + rowColor[wid];
		colorCount[9u + k] = atHeavy + rowColor[nRows * 3u + wid];
	}
	for (uint32_t c0 = a; c0 < (uint32_t)g.gx; c0 += 4u * 96u) { // warp-uniform trip count, four loads in flight per round trip
		uint32_t v[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const uint32_t cx = c0 + 96u * (uint32_t)u + 3u * lane;
			v[u] = cx < (uint32_t)g.gx ? cc[cx] : 0u;
		}
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const uint32_t cx = c0 + 96u * (uint32_t)u + 3u * lane;
			const bool heavy = (v[u] & SPH_CELL_HEAVY) != 0u, light = v[u] != 0u && !heavy;
			const uint32_t mask = __ballot_sync(0xffffffffu, light), maskH = __ballot_sync(0xffffffffu, heavy);
			if (light) list[at + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = row * (uint32_t)g.gx + cx;
			if (heavy) list[listStride - 1u - (atHeavy + (uint32_t)__popc(maskH & ((1u << lane) - 1u)))] = row * (uint32_t)g.gx + cx;
			at += (uint32_t)__popc(mask);
			atHeavy += (uint32_t)__popc(maskH);
		}
	}
}

// cell number idx of one colour: the light list from the front, then the heavy list from the back (the per-colour
// kernels sweep both; the order inside a colour is free)
__device__ __forceinline__ uint32_t colour_cell(const uint32_t *__restrict__ list, uint32_t listStride, uint32_t nLight, uint32_t idx) {
	return idx < nLight ? list[idx] : list[listStride - 1u - (idx - nLight)];
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
		const float2 p = pos[i], q = prev[i]; // issued before the ranking loop so their latency overlaps it
		uint32_t r = 0, s = lo;
		for (; s + 4 <= hi; s += 4) { // four independent loads in flight
			const uint32_t a0 = slotId[s], a1 = slotId[s + 1], a2 = slotId[s + 2], a3 = slotId[s + 3];
			r += (a0 < me ? 1u : 0u) + (a1 < me ? 1u : 0u) + (a2 < me ? 1u : 0u) + (a3 < me ? 1u : 0u);
		}
		for (; s < hi; ++s) r += (slotId[s] < me) ? 1u : 0u;
		const uint32_t dst = lo + r;
		posOut[dst] = p;
		prevOut[dst] = q;
		idOut[dst] = me;
		cellOut[dst] = c;
		occMin = min(occMin, hi - lo);
		occMax = max(occMax, hi - lo);
	}
	// Running min/max of the cell occupancy (demo4.cpp:64-67).  One atomic per warp on the same two words is
	// 65 000 serialised L2 atomics per step at 1M particles - they alone took 50 of this kernel's 59 us (ncu, r1e).
	// Both values converge after a few warps, so look first (a stale value only costs a redundant atomic).
	occMin = warp_min(occMin);
	occMax = warp_max(occMax);
	if (lane_id() == 0 && occMax) {
		if (occMin < __ldcg(&ctr->minCell)) atomicMin(&ctr->minCell, occMin);
		if (occMax > __ldcg(&ctr->maxCell)) atomicMax(&ctr->maxCell, occMax);
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
	// (looked at before the atomic for the same reason as in reorder_kernel; the sum goes through one word per block)
	__shared__ unsigned long long blockSum;
	if (threadIdx.x == 0) blockSum = 0ull;
	__syncthreads();
	cMin = warp_min(cMin);
	cMax = warp_max(cMax);
	cSum = warp_sum(cSum);
	if (lane_id() == 0 && cMax) {
		if (cMin < __ldcg(&ctr->minNbr)) atomicMin(&ctr->minNbr, cMin);
		if (cMax > __ldcg(&ctr->maxNbr)) atomicMax(&ctr->maxNbr, cMax);
		atomicAdd(&blockSum, (unsigned long long)cSum);
	}
	__syncthreads();
	if (threadIdx.x == 0 && blockSum) atomicAdd(&ctr->pairCandidates, blockSum);
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

// ---- 9-colour cell Gauss-Seidel: the reference's in-place pair sweeps, race-free ------------------
// demo4.cpp:223-255 updates BOTH particles of every pair in place, in particle-index order.  That
// order is what keeps the double-density relaxation stable (each pair sees the displacements of
// the pairs before it); a plain gather of all pair terms from the old state overshoots and blows
// up on stiff scenes ("why not Jacobi").  Cells whose (cx mod 3, cy mod 3) agree have
// disjoint 3x3 footprints, so the nine colours are swept in order (nine launches, or one launch with per-cell dependency flags: color_sweep_flow_kernel); inside a
// launch ONE WARP owns one cell, walks its particles by ascending id, and spreads each particle's
// candidate loop over its 32 lanes (candidate k belongs to lane k mod 32 for the whole sweep, so a
// staged candidate is only ever written by one lane).  The partner is updated at once, the
// particle's own change is summed per lane, combined by a xor butterfly and applied at the end
// of its loop (demo4.cpp:253).  The 3x3 block is staged in shared memory; blocks larger than the
// staging capacity take the same algorithm through L2 (__ldcg/__stcg).
#define SPH_SWEEP_WARPS 4

enum { SWEEP_DELTA = 0, SWEEP_VISCOSITY = 1 };

// CHECK = false: the caller already knows the pair is within h (stage 1 of the sweeps tested exactly these
// positions), so the range test of sph.h:488 is not repeated.
template <class M, bool CHECK = true>
__device__ __forceinline__ float2 sweep_delta_term(const PairParams &k, float2 xi, float2 ppi, float2 xj, bool &hit) {
	// SPHComputeDelta, sph.h:483-495, then * 0.5f (demo4.cpp:250-251); (x, y) pairs as packed operations (sph_math.cuh)
	const float2 rel = M::sub2(xj, xi);
	float r2 = M::norm2(rel);
	hit = CHECK ? (r2 < k.h2) : true;
	if (!hit) return make_float2(0.0f, 0.0f);
	float r, inv;
	M::len_inv(r2, r, inv);
	float t = M::sub(1.0f, M::mul(r, k.invH));
	float d = M::mul(k.dt2, M::add(M::mul(ppi.x, t), M::mul(ppi.y, M::mul(t, t))));
	return M::scale2(M::scale2(M::scale2(rel, inv), d), 0.5f); // (d * (rel * inv)) * 0.5
}

template <class M, bool CHECK = true>
__device__ __forceinline__ float2 sweep_viscosity_term(const PairParams &k, float2 xi, float2 vi, float2 xj, float2 vj, bool &hit) {
	// SPHComputeViscosityForce, sph.h:497-512, then * 0.5f * deltaTime (demo4.cpp:233-234)
	hit = false;
	const float2 rel = M::sub2(xj, xi);
	float r2 = M::norm2(rel);
	if (CHECK && !(r2 < k.h2)) return make_float2(0.0f, 0.0f);
	float r, inv;
	M::len_inv(r2, r, inv);
	float q = M::mul(r, k.invH);
	const float2 n = M::scale2(rel, inv);
	const float2 proj = M::mul2(M::sub2(vi, vj), n);
	float u = M::add(proj.x, proj.y); // (vi - vj) . n: two products, one sum
	if (!(u > 0.0f)) return make_float2(0.0f, 0.0f);
	hit = true;
	float f = M::mul(M::sub(1.0f, q), M::add(M::mul(k.sigma, u), M::mul(k.beta, M::mul(u, u))));
	return M::scale2(M::scale2(M::scale2(n, f), 0.5f), k.dt); // ((f * n) * 0.5) * dt
}

__device__ __forceinline__ float butterfly_sum(float v) {
#pragma unroll
	for (int o = 16; o; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
// Both components of a particle's own change with 6 shuffles instead of 10: after the first exchange
// lanes 0-15 carry x and lanes 16-31 carry y, each half then runs the remaining four butterfly stages.
// The additions are the ones lane 0 performs in butterfly_sum (v_l + v_(l^o), same order of stages), so
// the sums are bit-identical to it.  Valid in lane 0 only.
__device__ __forceinline__ float2 butterfly_sum2_lane0(float x, float y, uint32_t lane) {
	const bool lower = lane < 16u;
	const float other = __shfl_xor_sync(0xffffffffu, lower ? y : x, 16);
	float v = __fadd_rn(lower ? x : y, other);
#pragma unroll
	for (int o = 8; o; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
	const float ysum = __shfl_sync(0xffffffffu, v, 16);
	return make_float2(v, ysum);
}

// Per particle i of the cell the candidate loop runs in two stages.  Stage 1 tests every candidate
// against h (5 flops) and compacts the ones in range into a queue (ballot + popc); stage 2 evaluates
// only those, with all 32 lanes busy: the h-th candidate in range belongs to lane h mod 32.  About a
// third of the candidates are in range, so this removes most of the divergence of the expensive
// part (sqrt, 1/r, the pair term).
//
// Shared memory per warp: sPos[cap] (+ sVel[cap] for viscosity) + queue[cap] (uint16; two of them for viscosity, whose
// range tests run for two particles at a time while that costs no occupancy, sweep_paired).  A block with
// more than `cap` candidates is not staged: it reads and writes the sorted arrays through L2 and uses
// the whole per-warp area as its queue, which bounds it at sweep_queue_capacity(cap) candidates
// (the reference asserts at 1000, demo4.cpp:199).
#define SPH_PAIR_MAX_CAP 256u // up to this staging capacity the second queue is free: registers, not shared memory, limit the resident blocks
__host__ __device__ inline bool sweep_paired(uint32_t cap, int pass) { return pass == SWEEP_VISCOSITY && cap <= SPH_PAIR_MAX_CAP; }
__host__ __device__ inline uint32_t sweep_bytes_per_warp(uint32_t cap, int pass) {
	return cap * 8u * (pass == SWEEP_VISCOSITY ? 2u : 1u) + cap * 2u * (sweep_paired(cap, pass) ? 2u : 1u);
}
__host__ __device__ inline uint32_t sweep_queue_capacity(uint32_t cap, int pass) { return sweep_bytes_per_warp(cap, pass) / 2u; }

// One cell, one warp.  STAGED: the block's candidates live in shared memory (sPos/sVel, padded to a
// multiple of 32 with far-away sentinels so stage 1 needs no bounds test); otherwise they are read
// and written in the sorted arrays through L2.
struct SweepBlock {
	uint32_t lo0, lo1, lo2, off1, off2, T; // three slabs of the sorted arrays and their offsets among the candidates
	uint32_t ownLo, m, ownOff;             // the cell's own particles: first sorted index, count, first candidate slot
	__device__ __forceinline__ uint32_t gidx(uint32_t t) const { return t < off1 ? lo0 + t : (t < off2 ? lo1 + (t - off1) : lo2 + (t - off2)); }
};

// COHERENT: the staging loads bypass L1 (ld.global.cg) because another SM may have rewritten the
// block earlier in the SAME launch (color_sweep_flow_kernel); the per-colour kernels read through L1.
// beforeWriteBack() runs once, after the last particle of the cell and before the block is written back.
struct SweepNoHook {
	__device__ __forceinline__ void operator()() const {}
};
template <class M, int PASS, bool STAGED, bool COHERENT = false, class Hook = SweepNoHook>
__device__ __forceinline__ void sweep_cell(const PairParams &k, const SweepBlock &b, float2 *pos, float2 *vel, const float2 *__restrict__ press,
                                           float2 *sPos, float2 *sVel, uint16_t *queue, uint32_t lane, uint32_t ltMask, Hook beforeWriteBack = Hook(),
                                           uint16_t *queueB = nullptr) {
	float2 *state = (PASS == SWEEP_DELTA) ? pos : vel;
	const uint32_t Tpad = (b.T + 31u) & ~31u;
	if (STAGED) {
		if (COHERENT) {
			// four trips of loads in flight before the first one is consumed: the block comes from L2 (~1 us away
			// under load) and a warp has nothing else to do until it is staged
			for (uint32_t t0 = lane; t0 < Tpad; t0 += 128) {
				float2 p[4], v[4];
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					const uint32_t t = t0 + 32u * (uint32_t)u;
					p[u] = make_float2(3.0e18f, 3.0e18f); // never within h of anything
					v[u] = make_float2(0.0f, 0.0f);
					if (t < b.T) {
						const uint32_t j = b.gidx(t);
						p[u] = __ldcg(&pos[j]);
						if (PASS == SWEEP_VISCOSITY) v[u] = __ldcg(&vel[j]);
					}
				}
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					const uint32_t t = t0 + 32u * (uint32_t)u;
					if (t < Tpad) {
						sPos[t] = p[u];
						if (PASS == SWEEP_VISCOSITY) sVel[t] = v[u];
					}
				}
			}
		} else {
			for (uint32_t t = lane; t < Tpad; t += 32) {
				if (t < b.T) {
					const uint32_t j = b.gidx(t);
					sPos[t] = pos[j];
					if (PASS == SWEEP_VISCOSITY) sVel[t] = vel[j];
				} else {
					sPos[t] = make_float2(3.0e18f, 3.0e18f); // never within h of anything
				}
			}
		}
		__syncwarp();
	}
	// stage 2 of particle `ki` of the cell (slot si among the candidates): the pair terms of the queued candidates,
	// partner updated at once (demo4.cpp:233-234, 250-251), then the particle's own change (demo4.cpp:253)
	auto stage2 = [&](uint32_t ki, uint32_t si, float2 xi, float2 ppi, const uint16_t *q16, uint32_t nHit) {
		float2 vi = make_float2(0.0f, 0.0f);
		if (PASS == SWEEP_VISCOSITY) vi = STAGED ? sVel[si] : __ldcg(&vel[b.ownLo + ki]);
		float2 acc = make_float2(0.0f, 0.0f);
		for (uint32_t q = lane; q < nHit; q += 32) {
			const uint32_t t = q16[q];
			bool hit;
			if (PASS == SWEEP_DELTA) {
				float2 *slot = STAGED ? &sPos[t] : &pos[b.gidx(t)];
				const float2 xj = STAGED ? *slot : __ldcg(slot);
				const float2 hlf = sweep_delta_term<M, false>(k, xi, ppi, xj, hit); // queued = within h, nothing moved it since stage 1
				const float2 moved = Exact::add2(hlf, xj);
				if (STAGED) *slot = moved;
				else __stcg(slot, moved);
				acc = Exact::sub2(acc, hlf);
			} else {
				const uint32_t j = STAGED ? 0u : b.gidx(t);
				float2 *slot = STAGED ? &sVel[t] : &vel[j];
				const float2 vj = STAGED ? *slot : __ldcg(slot);
				const float2 xj = STAGED ? sPos[t] : __ldcg(&pos[j]);
				const float2 hlf = sweep_viscosity_term<M, false>(k, xi, vi, xj, vj, hit);
				if (hit) {
					const float2 moved = Exact::add2(hlf, vj);
					if (STAGED) *slot = moved;
					else __stcg(slot, moved);
					acc = Exact::sub2(acc, hlf);
				}
			}
		}
		const float2 own = butterfly_sum2_lane0(acc.x, acc.y, lane);
		__syncwarp();
		if (lane == 0) { // curPosition += dx (demo4.cpp:253): dx + cur
			if (STAGED) {
				float2 *slot = (PASS == SWEEP_DELTA) ? &sPos[si] : &sVel[si];
				*slot = make_float2(__fadd_rn(own.x, slot->x), __fadd_rn(own.y, slot->y));
			} else {
				float2 *slot = &state[b.ownLo + ki];
				const float2 cur = __ldcg(slot);
				__stcg(slot, make_float2(__fadd_rn(own.x, cur.x), __fadd_rn(own.y, cur.y)));
			}
		}
		__syncwarp();
	};
	// The viscosity pass never moves a position, so the range tests (stage 1) of two consecutive particles of the cell
	// can share one walk over the staged block - one load per 32 candidates for both - before their pair loops run one
	// after the other; the displacement pass must test each particle against the positions the previous one left.
	// (queueB: the second queue, where sweep_bytes_per_warp reserves one)
	const bool paired = STAGED && PASS == SWEEP_VISCOSITY && queueB != nullptr;
	for (uint32_t kBase = 0; kBase < b.m; kBase += 32) {
		float2 myPress = make_float2(0.0f, 0.0f);
		if (PASS == SWEEP_DELTA && kBase + lane < b.m) myPress = press[b.ownLo + kBase + lane];
		const uint32_t kEnd = min(b.m - kBase, 32u);
		for (uint32_t kk = 0; kk < kEnd; kk += (paired ? 2u : 1u)) {
			const uint32_t si = b.ownOff + kBase + kk; // this particle's slot among the candidates
			const bool two = paired && kk + 1u < kEnd;
			float2 xi, xi2 = make_float2(3.0e18f, 3.0e18f), ppi = make_float2(0.0f, 0.0f);
			if (STAGED) {
				xi = sPos[si];
				if (two) xi2 = sPos[si + 1u];
			} else {
				xi = __ldcg(&pos[b.ownLo + kBase + kk]);
			}
			if (PASS == SWEEP_DELTA) {
				ppi.x = __shfl_sync(0xffffffffu, myPress.x, (int)kk);
				ppi.y = __shfl_sync(0xffffffffu, myPress.y, (int)kk);
			}
			// stage 1: which candidates are within h (sph.h:488,502)
			uint32_t nHit = 0, nHit2 = 0;
			if (STAGED && PASS == SWEEP_VISCOSITY && two) {
#pragma unroll 2
				for (uint32_t tb = 0; tb < Tpad; tb += 32) {
					const uint32_t t = tb + lane;
					const float2 xj = sPos[t];
					const bool hit = M::norm2(M::sub2(xj, xi)) < k.h2, hit2 = M::norm2(M::sub2(xj, xi2)) < k.h2;
					const uint32_t mask = __ballot_sync(0xffffffffu, hit), mask2 = __ballot_sync(0xffffffffu, hit2);
					if (hit) queue[nHit + (uint32_t)__popc(mask & ltMask)] = (uint16_t)t;
					if (hit2) queueB[nHit2 + (uint32_t)__popc(mask2 & ltMask)] = (uint16_t)t;
					nHit += (uint32_t)__popc(mask);
					nHit2 += (uint32_t)__popc(mask2);
				}
			} else {
#pragma unroll 2
				for (uint32_t tb = 0; tb < Tpad; tb += 32) {
					const uint32_t t = tb + lane;
					float2 xj;
					if (STAGED) xj = sPos[t];
					else xj = (t < b.T) ? __ldcg(&pos[b.gidx(t)]) : make_float2(3.0e18f, 3.0e18f);
					const bool hit = M::norm2(M::sub2(xj, xi)) < k.h2;
					const uint32_t mask = __ballot_sync(0xffffffffu, hit);
					if (hit) queue[nHit + (uint32_t)__popc(mask & ltMask)] = (uint16_t)t;
					nHit += (uint32_t)__popc(mask);
				}
			}
			__syncwarp();
			stage2(kBase + kk, si, xi, ppi, queue, nHit);
			if (two) stage2(kBase + kk + 1u, si + 1u, xi2, ppi, queueB, nHit2);
		}
	}
	beforeWriteBack();
	if (STAGED) {
		for (uint32_t t = lane; t < b.T; t += 32) state[b.gidx(t)] = (PASS == SWEEP_DELTA) ? sPos[t] : sVel[t];
		__syncwarp();
	}
}

// the 3x3 block of local cell c
__device__ __forceinline__ SweepBlock sweep_block_of(const GridDesc &g, const uint32_t *__restrict__ cellStart, uint32_t c, int nRows) {
	const int yl = (int)(c / (uint32_t)g.gx), cx = (int)(c - (uint32_t)yl * (uint32_t)g.gx);
	const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.gx - 1);
	uint32_t lo[3], cnt[3];
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		const int y = yl - 1 + r;
		if (y < 0 || y >= nRows) {
			lo[r] = 0;
			cnt[r] = 0;
		} else {
			lo[r] = cellStart[(uint32_t)y * (uint32_t)g.gx + (uint32_t)x0];
			cnt[r] = cellStart[(uint32_t)y * (uint32_t)g.gx + (uint32_t)x1 + 1u] - lo[r];
		}
	}
	SweepBlock b;
	b.lo0 = lo[0];
	b.lo1 = lo[1];
	b.lo2 = lo[2];
	b.off1 = cnt[0];
	b.off2 = cnt[0] + cnt[1];
	b.T = b.off2 + cnt[2];
	b.ownLo = cellStart[c];
	b.m = cellStart[c + 1] - b.ownLo;
	b.ownOff = b.off1 + (b.ownLo - lo[1]);
	return b;
}

template <class M, int PASS>
__global__ void __launch_bounds__(SPH_SWEEP_WARPS * 32) color_sweep_kernel(GridDesc g, PairParams k, const uint32_t *__restrict__ cellStart,
                                                                          const uint32_t *__restrict__ colorList, uint32_t listStride,
                                                                          const uint32_t *__restrict__ colorCount, float2 *pos, float2 *vel,
                                                                          const float2 *__restrict__ press, uint32_t cap, Counters *__restrict__ ctr) {
	extern __shared__ __align__(16) unsigned char sweepSmem[];
	const uint32_t lane = lane_id(), w = threadIdx.x >> 5, ltMask = (1u << lane) - 1u;
	unsigned char *mine = sweepSmem + (size_t)w * sweep_bytes_per_warp(cap, PASS);
	float2 *sPos = reinterpret_cast<float2 *>(mine);
	float2 *sVel = sPos + cap; // viscosity only
	uint16_t *queueStaged = reinterpret_cast<uint16_t *>(sPos + cap * (PASS == SWEEP_VISCOSITY ? 2 : 1));
	uint16_t *queueWide = reinterpret_cast<uint16_t *>(mine);
	const uint32_t wideCap = sweep_queue_capacity(cap, PASS);
	const uint32_t nLight = colorCount[0], nList = nLight + colorCount[9]; // this colour's light and heavy cells alike
	const int nRows = g.rowHi - g.rowLo;
	for (uint32_t idx = blockIdx.x * SPH_SWEEP_WARPS + w; idx < nList; idx += gridDim.x * SPH_SWEEP_WARPS) {
		const uint32_t c = colour_cell(colorList, listStride, nLight, idx);
		const SweepBlock b = sweep_block_of(g, cellStart, c, nRows);
		if (((b.T + 31u) & ~31u) <= cap) {
			sweep_cell<M, PASS, true>(k, b, pos, vel, press, sPos, sVel, queueStaged, lane, ltMask, SweepNoHook(), sweep_paired(cap, PASS) ? queueStaged + cap : nullptr);
		} else if (b.T <= wideCap) {
			sweep_cell<M, PASS, false>(k, b, pos, vel, press, sPos, sVel, queueWide, lane, ltMask);
			__syncwarp();
		} else if (lane == 0) { // denser than anything the queue can hold: report, leave the cell alone
			atomicOr(&ctr->overflow, 4u);
		}
	}
}

// ---- the same sweep with a whole thread block per cell --------------------------------------------
// One warp per cell leaves most of the GPU idle when a colour has fewer cells than the device has warp
// slots (the reference's own scenes: 594 cells, 66 per colour), and the chain of a cell's particles is
// strictly serial.  Here SPH_TEAM_WARPS warps share one cell: the candidate tests and the pair terms of
// a particle are spread over the warps, the queue is assembled from per-trip ballots with a prefix sum,
// and warp 0 folds the stored pair terms in queue order - the additions, their order and the lane
// assignment are exactly those of color_sweep_kernel, so both kernels produce the same bits and the
// host may pick either by load.  Used by color_sweep_team_kernel (nine launches, every cell) and for the
// HEAVY cells of the one-launch sweep (color_sweep_flow_kernel).
#define SPH_TEAM_WARPS 8
#define SPH_TEAM_MAX_CAP 2048u // two trips per lane in the prefix over the trips' hit counts
__host__ __device__ inline uint32_t team_smem_bytes(uint32_t cap, int pass) {
	return cap * 8u * (pass == SWEEP_VISCOSITY ? 2u : 1u) /* sPos (+ sVel) */ + cap * 2u /* queue */ + cap * 8u /* pair terms */ + (cap / 32u + 2u) * 4u /* ballots */;
}
// the largest staging capacity (multiple of 32, <= SPH_TEAM_MAX_CAP) a team can run in `bytes` of shared memory
__host__ __device__ inline uint32_t team_capacity(uint32_t bytes, int pass) {
	const uint32_t per32 = 32u * (8u * (pass == SWEEP_VISCOSITY ? 2u : 1u) + 2u + 8u) + 4u;
	const uint32_t cap = bytes > 8u ? ((bytes - 8u) / per32) * 32u : 0u;
	return cap < SPH_TEAM_MAX_CAP ? cap : SPH_TEAM_MAX_CAP;
}

// One cell, one thread block of SPH_TEAM_WARPS warps (all threads call this; block barriers inside).
// COHERENT as in sweep_cell: the staging loads bypass L1.
template <class M, int PASS, bool COHERENT>
__device__ __forceinline__ void sweep_cell_team(const PairParams &k, const SweepBlock &b, float2 *pos, float2 *vel, const float2 *__restrict__ press,
                                                unsigned char *smem, uint32_t cap, Counters *__restrict__ ctr) {
	const uint32_t lane = lane_id(), w = threadIdx.x >> 5, ltMask = (1u << lane) - 1u;
	float2 *sPos = reinterpret_cast<float2 *>(smem);
	float2 *sVel = sPos + cap; // viscosity only
	float2 *sTerm = sPos + cap * (PASS == SWEEP_VISCOSITY ? 2 : 1);
	uint16_t *queue = reinterpret_cast<uint16_t *>(sTerm + cap);
	uint32_t *tripMask = reinterpret_cast<uint32_t *>(queue + cap);
	float2 *state = (PASS == SWEEP_DELTA) ? pos : vel;
	const uint32_t Tpad = (b.T + 31u) & ~31u, nTrips = Tpad >> 5;
	if (Tpad > cap) { // too large to stage: warp 0 takes the L2 path of the one-warp kernel (same arithmetic)
		if (w == 0) {
			const uint32_t wide = team_smem_bytes(cap, PASS) / 2u;
			if (b.T <= wide) sweep_cell<M, PASS, false, COHERENT>(k, b, pos, vel, press, sPos, sVel, reinterpret_cast<uint16_t *>(smem), lane, ltMask);
			else if (lane == 0) atomicOr(&ctr->overflow, 4u);
		}
		__syncthreads();
		return;
	}
	for (uint32_t t = threadIdx.x; t < Tpad; t += SPH_TEAM_WARPS * 32) {
		if (t < b.T) {
			const uint32_t j = b.gidx(t);
			sPos[t] = COHERENT ? __ldcg(&pos[j]) : pos[j];
			if (PASS == SWEEP_VISCOSITY) sVel[t] = COHERENT ? __ldcg(&vel[j]) : vel[j];
		} else {
			sPos[t] = make_float2(3.0e18f, 3.0e18f);
		}
	}
	__syncthreads();
	for (uint32_t kBase = 0; kBase < b.m; kBase += 32) {
		float2 myPress = make_float2(0.0f, 0.0f);
		if (PASS == SWEEP_DELTA && kBase + lane < b.m) myPress = press[b.ownLo + kBase + lane];
		const uint32_t kEnd = min(b.m - kBase, 32u);
		for (uint32_t kk = 0; kk < kEnd; ++kk) {
			const uint32_t si = b.ownOff + kBase + kk;
			const float2 xi = sPos[si];
			float2 vi = make_float2(0.0f, 0.0f), ppi = make_float2(0.0f, 0.0f);
			if (PASS == SWEEP_VISCOSITY) vi = sVel[si];
			if (PASS == SWEEP_DELTA) {
				ppi.x = __shfl_sync(0xffffffffu, myPress.x, (int)kk);
				ppi.y = __shfl_sync(0xffffffffu, myPress.y, (int)kk);
			}
			// stage 1a: every warp tests its share of the 32-candidate trips
			for (uint32_t tr = w; tr < nTrips; tr += SPH_TEAM_WARPS) {
				const float2 xj = sPos[tr * 32u + lane];
				const uint32_t mask = __ballot_sync(0xffffffffu, M::norm2(M::sub2(xj, xi)) < k.h2);
				if (lane == 0) tripMask[tr] = mask;
			}
			__syncthreads();
			// stage 1b: exclusive prefix of the trips' hit counts (every warp redundantly; a lane looks after trips
			// `lane` and `lane + 32`: cap <= 2048), then the queue in candidate order
			const uint32_t myMask0 = lane < nTrips ? tripMask[lane] : 0u, myMask1 = lane + 32u < nTrips ? tripMask[lane + 32u] : 0u;
			uint32_t incl0 = (uint32_t)__popc(myMask0), incl1 = (uint32_t)__popc(myMask1);
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t up0 = __shfl_up_sync(0xffffffffu, incl0, o), up1 = __shfl_up_sync(0xffffffffu, incl1, o);
				if ((int)lane >= o) {
					incl0 += up0;
					incl1 += up1;
				}
			}
			const uint32_t nHit0 = __shfl_sync(0xffffffffu, incl0, 31);
			const uint32_t nHit = nHit0 + __shfl_sync(0xffffffffu, incl1, 31);
			const uint32_t excl0 = incl0 - (uint32_t)__popc(myMask0), excl1 = nHit0 + incl1 - (uint32_t)__popc(myMask1);
			for (uint32_t tr = w; tr < nTrips; tr += SPH_TEAM_WARPS) {
				const uint32_t mask = tr < 32u ? __shfl_sync(0xffffffffu, myMask0, (int)tr) : __shfl_sync(0xffffffffu, myMask1, (int)(tr - 32u));
				const uint32_t base = tr < 32u ? __shfl_sync(0xffffffffu, excl0, (int)tr) : __shfl_sync(0xffffffffu, excl1, (int)(tr - 32u));
				if (mask & (1u << lane)) queue[base + (uint32_t)__popc(mask & ltMask)] = (uint16_t)(tr * 32u + lane);
			}
			__syncthreads();
			// stage 2: pair terms, partner updated at once; the terms are kept for warp 0
			for (uint32_t q = w * 32u + lane; q < nHit; q += SPH_TEAM_WARPS * 32) {
				const uint32_t t = queue[q];
				bool hit;
				float2 hlf;
				if (PASS == SWEEP_DELTA) {
					const float2 xj = sPos[t];
					hlf = sweep_delta_term<M, false>(k, xi, ppi, xj, hit);
					sPos[t] = Exact::add2(hlf, xj);
				} else {
					const float2 vj = sVel[t];
					hlf = sweep_viscosity_term<M, false>(k, xi, vi, sPos[t], vj, hit);
					if (hit) sVel[t] = Exact::add2(hlf, vj);
					else hlf = make_float2(0.0f, 0.0f); // x - (+0) == x bit for bit
				}
				sTerm[q] = hlf;
			}
			__syncthreads();
			// the particle's own change: warp 0 repeats the one-warp kernel's additions in its order
			if (w == 0) {
				float ax = 0.0f, ay = 0.0f;
				for (uint32_t q = lane; q < nHit; q += 32) {
					const float2 hlf = sTerm[q];
					ax = __fsub_rn(ax, hlf.x);
					ay = __fsub_rn(ay, hlf.y);
				}
				const float2 own = butterfly_sum2_lane0(ax, ay, lane);
				if (lane == 0) {
					float2 *slot = (PASS == SWEEP_DELTA) ? &sPos[si] : &sVel[si];
					*slot = make_float2(__fadd_rn(own.x, slot->x), __fadd_rn(own.y, slot->y));
				}
			}
			__syncthreads();
		}
	}
	for (uint32_t t = threadIdx.x; t < b.T; t += SPH_TEAM_WARPS * 32) state[b.gidx(t)] = (PASS == SWEEP_DELTA) ? sPos[t] : sVel[t];
	__syncthreads();
}

template <class M, int PASS>
__global__ void __launch_bounds__(SPH_TEAM_WARPS * 32) color_sweep_team_kernel(GridDesc g, PairParams k, const uint32_t *__restrict__ cellStart,
                                                                               const uint32_t *__restrict__ colorList, uint32_t listStride,
                                                                               const uint32_t *__restrict__ colorCount, float2 *pos, float2 *vel,
                                                                               const float2 *__restrict__ press, uint32_t cap, Counters *__restrict__ ctr) {
	extern __shared__ __align__(16) unsigned char teamSmem[];
	const uint32_t nLight = colorCount[0], nList = nLight + colorCount[9];
	const int nRows = g.rowHi - g.rowLo;
	for (uint32_t idx = blockIdx.x; idx < nList; idx += gridDim.x) {
		const uint32_t c = colour_cell(colorList, listStride, nLight, idx);
		const SweepBlock b = sweep_block_of(g, cellStart, c, nRows);
		sweep_cell_team<M, PASS, false>(k, b, pos, vel, press, teamSmem, cap, ctr);
	}
}

// ---- the nine colours in ONE launch: dependency-driven sweep --------------------------------------
// Nine launches per pass leave eighteen tails per step in which most of the GPU waits for the last
// few cells of a colour.  Here the occupied cells of all colours form one queue (colour 0's list,
// then colour 1's, ...) that persistent warps drain through an atomic ticket.  A cell of colour c may
// start once every occupied cell of a LOWER colour whose 3x3 footprint overlaps its own - the cells
// within two rows/columns - has finished; cells of a higher colour inside that range wait for it by
// the same rule.  Cells that are further apart never touch the same particles.  Every cell therefore
// reads exactly the state it reads in the nine-launch version and the results are bit-identical.
// `flow[0..3]` are ticket counters and `flow[4 + cell]` the done flags, (re)initialised with every grid build
// (color_rows_count_kernel).  Loads of particle state bypass L1 (another SM may have just rewritten it).
//
// Two queues: the LIGHT cells (one warp per cell) and the HEAVY cells (one block per cell, sweep_cell_team), both in
// colour-major, row-major order.  The first `teams` blocks of the grid (as many as the fullest colour has heavy cells,
// at most half the grid) drain the heavy queue as teams and then join the others on the light queue.
// Deadlock-free: tickets of either queue are handed out in queue order and a worker takes its tickets in order, so the
// unfinished cell that comes first in (colour, row, column) order only waits for finished cells, all cells ahead of it
// in its own queue are finished, i.e. their workers have moved on and one of them holds its ticket or draws it next;
// there is always at least one worker per non-empty queue (the host launches at least two blocks, all resident).
#define SPH_FLOW_WARPS 8
#define SPH_FLOW_MIN_BLOCKS 5 // 48 registers: at 40 ptxas rematerialises addresses inside the pair loops, +26 % instructions (ncu, r1b)

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// The lower-colour cells of the 5x5 neighbourhood of local cell c must be done (empty ones always are): lane l looks
// after cell (l%5-2, l/5-2).  Returns the flag this lane still has to wait for (nullptr: none).
__device__ __forceinline__ const uint32_t *flow_pending_flag(const GridDesc &g, const uint32_t *flags, uint32_t c, int nRows, uint32_t epoch, uint32_t lane) {
	const int yl = (int)(c / (uint32_t)g.gx), cx = (int)(c - (uint32_t)yl * (uint32_t)g.gx);
	const int color = (int)(((uint32_t)(yl + g.rowLo) % 3u) * 3u + (uint32_t)cx % 3u);
	const int nx = cx + (int)(lane % 5u) - 2, ny = yl + (int)(lane / 5u) - 2;
	if (lane < 25u && lane != 12u && nx >= 0 && nx < g.gx && ny >= 0 && ny < nRows) {
		const int ncolor = (int)(((uint32_t)(ny + g.rowLo) % 3u) * 3u + (uint32_t)nx % 3u);
		const uint32_t *f = flags + ((uint32_t)ny * (uint32_t)g.gx + (uint32_t)nx);
		if (ncolor < color && ld_acquire_gpu(f) < epoch) return f;
	}
	return nullptr;
}

// STRIP: the launch runs on a y-strip (ghost rows exist); a single-GPU launch does not pay for the test below.
template <class M, int PASS, bool STRIP>
__global__ void __launch_bounds__(SPH_FLOW_WARPS * 32, SPH_FLOW_MIN_BLOCKS)
    color_sweep_flow_kernel(GridDesc g, PairParams k, const uint32_t *__restrict__ cellStart, const uint32_t *__restrict__ colorList, uint32_t listStride,
                            const uint32_t *__restrict__ colorCount, float2 *pos, float2 *vel, const float2 *__restrict__ press, uint32_t cap,
                            Counters *__restrict__ ctr, uint32_t *flow, uint32_t epoch, uint32_t maxTeams) {
	extern __shared__ __align__(16) unsigned char sweepSmem[];
	const uint32_t lane = lane_id(), w = threadIdx.x >> 5, ltMask = (1u << lane) - 1u;
	const int nRows = g.rowHi - g.rowLo;
	__shared__ uint32_t counts[18]; // cells per class and colour: shared memory, registers would cost occupancy
	__shared__ uint32_t teamTicket;
	if (threadIdx.x < 18) counts[threadIdx.x] = colorCount[threadIdx.x];
	__syncthreads();
	// `epoch` counts the sweeps over this grid (1 = the displacement pass right after the grid build, 2 = the next
	// step's viscosity pass, more only through sph_run_pass): a cell is done once its flag >= epoch, so the flags
	// need no reset between passes; passes alternate between two pairs of ticket counters and zero the other pair.
	uint32_t *tickets = flow + (epoch & 1u);
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		flow[(epoch + 1u) & 1u] = 0u;
		flow[2u + ((epoch + 1u) & 1u)] = 0u;
	}
	uint32_t *flags = flow + SPH_FLOW_FLAGS;
	// queue position -> cell (SPH_KEY_NONE past the end); cls 0 = light (front of a colour's region), 1 = heavy (back)
	auto cell_of_ticket = [&](uint32_t t, uint32_t cls) -> uint32_t {
		uint32_t seen = 0, at = 0xffffffffu, base = 0;
#pragma unroll
		for (int cc = 0; cc < 9; ++cc) {
			const uint32_t n = counts[cls * 9u + (uint32_t)cc];
			if (t >= seen && t < seen + n) {
				at = t - seen;
				base = (uint32_t)cc * listStride;
			}
			seen += n;
		}
		if (at == 0xffffffffu) return SPH_KEY_NONE;
		return __ldg(&colorList[(size_t)base + (cls ? listStride - 1u - at : at)]);
	};

	// Strips: the viscosity sweep is the LAST sweep before the next exchange drops every ghost, so only the owned rows
	// have to come out right, and an error travels at most three rows inward per sweep (the halo bound of DESIGN.md
	// section 7): ghost cells more than three rows away from the owned rows are marked done without being swept.
	const uint32_t sweptLo = (STRIP && PASS == SWEEP_VISCOSITY) ? (uint32_t)(max(g.ownLo - 3, g.rowLo) - g.rowLo) * (uint32_t)g.gx : 0u;
	const uint32_t sweptHi = (STRIP && PASS == SWEEP_VISCOSITY) ? (uint32_t)(min(g.ownHi + 3, g.rowHi) - g.rowLo) * (uint32_t)g.gx : 0xffffffffu;

	// ---- heavy cells: the first `teams` blocks, one block per cell --------------------------------------------
	{
		uint32_t mostHeavy = 0;
#pragma unroll
		for (int cc = 0; cc < 9; ++cc) mostHeavy = max(mostHeavy, counts[9 + cc]);
		const uint32_t teams = mostHeavy ? max(1u, min(mostHeavy, min(maxTeams, gridDim.x - 1u))) : 0u; // (the host keeps a share of the blocks for the light queue)
		if (blockIdx.x < teams) {
			const uint32_t capTeam = team_capacity(SPH_FLOW_WARPS * sweep_bytes_per_warp(cap, PASS), PASS);
			uint32_t *teamTickets = flow + 2u + (epoch & 1u);
			for (;;) {
				if (threadIdx.x == 0) teamTicket = atomicAdd(teamTickets, 1u);
				__syncthreads();
				const uint32_t c = cell_of_ticket(teamTicket, 1u);
				if (c == SPH_KEY_NONE) break; // (uniform: every thread read the same ticket)
				if (STRIP && PASS == SWEEP_VISCOSITY && (c < sweptLo || c >= sweptHi)) { // a far ghost cell: nothing to sweep, nobody needs to wait for it
					__syncthreads(); // (everybody has read the ticket)
					if (threadIdx.x == 0) st_release_gpu(flags + c, epoch);
					continue;
				}
				const SweepBlock b = sweep_block_of(g, cellStart, c, nRows);
				if (w == 0) {
					const uint32_t *flag = flow_pending_flag(g, flags, c, nRows, epoch, lane);
					while (!__all_sync(0xffffffffu, flag == nullptr)) {
						__nanosleep(100);
						if (flag && ld_acquire_gpu(flag) >= epoch) flag = nullptr;
					}
				}
				__syncthreads(); // also: nobody draws the next ticket before everybody has read this one
				sweep_cell_team<M, PASS, true>(k, b, pos, vel, press, sweepSmem, capTeam, ctr);
				// done: every thread's write-back is ordered before the barrier inside sweep_cell_team, one thread releases
				__threadfence();
				__syncthreads();
				if (threadIdx.x == 0) st_release_gpu(flags + c, epoch);
			}
			__syncthreads();
		}
	}

	// ---- light cells: one warp per cell -------------------------------------------------------------------------
	unsigned char *mine = sweepSmem + (size_t)w * sweep_bytes_per_warp(cap, PASS);
	float2 *sPos = reinterpret_cast<float2 *>(mine);
	float2 *sVel = sPos + cap; // viscosity only
	uint16_t *queueStaged = reinterpret_cast<uint16_t *>(sPos + cap * (PASS == SWEEP_VISCOSITY ? 2 : 1));
	uint16_t *queueWide = reinterpret_cast<uint16_t *>(mine);
	const uint32_t wideCap = sweep_queue_capacity(cap, PASS);
	// The next ticket is drawn when a cell's arithmetic is over, just before its write-back, and read after the
	// done flag is up: the atomic's round trip overlaps the stores'.  Drawing it any earlier would hide it as well
	// but park cells: every ticket a warp holds without working on it widens the window of cells in flight, and
	// once that window exceeds a colour's list, cells start before the lower-colour neighbours they wait for
	// (measured: two tickets ahead = 8 % of the cells wait, profiles/).  Lane 1 increments a per-warp decoy word
	// alongside on purpose: on an address it can prove uniform, ptxas warp-aggregates the atomic (atom.inc and
	// run-time increments too) and broadcasts, i.e. waits for, its result on the spot.
	const uint32_t decoy = SPH_FLOW_FLAGS + g.nCells + (blockIdx.x * SPH_FLOW_WARPS + w); // one word per warp behind the flags
	auto draw_ticket = [&]() -> uint32_t {
		uint32_t t = 0;
		if (lane < 2u) asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(tickets + (lane == 0u ? 0u : decoy)) : "memory");
		return t; // valid in lane 0
	};
	uint32_t c = cell_of_ticket(__shfl_sync(0xffffffffu, draw_ticket(), 0), 0u);
	while (c != SPH_KEY_NONE) {
		if (STRIP && PASS == SWEEP_VISCOSITY && (c < sweptLo || c >= sweptHi)) { // a far ghost cell (see above)
			if (lane == 0) st_release_gpu(flags + c, epoch);
			c = cell_of_ticket(__shfl_sync(0xffffffffu, draw_ticket(), 0), 0u);
			continue;
		}
		const SweepBlock b = sweep_block_of(g, cellStart, c, nRows);
		// (the flags are fetched together with the loads above: one round trip to L2 for all of them)
		const uint32_t *flag = flow_pending_flag(g, flags, c, nRows, epoch, lane);
		while (!__all_sync(0xffffffffu, flag == nullptr)) {
			__nanosleep(100);
			if (flag && ld_acquire_gpu(flag) >= epoch) flag = nullptr;
		}
		__syncwarp();
		uint32_t nextTicket = 0;
		auto draw_next = [&]() { nextTicket = draw_ticket(); };
		if (((b.T + 31u) & ~31u) <= cap) {
			sweep_cell<M, PASS, true, true>(k, b, pos, vel, press, sPos, sVel, queueStaged, lane, ltMask, draw_next, sweep_paired(cap, PASS) ? queueStaged + cap : nullptr);
		} else if (b.T <= wideCap) { // (only when the lists were classified for a larger staging capacity than this launch has)
			sweep_cell<M, PASS, false, true>(k, b, pos, vel, press, sPos, sVel, queueWide, lane, ltMask, draw_next);
		} else { // denser than anything the queue can hold: report, leave the cell alone
			if (lane == 0) atomicOr(&ctr->overflow, 4u);
			draw_next();
		}
		// done: one lane releases for the warp (the __syncwarp orders every lane's write-back before it)
		__syncwarp();
		if (lane == 0) st_release_gpu(flags + c, epoch);
		c = cell_of_ticket(__shfl_sync(0xffffffffu, nextTicket, 0), 0u);
	}
}

// ---- phases 8+9: body collisions and velocity (demo4.cpp:412-450) ---------------------------------
// Also closes the step: publishes n = nSorted = nOut for the next one.
__global__ void __launch_bounds__(SPH_THREADS) collide_velocity_kernel(Counters *__restrict__ ctr, float2 *__restrict__ pos, const float2 *__restrict__ prev,
                                                                      float2 *__restrict__ vel, const DevBody *__restrict__ bodies, int nbodies,
                                                                      float invDt, int doCollide, int doVelocity, int commit) {
	const uint32_t n = commit ? ctr->nOut : ctr->n;
	// particle pairs, two pairs in flight per thread (see integrate_kernel)
	const uint32_t nPairs = (n + 1u) >> 1, T = gridDim.x * blockDim.x;
	for (uint32_t q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < nPairs; q0 += 2u * T) {
		const uint32_t q1 = q0 + T;
		const bool has1 = q1 < nPairs;
		const bool both0 = 2u * q0 + 1u < n, both1 = has1 && 2u * q1 + 1u < n;
		Pair2 p0 = load_pair(pos, q0, both0), p1 = {}, r0 = {}, r1 = {};
		if (has1) p1 = load_pair(pos, q1, both1);
		if (doVelocity) {
			r0 = load_pair(prev, q0, both0);
			if (has1) r1 = load_pair(prev, q1, both1);
		}
		if (doCollide) {
			p0.a = col::all(p0.a, bodies, nbodies);
			if (both0) p0.b = col::all(p0.b, bodies, nbodies);
			store_pair(pos, q0, both0, p0);
			if (has1) {
				p1.a = col::all(p1.a, bodies, nbodies);
				if (both1) p1.b = col::all(p1.b, bodies, nbodies);
				store_pair(pos, q1, both1, p1);
			}
		}
		if (doVelocity) {
			Pair2 v;
			v.a = make_float2(__fmul_rn(__fsub_rn(p0.a.x, r0.a.x), invDt), __fmul_rn(__fsub_rn(p0.a.y, r0.a.y), invDt));
			v.b = make_float2(__fmul_rn(__fsub_rn(p0.b.x, r0.b.x), invDt), __fmul_rn(__fsub_rn(p0.b.y, r0.b.y), invDt));
			store_pair(vel, q0, both0, v);
			if (has1) {
				v.a = make_float2(__fmul_rn(__fsub_rn(p1.a.x, r1.a.x), invDt), __fmul_rn(__fsub_rn(p1.a.y, r1.a.y), invDt));
				v.b = make_float2(__fmul_rn(__fsub_rn(p1.b.x, r1.b.x), invDt), __fmul_rn(__fsub_rn(p1.b.y, r1.b.y), invDt));
				store_pair(vel, q1, both1, v);
			}
		}
	}
	// closes the step: every block read nOut above, nobody in this kernel reads n / nSorted
	if (commit && blockIdx.x == 0 && threadIdx.x == 0) {
		ctr->n = n;
		ctr->nSorted = n;
	}
}

__global__ void commit_kernel(Counters *ctr) {
	ctr->n = ctr->nOut;
	ctr->nSorted = ctr->nOut;
}

// per-step reset of the statistics that are per-step in the reference (demo4.cpp:369-370); the full
// step does this inside integrate_kernel, single passes (sph_run_pass) use this
__global__ void begin_step_kernel(Counters *ctr) {
	ctr->minNbr = 0xffffffffu;
	ctr->maxNbr = 0;
	ctr->pairCandidates = 0ull;
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

// multi-GPU readback: the particles this rank OWNS (row of their cell inside the strip), compacted
// in arbitrary order together with their ids; colours as in render_kernel
__global__ void __launch_bounds__(SPH_THREADS) gather_owned_kernel(GridDesc g, const Counters *__restrict__ ctr, const uint32_t *__restrict__ id,
                                                                  const uint32_t *__restrict__ cellOf, const float2 *__restrict__ pos,
                                                                  const float2 *__restrict__ prev, const float2 *__restrict__ vel,
                                                                  const float2 *__restrict__ acc, const float2 *__restrict__ dens,
                                                                  const float2 *__restrict__ press, float restDensity, uint32_t *__restrict__ outCount,
                                                                  uint32_t *__restrict__ outId, ParticleRecord *__restrict__ outRec,
                                                                  float2 *__restrict__ outPos, float4 *__restrict__ outColor) {
	const uint32_t n = ctr->n, nSorted = ctr->nSorted;
	SPH_WARP_LOOP(i, n) {
		bool own = i < n;
		if (own && i < nSorted) {
			const int row = (int)(cellOf[i] >> 16);
			own = row >= g.ownLo && row < g.ownHi;
		}
		const uint32_t mask = __ballot_sync(0xffffffffu, own);
		if (!mask) continue;
		uint32_t base = 0;
		const int leader = __ffs(mask) - 1;
		if ((int)lane_id() == leader) base = atomicAdd(outCount, (uint32_t)__popc(mask));
		base = __shfl_sync(0xffffffffu, base, leader);
		if (!own) continue;
		const uint32_t k = base + (uint32_t)__popc(mask & ((1u << lane_id()) - 1u));
		outId[k] = id[i];
		const float2 d = dens[i], pr = press[i], v = vel[i];
		if (outRec) {
			ParticleRecord r;
			r.cur = pos[i];
			r.prev = prev[i];
			r.acc = acc[i];
			r.vel = v;
			r.rho = d.x;
			r.rhoNear = d.y;
			r.P = pr.x;
			r.PNear = pr.y;
			outRec[k] = r;
		}
		if (outPos) {
			const float rr = __fdiv_rn(pr.x, -10.0f);
			const float gg = __fdiv_rn(d.x, restDensity);
			const float bb = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))), 10.0f);
			outPos[k] = pos[i];
			outColor[k] = make_float4(fmaxf(fminf(rr, 1.0f), 0.0f), fmaxf(fminf(gg, 1.0f), 0.0f), fmaxf(fminf(bb, 1.0f), 0.0f), 1.0f);
		}
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

// Strips: AddParticle in bulk (sph_add_particles, emitters) and state injection (sph_write_particles).  Every rank is
// given the WHOLE list and keeps the particles whose cell row it owns; the creation index is the position in the list,
// the same on every rank.  `rec` (optional) carries the full ParticleData rows of an injected state; an injected state is
// kept for the whole local window [rowBegin, rowEnd) = owned + ghost rows, because the ghosts' velocities cannot be
// rebuilt from what a neighbour exchange carries.
__global__ void __launch_bounds__(SPH_THREADS) append_owned_kernel(GridDesc g, int rowBegin, int rowEnd, Counters *__restrict__ ctr, uint32_t capacity, uint32_t count,
                                                                  const float2 *__restrict__ posIn, const float2 *__restrict__ accIn,
                                                                  const ParticleRecord *__restrict__ rec, uint32_t firstId, float2 *__restrict__ pos,
                                                                  float2 *__restrict__ prev, float2 *__restrict__ vel, float2 *__restrict__ acc,
                                                                  float2 *__restrict__ dens, float2 *__restrict__ press, uint32_t *__restrict__ id) {
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const float2 p = rec ? rec[k].cur : posIn[k];
		int cx, cy;
		cell_of(g, p, cx, cy);
		if (cy < rowBegin || cy >= rowEnd) continue;
		const uint32_t slot = atomicAdd(&ctr->n, 1u);
		if (slot >= capacity) {
			atomicOr(&ctr->overflow, 1u);
			continue;
		}
		pos[slot] = p;
		prev[slot] = rec ? rec[k].prev : p; // ParticleData(pos), demo4.h:101-106
		vel[slot] = rec ? rec[k].vel : make_float2(0.0f, 0.0f);
		acc[slot] = rec ? rec[k].acc : (accIn ? accIn[k] : make_float2(0.0f, 0.0f));
		dens[slot] = rec ? make_float2(rec[k].rho, rec[k].rhoNear) : make_float2(0.0f, 0.0f);
		press[slot] = rec ? make_float2(rec[k].P, rec[k].PNear) : make_float2(0.0f, 0.0f);
		id[slot] = firstId + k;
	}
}

__global__ void clamp_count_kernel(Counters *ctr, uint32_t capacity) {
	if (ctr->n > capacity) ctr->n = capacity;
}

} // namespace sphb200

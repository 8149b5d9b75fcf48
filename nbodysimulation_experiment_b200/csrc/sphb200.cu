// libsphb200.so — host side of the B200 SPH hot path and its C ABI (include/sphb200.h).
//
// The host object mirrors Demo4::ParticleSimulation (demo4.h:137-226): it owns particle, body,
// emitter and grid storage (all of it in HBM), takes the same setup calls, and turns Update(dt)
// into one stream of kernel launches (sph_kernels.cuh) with no host round trip.  There is no CPU
// implementation of any phase in this library.
#include "../../include/sphb200.h"
#include "sph_kernels.cuh"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

using namespace sphb200;

namespace {

thread_local std::string g_createError;

template <class T>
struct DoubleBuf {
	T *buf[2] = { nullptr, nullptr };
	int cur = 0;
	T *in() const { return buf[cur]; }
	T *out() const { return buf[cur ^ 1]; }
	void flip() { cur ^= 1; }
};

struct HostEmitter { // Demo4::ParticleEmitter, demo4.h:123-133
	float px, py, dx, dy, radius, speed, rate, duration, elapsed, totalElapsed;
	int active;
};

// ---- NCCL, bound at run time (no link-time dependency: single-GPU users never load it) -------
// Minimal declarations of the stable NCCL 2.x C API (nccl.h:37-38,146,160,181,215,442,461,493,503).
typedef struct ncclComm *NcclComm;
typedef struct { char internal[128]; } NcclUniqueId;
struct NcclApi {
	void *lib = nullptr;
	int (*GetUniqueId)(NcclUniqueId *) = nullptr;
	int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
	int (*CommDestroy)(NcclComm) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
	int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
	std::string error;
	bool load() {
		if (lib) return true;
		for (const char *name : { "libnccl.so.2", "libnccl.so" }) {
			lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (lib) break;
		}
		if (!lib) {
			error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
			return false;
		}
#define BIND(field, sym)                                                    \
	field = reinterpret_cast<decltype(field)>(dlsym(lib, sym));             \
	if (!field) {                                                           \
		error = std::string("NCCL symbol missing: ") + sym;                 \
		return false;                                                       \
	}
		BIND(GetUniqueId, "ncclGetUniqueId")
		BIND(CommInitRank, "ncclCommInitRank")
		BIND(CommDestroy, "ncclCommDestroy")
		BIND(GetErrorString, "ncclGetErrorString")
		BIND(Send, "ncclSend")
		BIND(Recv, "ncclRecv")
		BIND(GroupStart, "ncclGroupStart")
		BIND(GroupEnd, "ncclGroupEnd")
		BIND(AllReduce, "ncclAllReduce")
#undef BIND
		return true;
	}
};
NcclApi g_nccl;
constexpr int kNcclUint8 = 1;  // ncclUint8, nccl.h:279
constexpr int kNcclUint32 = 3; // ncclUint32, nccl.h:281
constexpr int kNcclMax = 2;    // ncclMax, nccl.h:262
constexpr int kNcclSum = 0;    // ncclSum, nccl.h:260
constexpr int kHaloResizeEvery = 16; // steps between re-sizings of the exchange messages

enum Phase { PH_INTEGRATE, PH_VISCOSITY, PH_PREDICT, PH_SCAN, PH_REORDER, PH_DENSITY, PH_DELTA, PH_COLLIDE, PH_EXCHANGE, PH_COUNT };
static_assert(PH_COUNT == SPH_NUM_PHASES, "phase list");

struct StepGraphKey {
	uint32_t parity, sweepCap, nb, nbodies, haloMsgRecords, parts, flowEpoch;
	float2 force;
	PairParams k;
};
struct StepGraph {
	StepGraphKey key;
	uint32_t parityAfter = 0, flowEpochAfter = 0;
	cudaGraphExec_t exec = nullptr;
};

} // namespace

struct SphSim {
	SphConfig cfg;
	SphParams params;
	float2 gravity = { 0, 0 }, extForce = { 0, 0 };
	float omega = 1.0f;
	GridDesc grid;
	uint32_t capacity = 0;

	cudaStream_t stream = nullptr;
	cudaStream_t copyStream = nullptr;          // Render readback overlaps the next Update
	cudaEvent_t renderReady = nullptr, copyDone = nullptr;
	bool copyPending = false;
	Counters *dCtr = nullptr;
	Counters *hCtr = nullptr; // pinned mirror

	DoubleBuf<float2> pos, prev, vel, acc, dens, press;
	DoubleBuf<uint32_t> id, cellOf;
	uint32_t *cellNew = nullptr, *rank = nullptr, *slotId = nullptr;
	uint32_t *cellCount = nullptr, *cellStart = nullptr, *tileSums = nullptr;
	uint32_t nTiles = 0;
	// occupied cells per colour for the coloured Gauss-Seidel sweeps
	uint32_t *colorCount = nullptr, *colorList = nullptr;
	uint32_t *rowColor = nullptr;  // occupied cells per (local row, cx mod 3), then their offsets in the colour lists
	uint32_t *sweepFlow = nullptr; // [0..1] ticket counters, [2 + cell] done flags of the one-launch sweep (color_sweep_flow_kernel), then decoy words
	uint32_t flowEpoch = 0;        // sweeps launched over the current grid (the flags count passes, see color_sweep_flow_kernel)
	uint32_t listStride = 0, sweepCap = 512;
	bool sweepAdaptive = true;       // pick the staging capacity from the candidate-list maximum of recent steps
	Counters *hCtrLag = nullptr;     // pinned, refreshed asynchronously after every step
	cudaEvent_t lagEvent = nullptr;
	bool lagPending = false;

	std::vector<DevBody> bodies;
	DevBody *dBodies = nullptr;
	bool bodiesDirty = true;
	std::vector<HostEmitter> emitters;

	// lazily allocated readback staging
	ParticleRecord *dRecords = nullptr;
	float2 *dRenderPos = nullptr;
	float4 *dRenderCol = nullptr;
	int2 *dCellXY = nullptr;

	// y-strip decomposition
	StripDesc strip = {};
	HaloBuffer *sendBuf[2] = { nullptr, nullptr }, *recvBuf[2] = { nullptr, nullptr }; // [0] = lower neighbour, [1] = upper
	size_t haloBytes = 0;        // allocation per buffer
	uint32_t haloMsgRecords = 0; // records actually shipped per message (all ranks agree; re-sized every kHaloResizeEvery steps)
	uint32_t *dPeak = nullptr;   // [0] peak records packed since the last re-size, [1] all-reduced maximum
	uint64_t exchanges = 0;
	NcclComm comm = nullptr;
	uint32_t *dOwnedCount = nullptr, *dOwnedIds = nullptr;
	// periodic re-balancing of the strips (sph_set_rebalance)
	int rebalanceEvery = 0, rebalanceMaxShift = 2;
	bool allocFullGrid = false;         // cell arrays sized for any window of the grid, so a strip can move without reallocating
	uint32_t *dRowCounts = nullptr;     // gy row counts + world first-rows, all-reduced
	std::vector<uint32_t> hRowCounts;
	bool pendingRetarget = false;       // apply pendLo/pendHi between the viscosity pass and the grid build of this step
	int pendLo = 0, pendHi = 0;
	uint64_t rebalances = 0;
	// overlapped strip readback (sph_render_owned / sph_wait_render_owned)
	uint32_t *hOwnedCount = nullptr; // pinned
	uint64_t ownedShipped = 0, ownedLast = 0;
	uint32_t *ownedIdsDst = nullptr;
	void *ownedPosDst = nullptr, *ownedColDst = nullptr;
	size_t ownedPosStride = 0, ownedColStride = 0;
	bool ownedPending = false;

	uint64_t hostN = 0;        // particles this rank holds (exact on one GPU; an upper bound on strips)
	uint64_t nextId = 0;       // creation counter
	uint32_t accFrom = 0xFFFFFFFFu; // first array slot whose acceleration is live

	// statistics
	uint64_t steps = 0;
	bool steppedOnce = false;
	cudaEvent_t phaseEv[PH_COUNT + 1] = {};
	double phaseMs[PH_COUNT] = {};
	uint64_t phaseSteps = 0;
	float hostEmitterMs = 0.0f;
	cudaEvent_t marks[8] = {};

	// step graphs
	bool useGraphs = true;
	std::vector<StepGraph> graphs;

	std::string err;
};

namespace {

int fail(SphSim *s, int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (s) s->err = buf;
	else g_createError = buf;
	return code;
}

#define CU(s, call)                                                                                          \
	do {                                                                                                     \
		cudaError_t e__ = (call);                                                                            \
		if (e__ != cudaSuccess) return fail((s), SPH_ERR_CUDA, "%s -> %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

#define CHECK_HANDLE(h)                                   \
	do {                                                  \
		if (!(h)) return fail(nullptr, SPH_ERR_INVALID, "null handle"); \
	} while (0)

// strided host<->device copy; the contiguous case must not go through cudaMemcpy2D (a million 8-byte
// rows copy an order of magnitude slower than one flat transfer)
inline cudaError_t copy_strided(void *dst, size_t dstPitch, const void *src, size_t srcPitch, size_t width, size_t rows, cudaMemcpyKind kind,
                                cudaStream_t stream) {
	if (dstPitch == width && srcPitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, stream);
	return cudaMemcpy2DAsync(dst, dstPitch, src, srcPitch, width, rows, kind, stream);
}

// (integrate / collide+velocity take two particle pairs per thread, predict+key one pair: they are launched with
// blocks_for(n / 4) and blocks_for(n / 2))
inline unsigned blocks_for(uint64_t n) {
	uint64_t b = (n + SPH_THREADS - 1) / SPH_THREADS;
	if (b < 1) b = 1;
	if (b > 148u * 512u) b = 148u * 512u;
	return (unsigned)b;
}

template <class T>
cudaError_t alloc2(DoubleBuf<T> &b, size_t n) {
	cudaError_t e = cudaMalloc(&b.buf[0], n * sizeof(T));
	if (e != cudaSuccess) return e;
	e = cudaMalloc(&b.buf[1], n * sizeof(T));
	if (e != cudaSuccess) return e;
	cudaMemset(b.buf[0], 0, n * sizeof(T));
	cudaMemset(b.buf[1], 0, n * sizeof(T));
	return cudaSuccess;
}
template <class T>
void free2(DoubleBuf<T> &b) {
	cudaFree(b.buf[0]);
	cudaFree(b.buf[1]);
	b.buf[0] = b.buf[1] = nullptr;
}

PairParams pair_params(const SphSim *s, float dt) {
	PairParams k;
	k.h2 = s->params.kernel_height * s->params.kernel_height; // sph.h:470
	k.invH = s->params.inv_kernel_height;
	k.restDensity = s->params.rest_density;
	k.stiffness = s->params.stiffness;
	k.nearStiffness = s->params.near_stiffness;
	k.sigma = s->params.linear_viscosity;
	k.beta = s->params.quadratic_viscosity;
	k.dt = dt;
	k.dt2 = dt * dt;
	k.halfDt2 = (dt * dt) * 0.5f;
	k.omega = s->omega;
	return k;
}

void default_params(SphParams *p) { // SPHParameters(), sph.h:88-98
	const float radius = 0.05f;
	p->kernel_height = 6.0f * radius;
	p->cell_size = p->kernel_height;
	p->particle_spacing = p->kernel_height * 0.5f;
	p->inv_kernel_height = 1.0f / p->kernel_height;
	p->rest_density = 20.0f;
	p->stiffness = 0.6f;
	p->near_stiffness = p->stiffness * 10.0f;
	p->linear_viscosity = 0.5f;
	p->quadratic_viscosity = 0.3f;
}

int upload_bodies(SphSim *s) {
	if (!s->bodiesDirty) return SPH_OK;
	if (!s->bodies.empty())
		CU(s, cudaMemcpyAsync(s->dBodies, s->bodies.data(), s->bodies.size() * sizeof(DevBody), cudaMemcpyHostToDevice, s->stream));
	s->bodiesDirty = false;
	return SPH_OK;
}

constexpr size_t kMaxBodies = 100;  // kSPHMaxBodyCount, sph.h:71
constexpr size_t kMaxEmitters = 8;  // kSPHMaxEmitterCount, sph.h:72
constexpr size_t kMaxPolyVerts = 8; // kMaxScenarioPolygonCount, sph.h:161

int add_body(SphSim *s, const DevBody &b) {
	if (s->bodies.size() >= kMaxBodies) return fail(s, SPH_ERR_CAPACITY, "more than %zu bodies (demo4.cpp:87)", kMaxBodies);
	s->bodies.push_back(b);
	s->bodiesDirty = true;
	return SPH_OK;
}

__global__ void set_counts_kernel(Counters *ctr, uint32_t n, uint32_t nSorted) {
	ctr->n = n;
	ctr->nSorted = nSorted;
	ctr->nIn = n;
	ctr->nOut = nSorted;
}
// records packed this step vs the records a message ships: remember the peak, flag what does not fit
__global__ void note_peak_kernel(const HaloBuffer *down, const HaloBuffer *up, uint32_t *peak, uint32_t msgRecords, Counters *ctr) {
	const uint32_t most = max(down ? down->count : 0u, up ? up->count : 0u);
	if (most > peak[0]) peak[0] = most;
	if (most > msgRecords) atomicOr(&ctr->overflow, 2u);
}
__global__ void grow_count_kernel(Counters *ctr, uint32_t n) {
	ctr->n = n;
	ctr->nIn = n;
}
__global__ void fill_ids_kernel(uint32_t *id, uint32_t from, uint32_t count, uint32_t firstId) {
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) id[from + k] = firstId + k;
}
__global__ void reset_stats_kernel(Counters *ctr) {
	ctr->minNbr = 0xffffffffu;
	ctr->maxNbr = 0;
	ctr->minCell = 0xffffffffu;
	ctr->maxCell = 0;
	ctr->pairCandidates = 0;
	ctr->lost = 0;
	ctr->overflow = 0;
}
// predict alone (sph_run_pass(PREDICT)), demo4.cpp:330-339
__global__ void __launch_bounds__(SPH_THREADS) predict_only_kernel(const Counters *__restrict__ ctr, float2 *__restrict__ pos, float2 *__restrict__ prev,
                                                                  const float2 *__restrict__ vel, float dt) {
	const uint32_t n = ctr->n;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		float2 p = pos[i];
		const float2 v = vel[i];
		prev[i] = p;
		pos[i] = make_float2(__fadd_rn(__fmul_rn(v.x, dt), p.x), __fadd_rn(__fmul_rn(v.y, dt), p.y));
	}
}
// reorder payload that the full step does not need to move (velocity is recomputed, densities are
// rewritten) but a stand-alone grid pass must keep attached to its particle
__global__ void __launch_bounds__(SPH_THREADS) carry_kernel(const Counters *__restrict__ ctr, const uint32_t *__restrict__ idOld, const uint32_t *__restrict__ idNew,
                                                           const uint32_t *__restrict__ cellNew, const uint32_t *__restrict__ cellStart, GridDesc g,
                                                           const float2 *__restrict__ velIn, const float2 *__restrict__ accIn, const float2 *__restrict__ densIn,
                                                           const float2 *__restrict__ pressIn, float2 *__restrict__ velOut, float2 *__restrict__ accOut,
                                                           float2 *__restrict__ densOut, float2 *__restrict__ pressOut) {
	const uint32_t n = ctr->nIn;
	SPH_WARP_LOOP(i, n) {
		if (i >= n) continue;
		const uint32_t c = cellNew[i];
		if (c == SPH_KEY_NONE) continue;
		const uint32_t key = ((c >> 16) - (uint32_t)g.rowLo) * (uint32_t)g.gx + (c & 0xffffu);
		const uint32_t me = idOld[i];
		uint32_t dst = cellStart[key];
		while (idNew[dst] != me) ++dst; // ids are unique inside the cell's slab
		velOut[dst] = velIn[i];
		accOut[dst] = accIn[i];
		densOut[dst] = densIn[i];
		pressOut[dst] = pressIn[i];
	}
}

void record_phase(SphSim *s, int idx) {
	if (s->cfg.flags & SPH_FLAG_PHASE_TIMING) cudaEventRecord(s->phaseEv[idx], s->stream);
}

// ---- the grid build shared by sph_step and sph_run_pass(GRID) ------------------------------
// Message size: NCCL needs it on the host, the record counts only exist on the device.  All ranks
// therefore ship the same fixed number of records per message, re-agreed every kHaloResizeEvery steps
// as 1.5 x the largest count any rank packed since (one stream sync + a 4-byte all-reduce).  The
// first messages carry the whole buffer.  A count above the agreed size raises the overflow flag on
// the sender (note_peak_kernel) and on the receiver (header count > records shipped).
int maybe_resize_halo(SphSim *s) {
	const StripDesc &sd = s->strip;
	NcclApi &nc = g_nccl;
	if (s->exchanges > 0 && s->exchanges % kHaloResizeEvery == 0) {
		int rca = nc.AllReduce(s->dPeak, s->dPeak + 1, 1, kNcclUint32, kNcclMax, s->comm, s->stream);
		if (rca != 0) return fail(s, SPH_ERR_COMM, "NCCL all-reduce failed: %s", nc.GetErrorString(rca));
		uint32_t peak = 0;
		CU(s, cudaMemcpyAsync(&peak, s->dPeak + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
		CU(s, cudaMemsetAsync(s->dPeak, 0, sizeof(uint32_t), s->stream));
		CU(s, cudaStreamSynchronize(s->stream));
		const uint64_t want = (uint64_t)peak + peak / 2 + 4096;
		s->haloMsgRecords = (uint32_t)std::min<uint64_t>(sd.haloCap, want);
	}
	s->exchanges++;
	return SPH_OK;
}

// migration + halo in one neighbour exchange (fixed-size messages, count in the header), then file
// what arrived behind the local particles
int exchange_halos(SphSim *s) {
	const StripDesc &sd = s->strip;
	if (!s->comm) return fail(s, SPH_ERR_STATE, "sph_comm_init was not called on this multi-GPU handle");
	NcclApi &nc = g_nccl;
	const size_t msgBytes = sizeof(HaloBuffer) + (size_t)s->haloMsgRecords * sizeof(HaloRecord);
	note_peak_kernel<<<1, 1, 0, s->stream>>>(s->sendBuf[0], s->sendBuf[1], s->dPeak, s->haloMsgRecords, s->dCtr);
	int rc = nc.GroupStart();
	if (rc == 0 && sd.rank > 0) {
		rc = nc.Send(s->sendBuf[0], msgBytes, kNcclUint8, sd.rank - 1, s->comm, s->stream);
		if (rc == 0) rc = nc.Recv(s->recvBuf[0], msgBytes, kNcclUint8, sd.rank - 1, s->comm, s->stream);
	}
	if (rc == 0 && sd.rank + 1 < sd.world) {
		rc = nc.Send(s->sendBuf[1], msgBytes, kNcclUint8, sd.rank + 1, s->comm, s->stream);
		if (rc == 0) rc = nc.Recv(s->recvBuf[1], msgBytes, kNcclUint8, sd.rank + 1, s->comm, s->stream);
	}
	const int rcEnd = nc.GroupEnd();
	if (rc == 0) rc = rcEnd;
	if (rc != 0) return fail(s, SPH_ERR_COMM, "NCCL exchange failed: %s", nc.GetErrorString(rc));
	const unsigned nb = blocks_for(s->haloMsgRecords);
	const GridDesc &g = s->grid;
	if (sd.rank > 0)
		unpack_kernel<<<nb, SPH_THREADS, 0, s->stream>>>(g, s->dCtr, s->recvBuf[0], s->haloMsgRecords, s->capacity, 0u, nullptr, s->pos.in(), s->prev.in(), s->id.in(),
		                                                s->cellNew, s->rank, s->cellCount);
	if (sd.rank + 1 < sd.world)
		unpack_kernel<<<nb, SPH_THREADS, 0, s->stream>>>(g, s->dCtr, s->recvBuf[1], s->haloMsgRecords, s->capacity, 0u, sd.rank > 0 ? s->recvBuf[0] : nullptr, s->pos.in(),
		                                                s->prev.in(), s->id.in(), s->cellNew, s->rank, s->cellCount);
	CU(s, cudaGetLastError());
	return SPH_OK;
}

enum { GRID_FRONT = 1, GRID_EXCHANGE = 2, GRID_BACK = 4, GRID_ALL = 7 };
// parts: the strip exchange (NCCL) cannot be captured into a graph, so callers may enqueue the
// launches before it and after it separately
int launch_grid_build(SphSim *s, float dt, bool doPredict, bool carry, bool timed, int parts = GRID_ALL) {
	const GridDesc &g = s->grid;
	const unsigned nb = blocks_for(s->hostN);
	if (parts & GRID_FRONT) {
		CU(s, cudaMemsetAsync(s->cellCount, 0, (size_t)g.nCells * sizeof(uint32_t), s->stream));
		if (s->strip.world > 1) reset_halo_kernel<<<1, 1, 0, s->stream>>>(s->sendBuf[0], s->sendBuf[1]);
		predict_key_kernel<<<blocks_for((s->hostN + 1) / 2), SPH_THREADS, 0, s->stream>>>(g, s->strip, s->dCtr, s->pos.in(), s->prev.in(), s->vel.in(), s->id.in(), s->cellOf.in(), s->cellNew,
		                                                      s->rank, s->cellCount, dt, doPredict ? 1 : 0);
		// from the next grid on, authority follows the rows this grid is built for
		s->strip.authLo = g.ownLo;
		s->strip.authHi = g.ownHi;
		if (timed) record_phase(s, PH_PREDICT + 1);
	}
	if ((parts & GRID_EXCHANGE) && s->strip.world > 1) {
		int rc = exchange_halos(s);
		if (rc != SPH_OK) return rc;
	}
	if (!(parts & GRID_BACK)) return SPH_OK;
	if (timed) record_phase(s, PH_EXCHANGE + 1);
	scan_tiles_kernel<<<s->nTiles, SPH_THREADS, 0, s->stream>>>(s->cellCount, s->cellStart, s->tileSums, g.nCells);
	scan_sums_kernel<<<1, SPH_THREADS, 0, s->stream>>>(s->tileSums, s->nTiles, s->cellStart, g.nCells, s->dCtr);
	scan_add_kernel<<<s->nTiles, SPH_THREADS, 0, s->stream>>>(s->cellStart, s->tileSums, g.nCells);
	if (s->cfg.solver == SPH_SOLVER_COLORED_GS) { // occupied cells per colour, each list in row-major order
		const unsigned rowWarps = (unsigned)(g.rowHi - g.rowLo) * 3u, rowBlocks = (rowWarps + SPH_ROWLIST_WARPS - 1) / SPH_ROWLIST_WARPS;
		color_rows_count_kernel<<<rowBlocks, SPH_ROWLIST_WARPS * 32, 0, s->stream>>>(g, s->cellCount, s->rowColor, s->sweepFlow);
		color_rows_scan_kernel<<<1, 9 * 32, 0, s->stream>>>(g, s->rowColor, s->colorCount, s->sweepFlow);
		s->flowEpoch = 0; // the done flags are fresh: the next sweep over this grid is its first
		color_rows_fill_kernel<<<rowBlocks, SPH_ROWLIST_WARPS * 32, 0, s->stream>>>(g, s->cellCount, s->rowColor, s->colorList, s->listStride);
	}
	if (timed) record_phase(s, PH_SCAN + 1);
	scatter_ids_kernel<<<nb, SPH_THREADS, 0, s->stream>>>(g, s->dCtr, s->cellNew, s->rank, s->id.in(), s->cellStart, s->slotId);
	reorder_kernel<<<nb, SPH_THREADS, 0, s->stream>>>(g, s->dCtr, s->cellNew, s->id.in(), s->cellStart, s->slotId, s->pos.in(), s->prev.in(), s->pos.out(),
	                                                  s->prev.out(), s->id.out(), s->cellOf.out());
	if (carry) {
		carry_kernel<<<nb, SPH_THREADS, 0, s->stream>>>(s->dCtr, s->id.in(), s->id.out(), s->cellNew, s->cellStart, g, s->vel.in(), s->acc.in(), s->dens.in(),
		                                                s->press.in(), s->vel.out(), s->acc.out(), s->dens.out(), s->press.out());
		s->vel.flip();
		s->acc.flip();
		s->dens.flip();
		s->press.flip();
	}
	s->pos.flip();
	s->prev.flip();
	s->id.flip();
	s->cellOf.flip();
	if (timed) record_phase(s, PH_REORDER + 1);
	CU(s, cudaGetLastError());
	return SPH_OK;
}

template <class M>
void launch_viscosity(SphSim *s, const PairParams &k, unsigned nb) {
	viscosity_kernel<M><<<nb, SPH_THREADS, 0, s->stream>>>(s->grid, k, s->dCtr, s->pos.in(), s->vel.in(), s->cellOf.in(), s->cellStart, s->vel.out());
}
template <class M>
void launch_density(SphSim *s, const PairParams &k, unsigned nb) {
	density_kernel<M><<<nb, SPH_THREADS, 0, s->stream>>>(s->grid, k, s->dCtr, s->pos.in(), s->cellOf.in(), s->cellStart, s->dens.in(), s->press.in());
}
// In place on pos (delta) or vel (viscosity).  Three kernels with identical results:
//   * color_sweep_flow_kernel - one launch for all nine colours, persistent warps, per-cell dependency flags (large scenes);
//   * color_sweep_kernel      - nine launches, one warp per cell (SPH_FLAG_SWEEP_WARP);
//   * color_sweep_team_kernel - nine launches, one block per cell, for scenes whose colours have fewer cells than the
//                               GPU has warp slots (the reference's own scenes).
template <class M, int PASS>
void launch_sweeps(SphSim *s, const PairParams &k) {
	static bool attrSet = false;
	static int numSMs = 148;
	static int flowBlocksPerSM[97] = {};
	if (!attrSet) {
		cudaFuncSetAttribute(color_sweep_kernel<M, PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		cudaFuncSetAttribute(color_sweep_team_kernel<M, PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
		cudaFuncSetAttribute(color_sweep_flow_kernel<M, PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev);
		// resident blocks per SM of the persistent kernel for every staging capacity (32..3072 in steps of 32), asked
		// once here: the first steps of a simulation are never inside a graph capture
		for (uint32_t c32 = 1; c32 <= 96; ++c32) {
			int nbk = 0;
			const size_t bytes = (size_t)SPH_FLOW_WARPS * sweep_bytes_per_warp(c32 * 32u, PASS);
			if (bytes <= 200u * 1024u) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbk, color_sweep_flow_kernel<M, PASS>, SPH_FLOW_WARPS * 32, bytes);
			flowBlocksPerSM[c32] = nbk;
		}
		cudaGetLastError();
		attrSet = true;
	}
	// occupied cells of one colour <= min(cells of that colour, particles)
	const uint64_t cells = std::min<uint64_t>(s->listStride, std::max<uint64_t>(s->hostN, 1));
	bool team = s->cfg.world_size == 1 && s->hostN < 131072;
	if (s->cfg.flags & SPH_FLAG_SWEEP_TEAM) team = true;
	if (s->cfg.flags & (SPH_FLAG_SWEEP_WARP | SPH_FLAG_SWEEP_FLOW)) team = false;
	const bool flow = !team && !(s->cfg.flags & SPH_FLAG_SWEEP_WARP);
	if (flow) {
		// persistent: as many blocks as are resident at once (more would only draw a ticket and leave)
		const size_t smem = (size_t)SPH_FLOW_WARPS * sweep_bytes_per_warp(s->sweepCap, PASS);
		const int occBlocks = std::max(1, flowBlocksPerSM[std::min(96u, (s->sweepCap + 31u) / 32u)]);
		const uint64_t want = (9 * cells + SPH_FLOW_WARPS - 1) / SPH_FLOW_WARPS;
		const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)occBlocks * (uint64_t)numSMs));
		color_sweep_flow_kernel<M, PASS><<<blocks, SPH_FLOW_WARPS * 32, smem, s->stream>>>(s->grid, k, s->cellStart, s->colorList, s->listStride, s->colorCount, s->pos.in(),
		                                                                                 s->vel.in(), s->press.in(), s->sweepCap, s->dCtr, s->sweepFlow, ++s->flowEpoch);
		return;
	}
	for (int color = 0; color < 9; ++color) {
		const uint32_t *list = s->colorList + (size_t)color * s->listStride;
		if (team) {
			const uint32_t cap = std::min(s->sweepCap, 1024u);
			const unsigned blocks = (unsigned)std::min<uint64_t>(cells, 148u * 16u);
			color_sweep_team_kernel<M, PASS><<<blocks, SPH_TEAM_WARPS * 32, team_smem_bytes(cap, PASS), s->stream>>>(s->grid, k, s->cellStart, list, s->colorCount + color,
			                                                                                               s->pos.in(), s->vel.in(), s->press.in(), cap, s->dCtr);
		} else {
			const size_t smem = (size_t)SPH_SWEEP_WARPS * sweep_bytes_per_warp(s->sweepCap, PASS);
			const unsigned blocks = (unsigned)std::min<uint64_t>((cells + SPH_SWEEP_WARPS - 1) / SPH_SWEEP_WARPS, 148u * 64u);
			color_sweep_kernel<M, PASS><<<blocks, SPH_SWEEP_WARPS * 32, smem, s->stream>>>(s->grid, k, s->cellStart, list, s->colorCount + color, s->pos.in(), s->vel.in(),
			                                                                             s->press.in(), s->sweepCap, s->dCtr);
		}
	}
}

int run_viscosity(SphSim *s, const PairParams &k, unsigned nb) {
	const bool exact = s->cfg.fp_mode == SPH_FP_EXACT;
	if (s->cfg.solver == SPH_SOLVER_COLORED_GS) {
		if (exact) launch_sweeps<Exact, SWEEP_VISCOSITY>(s, k);
		else launch_sweeps<Fast, SWEEP_VISCOSITY>(s, k);
	} else {
		if (exact) launch_viscosity<Exact>(s, k, nb);
		else launch_viscosity<Fast>(s, k, nb);
		s->vel.flip();
	}
	return SPH_OK;
}

template <class M>
void launch_delta(SphSim *s, const PairParams &k, unsigned nb) {
	delta_kernel<M><<<nb, SPH_THREADS, 0, s->stream>>>(s->grid, k, s->dCtr, s->pos.in(), s->press.in(), s->cellOf.in(), s->cellStart, s->pos.out());
}

constexpr int kDefaultHaloRows = 7; // DESIGN.md "multi-GPU": an edge error moves <= 3 rows inward per sweep (density+displacement, viscosity) = 6, + 1

// (re)allocates everything sized by the local window of grid rows
int configure_strip(SphSim *s, int ownLo, int ownHi) {
	GridDesc &g = s->grid;
	if (ownLo < 0 || ownHi > g.gy || ownLo >= ownHi) return fail(s, SPH_ERR_INVALID, "strip rows [%d,%d) outside the grid (0..%d)", ownLo, ownHi, g.gy);
	const int world = s->strip.world, rank = s->strip.rank, halo = s->strip.halo;
	if (world > 1 && rank > 0 && rank + 1 < world && ownHi - ownLo < halo)
		return fail(s, SPH_ERR_INVALID, "interior strip of %d rows is thinner than the %d-row halo", ownHi - ownLo, halo);
	g.ownLo = ownLo;
	g.ownHi = ownHi;
	g.rowLo = world > 1 ? std::max(0, ownLo - halo) : 0;
	g.rowHi = world > 1 ? std::min(g.gy, ownHi + halo) : g.gy;
	if (world == 1 && (ownLo != 0 || ownHi != g.gy)) return fail(s, SPH_ERR_INVALID, "a single-GPU simulation owns the whole grid");
	g.nCells = (uint32_t)(g.rowHi - g.rowLo) * (uint32_t)g.gx;
	s->strip.authLo = ownLo;
	s->strip.authHi = ownHi;
	// everything sized by the window of rows: for the window itself, or (re-balancing on) for the whole grid
	const int allocRows = s->allocFullGrid ? g.gy : (g.rowHi - g.rowLo);
	const size_t allocCells = (size_t)allocRows * (size_t)g.gx;
	cudaFree(s->cellCount);
	cudaFree(s->cellStart);
	cudaFree(s->tileSums);
	cudaFree(s->colorList);
	cudaFree(s->sweepFlow);
	cudaFree(s->rowColor);
	s->cellCount = s->cellStart = s->tileSums = s->colorList = s->sweepFlow = s->rowColor = nullptr;
	s->nTiles = (g.nCells + SPH_SCAN_TILE - 1) / SPH_SCAN_TILE;
	CU(s, cudaMalloc(&s->cellCount, (allocCells + 1) * sizeof(uint32_t)));
	CU(s, cudaMalloc(&s->cellStart, (allocCells + 1) * sizeof(uint32_t)));
	CU(s, cudaMemset(s->cellStart, 0, (allocCells + 1) * sizeof(uint32_t)));
	CU(s, cudaMalloc(&s->tileSums, ((allocCells + SPH_SCAN_TILE - 1) / SPH_SCAN_TILE + 1) * sizeof(uint32_t)));
	s->listStride = (uint32_t)((g.gx + 2) / 3) * (uint32_t)((allocRows + 2) / 3 + 1);
	CU(s, cudaMalloc(&s->colorList, (size_t)9 * s->listStride * sizeof(uint32_t)));
	CU(s, cudaMalloc(&s->sweepFlow, (allocCells + 2 + 65536) * sizeof(uint32_t))); // + one decoy word per resident warp (color_sweep_flow_kernel)
	CU(s, cudaMemset(s->sweepFlow, 0xFF, (allocCells + 2 + 65536) * sizeof(uint32_t))); // no grid yet: every cell empty
	CU(s, cudaMemset(s->sweepFlow, 0, 2 * sizeof(uint32_t)));
	if (s->colorCount) CU(s, cudaMemset(s->colorCount, 0, 16 * sizeof(uint32_t)));
	s->flowEpoch = 0;
	CU(s, cudaMalloc(&s->rowColor, (size_t)allocRows * 3 * sizeof(uint32_t)));
	return SPH_OK;
}

// ---- periodic re-balancing of the strips ----------------------------------------------------------
// New boundaries from the global row histogram: boundary r goes where the prefix of the counts reaches r/world of the
// total, but (a) moves at most maxShift rows per re-balance, (b) stays `halo` rows inside the OLD ranges of the two
// ranks it separates - every row of a rank's new window then belongs (old ownership) to the rank itself or to a direct
// neighbour, so one neighbour exchange moves everything - and (c) leaves every strip at least minRows tall.  Pure
// integer arithmetic on data every rank holds identically: all ranks plan the same split.  strips.plan_bounds (Python)
// is the same function, tested on the CPU.
std::vector<int> plan_strip_bounds(const uint32_t *rowCounts, int gy, const std::vector<int> &oldB, int halo, int maxShift) {
	const int world = (int)oldB.size() - 1;
	const int minRows = 2 * halo + 4;
	uint64_t total = 0;
	for (int r = 0; r < gy; ++r) total += rowCounts[r];
	std::vector<int> nb = oldB;
	if (total == 0) return nb;
	uint64_t prefix = 0;
	int row = 0;
	for (int b = 1; b < world; ++b) {
		const uint64_t target = total * (uint64_t)b / (uint64_t)world;
		while (row < gy && prefix < target) prefix += rowCounts[row++]; // smallest `row` with sum(rows < row) >= target
		int want = row;
		want = std::max(want, oldB[b] - maxShift);
		want = std::min(want, oldB[b] + maxShift);
		want = std::max(want, oldB[b - 1] + halo);
		want = std::min(want, oldB[b + 1] - halo);
		nb[b] = want;
	}
	for (int b = 1; b < world; ++b) nb[b] = std::max(nb[b], nb[b - 1] + minRows);
	for (int b = world - 1; b >= 1; --b) nb[b] = std::min(nb[b], nb[b + 1] - minRows);
	for (int b = 1; b <= world; ++b)
		if (nb[b] - nb[b - 1] < minRows) return oldB; // the grid is too short for this many strips: leave it alone
	return nb;
}

// Called at the top of a step that is due: histogram of the owned rows, one all-reduce, the plan.  Sets
// pendingRetarget (the same on every rank) and this rank's new rows.
int plan_rebalance(SphSim *s) {
	const GridDesc &g = s->grid;
	const int world = s->strip.world, rank = s->strip.rank;
	NcclApi &nc = g_nccl;
	const size_t words = (size_t)g.gy + (size_t)world;
	if (!s->dRowCounts) CU(s, cudaMalloc(&s->dRowCounts, words * sizeof(uint32_t)));
	s->hRowCounts.resize(words);
	CU(s, cudaMemsetAsync(s->dRowCounts, 0, words * sizeof(uint32_t), s->stream));
	const int rows = g.ownHi - g.ownLo;
	row_counts_kernel<<<(rows + 255) / 256, 256, 0, s->stream>>>(g, s->cellStart, s->dRowCounts, rank);
	CU(s, cudaGetLastError());
	const int rca = nc.AllReduce(s->dRowCounts, s->dRowCounts, words, kNcclUint32, kNcclSum, s->comm, s->stream);
	if (rca != 0) return fail(s, SPH_ERR_COMM, "NCCL all-reduce (row histogram) failed: %s", nc.GetErrorString(rca));
	CU(s, cudaMemcpyAsync(s->hRowCounts.data(), s->dRowCounts, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	std::vector<int> oldB(world + 1);
	for (int r = 0; r < world; ++r) oldB[r] = (int)s->hRowCounts[(size_t)g.gy + r];
	oldB[world] = g.gy;
	const std::vector<int> nb = plan_strip_bounds(s->hRowCounts.data(), g.gy, oldB, s->strip.halo, s->rebalanceMaxShift);
	s->pendingRetarget = nb != oldB;
	s->pendLo = nb[rank];
	s->pendHi = nb[rank + 1];
	return SPH_OK;
}

// Between the viscosity pass (previous grid, old rows) and the grid build of the same step: the rank's rows and its
// window change; nothing is copied or reallocated (the cell arrays were sized for the whole grid).  predict_key_kernel
// then keeps / sends by the new rows while authority still follows the old ones (StripDesc::authLo/authHi), and the one
// neighbour exchange of the step carries the particles whose rows changed hands together with the usual halo.
void apply_retarget(SphSim *s) {
	GridDesc &g = s->grid;
	const int halo = s->strip.halo;
	g.ownLo = s->pendLo;
	g.ownHi = s->pendHi;
	g.rowLo = std::max(0, g.ownLo - halo);
	g.rowHi = std::min(g.gy, g.ownHi + halo);
	g.nCells = (uint32_t)(g.rowHi - g.rowLo) * (uint32_t)g.gx;
	s->nTiles = (g.nCells + SPH_SCAN_TILE - 1) / SPH_SCAN_TILE;
	for (StepGraph &c : s->graphs) cudaGraphExecDestroy(c.exec); // the grid description is baked into every launch
	s->graphs.clear();
	s->haloMsgRecords = s->strip.haloCap; // this exchange also carries the rows that change hands: ship whole buffers once
	s->pendingRetarget = false;
	s->rebalances++;
}

int run_delta(SphSim *s, const PairParams &k, unsigned nb) {
	const bool exact = s->cfg.fp_mode == SPH_FP_EXACT;
	if (s->cfg.solver == SPH_SOLVER_COLORED_GS) {
		if (exact) launch_sweeps<Exact, SWEEP_DELTA>(s, k);
		else launch_sweeps<Fast, SWEEP_DELTA>(s, k);
	} else {
		if (exact) launch_delta<Exact>(s, k, nb);
		else launch_delta<Fast>(s, k, nb);
		s->pos.flip();
	}
	return SPH_OK;
}

int append_particles(SphSim *s, size_t n, const float *posXY, const float *accXY, uint64_t *firstIndex) {
	if (n == 0) {
		if (firstIndex) *firstIndex = s->nextId;
		return SPH_OK;
	}
	if (s->hostN + n > s->capacity)
		return fail(s, SPH_ERR_CAPACITY, "particle capacity %u exceeded (have %llu, adding %zu)", s->capacity, (unsigned long long)s->hostN, n);
	const size_t at = (size_t)s->hostN, bytes = n * sizeof(float2);
	CU(s, cudaMemcpyAsync(s->pos.in() + at, posXY, bytes, cudaMemcpyHostToDevice, s->stream));
	CU(s, cudaMemcpyAsync(s->prev.in() + at, posXY, bytes, cudaMemcpyHostToDevice, s->stream)); // ParticleData(pos), demo4.h:101-106
	CU(s, cudaMemsetAsync(s->vel.in() + at, 0, bytes, s->stream));
	if (accXY) CU(s, cudaMemcpyAsync(s->acc.in() + at, accXY, bytes, cudaMemcpyHostToDevice, s->stream));
	else CU(s, cudaMemsetAsync(s->acc.in() + at, 0, bytes, s->stream));
	CU(s, cudaMemsetAsync(s->dens.in() + at, 0, bytes, s->stream));
	CU(s, cudaMemsetAsync(s->press.in() + at, 0, bytes, s->stream));
	fill_ids_kernel<<<blocks_for(n), SPH_THREADS, 0, s->stream>>>(s->id.in(), (uint32_t)at, (uint32_t)n, (uint32_t)s->nextId);
	if (firstIndex) *firstIndex = s->nextId;
	s->accFrom = std::min<uint32_t>(s->accFrom, (uint32_t)at);
	s->hostN += n;
	s->nextId += n;
	// n grows, nSorted stays: the newcomers are in nobody's neighbour list yet (demo4.cpp:148)
	grow_count_kernel<<<1, 1, 0, s->stream>>>(s->dCtr, (uint32_t)s->hostN);
	CU(s, cudaGetLastError());
	return SPH_OK;
}

} // namespace


// ============================================================================================
extern "C" {

int sph_abi_version(void) { return SPHB200_ABI_VERSION; }

int sph_config_default(SphConfig *cfg) {
	if (!cfg) return SPH_ERR_INVALID;
	memset(cfg, 0, sizeof(*cfg));
	cfg->struct_size = sizeof(SphConfig);
	cfg->domain_width = 10.0f;                    // sph.h:19
	cfg->domain_height = 10.0f / (16.0f / 9.0f);  // sph.h:18,20
	cfg->cell_size = 6.0f * 0.05f;                // sph.h:35-36,60
	cfg->max_particles = 10000;                   // sph.h:70
	cfg->device = 0;
	cfg->fp_mode = SPH_FP_EXACT;
	cfg->flags = 0;
	cfg->relaxation = 1.0f;
	cfg->solver = SPH_SOLVER_COLORED_GS;
	cfg->sweep_capacity = 0;
	cfg->rank = 0;
	cfg->world_size = 1;
	cfg->halo_capacity = 0;
	return SPH_OK;
}

int sph_last_error(SphHandle h, char *buf, size_t n) {
	if (!buf || n == 0) return SPH_ERR_INVALID;
	const std::string &e = h ? h->err : g_createError;
	snprintf(buf, n, "%s", e.c_str());
	return SPH_OK;
}

int sph_create(const SphConfig *cfg, SphHandle *out) {
	if (!cfg || !out) return fail(nullptr, SPH_ERR_INVALID, "null argument");
	if (cfg->struct_size != sizeof(SphConfig)) return fail(nullptr, SPH_ERR_INVALID, "SphConfig size mismatch (%u vs %zu)", cfg->struct_size, sizeof(SphConfig));
	if (!(cfg->cell_size > 0.0f) || !(cfg->domain_width > 0.0f) || !(cfg->domain_height > 0.0f)) return fail(nullptr, SPH_ERR_INVALID, "bad domain");
	if (cfg->max_particles == 0 || cfg->max_particles > 0xFFFFFF00ull) return fail(nullptr, SPH_ERR_INVALID, "bad max_particles");
	if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) return fail(nullptr, SPH_ERR_INVALID, "bad rank/world_size");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return fail(nullptr, SPH_ERR_CUDA, "no CUDA device: libsphb200 has no CPU fallback");
	if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, SPH_ERR_INVALID, "device %d of %d", cfg->device, ndev);

	SphSim *s = new SphSim();
	s->cfg = *cfg;
	default_params(&s->params);
	s->omega = cfg->relaxation > 0.0f ? cfg->relaxation : 1.0f;
	if (cfg->solver != SPH_SOLVER_COLORED_GS && cfg->solver != SPH_SOLVER_GATHER) {
		delete s;
		return fail(nullptr, SPH_ERR_INVALID, "unknown solver %d", cfg->solver);
	}
	s->sweepCap = cfg->sweep_capacity ? cfg->sweep_capacity : 512u;
	s->sweepAdaptive = cfg->sweep_capacity == 0;
	s->useGraphs = !(cfg->flags & SPH_FLAG_NO_GRAPHS);
	if (s->sweepCap < 32 || s->sweepCap > 3072) {
		delete s;
		return fail(nullptr, SPH_ERR_INVALID, "sweep_capacity %u outside 32..3072", s->sweepCap);
	}
	GridDesc &g = s->grid;
	g.halfW = cfg->domain_width * 0.5f;  // sph.h:21
	g.halfH = cfg->domain_height * 0.5f; // sph.h:22
	g.cell = cfg->cell_size;
	g.gx = (int)(cfg->domain_width / cfg->cell_size);  // sph.h:61
	g.gy = (int)(cfg->domain_height / cfg->cell_size); // sph.h:62
	if (g.gx < 1 || g.gy < 1 || g.gx > 65535 || g.gy > 65535) {
		delete s;
		return fail(nullptr, SPH_ERR_INVALID, "grid %d x %d outside 1..65535", g.gx, g.gy);
	}
	g.rowLo = g.ownLo = 0;
	g.rowHi = g.ownHi = g.gy;
	g.nCells = 0;
	s->capacity = (uint32_t)cfg->max_particles;

#define CUC(call)                                                                                        \
	do {                                                                                                 \
		cudaError_t e__ = (call);                                                                        \
		if (e__ != cudaSuccess) {                                                                        \
			int rc__ = fail(nullptr, SPH_ERR_CUDA, "%s -> %s", #call, cudaGetErrorString(e__));          \
			sph_destroy(s);                                                                              \
			return rc__;                                                                                 \
		}                                                                                                \
	} while (0)

	CUC(cudaSetDevice(cfg->device));
	CUC(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
	CUC(cudaStreamCreateWithFlags(&s->copyStream, cudaStreamNonBlocking));
	CUC(cudaEventCreateWithFlags(&s->renderReady, cudaEventDisableTiming));
	CUC(cudaEventCreateWithFlags(&s->copyDone, cudaEventDisableTiming));
	CUC(cudaMalloc(&s->dCtr, sizeof(Counters)));
	CUC(cudaMemset(s->dCtr, 0, sizeof(Counters)));
	CUC(cudaMallocHost(&s->hCtr, sizeof(Counters)));
	CUC(cudaMallocHost(&s->hCtrLag, sizeof(Counters)));
	memset(s->hCtrLag, 0, sizeof(Counters));
	CUC(cudaEventCreateWithFlags(&s->lagEvent, cudaEventDisableTiming));
	const size_t cap = s->capacity;
	CUC(alloc2(s->pos, cap));
	CUC(alloc2(s->prev, cap));
	CUC(alloc2(s->vel, cap));
	CUC(alloc2(s->acc, cap));
	CUC(alloc2(s->dens, cap));
	CUC(alloc2(s->press, cap));
	CUC(alloc2(s->id, cap));
	CUC(alloc2(s->cellOf, cap));
	CUC(cudaMalloc(&s->cellNew, cap * sizeof(uint32_t)));
	CUC(cudaMalloc(&s->rank, cap * sizeof(uint32_t)));
	CUC(cudaMalloc(&s->slotId, cap * sizeof(uint32_t)));
	CUC(cudaMalloc(&s->colorCount, 16 * sizeof(uint32_t)));
	CUC(cudaMemset(s->colorCount, 0, 16 * sizeof(uint32_t)));
	s->strip.rank = cfg->rank;
	s->strip.world = cfg->world_size;
	s->strip.halo = cfg->halo_rows > 0 ? cfg->halo_rows : kDefaultHaloRows;
	{
		// default strips: an even split of the grid rows; sph_set_strip rebalances before particles are added
		const int lo = (int)((int64_t)g.gy * cfg->rank / cfg->world_size), hi = (int)((int64_t)g.gy * (cfg->rank + 1) / cfg->world_size);
		int rcs = configure_strip(s, lo, hi);
		if (rcs != SPH_OK) {
			g_createError = s->err;
			sph_destroy(s);
			return rcs;
		}
	}
	if (cfg->world_size > 1) {
		s->strip.haloCap = (uint32_t)(cfg->halo_capacity ? cfg->halo_capacity : std::max<uint64_t>(cap / 4, 4096));
		s->haloBytes = sizeof(HaloBuffer) + (size_t)s->strip.haloCap * sizeof(HaloRecord);
		for (int d = 0; d < 2; ++d) {
			CUC(cudaMalloc(&s->sendBuf[d], s->haloBytes));
			CUC(cudaMalloc(&s->recvBuf[d], s->haloBytes));
			CUC(cudaMemset(s->sendBuf[d], 0, sizeof(HaloBuffer)));
			CUC(cudaMemset(s->recvBuf[d], 0, sizeof(HaloBuffer)));
		}
		s->strip.sendDown = s->sendBuf[0];
		s->strip.sendUp = s->sendBuf[1];
		s->haloMsgRecords = s->strip.haloCap;
		CUC(cudaMalloc(&s->dPeak, 2 * sizeof(uint32_t)));
		CUC(cudaMemset(s->dPeak, 0, 2 * sizeof(uint32_t)));
		s->hostN = cap; // launch bound: the live count is only known on the device
	}
	CUC(cudaMalloc(&s->dOwnedCount, sizeof(uint32_t)));
	CUC(cudaMalloc(&s->dBodies, kMaxBodies * sizeof(DevBody)));
	for (auto &e : s->phaseEv) CUC(cudaEventCreate(&e));
	for (auto &e : s->marks) CUC(cudaEventCreate(&e));
	reset_stats_kernel<<<1, 1, 0, s->stream>>>(s->dCtr);
	CUC(cudaStreamSynchronize(s->stream));
#undef CUC
	*out = s;
	return SPH_OK;
}

int sph_destroy(SphHandle s) {
	if (!s) return SPH_OK;
	if (s->stream) cudaStreamSynchronize(s->stream);
	if (s->copyStream) {
		cudaStreamSynchronize(s->copyStream);
		cudaStreamDestroy(s->copyStream);
	}
	if (s->renderReady) cudaEventDestroy(s->renderReady);
	if (s->copyDone) cudaEventDestroy(s->copyDone);
	free2(s->pos);
	free2(s->prev);
	free2(s->vel);
	free2(s->acc);
	free2(s->dens);
	free2(s->press);
	free2(s->id);
	free2(s->cellOf);
	cudaFree(s->cellNew);
	cudaFree(s->rank);
	cudaFree(s->slotId);
	cudaFree(s->cellCount);
	cudaFree(s->cellStart);
	cudaFree(s->tileSums);
	cudaFree(s->colorCount);
	cudaFree(s->colorList);
	cudaFree(s->sweepFlow);
	cudaFree(s->rowColor);
	cudaFree(s->dBodies);
	cudaFree(s->dRecords);
	cudaFree(s->dRenderPos);
	cudaFree(s->dRenderCol);
	cudaFree(s->dCellXY);
	for (int d = 0; d < 2; ++d) {
		cudaFree(s->sendBuf[d]);
		cudaFree(s->recvBuf[d]);
	}
	cudaFree(s->dPeak);
	cudaFree(s->dOwnedCount);
	cudaFree(s->dOwnedIds);
	if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
	cudaFree(s->dCtr);
	if (s->hCtr) cudaFreeHost(s->hCtr);
	for (StepGraph &g : s->graphs)
		if (g.exec) cudaGraphExecDestroy(g.exec);
	if (s->hCtrLag) cudaFreeHost(s->hCtrLag);
	if (s->hOwnedCount) cudaFreeHost(s->hOwnedCount);
	cudaFree(s->dRowCounts);
	if (s->lagEvent) cudaEventDestroy(s->lagEvent);
	for (auto &e : s->phaseEv)
		if (e) cudaEventDestroy(e);
	for (auto &e : s->marks)
		if (e) cudaEventDestroy(e);
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
	return SPH_OK;
}

// ---- parameters ---------------------------------------------------------------------------
int sph_set_params(SphHandle s, const SphParams *p) {
	CHECK_HANDLE(s);
	if (!p) return fail(s, SPH_ERR_INVALID, "null params");
	s->params = *p;
	s->params.inv_kernel_height = 1.0f / s->params.kernel_height; // copy-ctor, sph.h:100-110
	return SPH_OK;
}
int sph_get_params(SphHandle s, SphParams *out) {
	CHECK_HANDLE(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	*out = s->params;
	return SPH_OK;
}
int sph_set_gravity(SphHandle s, float gx, float gy) {
	CHECK_HANDLE(s);
	s->gravity = make_float2(gx, gy);
	return SPH_OK;
}
int sph_add_external_force(SphHandle s, float fx, float fy) {
	CHECK_HANDLE(s);
	s->extForce.x += fx; // externalForce += force, demo4.h:189-191
	s->extForce.y += fy;
	return SPH_OK;
}
int sph_clear_external_force(SphHandle s) {
	CHECK_HANDLE(s);
	s->extForce = make_float2(0, 0);
	return SPH_OK;
}
int sph_set_relaxation(SphHandle s, float omega) {
	CHECK_HANDLE(s);
	if (!(omega > 0.0f)) return fail(s, SPH_ERR_INVALID, "relaxation must be > 0");
	s->omega = omega;
	return SPH_OK;
}
int sph_grid_dims(SphHandle s, int32_t *gx, int32_t *gy) {
	CHECK_HANDLE(s);
	if (gx) *gx = s->grid.gx;
	if (gy) *gy = s->grid.gy;
	return SPH_OK;
}

// ---- bodies ----------------------------------------------------------------------------------
int sph_clear_bodies(SphHandle s) {
	CHECK_HANDLE(s);
	s->bodies.clear();
	s->bodiesDirty = true;
	return SPH_OK;
}
int sph_add_plane(SphHandle s, float nx, float ny, float d) {
	CHECK_HANDLE(s);
	DevBody b = {};
	b.type = BODY_PLANE;
	b.f[0] = nx; b.f[1] = ny; b.f[2] = d;
	return add_body(s, b);
}
int sph_add_circle(SphHandle s, float x, float y, float r) {
	CHECK_HANDLE(s);
	DevBody b = {};
	b.type = BODY_CIRCLE;
	b.f[0] = x; b.f[1] = y; b.f[2] = r;
	return add_body(s, b);
}
int sph_add_segment(SphHandle s, float ax, float ay, float bx, float by) {
	CHECK_HANDLE(s);
	DevBody b = {};
	b.type = BODY_SEGMENT;
	b.f[0] = ax; b.f[1] = ay; b.f[2] = bx; b.f[3] = by;
	return add_body(s, b);
}
int sph_add_polygon(SphHandle s, size_t n, const float *xy) {
	CHECK_HANDLE(s);
	if (!xy || n < 3 || n > kMaxPolyVerts) return fail(s, SPH_ERR_INVALID, "polygon needs 3..%zu vertices (sph.h:161), got %zu", kMaxPolyVerts, n);
	DevBody b = {};
	b.type = BODY_POLYGON;
	b.nverts = (int32_t)n;
	memcpy(b.f, xy, n * 2 * sizeof(float));
	return add_body(s, b);
}
int sph_body_count(SphHandle s, size_t *out) {
	CHECK_HANDLE(s);
	if (out) *out = s->bodies.size();
	return SPH_OK;
}

// ---- particles -------------------------------------------------------------------------------
int sph_clear_particles(SphHandle s) {
	CHECK_HANDLE(s);
	s->hostN = s->cfg.world_size > 1 ? s->capacity : 0;
	s->nextId = 0;
	s->accFrom = 0xFFFFFFFFu;
	s->steppedOnce = false;
	set_counts_kernel<<<1, 1, 0, s->stream>>>(s->dCtr, 0, 0);
	CU(s, cudaMemsetAsync(s->cellStart, 0, ((size_t)s->grid.nCells + 1) * sizeof(uint32_t), s->stream));
	CU(s, cudaMemsetAsync(s->colorCount, 0, 9 * sizeof(uint32_t), s->stream)); // no occupied cells: the sweeps have nothing to visit
	CU(s, cudaGetLastError());
	return SPH_OK;
}
int sph_clear_emitters(SphHandle s) {
	CHECK_HANDLE(s);
	s->emitters.clear();
	return SPH_OK;
}

int sph_add_particles(SphHandle s, size_t n, const float *posXY, const float *accXY, uint64_t *firstIndex) {
	CHECK_HANDLE(s);
	if (n && !posXY) return fail(s, SPH_ERR_INVALID, "null positions");
	if (s->cfg.world_size > 1) return fail(s, SPH_ERR_STATE, "host-side particle lists are single-GPU; use sph_add_volume_hashed with strips");
	int rc = append_particles(s, n, posXY, accXY, firstIndex);
	if (rc != SPH_OK) return rc;
	// the pageable source must stay valid until the copies ran
	CU(s, cudaStreamSynchronize(s->stream));
	return SPH_OK;
}

static inline void random_direction(float *x, float *y) { // Vec2RandomDirection, vecmath.h:317-322
	float d = rand() / (float)RAND_MAX;
	float angle = d * ((float)M_PI * 2.0f);
	*x = cosf(angle);
	*y = sinf(angle);
}

int sph_add_volume(SphHandle s, float cx, float cy, float fx, float fy, int countX, int countY, float spacing) {
	CHECK_HANDLE(s);
	if (countX <= 0 || countY <= 0) return SPH_OK; // zero-sized volumes add nothing (app.cpp:522 relies on it)
	const float h = s->params.kernel_height; // the reference uses the constant kSPHKernelHeight (demo4.cpp:176); SetParams keeps them equal (sph.h:113)
	std::vector<float> pos((size_t)countX * countY * 2), acc((size_t)countX * countY * 2);
	const float offX = (countX * spacing) * 0.5f, offY = (countY * spacing) * 0.5f; // demo4.cpp:170
	const float baseX = cx - offX, baseY = cy - offY;
	size_t k = 0;
	for (int yi = 0; yi < countY; ++yi)
		for (int xi = 0; xi < countX; ++xi, ++k) {
			float px = (float)xi * spacing, py = (float)yi * spacing; // demo4.cpp:173
			px = spacing * 0.5f + px;                                 // :174
			py = spacing * 0.5f + py;
			px = baseX + px;                                          // :175
			py = baseY + py;
			float jx, jy;
			random_direction(&jx, &jy);
			jx = (jx * h) * 0.01f; // * kSPHKernelHeight * kSPHVolumeParticleDistributionScale, :176
			jy = (jy * h) * 0.01f;
			pos[2 * k] = jx + px;
			pos[2 * k + 1] = jy + py;
			acc[2 * k] = fx;
			acc[2 * k + 1] = fy;
		}
	return sph_add_particles(s, k, pos.data(), acc.data(), nullptr);
}

int sph_add_volume_hashed(SphHandle s, float cx, float cy, float fx, float fy, int64_t countX, int64_t countY, float spacing, uint64_t seed) {
	CHECK_HANDLE(s);
	if (countX <= 0 || countY <= 0) return SPH_OK;
	const uint64_t total = (uint64_t)countX * (uint64_t)countY;
	if (s->nextId + total > 0xFFFFFF00ull) return fail(s, SPH_ERR_CAPACITY, "particle ids exceed 32 bits");
	const float offX = ((float)countX * spacing) * 0.5f, offY = ((float)countY * spacing) * 0.5f;
	const float baseX = cx - offX, baseY = cy - offY;
	const float jitter = s->params.kernel_height * 0.01f;
	// lattice rows that can land in this rank's strip (one cell of slack for the jitter)
	const GridDesc &g = s->grid;
	int64_t rowFirst = 0, rowCount = countY;
	if (s->cfg.world_size > 1) {
		const float yLo = (float)g.ownLo * g.cell - g.halfH - g.cell, yHi = (float)g.ownHi * g.cell - g.halfH + g.cell;
		int64_t a = (int64_t)floorf((yLo - baseY) / spacing) - 1, b = (int64_t)ceilf((yHi - baseY) / spacing) + 1;
		if (g.ownLo == 0) a = 0;           // clamped cells absorb everything below / above the grid
		if (g.ownHi == g.gy) b = countY;
		rowFirst = std::max<int64_t>(0, a);
		rowCount = std::min<int64_t>(countY, b) - rowFirst;
	}
	if (rowCount > 0) {
		const uint64_t work = (uint64_t)countX * (uint64_t)rowCount;
		volume_hashed_kernel<<<blocks_for(work), SPH_THREADS, 0, s->stream>>>(g, s->dCtr, s->capacity, s->pos.in(), s->prev.in(), s->vel.in(), s->acc.in(),
		                                                                    s->dens.in(), s->press.in(), s->id.in(), baseX, baseY, make_float2(fx, fy),
		                                                                    (long long)countX, (long long)rowFirst, (long long)rowCount, spacing, jitter,
		                                                                    seed, (uint32_t)s->nextId);
	}
	clamp_count_kernel<<<1, 1, 0, s->stream>>>(s->dCtr, s->capacity);
	CU(s, cudaMemcpyAsync(s->hCtr, s->dCtr, sizeof(Counters), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	if (s->hCtr->overflow & 1u) return fail(s, SPH_ERR_CAPACITY, "particle capacity %u exceeded by sph_add_volume_hashed", s->capacity);
	s->accFrom = s->cfg.world_size > 1 ? 0u : std::min<uint32_t>(s->accFrom, (uint32_t)s->hostN);
	if (s->cfg.world_size == 1) s->hostN = s->hCtr->n;
	s->nextId += total;
	grow_count_kernel<<<1, 1, 0, s->stream>>>(s->dCtr, s->hCtr->n);
	CU(s, cudaGetLastError());
	return SPH_OK;
}

int sph_add_emitter(SphHandle s, float px, float py, float dx, float dy, float radius, float speed, float rate, float duration) {
	CHECK_HANDLE(s);
	if (s->emitters.size() >= kMaxEmitters) return fail(s, SPH_ERR_CAPACITY, "more than %zu emitters (demo4.cpp:156)", kMaxEmitters);
	HostEmitter e = { px, py, dx, dy, radius, speed, rate, duration, 0.0f, 0.0f, 1 };
	s->emitters.push_back(e);
	return SPH_OK;
}

int sph_local_particle_count(SphHandle s, uint64_t *out) {
	CHECK_HANDLE(s);
	if (s->cfg.world_size > 1) return sph_read_owned(s, nullptr, nullptr, 0, nullptr, 0, nullptr, 0, out);
	if (out) *out = s->hostN;
	return SPH_OK;
}
int sph_particle_count(SphHandle s, uint64_t *out) {
	CHECK_HANDLE(s);
	if (out) *out = s->hostN;
	return SPH_OK;
}

// UpdateEmitter, demo4.cpp:257-284 — clock and rand() on the host, particles appended on the device
static int update_emitters(SphSim *s, float dt) {
	std::vector<float> pos, acc;
	const float spacing = s->params.particle_spacing;
	const float invDt = 1.0f / dt;
	const float h = s->params.kernel_height;
	for (HostEmitter &e : s->emitters) {
		if (!e.active) continue;
		const float rate = 1.0f / e.rate;
		e.elapsed += dt;
		e.totalElapsed += dt;
		if (e.elapsed >= rate) {
			e.elapsed = 0;
			const float ax = (e.dx * e.speed) * invDt, ay = (e.dy * e.speed) * invDt; // :266
			const float dirX = -1.0f * e.dy, dirY = 1.0f * e.dx;                       // Vec2Cross(1.0f, direction), :267
			const int count = (int)floor(e.radius / spacing);                          // :268
			const float offX = ((dirX * (float)count) * spacing) * 0.5f, offY = ((dirY * (float)count) * spacing) * 0.5f; // :270
			const float baseX = e.px - offX, baseY = e.py - offY;
			for (int k = 0; k < count; ++k) {
				float px = (dirX * (float)k) * spacing, py = (dirY * (float)k) * spacing; // :272
				px = (dirX * spacing) * 0.5f + px;                                         // :273
				py = (dirY * spacing) * 0.5f + py;
				px = baseX + px;                                                           // :274
				py = baseY + py;
				float jx, jy;
				random_direction(&jx, &jy);
				jx = (jx * h) * 0.01f; // :275
				jy = (jy * h) * 0.01f;
				pos.push_back(jx + px);
				pos.push_back(jy + py);
				acc.push_back(ax);
				acc.push_back(ay);
			}
		}
		if (e.totalElapsed >= e.duration) e.active = 0; // :280-282
	}
	if (pos.empty()) return SPH_OK;
	return sph_add_particles(s, pos.size() / 2, pos.data(), acc.data(), nullptr);
}

// ---- the reference's built-in scenes (sph.h:307-437) through LoadScenario (app.cpp:477-534) -------
namespace {
struct SceneBody { int type; float px, py, rot, a, b; };  // plane: (a,b) normal | circle: a radius | box: (a,b) half extents
struct SceneVolume { float px, py, w, h, fx, fy; };
struct SceneEmitter { float px, py, dx, dy, radius, speed, rate, duration; };
struct Scene {
	const char *name;
	float gx, gy;
	std::vector<SceneVolume> volumes;
	std::vector<SceneEmitter> emitters;
	std::vector<SceneBody> bodies;
	float spacing, nearStiffness;
};
enum { SC_PLANE, SC_CIRCLE, SC_BOX };

const std::vector<Scene> &scene_table() {
	static std::vector<Scene> table;
	if (!table.empty()) return table;
	const float W = 10.0f, H = W / (16.0f / 9.0f), hw = W * 0.5f, hh = H * 0.5f;      // sph.h:18-22
	const float radius = 0.05f, h = 6.0f * radius, k = 0.6f;                          // sph.h:35-36,41
	const float wallW = W * 0.05f, wallH = H * 0.85f, damW = W * 0.25f, damH = H * 0.95f; // sph.h:307-310
	const float blobW = W * 0.5f, blobH = H * 0.5f;                                   // sph.h:312-313
	const float deg = (float)M_PI / 180.0f;                                           // vecmath.h:9
	const SceneBody floorP = { SC_PLANE, 0, -hh, 0, 0, 1 }, ceilP = { SC_PLANE, 0, hh, 0, 0, -1 };
	const SceneBody leftP = { SC_PLANE, -hw, 0, 0, 1, 0 }, rightP = { SC_PLANE, hw, 0, 0, -1, 0 };
	const std::vector<SceneBody> walls = { floorP, ceilP, leftP, rightP };
	auto with = [&](std::vector<SceneBody> extra) {
		std::vector<SceneBody> v = walls;
		v.insert(v.end(), extra.begin(), extra.end());
		return v;
	};
	// the 8-argument SPHParameters constructor ignores its kernelHeight and restDensity arguments
	// (sph.h:113,117): only the spacing and the near stiffness differ between scenes
	table.push_back({ "Dambreak", 0, -10, { { -hw + damW * 0.5f, 0, damW, damH, 0, 0 } }, {},
	                  with({ { SC_BOX, -hw + damW + wallW * 0.5f + radius, H * 0.1f, 0.0f, wallW * 0.5f, wallH * 0.5f } }), h / 6.0f, k * 10.0f });
	table.push_back({ "Dambreak x 2", 0, -10, { { -hw + damW * 0.5f, 0, damW, damH, 0, 0 }, { hw - damW * 0.5f, 0, damW, damH, 0, 0 } }, {}, walls,
	                  h / 3.0f, k * 20.0f });
	table.push_back({ "Blob", 0, 0, { { 0, 0, blobW, blobH, 0, 0 } }, {}, walls, h / 3.0f, k * 10.0f });
	table.push_back({ "Blob x 2", 0, 0,
	                  { { -blobH * 0.75f, 0, blobH * 0.75f, blobH * 0.75f, 10, 0 }, { blobH * 0.75f, 0, blobH * 0.75f, blobH * 0.75f, -10, 0 } }, {}, walls,
	                  h / 3.0f, k * 10.0f });
	table.push_back({ "Liquid", 0, -2, {}, { { -3.5f, 0.0f, 1, 0, h * 3, 2.5f, 15.0f, 30.0f } }, walls, h / 4.0f, k * 10.0f });
	table.push_back({ "Glass", 0, -10, {}, { { -1.5f, 2.0f, 1, 0, h * 3, 2.5f, 15.0f, 25.0f } },
	                  with({ { SC_BOX, 0.0f, -2.0f, 0.0f, 1.0f, 0.2f }, { SC_BOX, -1.0f, -0.5f, 0.0f, 0.2f, 1.5f }, { SC_BOX, 1.0f, -0.5f, 0.0f, 0.2f, 1.5f } }),
	                  h / 4.0f, k * 6.0f });
	table.push_back({ "Fontain", 0, -10, {}, { { 0, -hh + 1.0f, 0, 1, h * 4, 8.0f, 15.0f, 25.0f } }, walls, h / 4.0f, k * 2.0f });
	table.push_back({ "Fun", 0, -10, {}, { { -4, 2, 1, 0, h * 4, 3.5f, 15.0f, 20.0f } },
	                  { floorP, leftP, rightP, { SC_BOX, -1.5f, 1.0f, deg * -2.5f, 3.5f, 0.1f }, { SC_BOX, 1.5f, -0.25f, deg * 2.5f, 3.5f, 0.1f },
	                    { SC_CIRCLE, -4.0f, -1.5f, 0, 0.5f, 0 }, { SC_BOX, 0, -hh + 0.5f, 0, 0.3f, 1.0f } },
	                  h / 4.0f, k * 6.0f });
	return table;
}
} // namespace

extern "C" int sph_scenario_count(void) { return (int)scene_table().size(); }
extern "C" const char *sph_scenario_name(int idx) {
	if (idx < 0 || idx >= (int)scene_table().size()) return nullptr;
	return scene_table()[(size_t)idx].name;
}
extern "C" int sph_load_scenario(SphHandle s, int idx, int seed) {
	CHECK_HANDLE(s);
	if (idx < 0 || idx >= (int)scene_table().size()) return fail(s, SPH_ERR_INVALID, "scenario %d of %zu", idx, scene_table().size());
	const Scene &sc = scene_table()[(size_t)idx];
	if (seed >= 0) srand((unsigned)seed);
	int rc;
#define TRY(call) do { rc = (call); if (rc != SPH_OK) return rc; } while (0)
	TRY(sph_reset_stats(s)); // app.cpp:480-485
	TRY(sph_clear_bodies(s));
	TRY(sph_clear_particles(s));
	TRY(sph_clear_emitters(s));
	TRY(sph_set_gravity(s, sc.gx, sc.gy));
	SphParams p;
	default_params(&p);
	p.particle_spacing = sc.spacing;
	p.near_stiffness = sc.nearStiffness;
	TRY(sph_set_params(s, &p));
	for (const SceneBody &b : sc.bodies) { // app.cpp:488-517
		if (b.type == SC_PLANE) {
			TRY(sph_add_plane(s, b.a, b.b, b.a * b.px + b.b * b.py)); // Vec2Dot(orientation.col1, position)
		} else if (b.type == SC_CIRCLE) {
			TRY(sph_add_circle(s, b.px, b.py, b.a));
		} else { // CreateBox (sph.h:204-216) rotated by Mat2FromAngle (vecmath.h:358-365), then translated
			const float sn = sinf(b.rot), cs = cosf(b.rot);
			const float lx[4] = { b.a, -b.a, -b.a, b.a }, ly[4] = { b.b, b.b, -b.b, -b.b };
			float xy[8];
			for (int v = 0; v < 4; ++v) {
				xy[2 * v] = (cs * lx[v] + (-sn) * ly[v]) + b.px; // Vec2MultMat2, vecmath.h:287-290
				xy[2 * v + 1] = (sn * lx[v] + cs * ly[v]) + b.py;
			}
			TRY(sph_add_polygon(s, 4, xy));
		}
	}
	const float spacing = s->params.particle_spacing;
	for (const SceneVolume &v : sc.volumes) { // app.cpp:519-527
		const int numX = (int)floor((v.w / spacing)), numY = (int)floor((v.h / spacing));
		TRY(sph_add_volume(s, v.px, v.py, v.fx, v.fy, numX, numY, spacing));
	}
	for (const SceneEmitter &e : sc.emitters) TRY(sph_add_emitter(s, e.px, e.py, e.dx, e.dy, e.radius, e.speed, e.rate, e.duration));
#undef TRY
	return SPH_OK;
}

// one Update() as a sequence of launches on s->stream (also what a step graph captures)
static int enqueue_step(SphSim *s, float dt, const PairParams &k, unsigned nb, float2 force, float invDt, int parts = GRID_ALL) {
	const bool exact = s->cfg.fp_mode == SPH_FP_EXACT;
	if (parts & GRID_FRONT) {
		record_phase(s, 0);
		integrate_kernel<<<blocks_for((s->hostN + 3) / 4), SPH_THREADS, 0, s->stream>>>(s->dCtr, s->vel.in(), s->acc.in(), s->accFrom, force, dt); // also opens the step
		s->accFrom = 0xFFFFFFFFu;
		record_phase(s, PH_INTEGRATE + 1);
		run_viscosity(s, k, nb);
		record_phase(s, PH_VISCOSITY + 1);
		if (s->pendingRetarget) apply_retarget(s); // (never inside a graph capture: a step that re-balances is not graphable)
	}
	int rc = launch_grid_build(s, dt, true, false, true, parts);
	if (rc != SPH_OK) return rc;
	if (!(parts & GRID_BACK)) return SPH_OK;
	if (exact) launch_density<Exact>(s, k, nb);
	else launch_density<Fast>(s, k, nb);
	record_phase(s, PH_DENSITY + 1);
	run_delta(s, k, nb);
	record_phase(s, PH_DELTA + 1);
	collide_velocity_kernel<<<blocks_for((s->hostN + 3) / 4), SPH_THREADS, 0, s->stream>>>(s->dCtr, s->pos.in(), s->prev.in(), s->vel.in(), s->dBodies, (int)s->bodies.size(), invDt, 1, 1, 1); // also closes the step
	record_phase(s, PH_COLLIDE + 1);
	return SPH_OK;
}

static uint32_t buffer_parity(const SphSim *s) {
	return (uint32_t)s->pos.cur | (uint32_t)s->prev.cur << 1 | (uint32_t)s->vel.cur << 2 | (uint32_t)s->acc.cur << 3 | (uint32_t)s->dens.cur << 4 |
	       (uint32_t)s->press.cur << 5 | (uint32_t)s->id.cur << 6 | (uint32_t)s->cellOf.cur << 7;
}
static void set_buffer_parity(SphSim *s, uint32_t p) {
	s->pos.cur = p & 1;
	s->prev.cur = (p >> 1) & 1;
	s->vel.cur = (p >> 2) & 1;
	s->acc.cur = (p >> 3) & 1;
	s->dens.cur = (p >> 4) & 1;
	s->press.cur = (p >> 5) & 1;
	s->id.cur = (p >> 6) & 1;
	s->cellOf.cur = (p >> 7) & 1;
}

// ---- the hot path -------------------------------------------------------------------------------
int sph_step(SphHandle s, float dt) {
	CHECK_HANDLE(s);
	if (!(dt > 0.0f)) return fail(s, SPH_ERR_INVALID, "dt must be > 0");
	if (s->cfg.world_size > 1 && !s->comm) return fail(s, SPH_ERR_STATE, "call sph_comm_init before stepping a multi-GPU handle");
	if (s->cfg.world_size > 1 && !s->emitters.empty()) return fail(s, SPH_ERR_STATE, "emitters are single-GPU");
	if (!s->emitters.empty()) {
		int rc = update_emitters(s, dt);
		if (rc != SPH_OK) return rc;
	}
	int rc = upload_bodies(s);
	if (rc != SPH_OK) return rc;
	if (s->sweepAdaptive && s->lagPending && cudaEventQuery(s->lagEvent) == cudaSuccess) {
		// shared-memory staging sized to twice the longest candidate list seen recently: a host that
		// enqueues hundreds of steps ahead of the device (graph replay costs ~7 us per step) only sees
		// old counts, and blocks that outgrow the staging fall back to the much slower L2 path
		const uint32_t longest = s->hCtrLag->maxNbr;
		if (longest) s->sweepCap = std::min(1024u, std::max(192u, ((2u * longest + 31u) / 32u) * 32u));
		s->lagPending = false;
	}
	const PairParams k = pair_params(s, dt);
	const unsigned nb = blocks_for(s->hostN);
	const float invDt = 1.0f / dt; // demo4.cpp:287
	const float2 force = make_float2(s->gravity.x + s->extForce.x, s->gravity.y + s->extForce.y); // gravity + externalForce, :306

	// Steady state (nothing appended since the last step, no per-phase timing): the ~14
	// launches of a step are replayed from a CUDA graph, which removes the launch gaps that dominate
	// small scenes.  Kernel arguments depend on which half of each double buffer is current, so graphs
	// are cached per buffer parity, staging capacity, dt, force and parameters.
	if (s->cfg.world_size > 1) {
		rc = maybe_resize_halo(s);
		if (rc != SPH_OK) return rc;
	}
	if (s->cfg.world_size > 1 && s->rebalanceEvery > 0 && s->steppedOnce && s->steps % (uint64_t)s->rebalanceEvery == 0) {
		rc = plan_rebalance(s);
		if (rc != SPH_OK) return rc;
	}
	const bool graphable = s->useGraphs && s->accFrom == 0xFFFFFFFFu && !(s->cfg.flags & SPH_FLAG_PHASE_TIMING) && s->steps >= 2 && !s->pendingRetarget;
	// One GPU: the whole step is one graph.  Strips: NCCL send/recv inside a captured stream dead-locked on
	// this stack (NCCL 2.28.9, driver 580), so the launches before and after the exchange are two graphs
	// and the exchange itself is enqueued plainly between them.
	auto run_part = [&](int parts) -> int {
		if (!graphable) return enqueue_step(s, dt, k, nb, force, invDt, parts);
		StepGraphKey key;
		memset(&key, 0, sizeof(key));
		key.parity = buffer_parity(s);
		key.sweepCap = s->sweepCap;
		key.nb = nb;
		key.nbodies = (uint32_t)s->bodies.size();
		key.haloMsgRecords = s->haloMsgRecords;
		key.parts = (uint32_t)parts;
		key.flowEpoch = s->flowEpoch; // the sweep launches carry the pass number over the current grid as an argument
		key.force = force;
		key.k = k;
		StepGraph *g = nullptr;
		for (StepGraph &c : s->graphs)
			if (memcmp(&c.key, &key, sizeof(key)) == 0) g = &c;
		if (!g) {
			if (s->graphs.size() >= 16) { // parameters keep changing: drop the oldest
				cudaGraphExecDestroy(s->graphs.front().exec);
				s->graphs.erase(s->graphs.begin());
			}
			cudaGraph_t graph = nullptr;
			CU(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
			const int rce = enqueue_step(s, dt, k, nb, force, invDt, parts);
			const cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
			if (rce != SPH_OK) return rce;
			if (ce != cudaSuccess) return fail(s, SPH_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
			StepGraph made;
			made.key = key;
			made.parityAfter = buffer_parity(s);
			made.flowEpochAfter = s->flowEpoch;
			CU(s, cudaGraphInstantiate(&made.exec, graph, 0));
			cudaGraphDestroy(graph);
			set_buffer_parity(s, key.parity); // capture only recorded the launches: the state is still "before"
			s->graphs.push_back(made);
			g = &s->graphs.back();
		}
		CU(s, cudaGraphLaunch(g->exec, s->stream));
		set_buffer_parity(s, g->parityAfter);
		s->flowEpoch = g->flowEpochAfter;
		return SPH_OK;
	};
	if (s->cfg.world_size == 1) {
		rc = run_part(GRID_ALL);
		if (rc != SPH_OK) return rc;
	} else {
		rc = run_part(GRID_FRONT);
		if (rc != SPH_OK) return rc;
		rc = enqueue_step(s, dt, k, nb, force, invDt, GRID_EXCHANGE);
		if (rc != SPH_OK) return rc;
		rc = run_part(GRID_BACK);
		if (rc != SPH_OK) return rc;
	}
	CU(s, cudaGetLastError());
	s->steps++;
	s->steppedOnce = true;
	if (s->sweepAdaptive && !s->lagPending) {
		CU(s, cudaMemcpyAsync(s->hCtrLag, s->dCtr, sizeof(Counters), cudaMemcpyDeviceToHost, s->stream));
		CU(s, cudaEventRecord(s->lagEvent, s->stream));
		s->lagPending = true;
	}

	if (s->cfg.flags & SPH_FLAG_PHASE_TIMING) {
		// event k+1 closes phase k; the exchange sits between predict and scan, so its event (index
		// PH_EXCHANGE+1) is the one that opens the scan
		CU(s, cudaEventSynchronize(s->phaseEv[PH_COLLIDE + 1]));
		const int open[PH_COUNT] = { 0, PH_INTEGRATE + 1, PH_VISCOSITY + 1, PH_EXCHANGE + 1, PH_SCAN + 1, PH_REORDER + 1, PH_DENSITY + 1, PH_DELTA + 1, PH_PREDICT + 1 };
		for (int p = 0; p < PH_COUNT; ++p) {
			float ms = 0.0f;
			cudaEventElapsedTime(&ms, s->phaseEv[open[p]], s->phaseEv[p + 1]);
			s->phaseMs[p] += ms;
		}
		s->phaseSteps++;
	}
	return SPH_OK;
}

int sph_sync(SphHandle s) {
	CHECK_HANDLE(s);
	CU(s, cudaStreamSynchronize(s->stream));
	CU(s, cudaStreamSynchronize(s->copyStream));
	s->copyPending = false;
	return SPH_OK;
}

int sph_run_pass(SphHandle s, int pass, float dt) {
	CHECK_HANDLE(s);
	int rc = upload_bodies(s);
	if (rc != SPH_OK) return rc;
	const bool exact = s->cfg.fp_mode == SPH_FP_EXACT;
	const PairParams k = pair_params(s, dt);
	const unsigned nb = blocks_for(s->hostN);
	switch (pass) {
		case SPH_PASS_INTEGRATE: {
			const float2 force = make_float2(s->gravity.x + s->extForce.x, s->gravity.y + s->extForce.y);
			integrate_kernel<<<blocks_for((s->hostN + 3) / 4), SPH_THREADS, 0, s->stream>>>(s->dCtr, s->vel.in(), s->acc.in(), s->accFrom, force, dt);
			s->accFrom = 0xFFFFFFFFu;
		} break;
		case SPH_PASS_VISCOSITY:
			run_viscosity(s, k, nb);
			break;
		case SPH_PASS_PREDICT:
			predict_only_kernel<<<nb, SPH_THREADS, 0, s->stream>>>(s->dCtr, s->pos.in(), s->prev.in(), s->vel.in(), dt);
			break;
		case SPH_PASS_GRID:
			if (s->cfg.world_size > 1) {
				rc = maybe_resize_halo(s);
				if (rc != SPH_OK) return rc;
			}
			begin_step_kernel<<<1, 1, 0, s->stream>>>(s->dCtr);
			rc = launch_grid_build(s, dt, false, true, false);
			if (rc != SPH_OK) return rc;
			commit_kernel<<<1, 1, 0, s->stream>>>(s->dCtr);
			break;
		case SPH_PASS_DENSITY:
			begin_step_kernel<<<1, 1, 0, s->stream>>>(s->dCtr);
			if (exact) launch_density<Exact>(s, k, nb);
			else launch_density<Fast>(s, k, nb);
			break;
		case SPH_PASS_DELTA:
			run_delta(s, k, nb);
			break;
		case SPH_PASS_COLLIDE:
			collide_velocity_kernel<<<blocks_for((s->hostN + 3) / 4), SPH_THREADS, 0, s->stream>>>(s->dCtr, s->pos.in(), s->prev.in(), s->vel.in(), s->dBodies, (int)s->bodies.size(), 1.0f / dt, 1, 0, 0);
			break;
		case SPH_PASS_VELOCITY:
			collide_velocity_kernel<<<blocks_for((s->hostN + 3) / 4), SPH_THREADS, 0, s->stream>>>(s->dCtr, s->pos.in(), s->prev.in(), s->vel.in(), s->dBodies, (int)s->bodies.size(), 1.0f / dt, 0, 1, 0);
			break;
		default: return fail(s, SPH_ERR_INVALID, "unknown pass %d", pass);
	}
	CU(s, cudaGetLastError());
	return SPH_OK;
}

// ---- statistics ----------------------------------------------------------------------------------
int sph_reset_stats(SphHandle s) {
	CHECK_HANDLE(s);
	reset_stats_kernel<<<1, 1, 0, s->stream>>>(s->dCtr);
	memset(s->phaseMs, 0, sizeof(s->phaseMs));
	s->phaseSteps = 0;
	s->steppedOnce = false;
	CU(s, cudaGetLastError());
	return SPH_OK;
}

int sph_get_stats(SphHandle s, SphStats *out) {
	CHECK_HANDLE(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	CU(s, cudaMemcpyAsync(s->hCtr, s->dCtr, sizeof(Counters), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	memset(out, 0, sizeof(*out));
	const Counters &c = *s->hCtr;
	// SPHStatistics(), sph.h:143-149: 500 / 0 / 500 / 0 until something was measured
	out->min_particle_neighbor_count = (c.maxNbr == 0) ? 500 : c.minNbr;
	out->max_particle_neighbor_count = c.maxNbr;
	out->min_cell_particle_count = (c.maxCell == 0) ? 500 : c.minCell;
	out->max_cell_particle_count = c.maxCell;
	out->pair_candidates = c.pairCandidates;
	out->steps = s->steps;
	if (s->phaseSteps) {
		const double inv = 1.0 / (double)s->phaseSteps;
		out->time_integration = (float)(s->phaseMs[PH_INTEGRATE] * inv);
		out->time_viscosity_forces = (float)(s->phaseMs[PH_VISCOSITY] * inv);
		out->time_predict = (float)(s->phaseMs[PH_PREDICT] * inv);
		out->time_update_grid = (float)((s->phaseMs[PH_SCAN] + s->phaseMs[PH_REORDER]) * inv);
		out->time_neighbor_search = 0.0f; // no neighbour lists are materialised
		out->time_density_and_pressure = (float)(s->phaseMs[PH_DENSITY] * inv);
		out->time_delta_positions = (float)(s->phaseMs[PH_DELTA] * inv);
		out->time_collisions = (float)(s->phaseMs[PH_COLLIDE] * inv);
	}
	if (c.lost)
		return fail(s, SPH_ERR_STATE, "%u particle-steps left the rows this rank and its two neighbours hold (moved more than the halo in one step)", c.lost);
	if (c.overflow)
		return fail(s, SPH_ERR_CAPACITY, "device reported a capacity overflow (flags %u: 1 = particles, 2 = halo buffer, 4 = more candidates in one 3x3 block than the sweep queue holds)",
		            c.overflow);
	return SPH_OK;
}

int sph_get_phase_ms(SphHandle s, float out[SPH_NUM_PHASES], uint64_t *steps) {
	CHECK_HANDLE(s);
	for (int p = 0; p < PH_COUNT; ++p) out[p] = s->phaseSteps ? (float)(s->phaseMs[p] / (double)s->phaseSteps) : 0.0f;
	if (steps) *steps = s->phaseSteps;
	return SPH_OK;
}

// ---- readback / injection -------------------------------------------------------------------------
int sph_read_particles(SphHandle s, void *dst, size_t stride) {
	CHECK_HANDLE(s);
	if (!dst || stride < sizeof(ParticleRecord)) return fail(s, SPH_ERR_INVALID, "stride must be >= 48");
	if (s->nextId == 0) return SPH_OK;
	if (!s->dRecords) CU(s, cudaMalloc(&s->dRecords, (size_t)s->capacity * sizeof(ParticleRecord)));
	const uint32_t count = (uint32_t)std::min<uint64_t>(s->nextId, s->capacity);
	gather_records_kernel<<<blocks_for(s->hostN), SPH_THREADS, 0, s->stream>>>(s->dCtr, s->id.in(), s->pos.in(), s->prev.in(), s->vel.in(), s->acc.in(),
	                                                                         s->dens.in(), s->press.in(), s->dRecords, 0u, count);
	CU(s, cudaGetLastError());
	CU(s, copy_strided(dst, stride, s->dRecords, sizeof(ParticleRecord), sizeof(ParticleRecord), count, cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	return SPH_OK;
}

int sph_write_particles(SphHandle s, const void *src, size_t stride) {
	CHECK_HANDLE(s);
	if (!src || stride < sizeof(ParticleRecord)) return fail(s, SPH_ERR_INVALID, "stride must be >= 48");
	if (s->cfg.world_size > 1) return fail(s, SPH_ERR_STATE, "state injection is single-GPU");
	if (s->hostN == 0) return SPH_OK;
	if (!s->dRecords) CU(s, cudaMalloc(&s->dRecords, (size_t)s->capacity * sizeof(ParticleRecord)));
	const uint32_t n = (uint32_t)s->hostN;
	CU(s, copy_strided(s->dRecords, sizeof(ParticleRecord), src, stride, sizeof(ParticleRecord), n, cudaMemcpyHostToDevice, s->stream));
	scatter_records_kernel<<<blocks_for(n), SPH_THREADS, 0, s->stream>>>(n, s->dRecords, s->id.in(), s->pos.in(), s->prev.in(), s->vel.in(), s->acc.in(),
	                                                                    s->dens.in(), s->press.in());
	set_counts_kernel<<<1, 1, 0, s->stream>>>(s->dCtr, n, 0);
	s->accFrom = 0;
	CU(s, cudaGetLastError());
	// re-file the grid from the injected positions (what the reference's update-grid loop would do)
	int rc = sph_run_pass(s, SPH_PASS_GRID, 1.0f);
	if (rc != SPH_OK) return rc;
	CU(s, cudaStreamSynchronize(s->stream));
	return SPH_OK;
}

int sph_render_particles(SphHandle s, void *positions, size_t posStride, void *colors, size_t colorStride) {
	CHECK_HANDLE(s);
	if ((positions && posStride < sizeof(float2)) || (colors && colorStride < sizeof(float4))) return fail(s, SPH_ERR_INVALID, "stride too small");
	if (s->nextId == 0) return SPH_OK;
	if (!s->dRenderPos) {
		CU(s, cudaMalloc(&s->dRenderPos, (size_t)s->capacity * sizeof(float2)));
		CU(s, cudaMalloc(&s->dRenderCol, (size_t)s->capacity * sizeof(float4)));
	}
	const uint32_t count = (uint32_t)std::min<uint64_t>(s->nextId, s->capacity);
	// the device-side snapshot may only be overwritten once the previous frame's copy has left it
	if (s->copyPending) CU(s, cudaStreamWaitEvent(s->stream, s->copyDone, 0));
	render_kernel<<<blocks_for(s->hostN), SPH_THREADS, 0, s->stream>>>(s->dCtr, s->id.in(), s->pos.in(), s->vel.in(), s->dens.in(), s->press.in(),
	                                                                 s->params.rest_density, s->dRenderPos, s->dRenderCol, 0u, count);
	CU(s, cudaGetLastError());
	// the copies run on their own stream, so the next sph_step overlaps them; sph_sync waits for both
	CU(s, cudaEventRecord(s->renderReady, s->stream));
	CU(s, cudaStreamWaitEvent(s->copyStream, s->renderReady, 0));
	if (positions) CU(s, copy_strided(positions, posStride, s->dRenderPos, sizeof(float2), sizeof(float2), count, cudaMemcpyDeviceToHost, s->copyStream));
	if (colors) CU(s, copy_strided(colors, colorStride, s->dRenderCol, sizeof(float4), sizeof(float4), count, cudaMemcpyDeviceToHost, s->copyStream));
	CU(s, cudaEventRecord(s->copyDone, s->copyStream));
	s->copyPending = true;
	return SPH_OK;
}

int sph_wait_render(SphHandle s) {
	CHECK_HANDLE(s);
	if (s->copyPending) CU(s, cudaEventSynchronize(s->copyDone));
	return SPH_OK;
}

int sph_read_cell_start(SphHandle s, uint32_t *out) {
	CHECK_HANDLE(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	CU(s, cudaMemcpyAsync(out, s->cellStart, ((size_t)s->grid.nCells + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	return SPH_OK;
}

int sph_read_cell_counts(SphHandle s, uint32_t *out) {
	CHECK_HANDLE(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	std::vector<uint32_t> start((size_t)s->grid.nCells + 1);
	int rc = sph_read_cell_start(s, start.data());
	if (rc != SPH_OK) return rc;
	for (size_t c = 0; c < s->grid.nCells; ++c) out[c] = start[c + 1] - start[c];
	return SPH_OK;
}

int sph_read_sorted_ids(SphHandle s, uint32_t *out) {
	CHECK_HANDLE(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	if (s->hostN == 0) return SPH_OK;
	CU(s, cudaMemcpyAsync(out, s->id.in(), (size_t)s->hostN * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	return SPH_OK;
}

int sph_read_cell_of_particle(SphHandle s, int32_t *out) {
	CHECK_HANDLE(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	if (s->nextId == 0) return SPH_OK;
	if (!s->dCellXY) CU(s, cudaMalloc(&s->dCellXY, (size_t)s->capacity * sizeof(int2)));
	const uint32_t count = (uint32_t)std::min<uint64_t>(s->nextId, s->capacity);
	CU(s, cudaMemsetAsync(s->dCellXY, 0xFF, (size_t)count * sizeof(int2), s->stream));
	cell_of_particle_kernel<<<blocks_for(s->hostN), SPH_THREADS, 0, s->stream>>>(s->dCtr, s->id.in(), s->cellOf.in(), s->dCellXY, 0u, count);
	CU(s, cudaGetLastError());
	CU(s, cudaMemcpyAsync(out, s->dCellXY, (size_t)count * sizeof(int2), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	return SPH_OK;
}

// ---- host memory, stream, timing marks ------------------------------------------------------------
int sph_host_alloc(void **out, size_t bytes) {
	if (!out) return SPH_ERR_INVALID;
	return cudaMallocHost(out, bytes) == cudaSuccess ? SPH_OK : SPH_ERR_CUDA;
}
int sph_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? SPH_OK : SPH_ERR_CUDA; }

int sph_get_stream(SphHandle s, void **out) {
	CHECK_HANDLE(s);
	if (out) *out = (void *)s->stream;
	return SPH_OK;
}
int sph_mark(SphHandle s, int slot) {
	CHECK_HANDLE(s);
	if (slot < 0 || slot >= 8) return fail(s, SPH_ERR_INVALID, "mark slot 0..7");
	CU(s, cudaEventRecord(s->marks[slot], s->stream));
	return SPH_OK;
}
int sph_elapsed_ms(SphHandle s, int a, int b, float *ms) {
	CHECK_HANDLE(s);
	if (a < 0 || a >= 8 || b < 0 || b >= 8 || !ms) return fail(s, SPH_ERR_INVALID, "mark slot 0..7");
	CU(s, cudaEventSynchronize(s->marks[b]));
	CU(s, cudaEventElapsedTime(ms, s->marks[a], s->marks[b]));
	return SPH_OK;
}

// ---- multi-GPU plumbing: not wired in this build ----------------------------------------------------
int sph_comm_unique_id(uint8_t id128[128]) {
	if (!id128) return SPH_ERR_INVALID;
	if (!g_nccl.load()) return fail(nullptr, SPH_ERR_COMM, "%s", g_nccl.error.c_str());
	NcclUniqueId id;
	const int rc = g_nccl.GetUniqueId(&id);
	if (rc != 0) return fail(nullptr, SPH_ERR_COMM, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
	memcpy(id128, id.internal, 128);
	return SPH_OK;
}
int sph_comm_init(SphHandle s, const uint8_t id128[128]) {
	CHECK_HANDLE(s);
	if (!id128) return fail(s, SPH_ERR_INVALID, "null id");
	if (s->cfg.world_size < 2) return fail(s, SPH_ERR_STATE, "sph_comm_init needs world_size > 1");
	if (s->comm) return fail(s, SPH_ERR_STATE, "communicator already initialised");
	if (!g_nccl.load()) return fail(s, SPH_ERR_COMM, "%s", g_nccl.error.c_str());
	NcclUniqueId id;
	memcpy(id.internal, id128, 128);
	CU(s, cudaSetDevice(s->cfg.device));
	const int rc = g_nccl.CommInitRank(&s->comm, s->cfg.world_size, id, s->cfg.rank);
	if (rc != 0) {
		s->comm = nullptr;
		return fail(s, SPH_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
	}
	return SPH_OK;
}
int sph_set_strip(SphHandle s, int32_t rowBegin, int32_t rowEnd) {
	CHECK_HANDLE(s);
	if (s->nextId != 0) return fail(s, SPH_ERR_STATE, "set the strip before adding particles");
	CU(s, cudaStreamSynchronize(s->stream));
	return configure_strip(s, rowBegin, rowEnd);
}

// the particles this rank owns, compacted (arbitrary order) with their creation ids; any output may be NULL
int sph_read_owned(SphHandle s, uint32_t *ids, void *records, size_t recStride, void *positions, size_t posStride, void *colors, size_t colStride,
                   uint64_t *count) {
	CHECK_HANDLE(s);
	if ((records && recStride < sizeof(ParticleRecord)) || (positions && posStride < sizeof(float2)) || (colors && colStride < sizeof(float4)))
		return fail(s, SPH_ERR_INVALID, "stride too small");
	const size_t cap = s->capacity;
	if (!s->dOwnedIds) CU(s, cudaMalloc(&s->dOwnedIds, cap * sizeof(uint32_t)));
	if (records && !s->dRecords) CU(s, cudaMalloc(&s->dRecords, cap * sizeof(ParticleRecord)));
	if ((positions || colors) && !s->dRenderPos) {
		CU(s, cudaMalloc(&s->dRenderPos, cap * sizeof(float2)));
		CU(s, cudaMalloc(&s->dRenderCol, cap * sizeof(float4)));
	}
	CU(s, cudaMemsetAsync(s->dOwnedCount, 0, sizeof(uint32_t), s->stream));
	gather_owned_kernel<<<blocks_for(s->hostN), SPH_THREADS, 0, s->stream>>>(s->grid, s->dCtr, s->id.in(), s->cellOf.in(), s->pos.in(), s->prev.in(), s->vel.in(),
	                                                                       s->acc.in(), s->dens.in(), s->press.in(), s->params.rest_density, s->dOwnedCount,
	                                                                       s->dOwnedIds, records ? s->dRecords : nullptr,
	                                                                       (positions || colors) ? s->dRenderPos : nullptr, s->dRenderCol);
	CU(s, cudaGetLastError());
	uint32_t n = 0;
	CU(s, cudaMemcpyAsync(&n, s->dOwnedCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	if (count) *count = n;
	if (n == 0) return SPH_OK;
	if (ids) CU(s, cudaMemcpyAsync(ids, s->dOwnedIds, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	if (records) CU(s, copy_strided(records, recStride, s->dRecords, sizeof(ParticleRecord), sizeof(ParticleRecord), n, cudaMemcpyDeviceToHost, s->stream));
	if (positions) CU(s, copy_strided(positions, posStride, s->dRenderPos, sizeof(float2), sizeof(float2), n, cudaMemcpyDeviceToHost, s->stream));
	if (colors) CU(s, copy_strided(colors, colStride, s->dRenderCol, sizeof(float4), sizeof(float4), n, cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	return SPH_OK;
}
// The strip counterpart of sph_render_particles: snapshot of the owned particles (ids, positions, colours) on the
// simulation's stream, device-to-host copies on the copy stream, so the next sph_step overlaps them.  The number of
// owned particles only exists on the device; the copies ship 1.1 x the previous frame's count (+4096) and
// sph_wait_render_owned fetches the rest in the rare frame that outgrew it.  The first frame is read synchronously.
int sph_render_owned(SphHandle s, uint32_t *ids, void *positions, size_t posStride, void *colors, size_t colStride) {
	CHECK_HANDLE(s);
	if (!ids || !positions || !colors) return fail(s, SPH_ERR_INVALID, "sph_render_owned needs ids, positions and colours");
	if (posStride < sizeof(float2) || colStride < sizeof(float4)) return fail(s, SPH_ERR_INVALID, "stride too small");
	if (s->ownedPending) return fail(s, SPH_ERR_STATE, "sph_wait_render_owned was not called for the previous frame");
	const size_t cap = s->capacity;
	if (!s->hOwnedCount) CU(s, cudaMallocHost(&s->hOwnedCount, sizeof(uint32_t)));
	s->ownedIdsDst = ids;
	s->ownedPosDst = positions;
	s->ownedColDst = colors;
	s->ownedPosStride = posStride;
	s->ownedColStride = colStride;
	if (s->ownedLast == 0) { // no estimate yet: synchronous read
		uint64_t n = 0;
		int rc = sph_read_owned(s, ids, nullptr, 0, positions, posStride, colors, colStride, &n);
		if (rc != SPH_OK) return rc;
		*s->hOwnedCount = (uint32_t)n;
		s->ownedShipped = n;
		s->ownedLast = std::max<uint64_t>(n, 1);
		s->ownedPending = true;
		return SPH_OK;
	}
	if (!s->dOwnedIds) CU(s, cudaMalloc(&s->dOwnedIds, cap * sizeof(uint32_t)));
	if (!s->dRenderPos) {
		CU(s, cudaMalloc(&s->dRenderPos, cap * sizeof(float2)));
		CU(s, cudaMalloc(&s->dRenderCol, cap * sizeof(float4)));
	}
	// the device-side snapshot may only be overwritten once the previous frame's copies have left it
	if (s->copyPending) CU(s, cudaStreamWaitEvent(s->stream, s->copyDone, 0));
	CU(s, cudaMemsetAsync(s->dOwnedCount, 0, sizeof(uint32_t), s->stream));
	gather_owned_kernel<<<blocks_for(s->hostN), SPH_THREADS, 0, s->stream>>>(s->grid, s->dCtr, s->id.in(), s->cellOf.in(), s->pos.in(), s->prev.in(), s->vel.in(),
	                                                                       s->acc.in(), s->dens.in(), s->press.in(), s->params.rest_density, s->dOwnedCount,
	                                                                       s->dOwnedIds, nullptr, s->dRenderPos, s->dRenderCol);
	CU(s, cudaGetLastError());
	CU(s, cudaEventRecord(s->renderReady, s->stream));
	CU(s, cudaStreamWaitEvent(s->copyStream, s->renderReady, 0));
	const uint64_t ship = std::min<uint64_t>(cap, s->ownedLast + s->ownedLast / 10 + 4096);
	CU(s, cudaMemcpyAsync(s->hOwnedCount, s->dOwnedCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->copyStream));
	CU(s, cudaMemcpyAsync(ids, s->dOwnedIds, (size_t)ship * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->copyStream));
	CU(s, copy_strided(positions, posStride, s->dRenderPos, sizeof(float2), sizeof(float2), (size_t)ship, cudaMemcpyDeviceToHost, s->copyStream));
	CU(s, copy_strided(colors, colStride, s->dRenderCol, sizeof(float4), sizeof(float4), (size_t)ship, cudaMemcpyDeviceToHost, s->copyStream));
	CU(s, cudaEventRecord(s->copyDone, s->copyStream));
	s->ownedShipped = ship;
	s->copyPending = true;
	s->ownedPending = true;
	return SPH_OK;
}

int sph_wait_render_owned(SphHandle s, uint64_t *count) {
	CHECK_HANDLE(s);
	if (!s->ownedPending) return fail(s, SPH_ERR_STATE, "no sph_render_owned frame in flight");
	if (s->copyPending) CU(s, cudaEventSynchronize(s->copyDone));
	const uint64_t n = std::min<uint64_t>(*s->hOwnedCount, s->capacity);
	if (n > s->ownedShipped) { // the strip grew by more than 10 % in one frame: fetch the tail (the snapshot is still intact)
		const size_t from = (size_t)s->ownedShipped, more = (size_t)(n - s->ownedShipped);
		CU(s, cudaMemcpyAsync(s->ownedIdsDst + from, s->dOwnedIds + from, more * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->copyStream));
		CU(s, copy_strided((char *)s->ownedPosDst + from * s->ownedPosStride, s->ownedPosStride, s->dRenderPos + from, sizeof(float2), sizeof(float2), more,
		                   cudaMemcpyDeviceToHost, s->copyStream));
		CU(s, copy_strided((char *)s->ownedColDst + from * s->ownedColStride, s->ownedColStride, s->dRenderCol + from, sizeof(float4), sizeof(float4), more,
		                   cudaMemcpyDeviceToHost, s->copyStream));
		CU(s, cudaStreamSynchronize(s->copyStream));
	}
	s->ownedLast = std::max<uint64_t>(n, 1);
	s->ownedPending = false;
	if (count) *count = n;
	return SPH_OK;
}

int sph_plan_strip_bounds(const uint32_t *rowCounts, int32_t gridY, const int32_t *oldBounds, int32_t world, int32_t haloRows, int32_t maxShiftRows,
                          int32_t *newBounds) {
	if (!rowCounts || !oldBounds || !newBounds || gridY <= 0 || world <= 0) return SPH_ERR_INVALID;
	const std::vector<int> oldB(oldBounds, oldBounds + world + 1);
	const std::vector<int> nb = plan_strip_bounds(rowCounts, gridY, oldB, haloRows, maxShiftRows > 0 ? maxShiftRows : 2);
	for (int b = 0; b <= world; ++b) newBounds[b] = nb[b];
	return SPH_OK;
}

int sph_set_rebalance(SphHandle s, int32_t everySteps, int32_t maxShiftRows) {
	CHECK_HANDLE(s);
	if (everySteps < 0 || maxShiftRows < 0) return fail(s, SPH_ERR_INVALID, "negative argument");
	if (s->cfg.world_size == 1 || everySteps == 0) {
		s->rebalanceEvery = 0;
		return SPH_OK;
	}
	if (s->nextId != 0 && !s->allocFullGrid) return fail(s, SPH_ERR_STATE, "enable re-balancing before adding particles (the cell arrays are re-sized)");
	s->rebalanceEvery = everySteps;
	s->rebalanceMaxShift = maxShiftRows > 0 ? maxShiftRows : 2;
	if (!s->allocFullGrid) {
		s->allocFullGrid = true;
		CU(s, cudaStreamSynchronize(s->stream));
		return configure_strip(s, s->grid.ownLo, s->grid.ownHi);
	}
	return SPH_OK;
}

int sph_get_strip(SphHandle s, int32_t *rowBegin, int32_t *rowEnd) {
	CHECK_HANDLE(s);
	if (rowBegin) *rowBegin = s->grid.ownLo;
	if (rowEnd) *rowEnd = s->grid.ownHi;
	return SPH_OK;
}

} // extern "C"

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
#include <chrono>
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
	int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
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
		BIND(AllGather, "ncclAllGather")
#undef BIND
		return true;
	}
};
NcclApi g_nccl;
constexpr int kNcclUint8 = 1;  // ncclUint8, nccl.h:279
constexpr int kNcclUint32 = 3; // ncclUint32, nccl.h:281
constexpr int kNcclMax = 2;    // ncclMax, nccl.h:262
constexpr int kNcclSum = 0;    // ncclSum, nccl.h:260
constexpr int kNcclMin = 3;    // ncclMin, nccl.h:263
constexpr int kHaloResizeEvery = 16; // NCCL transport only: steps between re-sizings of the exchange messages
constexpr unsigned long long kHaloWaitTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull; // a neighbour strip that is 20 s late is gone

// How the records packed by predict_key_kernel reach the neighbour strips.
enum Transport {
	TR_NONE = 0,
	TR_PEER,  // stored straight into the neighbour's mailboxes (CUDA IPC mapping of its memory, NVLink): the default
	TR_LOCAL, // the same with plain pointers: all strips live in this process (sph_comm_init_local / sph_step_group)
	TR_NCCL   // packed into local send buffers, shipped as fixed-size ncclSend/ncclRecv messages (fallback, SPH_FLAG_EXCHANGE_NCCL)
};

enum Phase { PH_INTEGRATE, PH_VISCOSITY, PH_PREDICT, PH_SCAN, PH_REORDER, PH_DENSITY, PH_DELTA, PH_COLLIDE, PH_EXCHANGE, PH_COUNT };
static_assert(PH_COUNT == SPH_NUM_PHASES, "phase list");

struct StepGraphKey {
	uint32_t parity, sweepCap, gridSweepCap, workHeavy, nb, nbodies, haloMsgRecords, parts, flowEpoch;
	uint32_t gridGen; // bumped whenever the grid description or the cell arrays change (configure_strip, apply_retarget)
	float2 force;
	PairParams k;
};
struct StepGraph {
	StepGraphKey key;
	uint32_t parityAfter = 0, flowEpochAfter = 0;
	uint32_t exchanges = 0; // strip exchanges the graph publishes (0 or 1)
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
	uint32_t listStride = 0, sweepCap = 256;
	uint32_t gridSweepCap = 256;     // the staging capacity the current grid's lists were classified for (light cells fit it)
	uint32_t workHeavy = 6000;       // m x T from which a cell is swept by a whole block (SweepClass)
	float heavyFactor = 6.0f;        // ... as a multiple of the average m x T
	float teamFrac = 0.5f;           // at most this share of the sweep's blocks work as teams (never more than 7/8: the light queue needs workers)
	float capFactor = 2.2f;          // per-warp staging capacity as a multiple of the average candidate list
	float capAvg = 0.0f;             // candidates per particle the adaptive capacity was last chosen for
	bool capMeasured = false;        // ... from a measurement (not from the scene's nominal density)
	bool sweepAdaptive = true;       // pick the staging capacity from the candidates per particle of recent steps
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
	int transport = TR_NONE;
	unsigned char *mail = nullptr;     // the four mailboxes [from the lower | upper neighbour][exchange parity]: ONE allocation, the unit CUDA IPC exports
	HaloBuffer *mailIn[2][2] = {};
	unsigned char *sendMem = nullptr;  // NCCL transport only: four local send buffers [to the lower | upper neighbour][parity]
	void *peerBase[2] = { nullptr, nullptr }; // CUDA IPC mappings of the neighbours' `mail` (TR_PEER)
	size_t haloBytes = 0;        // bytes per mailbox: header + haloCap records, rounded to 256
	uint32_t haloMsgRecords = 0; // NCCL transport: records shipped per message (all ranks agree; re-sized every kHaloResizeEvery steps)
	uint32_t *dSendCount = nullptr; // [0],[1] slot counters of the current exchange (StripDesc::sendCount), [2] peak count, [3] all-reduced peak
	uint64_t exchanges = 0;      // host mirror of Counters::xseq
	uint32_t *hPeak = nullptr;   // pinned; NCCL transport: all-reduced peak record count, fetched asynchronously
	cudaEvent_t peakEvent = nullptr;
	bool peakPending = false;
	std::vector<SphSim *> group; // TR_LOCAL: all strips of the simulation, in rank order (sph_step_group)
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
	float hostEmitterMs = 0.0f; // host time of UpdateEmitter (demo4.cpp:296-299), summed since sph_reset_stats
	uint64_t hostEmitterSteps = 0;
	cudaEvent_t marks[8] = {};

	// step graphs
	bool useGraphs = true;
	std::vector<StepGraph> graphs;
	uint32_t gridGen = 0;

	// launch configuration of the sweep kernels on THIS handle's device, per (fp mode, pass)
	struct SweepLaunch {
		bool ready = false;
		int flowBlocksPerSM[97] = {};
	} sweepLaunch[2][2];
	int numSMs = 148;
	size_t maxDynSmem = 200 * 1024;

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

// Every entry point works on the handle's own device, whatever the caller's current device is, and puts the
// caller's device back on return (two handles on different GPUs may live in one process).
struct DeviceScope {
	int prev = -1;
	bool changed = false;
	explicit DeviceScope(const SphSim *s) {
		if (cudaGetDevice(&prev) == cudaSuccess && prev != s->cfg.device) changed = cudaSetDevice(s->cfg.device) == cudaSuccess;
	}
	~DeviceScope() {
		if (changed) cudaSetDevice(prev);
	}
	DeviceScope(const DeviceScope &) = delete;
	DeviceScope &operator=(const DeviceScope &) = delete;
};
#define ENTER(h)     \
	CHECK_HANDLE(h); \
	DeviceScope deviceScope__(h)

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
// NCCL transport only.  Message size: NCCL needs it on the host, the record counts only exist on the device.  All
// ranks therefore ship the same fixed number of records per message, re-agreed every kHaloResizeEvery steps as
// 1.5 x the largest count any rank packed since.  The decision is taken WITHOUT stalling the step: the all-reduce and
// the copy of its result are enqueued behind step k, and the new size is applied at the first later step that finds
// the copy complete.  A count above the agreed size raises the overflow flag on the receiver (header count > records
// shipped).  (The peer transport has no message size: records are stored straight into the neighbour's mailbox.)
int maybe_resize_halo(SphSim *s) {
	const StripDesc &sd = s->strip;
	if (s->transport != TR_NCCL) return SPH_OK;
	NcclApi &nc = g_nccl;
	if (s->peakPending && s->exchanges % kHaloResizeEvery == 0) {
		// every rank must switch at the same exchange: the result was requested 15 exchanges ago, so this wait is over
		// before it starts (it is a wait, not a query, only to keep the ranks' decisions identical)
		CU(s, cudaEventSynchronize(s->peakEvent));
		const uint64_t peak = *s->hPeak, want = peak + peak / 2 + 4096;
		s->haloMsgRecords = (uint32_t)std::min<uint64_t>(sd.haloCap, want);
		s->peakPending = false;
	}
	if (!s->peakPending && s->exchanges > 0 && s->exchanges % kHaloResizeEvery == 1) {
		int rca = nc.AllReduce(s->dSendCount + 2, s->dSendCount + 3, 1, kNcclUint32, kNcclMax, s->comm, s->stream);
		if (rca != 0) return fail(s, SPH_ERR_COMM, "NCCL all-reduce failed: %s", nc.GetErrorString(rca));
		CU(s, cudaMemcpyAsync(s->hPeak, s->dSendCount + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
		CU(s, cudaMemsetAsync(s->dSendCount + 2, 0, sizeof(uint32_t), s->stream));
		CU(s, cudaEventRecord(s->peakEvent, s->stream));
		s->peakPending = true;
	}
	return SPH_OK;
}

// Migration + halo in one neighbour exchange.  `front`: after predict_key_kernel has stored this rank's records into
// the out-boxes, publish them (count, then the exchange number, released at system scope).  `back`: wait for the two
// neighbours' boxes of the same exchange and file what arrived behind the local particles.  With the NCCL transport
// the out-boxes are local and one grouped send/recv per neighbour moves them in between.
int exchange_front(SphSim *s) {
	publish_halo_kernel<<<1, 1, 0, s->stream>>>(s->strip, s->dCtr);
	s->exchanges++;
	CU(s, cudaGetLastError());
	return SPH_OK;
}
int exchange_nccl(SphSim *s) {
	const StripDesc &sd = s->strip;
	if (!s->comm) return fail(s, SPH_ERR_STATE, "sph_comm_init was not called on this multi-GPU handle");
	NcclApi &nc = g_nccl;
	const int par = (int)(s->exchanges & 1u); // exchange_front made this exchange current
	const size_t msgBytes = sizeof(HaloBuffer) + (size_t)s->haloMsgRecords * sizeof(HaloRecord);
	int rc = nc.GroupStart();
	if (rc == 0 && sd.rank > 0) {
		rc = nc.Send(sd.outDown[par], msgBytes, kNcclUint8, sd.rank - 1, s->comm, s->stream);
		if (rc == 0) rc = nc.Recv(s->mailIn[0][par], msgBytes, kNcclUint8, sd.rank - 1, s->comm, s->stream);
	}
	if (rc == 0 && sd.rank + 1 < sd.world) {
		rc = nc.Send(sd.outUp[par], msgBytes, kNcclUint8, sd.rank + 1, s->comm, s->stream);
		if (rc == 0) rc = nc.Recv(s->mailIn[1][par], msgBytes, kNcclUint8, sd.rank + 1, s->comm, s->stream);
	}
	const int rcEnd = nc.GroupEnd();
	if (rc == 0) rc = rcEnd;
	if (rc != 0) return fail(s, SPH_ERR_COMM, "NCCL exchange failed: %s", nc.GetErrorString(rc));
	return SPH_OK;
}
int exchange_back(SphSim *s) {
	const StripDesc &sd = s->strip;
	const GridDesc &g = s->grid;
	const bool lower = sd.rank > 0, upper = sd.rank + 1 < sd.world;
	wait_halo_kernel<<<1, 32, 0, s->stream>>>(s->dCtr, lower ? s->mailIn[0][0] : nullptr, lower ? s->mailIn[0][1] : nullptr, upper ? s->mailIn[1][0] : nullptr,
	                                          upper ? s->mailIn[1][1] : nullptr, kHaloWaitTimeoutNs);
	// the count lives in the mailbox: a fixed grid of grid-stride warps
	const uint32_t shipCap = s->transport == TR_NCCL ? s->haloMsgRecords : sd.haloCap;
	const unsigned nb = std::min(blocks_for(sd.haloCap), 148u * 4u);
	if (lower)
		unpack_kernel<<<nb, SPH_THREADS, 0, s->stream>>>(g, s->dCtr, s->mailIn[0][0], s->mailIn[0][1], shipCap, s->capacity, nullptr, nullptr, s->pos.in(), s->prev.in(),
		                                                s->id.in(), s->cellNew, s->rank, s->cellCount);
	if (upper)
		unpack_kernel<<<nb, SPH_THREADS, 0, s->stream>>>(g, s->dCtr, s->mailIn[1][0], s->mailIn[1][1], shipCap, s->capacity, lower ? s->mailIn[0][0] : nullptr,
		                                                lower ? s->mailIn[0][1] : nullptr, s->pos.in(), s->prev.in(), s->id.in(), s->cellNew, s->rank, s->cellCount);
	CU(s, cudaGetLastError());
	return SPH_OK;
}

enum { GRID_FRONT = 1, GRID_EXCHANGE = 2, GRID_BACK = 4, GRID_ALL = 7 };
// parts: FRONT = up to and including the publication of this rank's halo records, BACK = from the wait for the
// neighbours' records on.  The peer transport runs all of it as one graph; the NCCL transport enqueues its send/recv
// (GRID_EXCHANGE, not capturable on this stack) between two graphs; strips of one process (sph_step_group) run every
// strip's FRONT before any BACK, so that no kernel ever waits for work the host has not enqueued yet.
// exchange = false: a purely local re-filing (state injection on strips: every rank was given its whole window, ghost
// rows included, so nothing is sent and nothing is awaited)
int launch_grid_build(SphSim *s, float dt, bool doPredict, bool carry, bool timed, int parts = GRID_ALL, bool exchange = true) {
	const GridDesc &g = s->grid;
	const unsigned nb = blocks_for(s->hostN);
	const bool strips = s->strip.world > 1 && exchange;
	if (parts & GRID_FRONT) {
		if (strips && s->transport == TR_NONE) return fail(s, SPH_ERR_STATE, "sph_comm_init / sph_comm_init_local was not called on this multi-GPU handle");
		CU(s, cudaMemsetAsync(s->cellCount, 0, (size_t)g.nCells * sizeof(uint32_t), s->stream));
		StripDesc sd = s->strip;
		if (!exchange) sd.world = 1; // keep what lies in the window, send nothing
		predict_key_kernel<<<blocks_for((s->hostN + 1) / 2), SPH_THREADS, 0, s->stream>>>(g, sd, s->dCtr, s->pos.in(), s->prev.in(), s->vel.in(), s->id.in(), s->cellOf.in(), s->cellNew,
		                                                      s->rank, s->cellCount, dt, doPredict ? 1 : 0);
		// from the next grid on, authority follows the rows this grid is built for
		s->strip.authLo = g.ownLo;
		s->strip.authHi = g.ownHi;
		if (timed) record_phase(s, PH_PREDICT + 1);
		if (strips) {
			int rc = exchange_front(s);
			if (rc != SPH_OK) return rc;
		}
	}
	if ((parts & GRID_EXCHANGE) && strips && s->transport == TR_NCCL) {
		int rc = exchange_nccl(s);
		if (rc != SPH_OK) return rc;
	}
	if (!(parts & GRID_BACK)) return SPH_OK;
	if (strips) {
		int rc = exchange_back(s);
		if (rc != SPH_OK) return rc;
	}
	if (timed) record_phase(s, PH_EXCHANGE + 1);
	scan_tile_sums_kernel<<<s->nTiles, SPH_THREADS, 0, s->stream>>>(s->cellCount, s->tileSums, g.nCells);
	scan_apply_kernel<<<s->nTiles, SPH_THREADS, 0, s->stream>>>(s->cellCount, s->tileSums, s->cellStart, g.nCells, s->dCtr);
	if (s->cfg.solver == SPH_SOLVER_COLORED_GS) { // occupied cells per colour, each list in row-major order
		const unsigned rowWarps = (unsigned)(g.rowHi - g.rowLo) * 3u, rowBlocks = (rowWarps + SPH_ROWLIST_WARPS - 1) / SPH_ROWLIST_WARPS;
		// light / heavy split of the lists (sph_kernels.cuh, "LIGHT and HEAVY cells"): by this grid's staging capacity
		s->gridSweepCap = s->sweepCap;
		const SweepClass cls = { s->sweepCap, s->workHeavy };
		color_rows_count_kernel<<<rowBlocks, SPH_ROWLIST_WARPS * 32, 0, s->stream>>>(g, cls, s->cellCount, s->cellStart, s->rowColor, s->colorCount, s->sweepFlow);
		s->flowEpoch = 0; // the done flags are fresh: the next sweep over this grid is its first
		color_rows_fill_kernel<<<rowBlocks, SPH_ROWLIST_WARPS * 32, 0, s->stream>>>(g, s->cellCount, s->rowColor, s->colorList, s->listStride, s->colorCount);
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
template <class M>
constexpr int fp_index();
template <>
constexpr int fp_index<Exact>() { return 0; }
template <>
constexpr int fp_index<Fast>() { return 1; }

template <class M, int PASS>
void launch_sweeps(SphSim *s, const PairParams &k) {
	// per handle, not per process: function attributes and occupancy belong to the handle's device
	SphSim::SweepLaunch &cfg = s->sweepLaunch[fp_index<M>()][PASS];
	const int numSMs = s->numSMs;
	int *flowBlocksPerSM = cfg.flowBlocksPerSM;
	if (!cfg.ready) {
		const int smem = (int)s->maxDynSmem;
		cudaFuncSetAttribute(color_sweep_kernel<M, PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		cudaFuncSetAttribute(color_sweep_team_kernel<M, PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
		cudaFuncSetAttribute(color_sweep_flow_kernel<M, PASS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		cudaFuncSetAttribute(color_sweep_flow_kernel<M, PASS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		// resident blocks per SM of the persistent kernel for every staging capacity (32..3072 in steps of 32), asked
		// once here: the first steps of a simulation are never inside a graph capture
		for (uint32_t c32 = 1; c32 <= 96; ++c32) {
			int nbk = 0;
			const size_t bytes = (size_t)SPH_FLOW_WARPS * sweep_bytes_per_warp(c32 * 32u, PASS);
			if (bytes <= s->maxDynSmem) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbk, color_sweep_flow_kernel<M, PASS, true>, SPH_FLOW_WARPS * 32, bytes);
			flowBlocksPerSM[c32] = nbk;
		}
		cudaGetLastError();
		cfg.ready = true;
	}
	// occupied cells of one colour <= min(cells of that colour, particles)
	const uint64_t cells = std::min<uint64_t>(s->listStride, std::max<uint64_t>(s->hostN, 1));
	bool team = s->cfg.world_size == 1 && s->hostN < 131072;
	if (s->cfg.flags & SPH_FLAG_SWEEP_TEAM) team = true;
	if (s->cfg.flags & (SPH_FLAG_SWEEP_WARP | SPH_FLAG_SWEEP_FLOW)) team = false;
	const bool flow = !team && !(s->cfg.flags & SPH_FLAG_SWEEP_WARP);
	if (flow) {
		// persistent: as many blocks as are resident at once (more would only draw a ticket and leave)
		// (the lists of the grid being swept were classified for gridSweepCap: never launch with less, or its light
		// cells would not fit their warp's staging area and take the slow path through L2)
		const uint32_t cap = std::max(s->sweepCap, s->gridSweepCap);
		const size_t smem = (size_t)SPH_FLOW_WARPS * sweep_bytes_per_warp(cap, PASS);
		const int occBlocks = std::max(1, flowBlocksPerSM[std::min(96u, (cap + 31u) / 32u)]);
		const uint64_t want = (9 * cells + SPH_FLOW_WARPS - 1) / SPH_FLOW_WARPS;
		// at least two blocks: heavy cells are swept by the first blocks of the grid as teams, the others must be there for the light ones
		const unsigned blocks = (unsigned)std::max<uint64_t>(2, std::min<uint64_t>(want, (uint64_t)occBlocks * (uint64_t)numSMs));
		const unsigned maxTeams = std::max(1u, std::min(blocks - std::max(1u, blocks / 8u), (unsigned)((float)blocks * s->teamFrac)));
		const uint32_t epoch = ++s->flowEpoch;
		if (s->cfg.world_size > 1) // (the strip variant leaves far ghost cells out of the viscosity sweep)
			color_sweep_flow_kernel<M, PASS, true><<<blocks, SPH_FLOW_WARPS * 32, smem, s->stream>>>(s->grid, k, s->cellStart, s->colorList, s->listStride, s->colorCount, s->pos.in(),
			                                                                                       s->vel.in(), s->press.in(), cap, s->dCtr, s->sweepFlow, epoch, maxTeams);
		else
			color_sweep_flow_kernel<M, PASS, false><<<blocks, SPH_FLOW_WARPS * 32, smem, s->stream>>>(s->grid, k, s->cellStart, s->colorList, s->listStride, s->colorCount, s->pos.in(),
			                                                                                        s->vel.in(), s->press.in(), cap, s->dCtr, s->sweepFlow, epoch, maxTeams);
		return;
	}
	for (int color = 0; color < 9; ++color) {
		const uint32_t *list = s->colorList + (size_t)color * s->listStride;
		if (team) {
			const uint32_t cap = std::min(s->sweepCap, 1024u);
			const unsigned blocks = (unsigned)std::min<uint64_t>(cells, 148u * 16u);
			color_sweep_team_kernel<M, PASS><<<blocks, SPH_TEAM_WARPS * 32, team_smem_bytes(cap, PASS), s->stream>>>(s->grid, k, s->cellStart, list, s->listStride, s->colorCount + color,
			                                                                                               s->pos.in(), s->vel.in(), s->press.in(), cap, s->dCtr);
		} else {
			const size_t smem = (size_t)SPH_SWEEP_WARPS * sweep_bytes_per_warp(s->sweepCap, PASS);
			const unsigned blocks = (unsigned)std::min<uint64_t>((cells + SPH_SWEEP_WARPS - 1) / SPH_SWEEP_WARPS, 148u * 64u);
			color_sweep_kernel<M, PASS><<<blocks, SPH_SWEEP_WARPS * 32, smem, s->stream>>>(s->grid, k, s->cellStart, list, s->listStride, s->colorCount + color, s->pos.in(), s->vel.in(),
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
	// cached step graphs hold the old grid description and the pointers freed below
	for (StepGraph &c : s->graphs)
		if (c.exec) cudaGraphExecDestroy(c.exec);
	s->graphs.clear();
	s->gridGen++;
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
	CU(s, cudaMalloc(&s->sweepFlow, (allocCells + SPH_FLOW_FLAGS + 65536) * sizeof(uint32_t))); // + one decoy word per resident warp (color_sweep_flow_kernel)
	CU(s, cudaMemset(s->sweepFlow, 0xFF, (allocCells + SPH_FLOW_FLAGS + 65536) * sizeof(uint32_t))); // no grid yet: every cell empty
	CU(s, cudaMemset(s->sweepFlow, 0, SPH_FLOW_FLAGS * sizeof(uint32_t)));
	if (s->colorCount) CU(s, cudaMemset(s->colorCount, 0, 32 * sizeof(uint32_t)));
	s->flowEpoch = 0;
	CU(s, cudaMalloc(&s->rowColor, (size_t)allocRows * 3 * 2 * sizeof(uint32_t))); // light counts, then heavy counts
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
	// The two passes above may have pushed a boundary past what (a) and (b) allow (strips that START thinner than
	// minRows): a boundary that jumps over rows of a non-neighbour would strand their particles.  Never trade a
	// correct split for a better balanced one.
	for (int b = 1; b < world; ++b)
		if (std::abs(nb[b] - oldB[b]) > maxShift || nb[b] < oldB[b - 1] + halo || nb[b] > oldB[b + 1] - halo) return oldB;
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

// The same plan for strips that live in one process (sph_comm_init_local): the histograms are summed on the host.
int plan_rebalance_group(SphHandle *hs, int n) {
	const int gy = hs[0]->grid.gy;
	std::vector<uint32_t> total((size_t)gy, 0u);
	std::vector<int> oldB((size_t)n + 1);
	const size_t words = (size_t)gy + (size_t)n;
	for (int r = 0; r < n; ++r) {
		SphSim *s = hs[r];
		DeviceScope dev(s);
		const GridDesc &g = s->grid;
		if (!s->dRowCounts) CU(s, cudaMalloc(&s->dRowCounts, words * sizeof(uint32_t)));
		s->hRowCounts.resize(words);
		CU(s, cudaMemsetAsync(s->dRowCounts, 0, words * sizeof(uint32_t), s->stream));
		row_counts_kernel<<<(g.ownHi - g.ownLo + 255) / 256, 256, 0, s->stream>>>(g, s->cellStart, s->dRowCounts, r);
		CU(s, cudaGetLastError());
		CU(s, cudaMemcpyAsync(s->hRowCounts.data(), s->dRowCounts, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
		CU(s, cudaStreamSynchronize(s->stream));
		for (int row = 0; row < gy; ++row) total[(size_t)row] += s->hRowCounts[(size_t)row];
		oldB[(size_t)r] = g.ownLo;
	}
	oldB[(size_t)n] = gy;
	const std::vector<int> nb = plan_strip_bounds(total.data(), gy, oldB, hs[0]->strip.halo, hs[0]->rebalanceMaxShift);
	for (int r = 0; r < n; ++r) {
		hs[r]->pendingRetarget = nb != oldB;
		hs[r]->pendLo = nb[(size_t)r];
		hs[r]->pendHi = nb[(size_t)r + 1];
	}
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
	s->gridGen++;
	s->haloMsgRecords = s->strip.haloCap; // NCCL transport: this exchange also carries the rows that change hands, ship whole buffers once
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

// Strips: the same list on every rank, each keeps what lands in its rows (append_owned_kernel).  `n` on the device
// grows by what was kept; the host only knows the bound.
int append_particles_owned(SphSim *s, size_t n, const float *posXY, const float *accXY, const void *records, size_t recStride, uint64_t *firstIndex) {
	if (firstIndex) *firstIndex = s->nextId;
	if (n == 0) return SPH_OK;
	if (s->nextId + n > 0xFFFFFF00ull) return fail(s, SPH_ERR_CAPACITY, "particle ids exceed 32 bits");
	struct Staging { // device copies of the host lists, released on every way out
		float2 *pos = nullptr, *acc = nullptr;
		ParticleRecord *rec = nullptr;
		~Staging() {
			cudaFree(pos);
			cudaFree(acc);
			cudaFree(rec);
		}
	} d;
	if (records) {
		CU(s, cudaMalloc(&d.rec, n * sizeof(ParticleRecord)));
		CU(s, copy_strided(d.rec, sizeof(ParticleRecord), records, recStride, sizeof(ParticleRecord), n, cudaMemcpyHostToDevice, s->stream));
	} else {
		CU(s, cudaMalloc(&d.pos, n * sizeof(float2)));
		CU(s, cudaMemcpyAsync(d.pos, posXY, n * sizeof(float2), cudaMemcpyHostToDevice, s->stream));
		if (accXY) {
			CU(s, cudaMalloc(&d.acc, n * sizeof(float2)));
			CU(s, cudaMemcpyAsync(d.acc, accXY, n * sizeof(float2), cudaMemcpyHostToDevice, s->stream));
		}
	}
	const bool window = records != nullptr; // an injected state also fills the ghost rows (see append_owned_kernel)
	append_owned_kernel<<<blocks_for(n), SPH_THREADS, 0, s->stream>>>(s->grid, window ? s->grid.rowLo : s->grid.ownLo, window ? s->grid.rowHi : s->grid.ownHi, s->dCtr,
	                                                                s->capacity, (uint32_t)n, d.pos, d.acc, d.rec, (uint32_t)s->nextId, s->pos.in(),
	                                                                s->prev.in(), s->vel.in(), s->acc.in(), s->dens.in(), s->press.in(), s->id.in());
	clamp_count_kernel<<<1, 1, 0, s->stream>>>(s->dCtr, s->capacity);
	CU(s, cudaGetLastError());
	CU(s, cudaStreamSynchronize(s->stream)); // the pageable sources and the staging buffers must outlive the copies
	s->accFrom = 0u;
	s->nextId += n;
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
	s->sweepCap = cfg->sweep_capacity ? cfg->sweep_capacity : 256u;
	s->sweepAdaptive = cfg->sweep_capacity == 0;
	s->useGraphs = !(cfg->flags & SPH_FLAG_NO_GRAPHS);
	if (const char *e = getenv("SPHB200_HEAVY_FACTOR")) s->heavyFactor = std::max(1.0f, (float)atof(e)); // tuning knobs of the light / heavy split
	if (const char *e = getenv("SPHB200_TEAM_FRAC")) s->teamFrac = std::min(0.875f, std::max(0.01f, (float)atof(e)));
	if (const char *e = getenv("SPHB200_CAP_FACTOR")) s->capFactor = std::max(1.0f, (float)atof(e));
	{
		// the viscosity sweep stages positions + velocities + a 16-bit queue: 18 bytes per candidate and warp
		int smemOptin = 0, sms = 0;
		cudaDeviceGetAttribute(&smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
		if (sms > 0) s->numSMs = sms;
		s->maxDynSmem = (size_t)std::min(std::max(smemOptin - 2048, 48 * 1024), 200 * 1024);
		const uint32_t most = (uint32_t)(s->maxDynSmem / (SPH_FLOW_WARPS * 18u)) / 32u * 32u;
		if (s->sweepCap < 32 || s->sweepCap > most) {
			delete s;
			return fail(nullptr, SPH_ERR_INVALID, "sweep_capacity %u outside 32..%u (shared memory of device %d)", cfg->sweep_capacity, most, cfg->device);
		}
	}
	GridDesc &g = s->grid;
	g.halfW = cfg->domain_width * 0.5f;  // sph.h:21
	g.halfH = cfg->domain_height * 0.5f; // sph.h:22
	g.cell = cfg->cell_size;
	g.gx = (int)(cfg->domain_width / cfg->cell_size);  // sph.h:61
	g.gy = (int)(cfg->domain_height / cfg->cell_size); // sph.h:62
	if (g.gx < 1 || g.gy < 1 || g.gx > 65535 || g.gy > 65535) {
		const int gx = g.gx, gy = g.gy;
		delete s;
		return fail(nullptr, SPH_ERR_INVALID, "grid %d x %d outside 1..65535", gx, gy);
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
	CUC(cudaMalloc(&s->colorCount, 32 * sizeof(uint32_t)));
	CUC(cudaMemset(s->colorCount, 0, 32 * sizeof(uint32_t)));
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
		s->haloBytes = (sizeof(HaloBuffer) + (size_t)s->strip.haloCap * sizeof(HaloRecord) + 255u) & ~(size_t)255u;
		CUC(cudaMalloc(&s->mail, 4 * s->haloBytes));
		CUC(cudaMemset(s->mail, 0, 4 * s->haloBytes));
		for (int side = 0; side < 2; ++side)
			for (int par = 0; par < 2; ++par) s->mailIn[side][par] = reinterpret_cast<HaloBuffer *>(s->mail + (size_t)(side * 2 + par) * s->haloBytes);
		s->haloMsgRecords = s->strip.haloCap;
		CUC(cudaMalloc(&s->dSendCount, 4 * sizeof(uint32_t)));
		CUC(cudaMemset(s->dSendCount, 0, 4 * sizeof(uint32_t)));
		s->strip.sendCount = s->dSendCount;
		CUC(cudaMallocHost(&s->hPeak, sizeof(uint32_t)));
		*s->hPeak = 0;
		CUC(cudaEventCreateWithFlags(&s->peakEvent, cudaEventDisableTiming));
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
	DeviceScope deviceScope__(s);
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
	if (s->transport == TR_PEER)
		for (int d = 0; d < 2; ++d)
			if (s->peerBase[d]) cudaIpcCloseMemHandle(s->peerBase[d]);
	for (SphSim *o : s->group) // strips of one process: the others must not publish into freed mailboxes
		if (o != s) {
			o->transport = TR_NONE;
			o->group.clear();
		}
	cudaFree(s->mail);
	cudaFree(s->sendMem);
	cudaFree(s->dSendCount);
	if (s->hPeak) cudaFreeHost(s->hPeak);
	if (s->peakEvent) cudaEventDestroy(s->peakEvent);
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
	ENTER(s);
	if (!p) return fail(s, SPH_ERR_INVALID, "null params");
	s->params = *p;
	s->params.inv_kernel_height = 1.0f / s->params.kernel_height; // copy-ctor, sph.h:100-110
	return SPH_OK;
}
int sph_get_params(SphHandle s, SphParams *out) {
	ENTER(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	*out = s->params;
	return SPH_OK;
}
int sph_set_gravity(SphHandle s, float gx, float gy) {
	ENTER(s);
	s->gravity = make_float2(gx, gy);
	return SPH_OK;
}
int sph_add_external_force(SphHandle s, float fx, float fy) {
	ENTER(s);
	s->extForce.x += fx; // externalForce += force, demo4.h:189-191
	s->extForce.y += fy;
	return SPH_OK;
}
int sph_clear_external_force(SphHandle s) {
	ENTER(s);
	s->extForce = make_float2(0, 0);
	return SPH_OK;
}
int sph_set_relaxation(SphHandle s, float omega) {
	ENTER(s);
	if (!(omega > 0.0f)) return fail(s, SPH_ERR_INVALID, "relaxation must be > 0");
	s->omega = omega;
	return SPH_OK;
}
int sph_grid_dims(SphHandle s, int32_t *gx, int32_t *gy) {
	ENTER(s);
	if (gx) *gx = s->grid.gx;
	if (gy) *gy = s->grid.gy;
	return SPH_OK;
}

// ---- bodies ----------------------------------------------------------------------------------
int sph_clear_bodies(SphHandle s) {
	ENTER(s);
	s->bodies.clear();
	s->bodiesDirty = true;
	return SPH_OK;
}
int sph_add_plane(SphHandle s, float nx, float ny, float d) {
	ENTER(s);
	DevBody b = {};
	b.type = BODY_PLANE;
	b.f[0] = nx; b.f[1] = ny; b.f[2] = d;
	return add_body(s, b);
}
int sph_add_circle(SphHandle s, float x, float y, float r) {
	ENTER(s);
	DevBody b = {};
	b.type = BODY_CIRCLE;
	b.f[0] = x; b.f[1] = y; b.f[2] = r;
	return add_body(s, b);
}
int sph_add_segment(SphHandle s, float ax, float ay, float bx, float by) {
	ENTER(s);
	DevBody b = {};
	b.type = BODY_SEGMENT;
	b.f[0] = ax; b.f[1] = ay; b.f[2] = bx; b.f[3] = by;
	return add_body(s, b);
}
int sph_add_polygon(SphHandle s, size_t n, const float *xy) {
	ENTER(s);
	if (!xy || n < 3 || n > kMaxPolyVerts) return fail(s, SPH_ERR_INVALID, "polygon needs 3..%zu vertices (sph.h:161), got %zu", kMaxPolyVerts, n);
	DevBody b = {};
	b.type = BODY_POLYGON;
	b.nverts = (int32_t)n;
	memcpy(b.f, xy, n * 2 * sizeof(float));
	return add_body(s, b);
}
int sph_body_count(SphHandle s, size_t *out) {
	ENTER(s);
	if (out) *out = s->bodies.size();
	return SPH_OK;
}

// ---- particles -------------------------------------------------------------------------------
int sph_clear_particles(SphHandle s) {
	ENTER(s);
	s->hostN = s->cfg.world_size > 1 ? s->capacity : 0;
	s->nextId = 0;
	s->accFrom = 0xFFFFFFFFu;
	s->steppedOnce = false;
	set_counts_kernel<<<1, 1, 0, s->stream>>>(s->dCtr, 0, 0);
	CU(s, cudaMemsetAsync(s->cellStart, 0, ((size_t)s->grid.nCells + 1) * sizeof(uint32_t), s->stream));
	CU(s, cudaMemsetAsync(s->colorCount, 0, 18 * sizeof(uint32_t), s->stream)); // no occupied cells: the sweeps have nothing to visit
	CU(s, cudaGetLastError());
	return SPH_OK;
}
int sph_clear_emitters(SphHandle s) {
	ENTER(s);
	s->emitters.clear();
	return SPH_OK;
}

int sph_add_particles(SphHandle s, size_t n, const float *posXY, const float *accXY, uint64_t *firstIndex) {
	ENTER(s);
	if (n && !posXY) return fail(s, SPH_ERR_INVALID, "null positions");
	if (s->cfg.world_size > 1) return append_particles_owned(s, n, posXY, accXY, nullptr, 0, firstIndex); // every rank is given the whole list
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
	ENTER(s);
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
	ENTER(s);
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
	ENTER(s);
	if (s->emitters.size() >= kMaxEmitters) return fail(s, SPH_ERR_CAPACITY, "more than %zu emitters (demo4.cpp:156)", kMaxEmitters);
	HostEmitter e = { px, py, dx, dy, radius, speed, rate, duration, 0.0f, 0.0f, 1 };
	s->emitters.push_back(e);
	return SPH_OK;
}

int sph_local_particle_count(SphHandle s, uint64_t *out) {
	ENTER(s);
	if (s->cfg.world_size > 1) return sph_read_owned(s, nullptr, nullptr, 0, nullptr, 0, nullptr, 0, out);
	if (out) *out = s->hostN;
	return SPH_OK;
}
int sph_particle_count(SphHandle s, uint64_t *out) {
	ENTER(s);
	// strips: the particles of the whole simulation (every rank is told about every creation); what this rank holds is
	// sph_local_particle_count
	if (out) *out = s->cfg.world_size > 1 ? s->nextId : s->hostN;
	return SPH_OK;
}

// UpdateEmitter, demo4.cpp:257-284 — clock and rand() on the host, particles appended on the device
static void emit_particles(SphSim *s, float dt, std::vector<float> &pos, std::vector<float> &acc) {
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
}
static int update_emitters(SphSim *s, float dt) {
	std::vector<float> pos, acc;
	emit_particles(s, dt, pos, acc);
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
	ENTER(s);
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
namespace {
struct StepCtx {
	float dt, invDt;
	float2 force;
	PairParams k;
	unsigned nb;
	bool graphable;
};

// staging capacity and heavy-cell threshold for `avg` candidates per particle (step_prepare)
void set_sweep_class(SphSim *s, float avg) {
	s->capAvg = avg;
	// decisions are taken on the average rounded to a 10 % geometric grid: a measured 80.7 after a nominal 81 must not
	// produce a new threshold, hence a new step graph
	avg = expf(roundf(logf(std::max(avg, 1.0f)) / logf(1.1f)) * logf(1.1f));
	if (s->sweepAdaptive) {
		static const uint32_t steps[] = { 256, 384, 512, 768, 1024 };
		uint32_t cap = 1024;
		for (uint32_t c : steps)
			if ((float)c >= s->capFactor * avg) {
				cap = c;
				break;
			}
		s->sweepCap = cap;
	}
	s->workHeavy = (uint32_t)std::min(4.0e9f, std::max(2048.0f, s->heavyFactor * avg * avg / 9.0f));
}

// host work ahead of a step's launches: emitters (demo4.cpp:296-299), body upload, staging capacity, strip bookkeeping
int step_prepare(SphSim *s, float dt, StepCtx &c, bool planRebalance, bool emit = true) {
	if (!(dt > 0.0f)) return fail(s, SPH_ERR_INVALID, "dt must be > 0");
	if (s->cfg.world_size > 1 && s->transport == TR_NONE) return fail(s, SPH_ERR_STATE, "call sph_comm_init (or sph_comm_init_local) before stepping a multi-GPU handle");
	int rc;
	// (strips, one process per GPU: every rank runs the same emitter clocks and the same rand() sequence - seeded
	// alike, sph_load_scenario - so all ranks emit the same list and each keeps what lands in its rows; strips of one
	// process share ONE rand() stream: sph_step_group emits once, on the first strip, and hands the list to all)
	if (!s->emitters.empty() && emit) {
		const auto t0 = std::chrono::steady_clock::now();
		rc = update_emitters(s, dt);
		if (rc != SPH_OK) return rc;
		s->hostEmitterMs += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
		s->hostEmitterSteps++;
	}
	rc = upload_bodies(s);
	if (rc != SPH_OK) return rc;
	if (s->capAvg == 0.0f && s->params.particle_spacing > 0.0f) {
		// before anything was measured: the scene's nominal density, 3x3 cells of (cell / spacing)^2 particles each -
		// what a volume filled at particle_spacing (AddVolume, demo4.cpp:169-181) gives; a decision that changes
		// only steps later costs a new step graph in the middle of somebody's timed region
		const float perCell = s->grid.cell / s->params.particle_spacing;
		set_sweep_class(s, 9.0f * perCell * perCell);
	}
	if (s->lagPending && cudaEventQuery(s->lagEvent) == cudaSuccess) {
		// Per-warp staging capacity and heavy-cell threshold from the candidates per particle of a recent step (the
		// host enqueues steps ahead of the device and only sees old counts): the capacity holds 2.2 x the AVERAGE list -
		// the cells beyond it are the heavy ones, swept by whole blocks - in coarse steps and with hysteresis, because
		// every new capacity is a new step graph.  A cell is also heavy from 6 x the average work m x T (m ~ T / 9).
		const Counters &lag = *s->hCtrLag;
		const float avg = lag.nOut ? (float)((double)lag.pairCandidates / (double)lag.nOut) : 0.0f;
		// the first measurement replaces the nominal density whatever it says (a jittered lattice has ~25 % fewer
		// candidates than 9 full cells), later ones only outside a +-30 % band
		if (avg > 0.0f && (s->capAvg == 0.0f || !s->capMeasured || avg > 1.3f * s->capAvg || avg < 0.7f * s->capAvg)) {
			set_sweep_class(s, avg);
			s->capMeasured = true;
		}
		s->lagPending = false;
	}
	c.dt = dt;
	c.k = pair_params(s, dt);
	c.nb = blocks_for(s->hostN);
	c.invDt = 1.0f / dt; // demo4.cpp:287
	c.force = make_float2(s->gravity.x + s->extForce.x, s->gravity.y + s->extForce.y); // gravity + externalForce, :306
	if (s->cfg.world_size > 1) {
		rc = maybe_resize_halo(s);
		if (rc != SPH_OK) return rc;
	}
	if (planRebalance && s->cfg.world_size > 1 && s->rebalanceEvery > 0 && s->steppedOnce && s->steps % (uint64_t)s->rebalanceEvery == 0) {
		rc = plan_rebalance(s);
		if (rc != SPH_OK) return rc;
	}
	// Steady state (nothing appended since the last step, no per-phase timing): the launches of a step are replayed
	// from a CUDA graph, which removes the launch gaps that dominate small scenes.  Kernel arguments depend on which
	// half of each double buffer is current, so graphs are cached per buffer parity, staging capacity, dt, force and
	// parameters.  (The exchange's parity and sequence number live on the device: they are not part of the key.)
	c.graphable = s->useGraphs && s->accFrom == 0xFFFFFFFFu && !(s->cfg.flags & SPH_FLAG_PHASE_TIMING) && s->steps >= 2 && !s->pendingRetarget;
	return SPH_OK;
}

int step_run(SphSim *s, const StepCtx &c, int parts) {
	if (!c.graphable) return enqueue_step(s, c.dt, c.k, c.nb, c.force, c.invDt, parts);
	StepGraphKey key;
	memset(&key, 0, sizeof(key));
	key.parity = buffer_parity(s);
	key.sweepCap = s->sweepCap;
	key.gridSweepCap = s->gridSweepCap; // the viscosity sweep runs on the previous grid's lists
	key.workHeavy = s->workHeavy;
	key.nb = c.nb;
	key.nbodies = (uint32_t)s->bodies.size();
	key.haloMsgRecords = s->haloMsgRecords;
	key.parts = (uint32_t)parts;
	key.flowEpoch = s->flowEpoch; // the sweep launches carry the pass number over the current grid as an argument
	key.gridGen = s->gridGen;
	key.force = c.force;
	key.k = c.k;
	StepGraph *g = nullptr;
	for (StepGraph &cand : s->graphs)
		if (memcmp(&cand.key, &key, sizeof(key)) == 0) g = &cand;
	if (!g) {
		if (s->graphs.size() >= 16) { // parameters keep changing: drop the oldest
			cudaGraphExecDestroy(s->graphs.front().exec);
			s->graphs.erase(s->graphs.begin());
		}
		cudaGraph_t graph = nullptr;
		const uint32_t epochBefore = s->flowEpoch;
		const uint64_t exchangesBefore = s->exchanges;
		const int32_t authLo = s->strip.authLo, authHi = s->strip.authHi;
		CU(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
		const int rce = enqueue_step(s, c.dt, c.k, c.nb, c.force, c.invDt, parts);
		const cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
		StepGraph made;
		made.key = key;
		made.parityAfter = buffer_parity(s);
		made.flowEpochAfter = s->flowEpoch;
		made.exchanges = (uint32_t)(s->exchanges - exchangesBefore);
		// capture only recorded the launches: the host-side state is still "before"
		set_buffer_parity(s, key.parity);
		s->flowEpoch = epochBefore;
		s->exchanges = exchangesBefore;
		cudaError_t ci = cudaSuccess;
		if (rce == SPH_OK && ce == cudaSuccess) ci = cudaGraphInstantiate(&made.exec, graph, 0);
		if (graph) cudaGraphDestroy(graph);
		if (rce != SPH_OK || ce != cudaSuccess || ci != cudaSuccess) {
			s->strip.authLo = authLo;
			s->strip.authHi = authHi;
			cudaGetLastError();
			if (rce != SPH_OK) return rce;
			return fail(s, SPH_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ci));
		}
		s->graphs.push_back(made);
		g = &s->graphs.back();
	}
	CU(s, cudaGraphLaunch(g->exec, s->stream));
	if (parts & GRID_BACK) s->gridSweepCap = s->sweepCap; // what the captured grid build classified its lists for
	set_buffer_parity(s, g->parityAfter);
	s->flowEpoch = g->flowEpochAfter;
	s->exchanges += g->exchanges;
	if (parts & GRID_FRONT) s->accFrom = 0xFFFFFFFFu;
	return SPH_OK;
}

int step_finish(SphSim *s) {
	CU(s, cudaGetLastError());
	s->steps++;
	s->steppedOnce = true;
	if (!s->lagPending) {
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
} // namespace

int sph_step(SphHandle s, float dt) {
	ENTER(s);
	if (s->transport == TR_LOCAL) return fail(s, SPH_ERR_STATE, "strips of one process are stepped together: use sph_step_group");
	StepCtx c;
	int rc = step_prepare(s, dt, c, true);
	if (rc != SPH_OK) return rc;
	if (s->cfg.world_size == 1 || s->transport == TR_PEER) {
		// one GPU, or strips that store their halo records straight into the neighbours' memory: the whole step is one graph
		rc = step_run(s, c, GRID_ALL);
		if (rc != SPH_OK) return rc;
	} else {
		// NCCL transport: send/recv inside a captured stream dead-locked on this stack (NCCL 2.28.9, driver 580), so the
		// launches before and after the exchange are two graphs and the exchange itself is enqueued plainly between them
		rc = step_run(s, c, GRID_FRONT);
		if (rc != SPH_OK) return rc;
		rc = exchange_nccl(s);
		if (rc != SPH_OK) return rc;
		rc = step_run(s, c, GRID_BACK);
		if (rc != SPH_OK) return rc;
	}
	return step_finish(s);
}

// Strips that live in one process (sph_comm_init_local): every strip's launches up to the publication of its halo
// records are enqueued before any strip's wait for them, so a waiting kernel never depends on work the host has
// not enqueued yet - also when all strips share one GPU.
int sph_step_group(SphHandle *handles, int32_t n, float dt) {
	if (!handles || n < 1) return fail(nullptr, SPH_ERR_INVALID, "sph_step_group needs handles");
	for (int r = 0; r < n; ++r) {
		CHECK_HANDLE(handles[r]);
		if (handles[r]->transport != TR_LOCAL || (int)handles[r]->group.size() != n || handles[r]->group[(size_t)r] != handles[r])
			return fail(handles[r], SPH_ERR_STATE, "sph_step_group takes exactly the handles of one sph_comm_init_local call, in rank order");
	}
	SphSim *first = handles[0];
	if (first->rebalanceEvery > 0 && first->steppedOnce && first->steps % (uint64_t)first->rebalanceEvery == 0) {
		int rc = plan_rebalance_group(handles, n);
		if (rc != SPH_OK) return rc;
	}
	// emitters: ONE rand() stream per process, so the first strip's emitters run (UpdateEmitter, demo4.cpp:257-284) and every
	// strip is handed the list
	if (!first->emitters.empty()) {
		std::vector<float> pos, acc;
		const auto t0 = std::chrono::steady_clock::now();
		emit_particles(first, dt, pos, acc);
		for (int r = 0; r < n && !pos.empty(); ++r) {
			int rc = sph_add_particles(handles[r], pos.size() / 2, pos.data(), acc.data(), nullptr);
			if (rc != SPH_OK) return rc;
		}
		first->hostEmitterMs += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
		first->hostEmitterSteps++;
	}
	std::vector<StepCtx> ctx((size_t)n);
	for (int r = 0; r < n; ++r) {
		DeviceScope dev(handles[r]);
		int rc = step_prepare(handles[r], dt, ctx[(size_t)r], false, false);
		if (rc == SPH_OK) rc = step_run(handles[r], ctx[(size_t)r], GRID_FRONT);
		if (rc != SPH_OK) return rc;
	}
	for (int r = 0; r < n; ++r) {
		DeviceScope dev(handles[r]);
		int rc = step_run(handles[r], ctx[(size_t)r], GRID_BACK);
		if (rc == SPH_OK) rc = step_finish(handles[r]);
		if (rc != SPH_OK) return rc;
	}
	return SPH_OK;
}

int sph_sync(SphHandle s) {
	ENTER(s);
	CU(s, cudaStreamSynchronize(s->stream));
	CU(s, cudaStreamSynchronize(s->copyStream));
	s->copyPending = false;
	return SPH_OK;
}

int sph_run_pass(SphHandle s, int pass, float dt) {
	ENTER(s);
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
			if (s->transport == TR_LOCAL) return fail(s, SPH_ERR_STATE, "strips of one process exchange inside sph_step_group only");
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
	ENTER(s);
	reset_stats_kernel<<<1, 1, 0, s->stream>>>(s->dCtr);
	memset(s->phaseMs, 0, sizeof(s->phaseMs));
	s->phaseSteps = 0;
	s->hostEmitterMs = 0.0f;
	s->hostEmitterSteps = 0;
	s->steppedOnce = false;
	CU(s, cudaGetLastError());
	return SPH_OK;
}

int sph_get_stats(SphHandle s, SphStats *out) {
	ENTER(s);
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
	if (s->hostEmitterSteps) out->time_emitters = s->hostEmitterMs / (float)s->hostEmitterSteps; // host time of UpdateEmitter, ms per step (sph.h:132)
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
	if (c.overflow & 8u) return fail(s, SPH_ERR_COMM, "a neighbour strip did not publish its halo records within %llu s", kHaloWaitTimeoutNs / 1000000000ull);
	if (c.overflow)
		return fail(s, SPH_ERR_CAPACITY, "device reported a capacity overflow (flags %u: 1 = particles, 2 = halo buffer, 4 = more candidates in one 3x3 block than the sweep queue holds)",
		            c.overflow);
	return SPH_OK;
}

int sph_get_phase_ms(SphHandle s, float out[SPH_NUM_PHASES], uint64_t *steps) {
	ENTER(s);
	for (int p = 0; p < PH_COUNT; ++p) out[p] = s->phaseSteps ? (float)(s->phaseMs[p] / (double)s->phaseSteps) : 0.0f;
	if (steps) *steps = s->phaseSteps;
	return SPH_OK;
}

// ---- readback / injection -------------------------------------------------------------------------
int sph_read_particles(SphHandle s, void *dst, size_t stride) {
	ENTER(s);
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
	ENTER(s);
	if (!src || stride < sizeof(ParticleRecord)) return fail(s, SPH_ERR_INVALID, "stride must be >= 48");
	if (s->cfg.world_size > 1) {
		// strips: every rank is given the whole state (nextId rows, creation order) and keeps its rows; the ghost copies
		// arrive with the exchange of the re-filing pass below
		const uint64_t total = s->nextId;
		if (total == 0) return SPH_OK;
		set_counts_kernel<<<1, 1, 0, s->stream>>>(s->dCtr, 0, 0);
		s->nextId = 0;
		int rcw = append_particles_owned(s, (size_t)total, nullptr, nullptr, src, stride, nullptr);
		if (rcw != SPH_OK) return rcw;
		s->accFrom = 0;
		begin_step_kernel<<<1, 1, 0, s->stream>>>(s->dCtr);
		rcw = launch_grid_build(s, 1.0f, false, true, false, GRID_ALL, false); // local re-filing, no exchange
		if (rcw != SPH_OK) return rcw;
		commit_kernel<<<1, 1, 0, s->stream>>>(s->dCtr);
		CU(s, cudaGetLastError());
		CU(s, cudaStreamSynchronize(s->stream));
		return SPH_OK;
	}
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
	ENTER(s);
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
	ENTER(s);
	if (s->copyPending) CU(s, cudaEventSynchronize(s->copyDone));
	return SPH_OK;
}

int sph_read_cell_start(SphHandle s, uint32_t *out) {
	ENTER(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	CU(s, cudaMemcpyAsync(out, s->cellStart, ((size_t)s->grid.nCells + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	return SPH_OK;
}

int sph_read_cell_counts(SphHandle s, uint32_t *out) {
	ENTER(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	std::vector<uint32_t> start((size_t)s->grid.nCells + 1);
	int rc = sph_read_cell_start(s, start.data());
	if (rc != SPH_OK) return rc;
	for (size_t c = 0; c < s->grid.nCells; ++c) out[c] = start[c + 1] - start[c];
	return SPH_OK;
}

int sph_read_sorted_ids(SphHandle s, uint32_t *out) {
	ENTER(s);
	if (!out) return fail(s, SPH_ERR_INVALID, "null out");
	if (s->hostN == 0) return SPH_OK;
	CU(s, cudaMemcpyAsync(out, s->id.in(), (size_t)s->hostN * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	return SPH_OK;
}

int sph_read_cell_of_particle(SphHandle s, int32_t *out) {
	ENTER(s);
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
	ENTER(s);
	if (out) *out = (void *)s->stream;
	return SPH_OK;
}
int sph_mark(SphHandle s, int slot) {
	ENTER(s);
	if (slot < 0 || slot >= 8) return fail(s, SPH_ERR_INVALID, "mark slot 0..7");
	CU(s, cudaEventRecord(s->marks[slot], s->stream));
	return SPH_OK;
}
int sph_elapsed_ms(SphHandle s, int a, int b, float *ms) {
	ENTER(s);
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
// out-boxes of rank r: the lower neighbour's "from above" boxes and the upper neighbour's "from below" boxes
static void wire_outboxes(SphSim *s, unsigned char *lowerMail, unsigned char *upperMail) {
	for (int par = 0; par < 2; ++par) {
		s->strip.outDown[par] = lowerMail ? reinterpret_cast<HaloBuffer *>(lowerMail + (size_t)(1 * 2 + par) * s->haloBytes) : nullptr;
		s->strip.outUp[par] = upperMail ? reinterpret_cast<HaloBuffer *>(upperMail + (size_t)(0 * 2 + par) * s->haloBytes) : nullptr;
	}
}

int sph_comm_init(SphHandle s, const uint8_t id128[128]) {
	ENTER(s);
	if (!id128) return fail(s, SPH_ERR_INVALID, "null id");
	if (s->cfg.world_size < 2) return fail(s, SPH_ERR_STATE, "sph_comm_init needs world_size > 1");
	if (s->comm || s->transport != TR_NONE) return fail(s, SPH_ERR_STATE, "communicator already initialised");
	if (!g_nccl.load()) return fail(s, SPH_ERR_COMM, "%s", g_nccl.error.c_str());
	NcclApi &nc = g_nccl;
	NcclUniqueId id;
	memcpy(id.internal, id128, 128);
	CU(s, cudaSetDevice(s->cfg.device));
	const int world = s->cfg.world_size, rank = s->cfg.rank;
	int rc = nc.CommInitRank(&s->comm, world, id, rank);
	if (rc != 0) {
		s->comm = nullptr;
		return fail(s, SPH_ERR_COMM, "ncclCommInitRank: %s", nc.GetErrorString(rc));
	}
	// Everything that would otherwise connect lazily inside the first steps happens here: an all-gather (it also
	// carries the IPC handles of the mailboxes), an all-reduce, and one send/recv with each neighbour.
	struct Hello {
		cudaIpcMemHandle_t mem;
		uint64_t haloBytes;
		int32_t ipcOk, pad;
	};
	Hello mine = {};
	mine.haloBytes = s->haloBytes;
	const bool wantPeer = !(s->cfg.flags & SPH_FLAG_EXCHANGE_NCCL);
	mine.ipcOk = (wantPeer && cudaIpcGetMemHandle(&mine.mem, s->mail) == cudaSuccess) ? 1 : 0;
	cudaGetLastError();
	Hello *dHello = nullptr;
	std::vector<Hello> all((size_t)world);
	CU(s, cudaMalloc(&dHello, (size_t)(world + 1) * sizeof(Hello)));
	CU(s, cudaMemcpyAsync(dHello + world, &mine, sizeof(Hello), cudaMemcpyHostToDevice, s->stream));
	rc = nc.AllGather(dHello + world, dHello, sizeof(Hello), kNcclUint8, s->comm, s->stream);
	if (rc != 0) return fail(s, SPH_ERR_COMM, "NCCL all-gather failed: %s", nc.GetErrorString(rc));
	CU(s, cudaMemcpyAsync(all.data(), dHello, (size_t)world * sizeof(Hello), cudaMemcpyDeviceToHost, s->stream));
	CU(s, cudaStreamSynchronize(s->stream));
	for (int r = 0; r < world; ++r)
		if (all[(size_t)r].haloBytes != s->haloBytes) {
			cudaFree(dHello);
			return fail(s, SPH_ERR_INVALID, "rank %d was created with a different halo_capacity (%llu vs %llu bytes per mailbox)", r,
			            (unsigned long long)all[(size_t)r].haloBytes, (unsigned long long)s->haloBytes);
		}
	// map the neighbours' mailboxes
	uint32_t ok = wantPeer ? 1u : 0u;
	for (int d = 0; d < 2 && ok; ++d) {
		const int nbr = d == 0 ? rank - 1 : rank + 1;
		if (nbr < 0 || nbr >= world) continue;
		if (!all[(size_t)nbr].ipcOk || cudaIpcOpenMemHandle(&s->peerBase[d], all[(size_t)nbr].mem, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
			s->peerBase[d] = nullptr;
			ok = 0u;
		}
	}
	cudaGetLastError();
	// all ranks must agree on the transport: one refusal (no peer access between two GPUs, IPC not permitted) sends everybody to NCCL
	uint32_t *dOk = reinterpret_cast<uint32_t *>(dHello);
	CU(s, cudaMemcpyAsync(dOk, &ok, sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
	rc = nc.AllReduce(dOk, dOk, 1, kNcclUint32, kNcclMin, s->comm, s->stream);
	if (rc != 0) return fail(s, SPH_ERR_COMM, "NCCL all-reduce failed: %s", nc.GetErrorString(rc));
	CU(s, cudaMemcpyAsync(&ok, dOk, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
	// neighbour connections (used by the NCCL transport every step, by the peer transport never again)
	rc = nc.GroupStart();
	for (int d = 0; d < 2 && rc == 0; ++d) {
		const int nbr = d == 0 ? rank - 1 : rank + 1;
		if (nbr < 0 || nbr >= world) continue;
		rc = nc.Send(dHello + world, sizeof(Hello), kNcclUint8, nbr, s->comm, s->stream);
		if (rc == 0) rc = nc.Recv(dHello + (size_t)nbr, sizeof(Hello), kNcclUint8, nbr, s->comm, s->stream);
	}
	const int rcEnd = nc.GroupEnd();
	if (rc == 0) rc = rcEnd;
	if (rc != 0) return fail(s, SPH_ERR_COMM, "NCCL neighbour hand-shake failed: %s", nc.GetErrorString(rc));
	CU(s, cudaStreamSynchronize(s->stream));
	cudaFree(dHello);
	if (ok) {
		s->transport = TR_PEER;
		wire_outboxes(s, static_cast<unsigned char *>(s->peerBase[0]), static_cast<unsigned char *>(s->peerBase[1]));
	} else {
		for (int d = 0; d < 2; ++d)
			if (s->peerBase[d]) {
				cudaIpcCloseMemHandle(s->peerBase[d]);
				s->peerBase[d] = nullptr;
			}
		CU(s, cudaMalloc(&s->sendMem, 4 * s->haloBytes));
		CU(s, cudaMemset(s->sendMem, 0, 4 * s->haloBytes));
		s->transport = TR_NCCL;
		// local send buffers laid out like a mailbox block, so the same wiring applies
		wire_outboxes(s, rank > 0 ? s->sendMem : nullptr, rank + 1 < world ? s->sendMem : nullptr);
	}
	return SPH_OK;
}

// All strips of a simulation inside ONE process (one host thread driving several GPUs, or several strips on one GPU
// for tests): the mailboxes are wired with plain pointers (peer access enabled between distinct devices) and the
// strips are stepped together with sph_step_group.
int sph_comm_init_local(SphHandle *handles, int32_t n) {
	if (!handles || n < 2) return fail(nullptr, SPH_ERR_INVALID, "sph_comm_init_local needs >= 2 handles");
	for (int r = 0; r < n; ++r) {
		SphSim *s = handles[r];
		ENTER(s);
		if (s->cfg.world_size != n || s->cfg.rank != r) return fail(s, SPH_ERR_INVALID, "handle %d of the group has rank %d of %d", r, s->cfg.rank, s->cfg.world_size);
		if (s->transport != TR_NONE) return fail(s, SPH_ERR_STATE, "communicator already initialised");
		if (s->haloBytes != handles[0]->haloBytes) return fail(s, SPH_ERR_INVALID, "the strips were created with different halo capacities");
	}
	for (int r = 0; r + 1 < n; ++r) {
		const int a = handles[r]->cfg.device, b = handles[r + 1]->cfg.device;
		if (a == b) continue;
		for (int dir = 0; dir < 2; ++dir) {
			const int from = dir ? b : a, to = dir ? a : b;
			int can = 0;
			cudaDeviceCanAccessPeer(&can, from, to);
			if (!can) return fail(handles[r], SPH_ERR_COMM, "device %d cannot access device %d's memory", from, to);
			cudaSetDevice(from);
			const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(handles[r], SPH_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", from, to, cudaGetErrorString(e));
			cudaGetLastError();
		}
	}
	for (int r = 0; r < n; ++r) {
		SphSim *s = handles[r];
		s->transport = TR_LOCAL;
		s->group.assign(handles, handles + n);
		wire_outboxes(s, r > 0 ? handles[r - 1]->mail : nullptr, r + 1 < n ? handles[r + 1]->mail : nullptr);
	}
	return SPH_OK;
}
int sph_set_strip(SphHandle s, int32_t rowBegin, int32_t rowEnd) {
	ENTER(s);
	if (s->nextId != 0) return fail(s, SPH_ERR_STATE, "set the strip before adding particles");
	CU(s, cudaStreamSynchronize(s->stream));
	return configure_strip(s, rowBegin, rowEnd);
}

// the particles this rank owns, compacted (arbitrary order) with their creation ids; any output may be NULL
int sph_read_owned(SphHandle s, uint32_t *ids, void *records, size_t recStride, void *positions, size_t posStride, void *colors, size_t colStride,
                   uint64_t *count) {
	ENTER(s);
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
// owned particles only exists on the device; the copies ship 1.02 x the previous frame's count (+4096) and
// sph_wait_render_owned fetches the rest in the rare frame that outgrew it.  The first frame is read synchronously.
int sph_render_owned(SphHandle s, uint32_t *ids, void *positions, size_t posStride, void *colors, size_t colStride) {
	ENTER(s);
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
	const uint64_t ship = std::min<uint64_t>(cap, s->ownedLast + s->ownedLast / 50 + 4096); // a strip's population changes by far less than 2 % per frame
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
	ENTER(s);
	if (!s->ownedPending) return fail(s, SPH_ERR_STATE, "no sph_render_owned frame in flight");
	if (s->copyPending) CU(s, cudaEventSynchronize(s->copyDone));
	const uint64_t n = std::min<uint64_t>(*s->hOwnedCount, s->capacity);
	if (n > s->ownedShipped) { // the strip grew by more than 2 % in one frame: fetch the tail (the snapshot is still intact)
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
	ENTER(s);
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
	ENTER(s);
	if (rowBegin) *rowBegin = s->grid.ownLo;
	if (rowEnd) *rowEnd = s->grid.ownHi;
	return SPH_OK;
}

} // extern "C"

// TEST INFRASTRUCTURE ONLY (oracle/): headless C-ABI wrapper around the UNMODIFIED reference
// solver.  It #includes /root/reference/NBodySimulation/demo4.cpp where it lies (nothing is
// copied into this repo) and exposes Demo4::ParticleSimulation through plain C functions so
// that tests/ and tools/make_golden.py can drive it with ctypes.  Output goes to
// oracle/_ref/libsphref.so (git-ignored).  Only tests/, smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it; the product path never does.
//
// Build recipe: oracle/Makefile (target `ref`).  Needs /root/reference, so it is built in the
// authoring container only; the GPU box uses the prebuilt .so that travels with the snapshot.
//
// What is reproduced here rather than included: DemoApplication::LoadScenario
// (app.cpp:477-534), because app.cpp drags in the window/GL layer.  ref_load_scenario() makes
// the same BaseSimulation calls in the same order, including the `bodyCount` loop bound of
// app.cpp:522.

#define FPL_IMPLEMENTATION
#define FPL_NO_AUDIO
#define FPL_NO_VIDEO
#define FPL_NO_WINDOW
#define FPL_NO_ENTRYPOINT
#include <final_platform_layer.h>

#include <string>
#include <string.h>
#include <float.h>
#include <stdlib.h>
#include <stdint.h>
#include <chrono>

// MSVC-isms of sph.h:75 and sph.h:282
#define strcpy_s(dst, n, src) strncpy((dst), (src), (n))
#undef fplStaticAssert
#define fplStaticAssert(expr)

#include "demo4.cpp"

#define STB_TRUETYPE_IMPLEMENTATION
#include <stb_truetype.h>

using Demo4::ParticleSimulation;

namespace {
bool g_platformUp = false;

void EnsurePlatform() {
	if (!g_platformUp) {
		fplPlatformInit(fplInitFlags_None, nullptr);
		g_platformUp = true;
	}
}
}  // namespace

extern "C" {

void *ref_create(void) {
	EnsurePlatform();
	return new ParticleSimulation();
}

void ref_destroy(void *h) { delete static_cast<ParticleSimulation *>(h); }

int ref_scenario_count(void) { return (int)(sizeof(SPHScenarios) / sizeof(SPHScenarios[0])); }

const char *ref_scenario_name(int idx) { return SPHScenarios[idx].name; }

// Same call sequence as app.cpp:477-534.  `seed` < 0 keeps libc's current rand() state
// (glibc default seed 1 on a fresh process); otherwise srand(seed) first.
void ref_load_scenario(void *h, int idx, int seed) {
	ParticleSimulation *demo = static_cast<ParticleSimulation *>(h);
	if (seed >= 0) srand((unsigned)seed);
	SPHScenario *scenario = &SPHScenarios[idx];
	demo->ResetStats();
	demo->ClearBodies();
	demo->ClearParticles();
	demo->ClearEmitters();
	demo->SetGravity(scenario->gravity);
	demo->SetParams(scenario->parameters);
	for (size_t b = 0; b < scenario->bodyCount; ++b) {
		SPHScenarioBody *body = &scenario->bodies[b];
		switch (body->type) {
			case SPHScenarioBodyType_Plane: {
				float distance = Vec2Dot(body->orientation.col1, body->position);
				demo->AddPlane(body->orientation.col1, distance);
			} break;
			case SPHScenarioBodyType_Circle:
				demo->AddCircle(body->position, body->radius);
				break;
			case SPHScenarioBodyType_LineSegment: {
				Vec2f a = Vec2MultMat2(body->orientation, body->localVerts[0]) + body->position;
				Vec2f c = Vec2MultMat2(body->orientation, body->localVerts[1]) + body->position;
				demo->AddLineSegment(a, c);
			} break;
			case SPHScenarioBodyType_Polygon: {
				Vec2f verts[kMaxScenarioPolygonCount];
				for (size_t v = 0; v < body->vertexCount; ++v)
					verts[v] = Vec2MultMat2(body->orientation, body->localVerts[v]) + body->position;
				demo->AddPolygon(body->vertexCount, verts);
			} break;
			default: break;
		}
	}
	const float spacing = demo->GetParams().particleSpacing;
	// app.cpp:522 iterates to bodyCount (not volumeCount); volumes[] is only 8 long, and the
	// unused entries are zero-sized, so clamp to the array and keep the effect identical.
	size_t volumeLoop = scenario->bodyCount < kSPHMaxScenarioVolumeCount ? scenario->bodyCount : kSPHMaxScenarioVolumeCount;
	for (size_t v = 0; v < volumeLoop; ++v) {
		SPHScenarioVolume *volume = &scenario->volumes[v];
		int numX = (int)floor((volume->size.w / spacing));
		int numY = (int)floor((volume->size.h / spacing));
		demo->AddVolume(volume->position, volume->force, numX, numY, spacing);
	}
	for (size_t e = 0; e < scenario->emitterCount; ++e) {
		SPHScenarioEmitter *em = &scenario->emitters[e];
		demo->AddEmitter(em->position, em->direction, em->radius, em->speed, em->rate, em->duration);
	}
}

// ---- raw BaseSimulation surface (base.h:8-39) ----------------------------------------
void ref_reset_stats(void *h) { static_cast<ParticleSimulation *>(h)->ResetStats(); }
void ref_clear_bodies(void *h) { static_cast<ParticleSimulation *>(h)->ClearBodies(); }
void ref_clear_particles(void *h) { static_cast<ParticleSimulation *>(h)->ClearParticles(); }
void ref_clear_emitters(void *h) { static_cast<ParticleSimulation *>(h)->ClearEmitters(); }
void ref_add_plane(void *h, float nx, float ny, float d) { static_cast<ParticleSimulation *>(h)->AddPlane(Vec2f(nx, ny), d); }
void ref_add_circle(void *h, float x, float y, float r) { static_cast<ParticleSimulation *>(h)->AddCircle(Vec2f(x, y), r); }
void ref_add_segment(void *h, float ax, float ay, float bx, float by) { static_cast<ParticleSimulation *>(h)->AddLineSegment(Vec2f(ax, ay), Vec2f(bx, by)); }
void ref_add_polygon(void *h, int n, const float *xy) {
	Vec2f verts[kMaxScenarioPolygonCount];
	for (int i = 0; i < n; ++i) verts[i] = Vec2f(xy[2 * i], xy[2 * i + 1]);
	static_cast<ParticleSimulation *>(h)->AddPolygon((size_t)n, verts);
}
uint64_t ref_add_particle(void *h, float x, float y, float fx, float fy) {
	return static_cast<ParticleSimulation *>(h)->AddParticle(Vec2f(x, y), Vec2f(fx, fy));
}
void ref_add_volume(void *h, float cx, float cy, float fx, float fy, int nx, int ny, float spacing) {
	static_cast<ParticleSimulation *>(h)->AddVolume(Vec2f(cx, cy), Vec2f(fx, fy), nx, ny, spacing);
}
void ref_add_emitter(void *h, float px, float py, float dx, float dy, float radius, float speed, float rate, float duration) {
	static_cast<ParticleSimulation *>(h)->AddEmitter(Vec2f(px, py), Vec2f(dx, dy), radius, speed, rate, duration);
}
void ref_set_gravity(void *h, float gx, float gy) { static_cast<ParticleSimulation *>(h)->SetGravity(Vec2f(gx, gy)); }
void ref_add_external_force(void *h, float fx, float fy) { static_cast<ParticleSimulation *>(h)->AddExternalForces(Vec2f(fx, fy)); }
void ref_clear_external_force(void *h) { static_cast<ParticleSimulation *>(h)->ClearExternalForce(); }
void ref_set_multithreading(void *h, int on) { static_cast<ParticleSimulation *>(h)->SetMultiThreading(on != 0); }
int ref_worker_threads(void *h) { return (int)static_cast<ParticleSimulation *>(h)->GetWorkerThreadCount(); }
uint64_t ref_particle_count(void *h) { return static_cast<ParticleSimulation *>(h)->GetParticleCount(); }

// params as the 9 floats of sph.h:77-87, in declaration order
void ref_get_params(void *h, float *out9) {
	const SPHParameters &p = static_cast<ParticleSimulation *>(h)->GetParams();
	memcpy(out9, &p, 9 * sizeof(float));
}
void ref_set_params(void *h, const float *in9) {
	SPHParameters p;
	memcpy(&p, in9, 9 * sizeof(float));
	static_cast<ParticleSimulation *>(h)->SetParams(p);   // copy-ctor recomputes invKernelHeight (sph.h:100-110)
}

void ref_update(void *h, float dt) { static_cast<ParticleSimulation *>(h)->Update(dt); }

// Steps `steps` times and returns the wall seconds spent inside Update().
double ref_update_timed(void *h, float dt, int steps) {
	ParticleSimulation *s = static_cast<ParticleSimulation *>(h);
	auto t0 = std::chrono::steady_clock::now();
	for (int i = 0; i < steps; ++i) s->Update(dt);
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

// ---- individual passes (public members of Demo4::ParticleSimulation, demo4.h:180-184) --
void ref_neighbor_search(void *h, float dt) { ParticleSimulation *s = static_cast<ParticleSimulation *>(h); s->NeighborSearch(0, (int64_t)s->particleCount - 1, dt); }
void ref_density_pressure(void *h, float dt) { ParticleSimulation *s = static_cast<ParticleSimulation *>(h); s->DensityAndPressure(0, (int64_t)s->particleCount - 1, dt); }
void ref_viscosity(void *h, float dt) { ParticleSimulation *s = static_cast<ParticleSimulation *>(h); s->ViscosityForces(0, (int64_t)s->particleCount - 1, dt); }
void ref_delta_positions(void *h, float dt) { ParticleSimulation *s = static_cast<ParticleSimulation *>(h); s->DeltaPositions(0, (int64_t)s->particleCount - 1, dt); }

// ---- state access ----------------------------------------------------------------------
// 12 floats per particle in ParticleData order (demo4.h:81-99): cur, prev, acc, vel, rho, rhoNear, P, PNear
void ref_get_particles(void *h, float *out12) {
	ParticleSimulation *s = static_cast<ParticleSimulation *>(h);
	static_assert(sizeof(Demo4::ParticleData) == 48, "ParticleData layout");
	memcpy(out12, s->particleDatas, s->particleCount * sizeof(Demo4::ParticleData));
}
// Overwrites the dynamic state and re-files every particle in the grid through the
// reference's own Remove/Insert (demo4.cpp:37-76), like its "Update grid" loop does.
void ref_set_particles(void *h, const float *in12) {
	ParticleSimulation *s = static_cast<ParticleSimulation *>(h);
	memcpy(s->particleDatas, in12, s->particleCount * sizeof(Demo4::ParticleData));
	for (size_t i = 0; i < s->particleCount; ++i) {
		Vec2i c = SPHComputeCellIndex(s->particleDatas[i].curPosition);
		Vec2i *old = &s->particleIndexes[i].cellIndex;
		if (c.x != old->x || c.y != old->y) {
			s->RemoveParticleFromGrid(i);
			s->InsertParticleIntoGrid(i);
		}
	}
}
void ref_get_cell_of_particle(void *h, int32_t *out_xy) {
	ParticleSimulation *s = static_cast<ParticleSimulation *>(h);
	for (size_t i = 0; i < s->particleCount; ++i) {
		out_xy[2 * i] = s->particleIndexes[i].cellIndex.x;
		out_xy[2 * i + 1] = s->particleIndexes[i].cellIndex.y;
	}
}
void ref_grid_dims(int32_t *out2) { out2[0] = kSPHGridCountX; out2[1] = kSPHGridCountY; }
void ref_get_cell_counts(void *h, uint32_t *out) {
	ParticleSimulation *s = static_cast<ParticleSimulation *>(h);
	for (int c = 0; c < kSPHGridTotalCount; ++c) out[c] = (uint32_t)s->cells[c].count;
}
// Members of one cell in the reference's own storage order; returns the count.
uint32_t ref_get_cell_members(void *h, int cell, uint32_t *out) {
	ParticleSimulation *s = static_cast<ParticleSimulation *>(h);
	Demo4::Cell *c = &s->cells[cell];
	for (size_t k = 0; k < c->count; ++k) out[k] = (uint32_t)c->indices[k];
	return (uint32_t)c->count;
}
uint32_t ref_get_neighbor_count(void *h, uint64_t i) { return (uint32_t)static_cast<ParticleSimulation *>(h)->particleIndexes[i].neighborCount; }
uint32_t ref_get_neighbors(void *h, uint64_t i, uint32_t *out) {
	ParticleSimulation *s = static_cast<ParticleSimulation *>(h);
	Demo4::ParticleIndex *pi = &s->particleIndexes[i];
	for (size_t k = 0; k < pi->neighborCount; ++k) out[k] = (uint32_t)pi->neighbors[k];
	return (uint32_t)pi->neighborCount;
}
// all neighbour counts at once
void ref_get_neighbor_counts(void *h, uint32_t *out) {
	ParticleSimulation *s = static_cast<ParticleSimulation *>(h);
	for (size_t i = 0; i < s->particleCount; ++i) out[i] = (uint32_t)s->particleIndexes[i].neighborCount;
}
// stats: 4 counters then the 9 phase times (ms) of sph.h:125-141
void ref_get_stats(void *h, uint64_t *counters4, float *times9) {
	SPHStatistics &st = static_cast<ParticleSimulation *>(h)->GetStats();
	counters4[0] = st.minParticleNeighborCount;
	counters4[1] = st.maxParticleNeighborCount;
	counters4[2] = st.minCellParticleCount;
	counters4[3] = st.maxCellParticleCount;
	memcpy(times9, &st.time, 9 * sizeof(float));
}
// Render()'s colour rule (sph.h:683-695) for every particle, rgba
void ref_get_colors(void *h, float *out4) {
	ParticleSimulation *s = static_cast<ParticleSimulation *>(h);
	for (size_t i = 0; i < s->particleCount; ++i) {
		Demo4::ParticleData &d = s->particleDatas[i];
		Vec4f c = SPHGetParticleColor(s->params.restDensity, d.density, d.pressure, d.velocity);
		out4[4 * i] = c.r; out4[4 * i + 1] = c.g; out4[4 * i + 2] = c.b; out4[4 * i + 3] = c.a;
	}
}
// bodies as the product's flat record: type, then 17 floats (see include/sphb200.h SphBody)
int ref_body_count(void *h) { return (int)static_cast<ParticleSimulation *>(h)->bodyCount; }
void ref_get_body(void *h, int idx, int32_t *type, int32_t *nverts, float *out16) {
	Demo4::Body *b = &static_cast<ParticleSimulation *>(h)->bodies[idx];
	*type = (int32_t)b->type;
	*nverts = 0;
	memset(out16, 0, 16 * sizeof(float));
	switch (b->type) {
		case Demo4::BodyType_Plane: out16[0] = b->plane.normal.x; out16[1] = b->plane.normal.y; out16[2] = b->plane.distance; break;
		case Demo4::BodyType_Circle: out16[0] = b->circle.pos.x; out16[1] = b->circle.pos.y; out16[2] = b->circle.radius; break;
		case Demo4::BodyType_LineSegment: out16[0] = b->lineSegment.a.x; out16[1] = b->lineSegment.a.y; out16[2] = b->lineSegment.b.x; out16[3] = b->lineSegment.b.y; break;
		case Demo4::BodyType_Polygon:
			*nverts = (int32_t)b->polygon.vertexCount;
			for (size_t v = 0; v < b->polygon.vertexCount; ++v) { out16[2 * v] = b->polygon.verts[v].x; out16[2 * v + 1] = b->polygon.verts[v].y; }
			break;
		default: break;
	}
}
void ref_get_gravity(void *h, float *out2) { ParticleSimulation *s = static_cast<ParticleSimulation *>(h); out2[0] = s->gravity.x; out2[1] = s->gravity.y; }

// ---- the collision solvers of sph.h:514-681, one point at a time ------------------------
void ref_solve_plane(float *pxy, float nx, float ny, float d) { Vec2f p(pxy[0], pxy[1]); SPHSolvePlaneCollision(&p, Vec2f(nx, ny), d); pxy[0] = p.x; pxy[1] = p.y; }
void ref_solve_circle(float *pxy, float cx, float cy, float r) { Vec2f p(pxy[0], pxy[1]); SPHSolveCircleCollision(&p, Vec2f(cx, cy), r); pxy[0] = p.x; pxy[1] = p.y; }
void ref_solve_segment(float *pxy, float ax, float ay, float bx, float by) { Vec2f p(pxy[0], pxy[1]); SPHSolveLineSegmentCollision(&p, Vec2f(ax, ay), Vec2f(bx, by)); pxy[0] = p.x; pxy[1] = p.y; }
void ref_solve_polygon(float *pxy, int n, const float *xy) {
	Vec2f verts[kMaxScenarioPolygonCount];
	for (int i = 0; i < n; ++i) verts[i] = Vec2f(xy[2 * i], xy[2 * i + 1]);
	Vec2f p(pxy[0], pxy[1]);
	SPHSolvePolygonCollision(&p, (size_t)n, verts);
	pxy[0] = p.x; pxy[1] = p.y;
}
void ref_cell_index(float x, float y, int32_t *out2) { Vec2i c = SPHComputeCellIndex(Vec2f(x, y)); out2[0] = c.x; out2[1] = c.y; }

}  // extern "C"

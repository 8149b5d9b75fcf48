// The adapter a maintainer of the reference adds as NBodySimulation/demo_b200.h: a fifth
// BaseSimulation (base.h:8-39) whose 24 virtuals forward to libsphb200's C ABI.  Needs the REFERENCE's headers
// (base.h, sph.h, vecmath.h, render.h) on the include path; tests/test_adapter.py compiles it against them where
// /root/reference exists (under the g++ shim the oracle uses for vecmath.h) and links it with libsphb200.so.
#ifndef DEMO_B200_H
#define DEMO_B200_H
#include <vector>
#include "base.h"       // BaseSimulation, SPHParameters, SPHStatistics, Vec2f, Render::CommandBuffer
#include "sphb200.h"    // this repo: include/sphb200.h; link with -lsphb200

namespace DemoB200 {
const char *kDemoName = "Demo B200";

// Demo4::ParticleData layout (demo4.h:81-99): Render() hands &datas[0] with stride 48 to the GL executor
struct ParticleData { Vec2f curPosition, prevPosition, acceleration, velocity; float density, nearDensity, pressure, nearPressure; };

struct ParticleSimulation : BaseSimulation {
	SphHandle h = nullptr;
	SPHParameters params;                 // sph.h:77-123, same 9 floats as SphParams
	SPHStatistics stats;
	std::vector<ParticleData> datas;      // host mirror, creation order (what demo4 keeps in particleDatas)
	std::vector<Vec4f> colors;
	std::vector<uint32_t> cellCounts;
	bool multiThreading = true;

	ParticleSimulation() { SphConfig c; sph_config_default(&c); sph_create(&c, &h); }
	// BaseSimulation has no virtual destructor and app.cpp:381-383 deletes through the base pointer,
	// so the app calls Destroy() before `delete demo` (one added line in LoadDemo / ~DemoApplication).
	void Destroy() { sph_destroy(h); h = nullptr; }

	void ResetStats() override { sph_reset_stats(h); }
	void ClearBodies() override { sph_clear_bodies(h); }
	void ClearParticles() override { sph_clear_particles(h); }
	void ClearEmitters() override { sph_clear_emitters(h); }
	void AddPlane(const Vec2f &n, const float d) override { sph_add_plane(h, n.x, n.y, d); }
	void AddCircle(const Vec2f &p, const float r) override { sph_add_circle(h, p.x, p.y, r); }
	void AddLineSegment(const Vec2f &a, const Vec2f &b) override { sph_add_segment(h, a.x, a.y, b.x, b.y); }
	void AddPolygon(const size_t n, const Vec2f *v) override { sph_add_polygon(h, n, &v[0].x); }
	size_t AddParticle(const Vec2f &p, const Vec2f &f) override { uint64_t i; sph_add_particles(h, 1, &p.x, &f.x, &i); return (size_t)i; }
	void AddVolume(const Vec2f &c, const Vec2f &f, const int nx, const int ny, const float s) override { sph_add_volume(h, c.x, c.y, f.x, f.y, nx, ny, s); }
	void AddEmitter(const Vec2f &p, const Vec2f &d, const float radius, const float speed, const float rate, const float duration) override {
		sph_add_emitter(h, p.x, p.y, d.x, d.y, radius, speed, rate, duration);
	}
	void Update(const float dt) override { sph_step(h, dt); }          // returns once the step is enqueued
	void AddExternalForces(const Vec2f &f) override { sph_add_external_force(h, f.x, f.y); }
	void ClearExternalForce() override { sph_clear_external_force(h); }
	size_t GetParticleCount() override { uint64_t n; sph_particle_count(h, &n); return (size_t)n; }
	void SetGravity(const Vec2f &g) override { sph_set_gravity(h, g.x, g.y); }
	const SPHParameters &GetParams() override { sph_get_params(h, (SphParams *)&params); return params; }
	void SetParams(const SPHParameters &p) override { sph_set_params(h, (const SphParams *)&p); }
	SPHStatistics &GetStats() override {
		SphStats s; sph_get_stats(h, &s);
		stats.minParticleNeighborCount = s.min_particle_neighbor_count; stats.maxParticleNeighborCount = s.max_particle_neighbor_count;
		stats.minCellParticleCount = s.min_cell_particle_count;         stats.maxCellParticleCount = s.max_cell_particle_count;
		stats.time.integration = s.time_integration;   stats.time.viscosityForces = s.time_viscosity_forces;
		stats.time.predict = s.time_predict;           stats.time.updateGrid = s.time_update_grid;
		stats.time.neighborSearch = 0;                 stats.time.densityAndPressure = s.time_density_and_pressure;
		stats.time.deltaPositions = s.time_delta_positions; stats.time.collisions = s.time_collisions;
		return stats;
	}
	void SetMultiThreading(const bool v) override { multiThreading = v; }
	bool IsMultiThreadingSupported() override { return true; }
	bool IsMultiThreading() override { return multiThreading; }
	size_t GetWorkerThreadCount() override { return 148; }

	// demo4.cpp:453-532 with the particle arrays coming back from the device
	void Render(Render::CommandBuffer *cb, const float worldToScreenScale) override {
		const size_t n = GetParticleCount();
		datas.resize(n); colors.resize(n); cellCounts.resize(kSPHGridTotalCount);
		sph_render_particles(h, &datas[0].curPosition, sizeof(ParticleData), &colors[0], sizeof(Vec4f));
		sph_read_cell_counts(h, cellCounts.data());
		Render::PushRectangle(cb, Vec2f(-kSPHBoundaryHalfWidth, -kSPHBoundaryHalfHeight), Vec2f(kSPHBoundaryHalfWidth, kSPHBoundaryHalfHeight) * 2.0f, Vec4f(1, 0, 1, 1), false, 1.0f);
		for (int y = 0; y < kSPHGridCountY; ++y) for (int x = 0; x < kSPHGridCountX; ++x)
			if (cellCounts[SPHComputeCellOffset(x, y)] > 0)
				Render::PushRectangle(cb, kSPHGridOrigin + Vec2f((float)x, (float)y) * kSPHGridCellSize, Vec2f(kSPHGridCellSize), ColorLightGray, true);
		// (grid lines, bodies and emitters are drawn from host-side copies exactly as demo4.cpp:471-518 does)
		sph_wait_render(h);   // the pointers below must hold the frame before OpenGLDrawCommandBuffer runs (main.cpp:543)
		Render::PushVertexIndexArrayHeader(cb, sizeof(ParticleData), &datas[0], 0, nullptr, sizeof(Vec4f), &colors[0], 0, nullptr);
		Render::PushVertexIndexArrayDraw(cb, Render::PrimitiveType::Points, (uint32_t)n, kSPHParticleRenderRadius * 2.0f * worldToScreenScale, nullptr, {}, false);
	}
};
}
#endif

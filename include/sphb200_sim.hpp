// sphb200_sim.hpp — header-only C++ host class over the C ABI (sphb200.h) with the member names,
// argument meaning and call order of the reference's plugin interface `class BaseSimulation`
// (NBodySimulation/base.h:8-39) as implemented by Demo4::ParticleSimulation (demo4.h:137-226).
//
// It deliberately does NOT include any reference header, so it builds on its own (tests/ and the
// examples use it as is).  Inside the reference tree a maintainer wraps it in a 20-line subclass of
// BaseSimulation that converts Vec2f <-> (x, y): see INTEGRATION.md.
#ifndef SPHB200_SIM_HPP
#define SPHB200_SIM_HPP

#include <stdexcept>
#include <string>
#include <vector>

#include "sphb200.h"

namespace sphb200 {

struct Vec2 { // layout-compatible with the reference's Vec2f (vecmath.h:30-54)
	float x = 0.0f, y = 0.0f;
	Vec2() = default;
	Vec2(float ax, float ay) : x(ax), y(ay) {}
};

// Demo4::ParticleData (demo4.h:81-99): what Render() hands to the GL executor, stride 48
struct ParticleData {
	Vec2 curPosition, prevPosition, acceleration, velocity;
	float density, nearDensity, pressure, nearPressure;
};
static_assert(sizeof(ParticleData) == 48, "ParticleData must match demo4.h:81-99");

struct Color4 {
	float r, g, b, a;
};

class ParticleSimulation {
public:
	explicit ParticleSimulation(const SphConfig *config = nullptr) {
		SphConfig cfg;
		if (config) cfg = *config;
		else sph_config_default(&cfg);
		if (sph_create(&cfg, &handle_) != SPH_OK) {
			char buf[512];
			sph_last_error(nullptr, buf, sizeof(buf));
			throw std::runtime_error(std::string("sph_create: ") + buf);
		}
		sph_get_params(handle_, &params_);
	}
	// BaseSimulation has no virtual destructor and the app deletes through the base pointer
	// (app.cpp:381-383), so an adapter must call Destroy() itself; here the destructor does it.
	~ParticleSimulation() { Destroy(); }
	ParticleSimulation(const ParticleSimulation &) = delete;
	ParticleSimulation &operator=(const ParticleSimulation &) = delete;
	void Destroy() {
		if (handle_) sph_destroy(handle_);
		handle_ = nullptr;
	}

	// ---- base.h:10-13 ----
	void ResetStats() { check(sph_reset_stats(handle_)); }
	void ClearBodies() { check(sph_clear_bodies(handle_)); }
	void ClearParticles() { check(sph_clear_particles(handle_)); }
	void ClearEmitters() { check(sph_clear_emitters(handle_)); }
	// ---- base.h:15-18 ----
	void AddPlane(const Vec2 &normal, const float distance) { check(sph_add_plane(handle_, normal.x, normal.y, distance)); }
	void AddCircle(const Vec2 &pos, const float radius) { check(sph_add_circle(handle_, pos.x, pos.y, radius)); }
	void AddLineSegment(const Vec2 &a, const Vec2 &b) { check(sph_add_segment(handle_, a.x, a.y, b.x, b.y)); }
	void AddPolygon(const size_t vertexCount, const Vec2 *verts) { check(sph_add_polygon(handle_, vertexCount, &verts[0].x)); }
	// ---- base.h:20-22 ----
	size_t AddParticle(const Vec2 &position, const Vec2 &force) {
		uint64_t first = 0;
		check(sph_add_particles(handle_, 1, &position.x, &force.x, &first));
		return (size_t)first;
	}
	void AddVolume(const Vec2 &center, const Vec2 &force, const int countX, const int countY, const float spacing) {
		check(sph_add_volume(handle_, center.x, center.y, force.x, force.y, countX, countY, spacing));
	}
	void AddEmitter(const Vec2 &position, const Vec2 &direction, const float radius, const float speed, const float rate, const float duration) {
		check(sph_add_emitter(handle_, position.x, position.y, direction.x, direction.y, radius, speed, rate, duration));
	}
	// ---- base.h:24-25 ----
	void Update(const float deltaTime) { check(sph_step(handle_, deltaTime)); }
	// Render(): fills the host mirror the GL executor reads (vertices = &particleDatas()[0], stride 48;
	// colors = &particleColors()[0], stride 16; demo4.cpp:525-531) and the per-cell occupancy the grid
	// fill uses (demo4.cpp:459-469).  Returns the particle count.
	size_t Render() {
		uint64_t n = 0;
		check(sph_particle_count(handle_, &n));
		datas_.resize(n);
		colors_.resize(n);
		if (n) {
			check(sph_render_particles(handle_, &datas_[0].curPosition, sizeof(ParticleData), colors_.data(), sizeof(Color4)));
			check(sph_wait_render(handle_));
		}
		int32_t gx = 0, gy = 0;
		sph_grid_dims(handle_, &gx, &gy);
		cellCounts_.resize((size_t)gx * gy);
		check(sph_read_cell_counts(handle_, cellCounts_.data()));
		return (size_t)n;
	}
	const std::vector<ParticleData> &particleDatas() const { return datas_; }
	const std::vector<Color4> &particleColors() const { return colors_; }
	const std::vector<uint32_t> &cellCounts() const { return cellCounts_; }
	// ---- base.h:27-28 ----
	void AddExternalForces(const Vec2 &force) { check(sph_add_external_force(handle_, force.x, force.y)); }
	void ClearExternalForce() { check(sph_clear_external_force(handle_)); }
	// ---- base.h:30-38 ----
	size_t GetParticleCount() {
		uint64_t n = 0;
		check(sph_particle_count(handle_, &n));
		return (size_t)n;
	}
	void SetGravity(const Vec2 &gravity) { check(sph_set_gravity(handle_, gravity.x, gravity.y)); }
	const SphParams &GetParams() {
		check(sph_get_params(handle_, &params_));
		return params_;
	}
	SphStats &GetStats() {
		check(sph_get_stats(handle_, &stats_));
		return stats_;
	}
	void SetParams(const SphParams &params) { check(sph_set_params(handle_, &params)); }
	void SetMultiThreading(const bool value) { multiThreading_ = value; } // one mode on the GPU; kept for the interface
	bool IsMultiThreadingSupported() { return true; }
	bool IsMultiThreading() { return multiThreading_; }
	size_t GetWorkerThreadCount() { return 148; } // SMs of a B200

	// DemoApplication::LoadScenario (app.cpp:477-534) for SPHScenarios[index] (sph.h:315-437)
	void LoadScenario(int index, int seed = -1) { check(sph_load_scenario(handle_, index, seed)); }
	SphHandle handle() const { return handle_; }

private:
	void check(int rc) {
		if (rc == SPH_OK) return;
		char buf[512];
		sph_last_error(handle_, buf, sizeof(buf));
		throw std::runtime_error(std::string("sphb200: ") + buf);
	}
	SphHandle handle_ = nullptr;
	SphParams params_{};
	SphStats stats_{};
	bool multiThreading_ = true;
	std::vector<ParticleData> datas_;
	std::vector<Color4> colors_;
	std::vector<uint32_t> cellCounts_;
};

} // namespace sphb200
#endif

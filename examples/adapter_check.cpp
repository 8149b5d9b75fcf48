// Compile-and-link check of examples/demo_b200.h against the reference's own headers (tests/test_adapter.py):
// instantiates the adapter through a BaseSimulation pointer and, when a GPU is present, replays LoadScenario's call
// sequence for the "Blob" scene (app.cpp:477-534, sph.h:363-374) and a few frames of the app loop (app.cpp:228-236).
// Built only where /root/reference exists; never shipped.
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
#include <stdio.h>

#define strcpy_s(dst, n, src) strncpy((dst), (src), (n))
#undef fplStaticAssert
#define fplStaticAssert(expr)

#include "base.h"
#include "demo_b200.h"

#define STB_TRUETYPE_IMPLEMENTATION
#include <stb_truetype.h>

static_assert(sizeof(DemoB200::ParticleData) == 48, "ParticleData stride (demo4.h:81-99)");
static_assert(sizeof(SPHParameters) == sizeof(SphParams), "SPHParameters and SphParams are the same nine floats");

int main(int argc, char **argv) {
	DemoB200::ParticleSimulation *sim = new DemoB200::ParticleSimulation();
	BaseSimulation *demo = sim; // what DemoApplication holds (app.h:125)
	if (!sim->h) { // no GPU here: the library refused (there is no CPU fallback); the adapter compiled and linked
		char msg[256];
		sph_last_error(nullptr, msg, sizeof(msg));
		printf("adapter linked; no device: %s\n", msg);
		delete sim;
		return argc > 1 ? 2 : 0;
	}
	// LoadScenario (app.cpp:477-534) for SPHScenarios[2] "Blob"
	const SPHScenario *scenario = &SPHScenarios[2];
	demo->ResetStats();
	demo->ClearBodies();
	demo->ClearParticles();
	demo->ClearEmitters();
	demo->SetGravity(scenario->gravity);
	demo->SetParams(scenario->parameters);
	for (size_t i = 0; i < scenario->bodyCount; ++i) {
		const SPHScenarioBody *body = &scenario->bodies[i];
		if (body->type == SPHScenarioBodyType::SPHScenarioBodyType_Plane) {
			Vec2f normal = body->orientation.col1;
			demo->AddPlane(normal, Vec2Dot(normal, body->position));
		}
	}
	const float spacing = demo->GetParams().particleSpacing;
	for (size_t i = 0; i < scenario->volumeCount; ++i) {
		const SPHScenarioVolume *v = &scenario->volumes[i];
		demo->AddVolume(v->position, v->force, (int)floor(v->size.w / spacing), (int)floor(v->size.h / spacing), spacing);
	}
	const size_t n = demo->GetParticleCount();
	for (int frame = 0; frame < 8; ++frame) demo->Update(1.0f / 60.0f);
	SPHStatistics &st = demo->GetStats();
	sim->datas.resize(n);
	sim->colors.resize(n);
	sph_render_particles(sim->h, &sim->datas[0].curPosition, sizeof(DemoB200::ParticleData), &sim->colors[0], sizeof(Vec4f));
	sph_wait_render(sim->h);
	float ymin = 1e9f, ymax = -1e9f;
	for (size_t i = 0; i < n; ++i) {
		ymin = fminf(ymin, sim->datas[i].curPosition.y);
		ymax = fmaxf(ymax, sim->datas[i].curPosition.y);
	}
	printf("adapter ran: %zu particles, 8 frames, neighbours %zu..%zu, y in [%f, %f], alpha %f\n", n, st.minParticleNeighborCount,
	       st.maxParticleNeighborCount, ymin, ymax, sim->colors[0].a);
	const bool ok = n == 1400 && st.maxParticleNeighborCount > 0 && ymin > -kSPHBoundaryHalfHeight && ymax < kSPHBoundaryHalfHeight && sim->colors[0].a == 1.0f;
	sim->Destroy();
	delete sim;
	return ok ? 0 : 1;
}

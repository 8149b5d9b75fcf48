// The call sequence of the reference's app loop (app.cpp:215-289, 380-410, 477-534) against the
// C++ host class: LoadDemo -> LoadScenario -> { Update(1/60); GetStats(); Render(); } per frame.
//   g++ -std=c++17 -I include examples/demo_host.cpp -L nbodysimulation_experiment_b200 -lsphb200 -o demo_host
#include <cstdio>
#include <cstdlib>
#include <exception>

#include "sphb200_sim.hpp"

int main(int argc, char **argv) {
	const int scenario = argc > 1 ? atoi(argv[1]) : 0, frames = argc > 2 ? atoi(argv[2]) : 64;
	try {
		sphb200::ParticleSimulation demo;          // LoadDemo: new ParticleSimulation (app.cpp:398-402)
		demo.SetMultiThreading(true);              // app.cpp:408
		demo.LoadScenario(scenario, 1);            // app.cpp:409
		size_t n = 0;
		for (int f = 0; f < frames; ++f) {
			demo.Update(1.0f / 60.0f);             // app.cpp:231-233
			demo.GetStats();                       // app.cpp:240
			n = demo.Render();                     // app.cpp:286-289
		}
		const SphStats &st = demo.GetStats();
		double cx = 0, cy = 0;
		for (const auto &p : demo.particleDatas()) { cx += p.curPosition.x; cy += p.curPosition.y; }
		size_t occupied = 0;
		for (uint32_t c : demo.cellCounts()) occupied += c > 0;
		printf("scenario %d: %zu particles, %d frames, centre of mass (%.4f, %.4f), candidates/particle %llu..%llu, %zu occupied cells\n", scenario, n, frames,
		       n ? cx / n : 0.0, n ? cy / n : 0.0, (unsigned long long)st.min_particle_neighbor_count, (unsigned long long)st.max_particle_neighbor_count, occupied);
	} catch (const std::exception &e) {
		fprintf(stderr, "%s\n", e.what());
		return 3;
	}
	return 0;
}

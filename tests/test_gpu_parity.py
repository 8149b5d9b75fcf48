"""GPU parity tests: every call goes through the C ABI (libsphb200.so via ctypes) and is compared
with the CPU oracle (oracle/sph_oracle.c) and with the golden dumps of the reference itself.

Bars (stated where used):
  * integer work — cell of each particle, per-cell counts, candidate sets, sorted order: bit-exact;
  * EXACT fp mode vs the oracle's gather mode: bit-exact (same operations, same order);
  * vs the reference's own numbers from the same state: density within 2e-5 relative (summation
    order is the only difference); collisions bit-exact;
  * FAST fp mode vs EXACT after one pass: 2e-5 relative.
"""
import glob
import os

import numpy as np
import pytest

from oracle_lib import MODE_COLORED, MODE_JACOBI, CpuSim, point_solvers

pytestmark = pytest.mark.gpu

DT = float(np.float32(1.0) / np.float32(60.0))
GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pkg():
    import nbodysimulation_experiment_b200 as p

    return p


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bits_equal(a, b, what):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, what
    same = (bits(a) == bits(b)) | ((a == 0) & (b == 0))  # +0 / -0 compare equal in the reference too
    if not same.all():
        bad = np.argwhere(~same)
        raise AssertionError(f"{what}: {len(bad)} of {a.size} values differ, first at {bad[0]}: {a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}, "
                             f"max abs diff {np.nanmax(np.abs(a - b)):.3e}")


def candidate_sets_from_grid(ids, start, gx, gy, cx, cy):
    """candidate list of a particle in cell (cx,cy) read off the GPU's sorted grid, reference order"""
    out = []
    for y in range(cy - 1, cy + 2):
        for x in range(cx - 1, cx + 2):
            if 0 <= x < gx and 0 <= y < gy:
                c = y * gx + x
                out.append(ids[start[c]:start[c + 1]])
    return np.concatenate(out) if out else np.zeros(0, np.uint32)


# ---- whole steps, reference scenes, GPU exact == oracle (same solver), bit for bit -----------------
# solver "gs" = coloured Gauss-Seidel sweeps (default), "gather" = Jacobi gather with relaxation omega
@pytest.mark.parametrize("scene,steps,solver,omega", [
    (0, 64, "gs", 1.0), (1, 64, "gs", 1.0), (2, 64, "gs", 1.0), (3, 64, "gs", 1.0),
    (0, 8, "gather", 1.0), (0, 40, "gather", 0.5), (1, 40, "gather", 1.0), (2, 64, "gather", 1.0), (3, 64, "gather", 1.0)])
def test_scene_steps_bitwise_vs_oracle(pkg, scene, steps, solver, omega):
    gpu = pkg.ParticleSimulation(relaxation=omega, solver=pkg.SPH_SOLVER_COLORED_GS if solver == "gs" else pkg.SPH_SOLVER_GATHER)
    gpu.LoadScenario(scene, seed=1)
    cpu = CpuSim("oracle", mode=MODE_COLORED if solver == "gs" else MODE_JACOBI, threads=8)
    cpu.set_relaxation(omega)
    cpu.load_scenario(scene, 1)
    n = cpu.n
    assert gpu.GetParticleCount() == n
    assert_bits_equal(gpu.particles(), cpu.particles(), "initial state")
    assert np.array_equal(gpu.params_array(), cpu.params())
    for s in range(steps):
        gpu.Update(DT)
        cpu.advance(DT)
        if s in (0, 1, steps - 1):
            assert_bits_equal(gpu.particles(), cpu.particles(), f"scene {scene} state after step {s + 1}")
    # integer side
    assert np.array_equal(gpu.cell_of_particle(), cpu.cell_of_particle())
    assert np.array_equal(gpu.cell_counts(), cpu.cell_counts())
    ids, start = gpu.sorted_ids(), gpu.cell_start()
    gx, gy = gpu.grid_dims()
    assert sorted(ids.tolist()) == list(range(n))  # a permutation: nothing lost, nothing duplicated
    for c in range(gx * gy):
        seg = ids[start[c]:start[c + 1]]
        assert np.array_equal(seg, cpu.cell_members(c))  # same members, same (ascending id) order
    cells = cpu.cell_of_particle()
    for i in range(0, n, 37):
        assert np.array_equal(candidate_sets_from_grid(ids, start, gx, gy, *cells[i]), cpu.neighbors(i))
    st, (cst, _) = gpu.GetStats(), cpu.stats()
    assert (st.min_particle_neighbor_count, st.max_particle_neighbor_count) == (int(cst[0]), int(cst[1]))
    assert st.pair_candidates == int(cpu.neighbor_counts().astype(np.uint64).sum())
    pos, col = gpu.Render()
    assert_bits_equal(pos, cpu.particles()[:, 0:2], "render positions")
    assert_bits_equal(col, cpu.colors(), "render colours")
    gpu.close()
    cpu.close()


@pytest.mark.parametrize("scene", [4, 5, 6, 7])
def test_emitter_scenes_bitwise_vs_oracle(pkg, scene):
    """Emitters change N every few steps (host rand() cadence, demo4.cpp:257-284); polygons and a
    circle in scenes 5 and 7.  Runs are sequential because both sides draw from libc rand()."""
    steps = 150
    gpu = pkg.ParticleSimulation()  # default solver: coloured Gauss-Seidel
    gpu.LoadScenario(scene, seed=9)
    for _ in range(steps):
        gpu.Update(DT)
    a = gpu.particles()
    ga = gpu.cell_counts()
    gpu.close()
    cpu = CpuSim("oracle", mode=MODE_COLORED)
    cpu.load_scenario(scene, 9)
    cpu.advance(DT, steps)
    assert a.shape[0] == cpu.n and cpu.n > 300
    assert_bits_equal(a, cpu.particles(), f"emitter scene {scene}")
    assert np.array_equal(ga, cpu.cell_counts())
    cpu.close()


# ---- against the reference's own dumps -------------------------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN_DIR, "scene*.npz"))), ids=lambda p: os.path.basename(p))
def test_grid_and_density_from_reference_state(pkg, path):
    """Inject a state the REFERENCE produced; the GPU's cell assignment, per-cell counts and candidate
    sets must equal the reference's exactly, and density/pressure must match to summation order."""
    g = np.load(path)
    last = g[f"state{int(g['steps'][-1])}"]
    n = last.shape[0]
    gpu = pkg.ParticleSimulation()
    p = pkg.SphParams(*[float(x) for x in g["params"]])
    gpu.SetParams(p)
    gpu.AddParticles(last[:, 0:2])
    gpu.put_particles(last)  # re-files the grid
    assert np.array_equal(gpu.cell_of_particle(), g["refiled_cell_of_particle"])
    assert np.array_equal(gpu.cell_counts(), g["refiled_cell_counts"])
    ids, start = gpu.sorted_ids(), gpu.cell_start()
    gx, gy = gpu.grid_dims()
    cells = g["refiled_cell_of_particle"]
    sums = np.zeros(n, np.uint64)
    xors = np.zeros(n, np.uint32)
    lens = np.zeros(n, np.uint32)
    for i in range(n):
        nb = candidate_sets_from_grid(ids, start, gx, gy, *cells[i])
        lens[i] = len(nb)
        sums[i] = nb.astype(np.uint64).sum()
        xors[i] = np.bitwise_xor.reduce(nb * np.uint32(2654435761)) if len(nb) else 0
    assert np.array_equal(lens, g["refiled_neighbor_counts"])
    assert np.array_equal(sums, g["refiled_cand_sum"]) and np.array_equal(xors, g["refiled_cand_xor"])
    gpu.RunPass(pkg._lib.PASS_DENSITY, DT)
    d = gpu.particles()[:, 8:12]
    np.testing.assert_allclose(d, g["density_from_last"], rtol=2e-5, atol=2e-5)
    assert_bits_equal(gpu.particles()[:, 0:8], last[:, 0:8], "injected state survives the grid pass")
    gpu.close()


# ---- single passes --------------------------------------------------------------------------------
def random_bodies(rng, sim_add):
    for _ in range(3):
        ang = rng.uniform(0, 2 * np.pi)
        sim_add("plane", (np.float32(np.cos(ang)), np.float32(np.sin(ang)), np.float32(rng.uniform(-3, -1))))
    for _ in range(3):
        sim_add("circle", tuple(np.float32(v) for v in (rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(0.1, 1.0))))
    for _ in range(3):
        sim_add("segment", tuple(np.float32(v) for v in rng.uniform(-3, 3, 4)))
    for _ in range(4):
        k = rng.integers(3, 9)
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        c = rng.uniform(-2, 2, 2)
        verts = (np.stack([np.cos(ang), np.sin(ang)], 1) * rng.uniform(0.2, 1.2) + c).astype(np.float32)
        sim_add("polygon", verts if rng.integers(2) else verts[::-1].copy())


def test_collision_pass_bitwise(pkg):
    """All four solvers of sph.h:514-681, bodies applied in insertion order, on 20 000 random points."""
    rng = np.random.default_rng(5)
    gpu = pkg.ParticleSimulation(max_particles=20000)
    cpu = CpuSim("oracle", mode=MODE_JACOBI)
    solvers = point_solvers("oracle")
    bodies = []

    def add(kind, args):
        bodies.append((kind, args))
        if kind == "plane":
            gpu.AddPlane(args[:2], args[2]); cpu.add_plane(*map(float, args))
        elif kind == "circle":
            gpu.AddCircle(args[:2], args[2]); cpu.add_circle(*map(float, args))
        elif kind == "segment":
            gpu.AddLineSegment(args[:2], args[2:]); cpu.add_segment(*map(float, args))
        else:
            gpu.AddPolygon(args); cpu.polygon(args)

    random_bodies(rng, add)
    pts = rng.uniform(-4.5, 4.5, (20000, 2)).astype(np.float32)
    pts[:, 1] *= 0.6
    gpu.AddParticles(pts)
    for x, y in pts:
        cpu.add_particle(float(x), float(y), 0.0, 0.0)
    gpu.RunPass(pkg._lib.PASS_COLLIDE, DT)
    cpu.pass_collide()
    assert_bits_equal(gpu.particles()[:, 0:2], cpu.particles()[:, 0:2], "collide pass")
    # and a handful through the single-point solvers, body by body
    moved = gpu.particles()[:, 0:2]
    for i in range(0, 20000, 997):
        p = pts[i].copy()
        for kind, args in bodies:
            p = solvers[kind](p, *args) if kind != "polygon" else solvers[kind](p, args)
        assert_bits_equal(moved[i], p, f"point {i}")
    gpu.close()


@pytest.mark.parametrize("solver", ["gs", "gs-flow", "gs-warp", "gather"])
def test_viscosity_and_delta_passes_bitwise(pkg, solver):
    """Per-pass parity from an injected state (a reference dump, mid-splash).  The passes run twice over the same
    grid: the one-launch sweep numbers its passes over a grid (done flags count them), so a third and fourth pass
    must still wait for their neighbours."""
    g = np.load(os.path.join(GOLDEN_DIR, "scene1.npz"))
    state = g["state32"]
    flags = {"gs-flow": pkg._lib.SPH_FLAG_SWEEP_FLOW, "gs-warp": pkg._lib.SPH_FLAG_SWEEP_WARP}.get(solver, 0)
    gpu = pkg.ParticleSimulation(solver=pkg.SPH_SOLVER_GATHER if solver == "gather" else pkg.SPH_SOLVER_COLORED_GS, flags=flags)
    cpu = CpuSim("oracle", mode=MODE_JACOBI if solver == "gather" else MODE_COLORED)
    gpu.LoadScenario(1, seed=1)
    cpu.load_scenario(1, 1)
    gpu.put_particles(state)
    cpu.put_particles(state)
    cpu.pass_neighbor_search()
    for which, run in ((pkg._lib.PASS_VISCOSITY, lambda: cpu.pass_viscosity(DT)), (pkg._lib.PASS_DENSITY, cpu.pass_density),
                       (pkg._lib.PASS_DELTA, lambda: cpu.pass_delta(DT)), (pkg._lib.PASS_VISCOSITY, lambda: cpu.pass_viscosity(DT)),
                       (pkg._lib.PASS_DELTA, lambda: cpu.pass_delta(DT))):
        gpu.RunPass(which, DT)
        run()
        assert_bits_equal(gpu.particles(), cpu.particles(), f"pass {which}")
    gpu.close()


@pytest.mark.parametrize("solver", [0, 1])
def test_fast_mode_close_to_exact(pkg, solver):
    runs = []
    for mode in (pkg.SPH_FP_EXACT, pkg.SPH_FP_FAST):
        s = pkg.ParticleSimulation(fp_mode=mode, solver=solver)
        s.LoadScenario(2, seed=1)
        s.Update(DT)
        s.Update(DT)
        runs.append(s.particles())
        s.close()
    assert np.array_equal(runs[0][:, 0:2] != runs[0][:, 0:2], runs[1][:, 0:2] != runs[1][:, 0:2])  # no NaNs appear
    np.testing.assert_allclose(runs[1][:, 8:12], runs[0][:, 8:12], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(runs[1][:, 0:2], runs[0][:, 0:2], rtol=0, atol=2e-5)


def test_runs_are_bit_reproducible(pkg):
    runs = []
    for _ in range(2):
        s = pkg.ParticleSimulation()
        s.LoadScenario(0, seed=1)
        for _ in range(5):
            s.Update(DT)
        runs.append((s.particles(), s.sorted_ids()))
        s.close()
    assert_bits_equal(runs[0][0], runs[1][0], "two runs")
    assert np.array_equal(runs[0][1], runs[1][1])


# ---- larger synthetic scenes ----------------------------------------------------------------------
def test_hashed_block_bitwise_vs_threaded_oracle(pkg):
    """65 536 particles from the device-side generator; oracle gather mode on 8 threads."""
    from nbodysimulation_experiment_b200 import scenes

    gpu = scenes.fill_block(scenes.block_scene(256, spacing=0.1, gravity=(0.0, -2.0), solver=pkg.SPH_SOLVER_GATHER))
    n = gpu.GetParticleCount()
    assert n == 256 * 256
    init = gpu.particles()
    w, h = gpu.scene["width"], gpu.scene["height"]
    cpu = CpuSim("oracle", width=w, height=h, cell=scenes.KERNEL_HEIGHT, mode=MODE_JACOBI, threads=8)
    assert cpu.dims() == gpu.grid_dims()
    cpu.put_params(gpu.params_array())
    cpu.set_gravity(0.0, -2.0)
    for nx, ny, d in ((0.0, 1.0, -h / 2), (0.0, -1.0, -h / 2), (1.0, 0.0, -w / 2), (-1.0, 0.0, -w / 2)):
        cpu.add_plane(nx, ny, float(np.float32(d)))
    for x, y in init[:, 0:2]:
        cpu.add_particle(float(x), float(y), 0.0, 0.0)
    for s in range(6):
        gpu.Update(DT)
        cpu.advance(DT)
    assert_bits_equal(gpu.particles(), cpu.particles(), "65k block after 6 steps")
    assert np.array_equal(gpu.cell_counts(), cpu.cell_counts())
    gpu.close()
    cpu.close()


def test_million_particle_invariants(pkg):
    """Full size of BASELINE.json's 1M config: properties that need no oracle run."""
    from nbodysimulation_experiment_b200 import scenes

    gpu = scenes.fill_block(scenes.block_scene(1024, spacing=0.1, gravity=(0.0, -0.5)))
    n = gpu.GetParticleCount()
    assert n == 1024 * 1024
    for _ in range(10):
        gpu.Update(DT)
    p = gpu.particles()
    ids, start = gpu.sorted_ids(), gpu.cell_start()
    gx, gy = gpu.grid_dims()
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))  # permutation
    assert start[0] == 0 and start[-1] == n and np.all(np.diff(start.astype(np.int64)) >= 0)
    cells = gpu.cell_of_particle()
    keys = (cells[:, 1].astype(np.int64) * gx + cells[:, 0])[ids]
    assert np.all(np.diff(keys) >= 0)  # sorted by cell
    same = np.diff(keys) == 0
    assert np.all(np.diff(ids.astype(np.int64))[same] > 0)  # ascending id inside a cell
    assert np.isfinite(p).all()
    w, h = gpu.scene["width"], gpu.scene["height"]
    assert (np.abs(p[:, 0]) <= w / 2).all() and (np.abs(p[:, 1]) <= h / 2).all()  # planes keep everything inside
    st = gpu.GetStats()
    assert st.max_particle_neighbor_count >= st.min_particle_neighbor_count > 0
    assert 0 < st.min_cell_particle_count <= st.max_cell_particle_count  # running min/max of demo4.cpp:64-67
    # cell of each particle recomputed on the host with the reference formula (sph.h:450-463): the grid is re-filed
    # from the final positions (the step's own grid was filed from the predicted ones, demo4.cpp:342-356)
    gpu.RunPass(pkg._lib.PASS_GRID, DT)
    hw, hh, cell = np.float32(w) * np.float32(0.5), np.float32(h) * np.float32(0.5), np.float32(scenes.KERNEL_HEIGHT)
    hx = np.clip(((p[:, 0] + hw) / cell).astype(np.int32), 0, gx - 1)  # float32 arithmetic, truncation toward zero, clamp
    hy = np.clip(((p[:, 1] + hh) / cell).astype(np.int32), 0, gy - 1)
    cells = gpu.cell_of_particle()
    assert np.array_equal(cells[:, 0], hx) and np.array_equal(cells[:, 1], hy)
    counts = np.bincount(hy.astype(np.int64) * gx + hx, minlength=gx * gy)
    assert np.array_equal(gpu.cell_counts(), counts.astype(np.uint32))
    gpu.close()


# ---- the coloured Gauss-Seidel sweeps -------------------------------------------------------------
@pytest.mark.parametrize("cap,flags", [(96, 0), (160, 0), (512, 0), (96, 16), (160, 8), (512, 16)])
def test_colored_sweep_staging_capacity_does_not_change_results(pkg, cap, flags):
    """Blocks that do not fit the shared-memory staging take the L2 path: same bits either way.
    Scene 0 has 250-360 candidates per block, so cap=96/160 forces the L2 path everywhere.
    flags: 0 = block per cell (default for small scenes), 8 = nine launches warp per cell, 16 = one-launch flow kernel."""
    gpu = pkg.ParticleSimulation(sweep_capacity=cap, flags=flags)
    gpu.LoadScenario(0, seed=1)
    cpu = CpuSim("oracle", mode=MODE_COLORED)
    cpu.load_scenario(0, 1)
    for _ in range(6):
        gpu.Update(DT)
        cpu.advance(DT)
    assert_bits_equal(gpu.particles(), cpu.particles(), f"sweep capacity {cap}")
    gpu.close()
    cpu.close()


@pytest.mark.parametrize("sweep", ["team", "warp", "flow"])
def test_colored_block_bitwise_vs_oracle(pkg, sweep):
    """16 384 particles from the device-side generator, dense regime (spacing h/6), 12 steps, each of the
    three sweep kernels."""
    from nbodysimulation_experiment_b200 import scenes

    flags = {"team": pkg._lib.SPH_FLAG_SWEEP_TEAM, "warp": pkg._lib.SPH_FLAG_SWEEP_WARP, "flow": pkg._lib.SPH_FLAG_SWEEP_FLOW}[sweep]
    gpu = scenes.fill_block(scenes.block_scene(128, spacing=0.05, gravity=(0.0, -8.3), flags=flags))
    n = gpu.GetParticleCount()
    assert n == 128 * 128
    init = gpu.particles()
    w, h = gpu.scene["width"], gpu.scene["height"]
    cpu = CpuSim("oracle", width=w, height=h, cell=scenes.KERNEL_HEIGHT, mode=MODE_COLORED)
    cpu.put_params(gpu.params_array())
    cpu.set_gravity(0.0, float(np.float32(-8.3)))
    for nx, ny, d in ((0.0, 1.0, -h / 2), (0.0, -1.0, -h / 2), (1.0, 0.0, -w / 2), (-1.0, 0.0, -w / 2)):
        cpu.add_plane(nx, ny, float(np.float32(d)))
    for x, y in init[:, 0:2]:
        cpu.add_particle(float(x), float(y), 0.0, 0.0)
    for _ in range(12):
        gpu.Update(DT)
        cpu.advance(DT)
    assert_bits_equal(gpu.particles(), cpu.particles(), "16k dense block after 12 steps")
    gpu.close()
    cpu.close()


def test_colored_solver_tracks_reference_energy(pkg):
    """Statistical parity with the REFERENCE's own single-thread run of its default scene (golden dump): the coloured
    sweep is a different, equally legitimate in-place order, so trajectories differ particle by particle (like the
    reference's MT vs ST runs) but the bulk must agree - within the reference's own MT-vs-ST spread (envelope_lib;
    all four volume scenes and K up to 256: tests/test_gpu_envelope.py)."""
    import envelope_lib as E

    g = np.load(os.path.join(GOLDEN_DIR, "scene0.npz"))
    gpu = pkg.ParticleSimulation()
    gpu.LoadScenario(0, seed=1)
    assert_bits_equal(gpu.particles(), g["init"], "initial state equals the reference's")
    for _ in range(8):
        gpu.Update(DT)
    E.check_aggregates(E.load(), 0, 8, gpu.particles())
    gpu.close()


# ---- edge cases the reference guards with asserts ---------------------------
def test_empty_and_single_particle(pkg):
    sim = pkg.ParticleSimulation()
    sim.AddPlane((0.0, 1.0), -2.8125)
    sim.Update(DT)  # nothing to do, must not fail
    assert sim.GetParticleCount() == 0 and sim.particles().shape == (0, 12)
    sim.SetGravity((0.0, -10.0))
    sim.AddParticle((0.5, 0.25), (3.0, 0.0))
    cpu = CpuSim("oracle", mode=MODE_COLORED)
    cpu.add_plane(0.0, 1.0, -2.8125)
    cpu.set_gravity(0.0, -10.0)
    cpu.add_particle(0.5, 0.25, 3.0, 0.0)
    for _ in range(30):
        sim.Update(DT)
        cpu.advance(DT)
    assert_bits_equal(sim.particles(), cpu.particles(), "one particle: acceleration consumed once, then free fall")
    st = sim.GetStats()
    assert (st.min_particle_neighbor_count, st.max_particle_neighbor_count) == (1, 1)  # itself (demo4.cpp:183-206)
    sim.close()


def test_capacities_are_errors_not_asserts(pkg):
    sim = pkg.ParticleSimulation(max_particles=100)
    sim.AddParticles(np.zeros((100, 2), np.float32))
    with pytest.raises(pkg.SphError) as e:
        sim.AddParticle((0.0, 0.0))
    assert e.value.code == -2
    for _ in range(100):  # kSPHMaxBodyCount, sph.h:71
        sim.AddCircle((0.0, 0.0), 0.1)
    with pytest.raises(pkg.SphError) as e:
        sim.AddCircle((0.0, 0.0), 0.1)
    assert e.value.code == -2
    with pytest.raises(pkg.SphError):
        sim.AddPolygon(np.zeros((9, 2), np.float32))  # kMaxScenarioPolygonCount, sph.h:161
    for _ in range(8):  # kSPHMaxEmitterCount, sph.h:72
        sim.AddEmitter((0, 0), (1, 0), 0.3, 1.0, 15.0, 1.0)
    with pytest.raises(pkg.SphError) as e:
        sim.AddEmitter((0, 0), (1, 0), 0.3, 1.0, 15.0, 1.0)
    assert e.value.code == -2
    sim.close()


def test_out_of_domain_positions_clamp_into_edge_cells(pkg):
    """sph.h:459-460: the grid (9.9 x 5.4) is smaller than the domain and everything outside is
    clamped into the edge cells; also more than the reference's 500 particles per cell."""
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.uniform(-8, 8, (3000, 2)), rng.uniform(-0.1, 0.1, (700, 2)) + [1.15, 1.16]]).astype(np.float32)
    sim = pkg.ParticleSimulation(solver=pkg.SPH_SOLVER_GATHER)
    cpu = CpuSim("oracle", mode=MODE_JACOBI)
    sim.AddParticles(pts)
    for x, y in pts:
        cpu.add_particle(float(x), float(y), 0.0, 0.0)
    sim.RunPass(pkg._lib.PASS_GRID, DT)
    cpu.pass_update_grid()
    cpu.pass_neighbor_search()
    assert np.array_equal(sim.cell_of_particle(), cpu.cell_of_particle())
    assert np.array_equal(sim.cell_counts(), cpu.cell_counts())
    assert sim.cell_counts().max() > 500
    sim.RunPass(pkg._lib.PASS_DENSITY, DT)
    cpu.pass_density()
    assert_bits_equal(sim.particles()[:, 8:12], cpu.particles()[:, 8:12], "density with a 700-particle cell")
    sim.close()


def test_parameters_and_external_force(pkg):
    """SetParams (10x viscosity, different stiffness), AddExternalForces / ClearExternalForce."""
    for solver, mode in ((pkg.SPH_SOLVER_COLORED_GS, MODE_COLORED), (pkg.SPH_SOLVER_GATHER, MODE_JACOBI)):
        sim = pkg.ParticleSimulation(solver=solver)
        cpu = CpuSim("oracle", mode=mode)
        sim.LoadScenario(3, seed=4)
        cpu.load_scenario(3, 4)
        p = sim.GetParams()
        p.linear_viscosity, p.quadratic_viscosity, p.stiffness, p.near_stiffness = 5.0, 3.0, 0.4, 8.0
        sim.SetParams(p)
        cpu.put_params(sim.params_array())
        sim.AddExternalForces((1.5, 0.5))
        cpu.add_external_force(1.5, 0.5)
        for k in range(20):
            if k == 10:
                sim.ClearExternalForce()
                cpu.clear_external_force()
            sim.Update(DT)
            cpu.advance(DT)
        assert_bits_equal(sim.particles(), cpu.particles(), f"solver {solver}")
        sim.close()


@pytest.mark.parametrize("scene,steps", [(0, 24), (3, 40), (5, 150)])
def test_block_per_cell_and_warp_per_cell_sweeps_agree_with_the_oracle(pkg, scene, steps):
    """The two sweep kernels (one warp / one block per cell) must produce the same bits; scene 5 has
    emitters and polygons.  Each is compared with the oracle, hence with each other."""
    cpu = CpuSim("oracle", mode=MODE_COLORED)
    cpu.load_scenario(scene, 6)
    cpu.advance(DT, steps)
    want = cpu.particles()
    for flags in (pkg._lib.SPH_FLAG_SWEEP_TEAM, pkg._lib.SPH_FLAG_SWEEP_WARP, pkg._lib.SPH_FLAG_SWEEP_FLOW,
                  pkg._lib.SPH_FLAG_SWEEP_FLOW | pkg._lib.SPH_FLAG_NO_GRAPHS):
        s = pkg.ParticleSimulation(flags=flags)
        s.LoadScenario(scene, seed=6)
        for _ in range(steps):
            s.Update(DT)
        assert_bits_equal(s.particles(), want, f"scene {scene}, sweep flags {flags}")
        s.close()
    cpu.close()


@pytest.mark.parametrize("nx,spacing,g,steps", [(512, 0.1, -10.0, 60), (1024, 0.1, -0.5219, 300), (384, 0.05, -3.0, 16)])
def test_one_launch_flow_sweep_equals_nine_launch_sweep_at_scale(pkg, nx, spacing, g, steps):
    """The dependency-driven one-launch sweep (persistent warps, per-cell done flags) against the nine
    per-colour launches at sizes where every SM is busy and cells of different colours really run
    concurrently: 262 144 particles in a violent collapse (g = -10), the 1M bench scene for the length of the bench
    run (late in it heavy cells keep their neighbours waiting, profiles/r1_final_kernels_dambreak1m_step260.txt), a
    dense block.
    Bit-identical state, and two flow runs agree with each other (no ordering left to chance)."""
    from nbodysimulation_experiment_b200 import scenes

    runs = []
    for flags in (pkg._lib.SPH_FLAG_SWEEP_WARP, pkg._lib.SPH_FLAG_SWEEP_FLOW, pkg._lib.SPH_FLAG_SWEEP_FLOW | pkg._lib.SPH_FLAG_NO_GRAPHS):
        sim = scenes.fill_block(scenes.block_scene(nx, spacing=spacing, gravity=(0.0, g), flags=flags))
        assert sim.GetParticleCount() == nx * nx
        for _ in range(steps):
            sim.Update(DT)
        runs.append(sim.particles())
        sim.GetStats()  # raises on a capacity / overflow flag
        sim.close()
    assert np.isfinite(runs[0]).all()
    assert_bits_equal(runs[1], runs[0], f"flow vs nine launches, {nx}x{nx}")
    assert_bits_equal(runs[2], runs[0], f"flow (plain launches) vs nine launches, {nx}x{nx}")


def test_graph_replay_is_bit_identical_to_plain_launches(pkg):
    runs = []
    for flags in (0, pkg._lib.SPH_FLAG_NO_GRAPHS):
        s = pkg.ParticleSimulation(flags=flags)
        s.LoadScenario(1, seed=2)
        for _ in range(25):
            s.Update(DT)
        runs.append(s.particles())
        s.close()
    assert_bits_equal(runs[0], runs[1], "graph vs plain")


def test_overlapped_owned_readback_equals_the_synchronous_one(pkg):
    """sph_render_owned / sph_wait_render_owned (the strip counterpart of Render, copies overlapped with the next
    step) against sph_read_owned, on one GPU where a rank owns everything; emitters make the count grow between
    frames, including by more than the 10 % head-room a frame ships (the tail fetch)."""
    sim = pkg.ParticleSimulation()
    sim.LoadScenario(5, seed=3)  # emitters + polygons
    bufs = [sim.owned_buffers(records=False, render=True, pinned=True) for _ in range(2)]
    sync = sim.owned_buffers(records=False, render=True, pinned=False)
    for k in range(60):
        sim.Update(DT)
        if k == 30:  # a burst: +40 % particles in one frame
            n = sim.GetParticleCount()
            sim.AddParticles(np.random.default_rng(k).uniform(-1.0, 1.0, (max(n * 2 // 5, 50), 2)).astype(np.float32))
        sim.render_owned(bufs[k % 2])
        got = sim.wait_render_owned()
        want = sim.read_owned(records=False, render=True, buffers=sync)  # same state: nothing stepped in between
        assert len(got["ids"]) == len(want["ids"]) == sim.GetParticleCount()
        a, b = np.argsort(got["ids"]), np.argsort(want["ids"])
        assert np.array_equal(got["ids"][a], want["ids"][b])
        assert_bits_equal(got["positions"][a], want["positions"][b], f"frame {k} positions")
        assert_bits_equal(got["colors"][a], want["colors"][b], f"frame {k} colours")
    assert sim.wait_render_owned() is None
    sim.close()


def test_two_gpu_strips_match_one_gpu_bitwise():
    """Runs tools/mgpu_check.py under torchrun when the box has two GPUs (the round-end box has one)."""
    import subprocess
    import sys
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for port, extra in (("29517", []), ("29518", ["--rebalance", "8", "--steps", "100"])):  # static strips, then re-balanced every 8 steps
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", port,
                              os.path.join(root, "tools", "mgpu_check.py"), "--nx", "256", "--steps", "40"] + extra, capture_output=True, text=True, timeout=240)
        assert out.returncode == 0 and "MGPU PARITY OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]

"""y-strip decomposition on the GPU, runnable on a ONE-GPU box: the strips of a scene live in one
process and share the device (sph_comm_init_local / sph_step_group).  Kernels, mailbox protocol (records stored by
predict_key_kernel into the neighbour's mailbox, publish / wait by sequence number, unpack) and the keep/send rule are
exactly what one-process-per-GPU runs use; only the mapping of the neighbour's memory differs (plain pointers instead
of CUDA IPC).  The bar everywhere: the N-strip result equals the single-GPU result BIT FOR BIT, particle by particle
(within-cell order is by global id, so every floating-point sum has the same order on any number of strips).

From 3 strips on, interior ranks have two neighbours: the second unpack lands behind the first one's records, and a
strip's authority ends on both sides - neither is exercised by 2 strips.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DT = float(np.float32(1.0) / np.float32(60.0))


def run_strips(world, nx, steps, gravity=-10.0, spacing=0.1, halo_rows=0, rebalance=0, max_shift=2, flags=0, bodies=False, ny=None):
    """-> (records merged by creation id, final strips) of `world` strips sharing cuda:0"""
    from nbodysimulation_experiment_b200 import StripGroup, scenes, strips

    build = scenes.bodies_scene if bodies else scenes.block_scene
    kw = dict(spacing=spacing, gravity=(0.0, gravity), halo_rows=halo_rows, flags=flags)
    if ny is not None:
        kw["ny"] = ny
    n_total = nx * (ny or nx)
    sims = [build(nx, rank=r, world_size=world, capacity=2 * n_total + 1024, halo_capacity=n_total, **kw) for r in range(world)]
    rows = scenes.block_strips(sims[0], world)
    for r, s in enumerate(sims):
        s.set_strip(*rows[r])
        if rebalance:
            s.set_rebalance(rebalance, max_shift)
    group = StripGroup(sims)
    for s in sims:
        scenes.fill_block(s)
    assert sum(s.local_particle_count() for s in sims) == n_total
    for _ in range(steps):
        group.Update(DT)
    parts = []
    for s in sims:
        got = s.read_owned(records=True)
        s.GetStats()  # raises on overflow / lost / timeout flags
        parts.append((got["ids"].copy(), got["records"].copy()))
    final = [s.get_strip() for s in sims]
    merged = strips.merge_owned(parts, n_total)  # raises unless ownership is a partition
    group.close()
    return merged, final, rows


def run_single(nx, steps, gravity=-10.0, spacing=0.1, bodies=False, ny=None):
    from nbodysimulation_experiment_b200 import scenes

    build = scenes.bodies_scene if bodies else scenes.block_scene
    kw = dict(spacing=spacing, gravity=(0.0, gravity))
    if ny is not None:
        kw["ny"] = ny
    one = scenes.fill_block(build(nx, **kw))
    for _ in range(steps):
        one.Update(DT)
    out = one.particles()
    one.GetStats()
    one.close()
    return out


def assert_same_bits(multi, ref, what):
    same = (multi.view(np.uint32) == ref.view(np.uint32)) | ((multi == 0) & (ref == 0))
    bad = np.flatnonzero(~same.all(1))
    assert len(bad) == 0, f"{what}: {len(bad)} of {len(ref)} particles differ, max abs diff {np.abs(multi - ref).max():.3e}"


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_strips_on_one_gpu_match_the_single_gpu_run_bitwise(world):
    """192 x 384 block (a tall column: 129 grid rows, so that 8 strips are still taller than the halo), 60 steps of
    the violent g = -10 collapse: particles migrate across every boundary, interior strips unpack two neighbours."""
    nx, ny, steps = 192, 384, 60
    ref = run_single(nx, steps, ny=ny)
    multi, final, rows = run_strips(world, nx, steps, ny=ny)
    assert_same_bits(multi, ref, f"{world} strips {rows}")


def test_three_strips_rebalanced_match_the_single_gpu_run_bitwise():
    """Re-balancing every 8 steps (sph_set_rebalance; the group sums the row histograms on the host): boundaries move
    while the column collapses, rows change hands inside the step's one exchange, results stay identical."""
    nx, ny, steps = 256, 512, 96
    ref = run_single(nx, steps, ny=ny)
    multi, final, rows = run_strips(3, nx, steps, ny=ny, rebalance=8, max_shift=3)
    assert_same_bits(multi, ref, f"3 strips {rows} -> {final}")
    assert final != rows, "the collapse must have moved at least one boundary"


def test_strips_with_and_without_graphs_agree():
    from nbodysimulation_experiment_b200 import _lib

    a, _, _ = run_strips(3, 192, 24, ny=384)
    b, _, _ = run_strips(3, 192, 24, ny=384, flags=_lib.SPH_FLAG_NO_GRAPHS)
    assert_same_bits(a, b, "graph replay vs plain launches on strips")


def test_bodies_scene_on_four_strips_matches_one_gpu():
    """configs[4] in small: dense block, 10x viscosity, circles and boxes, four strips."""
    nx, steps = 160, 40
    ref = run_single(nx, steps, gravity=-3.0, spacing=0.05, bodies=True)
    multi, _, rows = run_strips(4, nx, steps, gravity=-3.0, spacing=0.05, bodies=True)
    assert_same_bits(multi, ref, f"bodies scene on 4 strips {rows}")


def test_halo_depth_below_the_bound_is_detected_as_a_difference():
    """The halo bound of DESIGN.md (an edge error travels <= 3 rows per coloured sweep, two sweeps per exchange: 6 rows,
    default 7) is not slack: with 2 ghost rows the strips visibly diverge from the single-GPU run.  Keeps the parity
    tests above honest - they would catch a halo that is too thin."""
    nx, ny, steps = 192, 384, 40
    ref = run_single(nx, steps, ny=ny)
    multi, _, _ = run_strips(2, nx, steps, ny=ny, halo_rows=2)
    same = (multi.view(np.uint32) == ref.view(np.uint32)) | ((multi == 0) & (ref == 0))
    assert not same.all()


def test_group_refuses_plain_step_and_wrong_order():
    from nbodysimulation_experiment_b200 import SphError, StripGroup, scenes

    sims = [scenes.block_scene(64, rank=r, world_size=2, capacity=20000, halo_capacity=8192) for r in range(2)]
    with pytest.raises(SphError):
        sims[0].Update(DT)  # no communicator yet
    group = StripGroup(sims)
    with pytest.raises(SphError):
        sims[0].Update(DT)  # strips of one process are stepped together
    with pytest.raises(SphError):
        StripGroup(sims)  # already wired
    group.close()


# ---- the reference's own scenes on strips: host-side particle lists, emitters, state injection --------------------------
def reference_scene_on_strips(scene, steps, world, halo_rows=0, inject=None):
    """LoadScenario on every strip (the whole list goes to every rank, each keeps its rows), `steps` updates -> records
    merged by creation id"""
    from nbodysimulation_experiment_b200 import ParticleSimulation, StripGroup, strips

    sims = [ParticleSimulation(rank=r, world_size=world, max_particles=20000, halo_capacity=20000, halo_rows=halo_rows) for r in range(world)]
    for s in sims:
        s.LoadScenario(scene, seed=1)
        if inject is not None:
            s.put_particles(inject)
    group = StripGroup(sims)
    for _ in range(steps):
        group.Update(DT)
    parts = []
    for s in sims:
        got = s.read_owned(records=True)
        s.GetStats()
        parts.append((got["ids"].copy(), got["records"].copy()))
    total = sims[0].GetParticleCount()
    merged = strips.merge_owned(parts, total)
    group.close()
    return merged


def reference_scene_single(scene, steps, inject=None):
    from nbodysimulation_experiment_b200 import ParticleSimulation

    one = ParticleSimulation()
    one.LoadScenario(scene, seed=1)
    if inject is not None:
        one.put_particles(inject)
    for _ in range(steps):
        one.Update(DT)
    out = one.particles()
    one.GetStats()
    one.close()
    return out


@pytest.mark.parametrize("scene,steps,world,halo", [(0, 48, 2, 0), (3, 60, 2, 0), (5, 150, 2, 0), (7, 120, 2, 0), (4, 120, 3, 6)])
def test_reference_scenes_on_strips_match_the_single_gpu_run_bitwise(scene, steps, world, halo):
    """sph_add_volume / emitters (demo4.cpp:169-181, 257-284) on strips: volumes with the libc rand() jitter, emitters
    that change N every few frames, boxes and a circle - the same bits as one GPU.  (18 grid rows: two strips of 9 rows
    with the default 7-row halo, or three of 6 with a 6-row halo.)"""
    ref = reference_scene_single(scene, steps)
    multi = reference_scene_on_strips(scene, steps, world, halo_rows=halo)
    assert len(multi) == len(ref)
    assert_same_bits(multi, ref, f"scene {scene} on {world} strips")


def test_state_injection_on_strips():
    """sph_write_particles on strips (demo4.cpp:223-255 from an injected reference state): every rank is given the whole
    state and keeps its window; then 12 steps equal the single-GPU run from the same injected state."""
    import os

    state = np.load(os.path.join(os.path.dirname(__file__), "golden", "scene0.npz"))["state8"]
    ref = reference_scene_single(0, 12, inject=state)
    multi = reference_scene_on_strips(0, 12, 2, inject=state)
    assert_same_bits(multi, ref, "scene 0 from the reference's state after 8 steps, 2 strips")

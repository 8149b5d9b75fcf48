"""CPU, authoring container only: the oracle restatement against the LIVE reference library
(oracle/_ref/libsphref.so = the reference's demo4.cpp compiled headless).  Skipped where the
reference could not be built (it needs /root/reference)."""
import numpy as np
import pytest

from oracle_lib import MODE_GS_INDEX, CpuSim, have_ref, point_solvers

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libsphref.so not built (needs /root/reference)")
DT = float(np.float32(1.0) / np.float32(60.0))


@pytest.mark.parametrize("scene", [0, 1, 2, 3])
def test_volume_scenes_bitwise(scene):
    a, b = CpuSim("ref"), CpuSim("oracle", mode=MODE_GS_INDEX)
    a.load_scenario(scene, 5)
    b.load_scenario(scene, 5)  # both consume the same rand() values: each call re-seeds
    assert np.array_equal(a.particles(), b.particles())
    for _ in range(12 if scene == 0 else 40):
        a.advance(DT)
        b.advance(DT)
    assert np.array_equal(a.particles(), b.particles())
    gx, gy = a.dims()
    for c in range(gx * gy):
        assert np.array_equal(a.cell_members(c), b.cell_members(c))  # same storage order, not just same set
    for i in range(0, a.n, 61):
        assert np.array_equal(a.neighbors(i), b.neighbors(i))
    assert np.array_equal(a.stats()[0], b.stats()[0])


@pytest.mark.parametrize("scene", [4, 5, 6, 7])
def test_emitter_scenes_bitwise(scene):
    # emitters draw rand() every few steps, so the two runs cannot interleave
    out = []
    for kind in ("ref", "oracle"):
        s = CpuSim(kind)
        s.load_scenario(scene, 3)
        for _ in range(240):
            s.advance(DT)
        out.append((s.particles(), s.neighbor_counts(), s.stats()[0]))
        s.close()
    assert out[0][0].shape == out[1][0].shape and out[0][0].shape[0] > 300
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


def test_collision_solvers_fuzz():
    A, B = point_solvers("ref"), point_solvers("oracle")
    rng = np.random.default_rng(11)
    for _ in range(4000):
        p = rng.uniform(-2, 2, 2).astype(np.float32)
        a3 = rng.uniform(-1, 1, 3).astype(np.float32)
        assert np.array_equal(A["circle"](p, a3[0], a3[1], abs(a3[2])), B["circle"](p, a3[0], a3[1], abs(a3[2])))
        n = (a3[:2] / np.linalg.norm(a3[:2])).astype(np.float32)
        assert np.array_equal(A["plane"](p, n[0], n[1], a3[2]), B["plane"](p, n[0], n[1], a3[2]))
        q = rng.uniform(-1, 1, 4).astype(np.float32)
        assert np.array_equal(A["segment"](p, *q), B["segment"](p, *q))
        ang = np.sort(rng.uniform(0, 2 * np.pi, rng.integers(3, 9)))
        verts = (np.stack([np.cos(ang), np.sin(ang)], 1) * rng.uniform(0.2, 1.5)).astype(np.float32)
        assert np.array_equal(A["polygon"](p, verts), B["polygon"](p, verts))
        assert np.array_equal(A["polygon"](p, verts[::-1]), B["polygon"](p, verts[::-1]))
    # a particle exactly on a circle's centre is left alone (sph.h:534)
    assert np.array_equal(A["circle"]([0.25, -0.5], 0.25, -0.5, 1.0), np.array([0.25, -0.5], np.float32))
    assert np.array_equal(B["circle"]([0.25, -0.5], 0.25, -0.5, 1.0), np.array([0.25, -0.5], np.float32))

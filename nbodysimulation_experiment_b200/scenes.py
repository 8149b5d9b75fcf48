"""Synthetic large scenes of BASELINE.json `configs`: a square block of particles
resting in the lower-left corner of a 16:9 domain four block-widths wide, like the reference's
dam (sph.h:307-310, 316-329), with the reference's constants h = cell = 0.3, dt = 1/60.

The particles come from `sph_add_volume_hashed` (AddVolume's lattice, demo4.cpp:169-181, with a
counter-hash jitter generated on the device), so nothing is built on the host.
"""
import numpy as np

from . import _lib, strips
from .simulation import ParticleSimulation

KERNEL_HEIGHT = float(np.float32(6.0) * np.float32(0.05))  # sph.h:35-36


def _domain(nx, spacing):
    width = 4.0 * nx * spacing
    return width, width * 9.0 / 16.0


def block_scene(nx, ny=None, spacing=0.1, gravity=(0.0, -10.0), seed=1337, fp_mode=_lib.SPH_FP_EXACT, flags=0,
                relaxation=1.0, device=0, rank=0, world_size=1, capacity=None, near_stiffness=None,
                solver=_lib.SPH_SOLVER_COLORED_GS, sweep_capacity=0, halo_rows=0, halo_capacity=0, **params):
    """c3 / c4: nx x ny block (default square), 4 boundary planes, dam-break under gravity."""
    ny = nx if ny is None else ny
    width, height = _domain(nx, spacing)
    n = nx * ny
    if capacity is None:
        # a strip holds what it owns (the dam collapses into the lower strips: allow 3x the even share)
        # plus the ghost copies of both neighbours
        capacity = n + 1024 if world_size == 1 else min(n, int(n / world_size * 3.0)) + int(n / world_size * 0.75) + 65536
    if world_size > 1 and not halo_capacity:
        # records per exchange message buffer: the block's rows that fit the halo, x6 for the compression
        # of a collapsing column (the messages themselves shrink to 1.5 x the observed peak after 16 steps)
        rows_per_rank = max(ny * spacing / KERNEL_HEIGHT / world_size, 1.0)
        halo_capacity = strips.halo_capacity_estimate(n / world_size, rows_per_rank, halo_rows or strips.DEFAULT_HALO_ROWS, safety=6.0)
    sim = ParticleSimulation(domain_width=width, domain_height=height, cell_size=KERNEL_HEIGHT, max_particles=capacity,
                             device=device, fp_mode=fp_mode, flags=flags, relaxation=relaxation, rank=rank, world_size=world_size,
                             solver=solver, sweep_capacity=sweep_capacity, halo_rows=halo_rows, halo_capacity=halo_capacity)
    p = sim.GetParams()
    p.particle_spacing = spacing
    if near_stiffness is not None:
        p.near_stiffness = near_stiffness
    for k, v in params.items():
        setattr(p, k, v)
    sim.SetParams(p)
    sim.SetGravity(gravity)
    hw, hh = width * 0.5, height * 0.5
    sim.AddPlane((0.0, 1.0), -hh)   # floor: n.p = d with p = (0,-hh)
    sim.AddPlane((0.0, -1.0), -hh)  # ceiling
    sim.AddPlane((1.0, 0.0), -hw)   # left wall
    sim.AddPlane((-1.0, 0.0), -hw)  # right wall
    cx = -hw + nx * spacing * 0.5 + 0.05
    cy = -hh + ny * spacing * 0.5 + 0.05
    sim.scene = {"nx": nx, "ny": ny, "spacing": spacing, "center": (cx, cy), "seed": seed, "width": width, "height": height}
    return sim


def block_strips(sim, world_size):
    """Row ranges [(lo, hi)] that split the block's rows evenly (the last strip also owns the empty
    rows above the block).  The starting split; sph_set_rebalance moves it with the fluid."""
    sc = sim.scene
    gx, gy = sim.grid_dims()
    rows = min(gy, int(np.ceil((sc["ny"] * sc["spacing"] + 0.1) / KERNEL_HEIGHT)) + 1)
    return strips.split_rows(rows, gy, world_size)


def fill_block(sim):
    sc = sim.scene
    sim.AddVolumeHashed(sc["center"], (0.0, 0.0), sc["nx"], sc["ny"], sc["spacing"], sc["seed"])
    return sim


def bodies_scene(nx, spacing=0.05, seed=1337, **kw):
    """BASELINE.json configs[4]: a dense block (spacing h/6, the reference's scene-0
    regime) with 10x viscosity next to rigid bodies a la "Fun" (sph.h:420-436): three circles along its
    free side, a tilted box above it and a box resting on the floor.  The block starts compressed
    (rho ~ 27 > rho0), expands into the bodies within a few dozen steps and piles up against them, which
    skews the candidate counts."""
    sim = block_scene(nx, spacing=spacing, seed=seed, linear_viscosity=5.0, quadratic_viscosity=3.0, **kw)
    w, h = sim.scene["width"], sim.scene["height"]
    s = nx * spacing  # block edge
    x0, y0 = -w * 0.5 + 0.05, -h * 0.5 + 0.05  # lower-left corner of the block
    r = 0.08 * s
    for k in range(3):
        sim.AddCircle((x0 + s + r + 0.6, y0 + (0.15 + 0.3 * k) * s), r)

    def box(cx, cy, ang_deg, ex, ey):
        a = np.deg2rad(ang_deg)
        c, sn = np.cos(a), np.sin(a)
        local = np.array([(ex, ey), (-ex, ey), (-ex, -ey), (ex, -ey)])
        verts = np.stack([c * local[:, 0] - sn * local[:, 1] + cx, sn * local[:, 0] + c * local[:, 1] + cy], 1)
        sim.AddPolygon(verts.astype(np.float32))

    box(x0 + 0.5 * s, y0 + s + 0.04 * s + 0.8, -2.5, 0.45 * s, 0.02 * s)  # lid, tilted like the reference's ramps
    box(x0 + 1.5 * s, y0 + 0.05 * s, 0.0, 0.03 * s, 0.05 * s)             # post on the floor, cf. sph.h:434
    return sim

"""ctypes binding of libsphb200.so — the C ABI declared in include/sphb200.h.

Loading fails loudly when the library is missing; nothing here computes anything on the CPU.
"""
import ctypes as C
import os

from . import build as _build

c_f = C.c_float
c_vp = C.c_void_p
c_u64 = C.c_uint64
c_sz = C.c_size_t
c_i32 = C.c_int32

SPH_OK = 0
SPH_FP_EXACT = 0
SPH_FP_FAST = 1
SPH_FLAG_PHASE_TIMING = 1
SPH_FLAG_NO_GRAPHS = 2
SPH_FLAG_SWEEP_TEAM = 4
SPH_FLAG_SWEEP_WARP = 8
SPH_FLAG_SWEEP_FLOW = 16
SPH_FLAG_EXCHANGE_NCCL = 32
SPH_SOLVER_COLORED_GS = 0
SPH_SOLVER_GATHER = 1
SPH_NUM_PHASES = 9
PHASE_NAMES = ("integrate", "viscosity", "predict_key", "scan", "reorder", "density", "delta", "collide_velocity", "exchange")

PASS_INTEGRATE, PASS_VISCOSITY, PASS_PREDICT, PASS_GRID, PASS_DENSITY, PASS_DELTA, PASS_COLLIDE, PASS_VELOCITY = range(1, 9)


class SphConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("domain_width", c_f),
        ("domain_height", c_f),
        ("cell_size", c_f),
        ("max_particles", c_u64),
        ("device", c_i32),
        ("fp_mode", c_i32),
        ("flags", C.c_uint32),
        ("relaxation", c_f),
        ("solver", c_i32),
        ("sweep_capacity", C.c_uint32),
        ("rank", c_i32),
        ("world_size", c_i32),
        ("halo_capacity", c_u64),
        ("halo_rows", c_i32),
        ("reserved0", c_i32),
    ]


class SphParams(C.Structure):
    _fields_ = [(n, c_f) for n in (
        "kernel_height", "cell_size", "particle_spacing", "inv_kernel_height", "rest_density",
        "stiffness", "near_stiffness", "linear_viscosity", "quadratic_viscosity")]


class SphStats(C.Structure):
    _fields_ = [
        ("min_particle_neighbor_count", c_u64), ("max_particle_neighbor_count", c_u64),
        ("min_cell_particle_count", c_u64), ("max_cell_particle_count", c_u64),
        ("time_emitters", c_f), ("time_integration", c_f), ("time_viscosity_forces", c_f),
        ("time_predict", c_f), ("time_update_grid", c_f), ("time_neighbor_search", c_f),
        ("time_density_and_pressure", c_f), ("time_delta_positions", c_f), ("time_collisions", c_f),
        ("steps", c_u64), ("pair_candidates", c_u64),
    ]


# name -> (restype, argtypes); mirrors include/sphb200.h one to one
SIGNATURES = {
    "sph_abi_version": (C.c_int, []),
    "sph_config_default": (C.c_int, [C.POINTER(SphConfig)]),
    "sph_create": (C.c_int, [C.POINTER(SphConfig), C.POINTER(c_vp)]),
    "sph_destroy": (C.c_int, [c_vp]),
    "sph_last_error": (C.c_int, [c_vp, C.c_char_p, c_sz]),
    "sph_set_params": (C.c_int, [c_vp, C.POINTER(SphParams)]),
    "sph_get_params": (C.c_int, [c_vp, C.POINTER(SphParams)]),
    "sph_set_gravity": (C.c_int, [c_vp, c_f, c_f]),
    "sph_add_external_force": (C.c_int, [c_vp, c_f, c_f]),
    "sph_clear_external_force": (C.c_int, [c_vp]),
    "sph_set_relaxation": (C.c_int, [c_vp, c_f]),
    "sph_grid_dims": (C.c_int, [c_vp, C.POINTER(c_i32), C.POINTER(c_i32)]),
    "sph_clear_bodies": (C.c_int, [c_vp]),
    "sph_add_plane": (C.c_int, [c_vp, c_f, c_f, c_f]),
    "sph_add_circle": (C.c_int, [c_vp, c_f, c_f, c_f]),
    "sph_add_segment": (C.c_int, [c_vp, c_f, c_f, c_f, c_f]),
    "sph_add_polygon": (C.c_int, [c_vp, c_sz, c_vp]),
    "sph_body_count": (C.c_int, [c_vp, C.POINTER(c_sz)]),
    "sph_clear_particles": (C.c_int, [c_vp]),
    "sph_clear_emitters": (C.c_int, [c_vp]),
    "sph_add_particles": (C.c_int, [c_vp, c_sz, c_vp, c_vp, C.POINTER(c_u64)]),
    "sph_add_volume": (C.c_int, [c_vp, c_f, c_f, c_f, c_f, C.c_int, C.c_int, c_f]),
    "sph_add_volume_hashed": (C.c_int, [c_vp, c_f, c_f, c_f, c_f, C.c_int64, C.c_int64, c_f, c_u64]),
    "sph_add_emitter": (C.c_int, [c_vp, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f]),
    "sph_scenario_count": (C.c_int, []),
    "sph_scenario_name": (C.c_char_p, [C.c_int]),
    "sph_load_scenario": (C.c_int, [c_vp, C.c_int, C.c_int]),
    "sph_particle_count": (C.c_int, [c_vp, C.POINTER(c_u64)]),
    "sph_local_particle_count": (C.c_int, [c_vp, C.POINTER(c_u64)]),
    "sph_step": (C.c_int, [c_vp, c_f]),
    "sph_sync": (C.c_int, [c_vp]),
    "sph_run_pass": (C.c_int, [c_vp, C.c_int, c_f]),
    "sph_reset_stats": (C.c_int, [c_vp]),
    "sph_get_stats": (C.c_int, [c_vp, C.POINTER(SphStats)]),
    "sph_read_particles": (C.c_int, [c_vp, c_vp, c_sz]),
    "sph_write_particles": (C.c_int, [c_vp, c_vp, c_sz]),
    "sph_render_particles": (C.c_int, [c_vp, c_vp, c_sz, c_vp, c_sz]),
    "sph_wait_render": (C.c_int, [c_vp]),
    "sph_read_cell_counts": (C.c_int, [c_vp, c_vp]),
    "sph_read_cell_of_particle": (C.c_int, [c_vp, c_vp]),
    "sph_read_sorted_ids": (C.c_int, [c_vp, c_vp]),
    "sph_read_cell_start": (C.c_int, [c_vp, c_vp]),
    "sph_host_alloc": (C.c_int, [C.POINTER(c_vp), c_sz]),
    "sph_host_free": (C.c_int, [c_vp]),
    "sph_get_stream": (C.c_int, [c_vp, C.POINTER(c_vp)]),
    "sph_mark": (C.c_int, [c_vp, C.c_int]),
    "sph_elapsed_ms": (C.c_int, [c_vp, C.c_int, C.c_int, C.POINTER(c_f)]),
    "sph_get_phase_ms": (C.c_int, [c_vp, C.POINTER(c_f * SPH_NUM_PHASES), C.POINTER(c_u64)]),
    "sph_comm_unique_id": (C.c_int, [c_vp]),
    "sph_comm_init": (C.c_int, [c_vp, c_vp]),
    "sph_comm_init_local": (C.c_int, [C.POINTER(c_vp), c_i32]),
    "sph_step_group": (C.c_int, [C.POINTER(c_vp), c_i32, c_f]),
    "sph_set_strip": (C.c_int, [c_vp, c_i32, c_i32]),
    "sph_get_strip": (C.c_int, [c_vp, C.POINTER(c_i32), C.POINTER(c_i32)]),
    "sph_set_rebalance": (C.c_int, [c_vp, C.c_int32, C.c_int32]),
    "sph_plan_strip_bounds": (C.c_int, [c_vp, C.c_int32, c_vp, C.c_int32, C.c_int32, C.c_int32, c_vp]),
    "sph_read_owned": (C.c_int, [c_vp, c_vp, c_vp, c_sz, c_vp, c_sz, c_vp, c_sz, C.POINTER(c_u64)]),
    "sph_render_owned": (C.c_int, [c_vp, c_vp, c_vp, c_sz, c_vp, c_sz]),
    "sph_wait_render_owned": (C.c_int, [c_vp, C.POINTER(c_u64)]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Returns the loaded library; raises if it is not built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m nbodysimulation_experiment_b200.build` "
            "(needs nvcc). This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = symbol missing from the .so
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

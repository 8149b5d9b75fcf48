"""B200-native hot path of f1nalspace/nbodysimulation_experiment's Demo 4 SPH solver.

Only what the path needs lives here: `csrc/` (hand-written sm_100a CUDA + the C ABI of
include/sphb200.h, built into libsphb200.so) and the host-side mirror of the reference's
`BaseSimulation` interface (`ParticleSimulation`).  There is no CPU or PyTorch fallback.
"""
from . import _lib
from ._lib import (SPH_FLAG_PHASE_TIMING, SPH_FP_EXACT, SPH_FP_FAST, SPH_SOLVER_COLORED_GS, SPH_SOLVER_GATHER, SphConfig,
                   SphParams, SphStats)
from .scenes import block_scene, bodies_scene
from .simulation import ParticleSimulation, SphError, StripGroup, bind_host_to_gpu, pinned_empty

__all__ = [
    "ParticleSimulation", "StripGroup", "bind_host_to_gpu", "SphError", "SphConfig", "SphParams", "SphStats", "pinned_empty",
    "SPH_FP_EXACT", "SPH_FP_FAST", "SPH_FLAG_PHASE_TIMING", "SPH_SOLVER_COLORED_GS", "SPH_SOLVER_GATHER", "block_scene", "bodies_scene",
]

"""Host-side mirror of the reference's plugin interface for the SPH hot path.

`ParticleSimulation` keeps the method names, argument meaning and call order of
`class BaseSimulation` (/root/reference/NBodySimulation/base.h:8-39) as implemented by
`Demo4::ParticleSimulation` (demo4.h:137-226), so tests read like calls into the reference.  Every
method forwards to one entry point of the C ABI (include/sphb200.h); no arithmetic of the hot path
happens in Python.  (The C++ adapter a maintainer would add to the reference itself is in
INTEGRATION.md.)
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import SphConfig, SphParams, SphStats

PARAM_FIELDS = [n for n, _ in SphParams._fields_]


class SphError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"sphb200 error {code}: {text}")
        self.code = code


class ParticleSimulation:
    """One simulation in HBM.  Defaults are the reference's compile-time world (sph.h:18-72)."""

    def __init__(self, domain_width=None, domain_height=None, cell_size=None, max_particles=None, device=0,
                 fp_mode=_lib.SPH_FP_EXACT, flags=0, relaxation=1.0, rank=0, world_size=1, halo_capacity=0,
                 solver=_lib.SPH_SOLVER_COLORED_GS, sweep_capacity=0, halo_rows=0):
        self._lib = _lib.load()
        cfg = SphConfig()
        self._check(self._lib.sph_config_default(C.byref(cfg)), None)
        if domain_width is not None:
            cfg.domain_width = domain_width
        if domain_height is not None:
            cfg.domain_height = domain_height
        if cell_size is not None:
            cfg.cell_size = cell_size
        if max_particles is not None:
            cfg.max_particles = int(max_particles)
        cfg.device = device
        cfg.fp_mode = fp_mode
        cfg.flags = flags
        cfg.relaxation = relaxation
        cfg.solver = solver
        cfg.sweep_capacity = sweep_capacity
        cfg.rank = rank
        cfg.world_size = world_size
        cfg.halo_capacity = halo_capacity
        cfg.halo_rows = halo_rows
        self.config = cfg
        self._h = _lib.c_vp()
        self._check(self._lib.sph_create(C.byref(cfg), C.byref(self._h)), None)
        self._multithreading = True
        self.bodies = []

    # -- plumbing ------------------------------------------------------------------------
    def _check(self, rc, handle="self"):
        if rc == _lib.SPH_OK:
            return
        buf = C.create_string_buffer(512)
        self._lib.sph_last_error(self._h if handle == "self" else None, buf, 512)
        raise SphError(rc, buf.value.decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- BaseSimulation surface (base.h:10-38), same names ----------------------------------
    def ResetStats(self):
        self._check(self._lib.sph_reset_stats(self._h))

    def ClearBodies(self):
        self._check(self._lib.sph_clear_bodies(self._h))
        self.bodies = []

    def ClearParticles(self):
        self._check(self._lib.sph_clear_particles(self._h))

    def ClearEmitters(self):
        self._check(self._lib.sph_clear_emitters(self._h))

    # (`self.bodies` mirrors what was added through this object, in insertion order, as float32 values: tests hand the
    # same bodies to the CPU oracle; LoadScenario adds its bodies inside the library and does not record them)
    def AddPlane(self, normal, distance):
        self._check(self._lib.sph_add_plane(self._h, normal[0], normal[1], distance))
        self.bodies.append(("plane", float(np.float32(normal[0])), float(np.float32(normal[1])), float(np.float32(distance))))

    def AddCircle(self, pos, radius):
        self._check(self._lib.sph_add_circle(self._h, pos[0], pos[1], radius))
        self.bodies.append(("circle", float(np.float32(pos[0])), float(np.float32(pos[1])), float(np.float32(radius))))

    def AddLineSegment(self, a, b):
        self._check(self._lib.sph_add_segment(self._h, a[0], a[1], b[0], b[1]))
        self.bodies.append(("segment", float(np.float32(a[0])), float(np.float32(a[1])), float(np.float32(b[0])), float(np.float32(b[1]))))

    def AddPolygon(self, verts):
        v = np.ascontiguousarray(verts, np.float32).reshape(-1)
        self._check(self._lib.sph_add_polygon(self._h, len(v) // 2, v.ctypes.data))
        self.bodies.append(("polygon", v.copy()))

    def AddParticle(self, position, force=(0.0, 0.0)):
        return self.AddParticles(np.array([position], np.float32), np.array([force], np.float32))

    def AddParticles(self, positions, forces=None):
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 2)
        f = None if forces is None else np.ascontiguousarray(forces, np.float32).reshape(-1, 2)
        first = C.c_uint64()
        self._check(self._lib.sph_add_particles(self._h, len(p), p.ctypes.data, None if f is None else f.ctypes.data, C.byref(first)))
        return first.value

    def AddVolume(self, center, force, countX, countY, spacing):
        self._check(self._lib.sph_add_volume(self._h, center[0], center[1], force[0], force[1], countX, countY, spacing))

    def AddVolumeHashed(self, center, force, countX, countY, spacing, seed=1337):
        self._check(self._lib.sph_add_volume_hashed(self._h, center[0], center[1], force[0], force[1], countX, countY, spacing, seed))

    def AddEmitter(self, position, direction, radius, speed, rate, duration):
        self._check(self._lib.sph_add_emitter(self._h, position[0], position[1], direction[0], direction[1], radius, speed, rate, duration))

    def Update(self, deltaTime):
        self._check(self._lib.sph_step(self._h, deltaTime))

    def Render(self, positions=None, colors=None, wait=True):
        """The particle section of Render() (demo4.cpp:520-531): positions + colours, creation order.
        wait=False returns once the copy is enqueued (pinned buffers): it overlaps the next Update and
        is complete after WaitRender()."""
        n = self.GetParticleCount()
        if positions is None:
            positions = np.empty((n, 2), np.float32)
        if colors is None:
            colors = np.empty((n, 4), np.float32)
        self._check(self._lib.sph_render_particles(self._h, positions.ctypes.data, positions.strides[0], colors.ctypes.data, colors.strides[0]))
        if wait:
            self.WaitRender()
        return positions, colors

    def WaitRender(self):
        self._check(self._lib.sph_wait_render(self._h))

    def AddExternalForces(self, force):
        self._check(self._lib.sph_add_external_force(self._h, force[0], force[1]))

    def ClearExternalForce(self):
        self._check(self._lib.sph_clear_external_force(self._h))

    def GetParticleCount(self):
        out = C.c_uint64()
        self._check(self._lib.sph_particle_count(self._h, C.byref(out)))
        return out.value

    def SetGravity(self, gravity):
        self._check(self._lib.sph_set_gravity(self._h, gravity[0], gravity[1]))

    def GetParams(self):
        p = SphParams()
        self._check(self._lib.sph_get_params(self._h, C.byref(p)))
        return p

    def SetParams(self, params):
        if not isinstance(params, SphParams):
            arr = np.asarray(params, np.float32)
            p = SphParams(*[float(x) for x in arr])
        else:
            p = params
        self._check(self._lib.sph_set_params(self._h, C.byref(p)))

    def GetStats(self):
        st = SphStats()
        self._check(self._lib.sph_get_stats(self._h, C.byref(st)))
        return st

    def SetMultiThreading(self, value):  # the GPU path has one mode; kept for interface parity
        self._multithreading = bool(value)

    def IsMultiThreadingSupported(self):
        return True

    def IsMultiThreading(self):
        return self._multithreading

    def GetWorkerThreadCount(self):
        return 148  # one worker per SM of the B200 the kernels are sized for

    # -- beyond base.h: scenes, passes, readback ---------------------------------------------
    def LoadScenario(self, index, seed=-1):
        """DemoApplication::LoadScenario (app.cpp:477-534) for SPHScenarios[index] (sph.h:315-437)."""
        self._check(self._lib.sph_load_scenario(self._h, index, seed))

    def SetRelaxation(self, omega):
        self._check(self._lib.sph_set_relaxation(self._h, omega))

    def RunPass(self, which, dt=1.0 / 60.0):
        self._check(self._lib.sph_run_pass(self._h, which, dt))

    def Sync(self):
        self._check(self._lib.sph_sync(self._h))

    # -- multi-GPU plumbing (one process per GPU) ---------------------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id; create on rank 0 and broadcast (e.g. with torch.distributed)."""
        buf = (C.c_uint8 * 128)()
        rc = _lib.load().sph_comm_unique_id(C.cast(buf, _lib.c_vp))
        if rc != 0:
            raise SphError(rc, "sph_comm_unique_id failed (is libnccl.so.2 loadable?)")
        return bytes(buf)

    def comm_init(self, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._lib.sph_comm_init(self._h, C.cast(buf, _lib.c_vp)))

    def set_strip(self, row_begin, row_end):
        self._check(self._lib.sph_set_strip(self._h, row_begin, row_end))

    def set_rebalance(self, every_steps, max_shift_rows=2):
        """Re-balance the strips every `every_steps` steps (0 = off); call before adding particles."""
        self._check(self._lib.sph_set_rebalance(self._h, int(every_steps), int(max_shift_rows)))

    def get_strip(self):
        a, b = C.c_int32(), C.c_int32()
        self._check(self._lib.sph_get_strip(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def local_particle_count(self):
        out = C.c_uint64()
        self._check(self._lib.sph_local_particle_count(self._h, C.byref(out)))
        return out.value

    def read_owned(self, records=True, render=False, capacity=None, buffers=None):
        """ids (+ ParticleData records, + positions/colours) of the particles this rank owns.
        `buffers` = dict from owned_buffers() to reuse (pinned) host memory across calls."""
        if buffers is None:
            buffers = self.owned_buffers(records=records, render=render, capacity=capacity, pinned=False)
        ids, rec, pos, col = buffers["ids"], buffers.get("records"), buffers.get("positions"), buffers.get("colors")
        n = C.c_uint64()
        self._check(self._lib.sph_read_owned(self._h, ids.ctypes.data, rec.ctypes.data if (records and rec is not None) else None, 48,
                                             pos.ctypes.data if (render and pos is not None) else None, 8,
                                             col.ctypes.data if (render and col is not None) else None, 16, C.byref(n)))
        k = n.value
        out = {"ids": ids[:k]}
        if records and rec is not None:
            out["records"] = rec[:k]
        if render and pos is not None:
            out["positions"], out["colors"] = pos[:k], col[:k]
        return out

    def render_owned(self, buffers):
        """Overlapped strip readback (sph_render_owned): enqueue the snapshot and its copies into `buffers`
        (owned_buffers(render=True, pinned=True)); the arrays are complete after wait_render_owned()."""
        ids, pos, col = buffers["ids"], buffers["positions"], buffers["colors"]
        self._check(self._lib.sph_render_owned(self._h, ids.ctypes.data, pos.ctypes.data, pos.strides[0], col.ctypes.data, col.strides[0]))
        self._owned_in_flight = buffers

    def wait_render_owned(self):
        """-> dict(ids, positions, colors) views of the frame enqueued by render_owned, or None if there is none."""
        buffers = getattr(self, "_owned_in_flight", None)
        if buffers is None:
            return None
        n = C.c_uint64()
        self._check(self._lib.sph_wait_render_owned(self._h, C.byref(n)))
        self._owned_in_flight = None
        k = n.value
        return {"ids": buffers["ids"][:k], "positions": buffers["positions"][:k], "colors": buffers["colors"][:k]}

    def owned_buffers(self, records=True, render=False, capacity=None, pinned=True):
        """Host arrays for read_owned (page-locked when pinned=True; keep the dict alive while in use)."""
        cap = int(capacity or self.config.max_particles)
        make = (lambda shape, dt: pinned_empty(shape, dt)) if pinned else (lambda shape, dt: (np.empty(shape, dt), None))
        out, owners = {}, []
        for name, shape, dt, want in (("ids", (cap,), np.uint32, True), ("records", (cap, 12), np.float32, records),
                                      ("positions", (cap, 2), np.float32, render), ("colors", (cap, 4), np.float32, render)):
            if want:
                arr, own = make(shape, dt)
                out[name] = arr
                owners.append(own)
        out["_owners"] = owners
        return out

    def grid_dims(self):
        gx, gy = C.c_int32(), C.c_int32()
        self._check(self._lib.sph_grid_dims(self._h, C.byref(gx), C.byref(gy)))
        return gx.value, gy.value

    def params_array(self):
        p = self.GetParams()
        return np.array([getattr(p, n) for n in PARAM_FIELDS], np.float32)

    def particles(self):
        """(n, 12) float32 in ParticleData order (demo4.h:81-99), creation order."""
        n = self.GetParticleCount()
        out = np.zeros((n, 12), np.float32)
        if n:
            self._check(self._lib.sph_read_particles(self._h, out.ctypes.data, 48))
        return out

    def put_particles(self, arr):
        arr = np.ascontiguousarray(arr, np.float32)
        assert arr.shape == (self.GetParticleCount(), 12)
        self._check(self._lib.sph_write_particles(self._h, arr.ctypes.data, 48))

    def cell_counts(self):
        gx, gy = self.grid_dims()
        out = np.zeros(gx * gy, np.uint32)
        self._check(self._lib.sph_read_cell_counts(self._h, out.ctypes.data))
        return out

    def cell_start(self):
        gx, gy = self.grid_dims()
        out = np.zeros(gx * gy + 1, np.uint32)
        self._check(self._lib.sph_read_cell_start(self._h, out.ctypes.data))
        return out

    def sorted_ids(self):
        out = np.zeros(self.GetParticleCount(), np.uint32)
        self._check(self._lib.sph_read_sorted_ids(self._h, out.ctypes.data))
        return out

    def cell_of_particle(self):
        out = np.zeros((self.GetParticleCount(), 2), np.int32)
        self._check(self._lib.sph_read_cell_of_particle(self._h, out.ctypes.data))
        return out

    def phase_ms(self):
        arr = (C.c_float * _lib.SPH_NUM_PHASES)()
        steps = C.c_uint64()
        self._check(self._lib.sph_get_phase_ms(self._h, C.byref(arr), C.byref(steps)))
        return dict(zip(_lib.PHASE_NAMES, [float(x) for x in arr])), steps.value

    def mark(self, slot):
        self._check(self._lib.sph_mark(self._h, slot))

    def elapsed_ms(self, a, b):
        ms = C.c_float()
        self._check(self._lib.sph_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def stream_ptr(self):
        out = _lib.c_vp()
        self._check(self._lib.sph_get_stream(self._h, C.byref(out)))
        return out.value


def pinned_empty(shape, dtype=np.float32):
    """numpy view of page-locked host memory from sph_host_alloc (freed with the returned owner)."""
    lib = _lib.load()
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = _lib.c_vp()
    if lib.sph_host_alloc(C.byref(ptr), max(nbytes, 1)) != 0:
        raise SphError(-3, "sph_host_alloc failed")
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, p):
            self.p = p

        def free(self):
            if self.p:
                lib.sph_host_free(self.p)
                self.p = None

        def __del__(self):  # page-locked memory must not outlive its last reference
            try:
                self.free()
            except Exception:
                pass

    return arr, _Owner(ptr)


class StripGroup:
    """All y-strips of one simulation inside ONE process (sph_comm_init_local / sph_step_group): one host thread
    drives several GPUs the way the reference's single-threaded app loop would (app.cpp:228-236), or - for tests -
    several strips share one GPU.  `sims` are ParticleSimulation objects created with rank = 0..n-1, world_size = n."""

    def __init__(self, sims):
        self.sims = list(sims)
        self._lib = _lib.load()
        self._arr = (_lib.c_vp * len(self.sims))(*[s._h for s in self.sims])
        rc = self._lib.sph_comm_init_local(self._arr, len(self.sims))
        if rc != _lib.SPH_OK:
            self._raise(rc)

    def _raise(self, rc):
        buf = C.create_string_buffer(512)
        for s in self.sims:  # the failing handle carries the text
            self._lib.sph_last_error(s._h, buf, 512)
            if buf.value:
                break
        else:
            self._lib.sph_last_error(None, buf, 512)
        raise SphError(rc, buf.value.decode(errors="replace"))

    def Update(self, deltaTime):
        rc = self._lib.sph_step_group(self._arr, len(self.sims), deltaTime)
        if rc != _lib.SPH_OK:
            self._raise(rc)

    def Sync(self):
        for s in self.sims:
            s.Sync()

    def close(self):
        for s in self.sims:
            s.close()


def bind_host_to_gpu(device):
    """Pins the calling process to the CPU cores next to GPU `device` (NVML's ideal affinity), so that page-locked
    readback buffers allocated afterwards land on that GPU's NUMA node.  Host-side placement only; returns True
    if the affinity was applied."""
    try:
        import os

        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = int(vis.split(",")[device]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else device
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx))
        return True
    except Exception:
        return False

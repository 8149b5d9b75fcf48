"""ctypes bindings for the two CHECKER libraries under oracle/_ref/ (test infrastructure only).

* ``libsphoracle.so`` - the in-repo CPU restatement (oracle/sph_oracle.c), prefix ``oracle_``.
* ``libsphref.so``    - the unmodified reference solver compiled headless (oracle/ref_harness),
  prefix ``ref_``.  Only present where /root/reference was available at build time.

Both expose the same verbs, so `CpuSim` wraps either one behind one Python class.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_ref", "libsphoracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsphref.so")

MODE_GS_INDEX = 0
MODE_JACOBI = 1
MODE_COLORED = 2
MODE_HYBRID = 3

f32 = C.c_float
vp = C.c_void_p


def build_oracle():
    """Compile the checker (never the product).  Safe to call repeatedly."""
    # make's chatter goes to stderr: bench.py promises exactly one JSON line on stdout
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle", "ref"], check=True, stdout=sys.stderr)


def _load(path):
    if not os.path.exists(path):
        build_oracle()
    return C.CDLL(path)


def have_ref():
    return os.path.exists(REF_SO)


_sigs_common = {
    # name: (restype, argtypes) with the handle first where applicable
    "reset_stats": (None, [vp]),
    "clear_bodies": (None, [vp]),
    "clear_particles": (None, [vp]),
    "clear_emitters": (None, [vp]),
    "add_plane": (None, [vp, f32, f32, f32]),
    "add_circle": (None, [vp, f32, f32, f32]),
    "add_segment": (None, [vp, f32, f32, f32, f32]),
    "add_polygon": (None, [vp, C.c_int, vp]),
    "add_particle": (C.c_uint64, [vp, f32, f32, f32, f32]),
    "add_volume": (None, [vp, f32, f32, f32, f32, C.c_int, C.c_int, f32]),
    "add_emitter": (None, [vp, f32, f32, f32, f32, f32, f32, f32, f32]),
    "set_gravity": (None, [vp, f32, f32]),
    "add_external_force": (None, [vp, f32, f32]),
    "clear_external_force": (None, [vp]),
    "particle_count": (C.c_uint64, [vp]),
    "get_params": (None, [vp, vp]),
    "set_params": (None, [vp, vp]),
    "load_scenario": (None, [vp, C.c_int, C.c_int]),
    "get_particles": (None, [vp, vp]),
    "set_particles": (None, [vp, vp]),
    "get_cell_of_particle": (None, [vp, vp]),
    "get_cell_counts": (None, [vp, vp]),
    "get_cell_members": (C.c_uint32, [vp, C.c_int, vp]),
    "get_neighbor_counts": (None, [vp, vp]),
    "get_neighbors": (C.c_uint32, [vp, C.c_uint64, vp]),
    "get_stats": (None, [vp, vp, vp]),
    "get_colors": (None, [vp, vp]),
    "body_count": (C.c_int, [vp]),
    "get_body": (None, [vp, C.c_int, vp, vp, vp]),
    "get_gravity": (None, [vp, vp]),
}


class CpuSim:
    """One simulation inside either checker library.

    kind="oracle": in-repo restatement (any domain size); kind="ref": the reference's own code
    (fixed 10 x 5.625 domain, <= 10 000 particles).
    """

    def __init__(self, kind="oracle", width=10.0, height=5.625, cell=None, mode=MODE_GS_INDEX, threads=1):
        self.kind = kind
        if kind == "oracle":
            self.lib = _load(ORACLE_SO)
            self.pre = "oracle_"
            if cell is None:
                cell = float(np.float32(6.0) * np.float32(0.05))
            self.lib.oracle_create.restype = vp
            self.lib.oracle_create.argtypes = [f32, f32, f32]
            self.h = vp(self.lib.oracle_create(width, height, cell))
            self._bind("set_mode", None, [vp, C.c_int])
            self._bind("set_threads", None, [vp, C.c_int])
            self._bind("set_relaxation", None, [vp, f32])
            self._bind("step", None, [vp, f32])
            self._bind("step_timed", C.c_double, [vp, f32, C.c_int])
            self._bind("grid_dims", None, [vp, vp])
            self._bind("add_particles", C.c_uint64, [vp, C.c_uint64, vp, vp])
            for p in ("update_grid", "neighbor_search", "density", "collide"):
                self._bind("pass_" + p, None, [vp])
            for p in ("viscosity", "delta"):
                self._bind("pass_" + p, None, [vp, f32])
            self.set_mode(mode)
            self.set_threads(threads)
        elif kind == "ref":
            self.lib = _load(REF_SO)
            self.pre = "ref_"
            self.lib.ref_create.restype = vp
            self.h = vp(self.lib.ref_create())
            self._bind("set_multithreading", None, [vp, C.c_int])
            self._bind("update", None, [vp, f32])
            self._bind("update_timed", C.c_double, [vp, f32, C.c_int])
            self._bind("worker_threads", C.c_int, [vp])
            for p in ("neighbor_search", "density_pressure", "viscosity", "delta_positions"):
                self._bind(p, None, [vp, f32])
            self.lib.ref_grid_dims.argtypes = [vp]
            self.set_multithreading(1 if threads > 1 else 0)
        else:
            raise ValueError(kind)
        for name, (res, args) in _sigs_common.items():
            self._bind(name, res, args)

    def _bind(self, name, res, args):
        fn = getattr(self.lib, self.pre + name)
        fn.restype = res
        fn.argtypes = args
        h = self.h

        def call(*a, _fn=fn):
            return _fn(h, *a)

        setattr(self, name, call)

    def close(self):
        if self.h:
            getattr(self.lib, self.pre + "destroy").argtypes = [vp]
            getattr(self.lib, self.pre + "destroy")(self.h)
            self.h = None

    # -- uniform verbs -----------------------------------------------------------------
    @property
    def n(self):
        return int(self.particle_count())

    def dims(self):
        out = np.zeros(2, np.int32)
        if self.kind == "oracle":
            self.grid_dims(out.ctypes.data)
        else:
            self.lib.ref_grid_dims(out.ctypes.data)
        return int(out[0]), int(out[1])

    def advance(self, dt, steps=1):
        dt = float(np.float32(dt))
        for _ in range(steps):
            (self.step if self.kind == "oracle" else self.update)(dt)

    def advance_timed(self, dt, steps):
        dt = float(np.float32(dt))
        return (self.step_timed if self.kind == "oracle" else self.update_timed)(dt, steps)

    def particles(self):
        """(n, 12) float32: cur, prev, acc, vel, rho, rhoNear, P, PNear (demo4.h:81-99)."""
        out = np.zeros((self.n, 12), np.float32)
        if self.n:
            self.get_particles(out.ctypes.data)
        return out

    def put_particles(self, arr):
        arr = np.ascontiguousarray(arr, np.float32)
        assert arr.shape == (self.n, 12)
        self.set_particles(arr.ctypes.data)

    def params(self):
        out = np.zeros(9, np.float32)
        self.get_params(out.ctypes.data)
        return out

    def put_params(self, p9):
        p9 = np.ascontiguousarray(p9, np.float32)
        self.set_params(p9.ctypes.data)

    def gravity(self):
        out = np.zeros(2, np.float32)
        self.get_gravity(out.ctypes.data)
        return out

    def cell_of_particle(self):
        out = np.zeros((self.n, 2), np.int32)
        if self.n:
            self.get_cell_of_particle(out.ctypes.data)
        return out

    def cell_counts(self):
        gx, gy = self.dims()
        out = np.zeros(gx * gy, np.uint32)
        self.get_cell_counts(out.ctypes.data)
        return out

    def cell_members(self, cell):
        buf = np.zeros(max(self.n, 1), np.uint32)
        k = self.get_cell_members(int(cell), buf.ctypes.data)
        return buf[:k].copy()

    def neighbor_counts(self):
        out = np.zeros(self.n, np.uint32)
        if self.n:
            self.get_neighbor_counts(out.ctypes.data)
        return out

    def neighbors(self, i):
        buf = np.zeros(max(self.n, 1), np.uint32)
        k = self.get_neighbors(int(i), buf.ctypes.data)
        return buf[:k].copy()

    def stats(self):
        c = np.zeros(4, np.uint64)
        t = np.zeros(9, np.float32)
        self.get_stats(c.ctypes.data, t.ctypes.data)
        return c, t

    def colors(self):
        out = np.zeros((self.n, 4), np.float32)
        if self.n:
            self.get_colors(out.ctypes.data)
        return out

    def bodies(self):
        res = []
        for i in range(self.body_count()):
            t = C.c_int32()
            nv = C.c_int32()
            f = np.zeros(16, np.float32)
            self.get_body(i, C.addressof(t), C.addressof(nv), f.ctypes.data)
            res.append((t.value, nv.value, f))
        return res

    def add_particle_array(self, pos_xy, acc_xy=None):
        """bulk AddParticle (oracle only): creation order = row order"""
        pos = np.ascontiguousarray(pos_xy, np.float32).reshape(-1, 2)
        acc = None if acc_xy is None else np.ascontiguousarray(acc_xy, np.float32).reshape(-1, 2)
        return self.add_particles(len(pos), pos.ctypes.data, None if acc is None else acc.ctypes.data)

    def add_bodies(self, bodies):
        """bodies as recorded by nbodysimulation_experiment_b200.ParticleSimulation.bodies, in insertion order"""
        for b in bodies:
            if b[0] == "plane":
                self.add_plane(b[1], b[2], b[3])
            elif b[0] == "circle":
                self.add_circle(b[1], b[2], b[3])
            elif b[0] == "segment":
                self.add_segment(b[1], b[2], b[3], b[4])
            else:
                self.polygon(b[1])

    def polygon(self, xy):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1)
        self.add_polygon(len(xy) // 2, xy.ctypes.data)


def point_solvers(kind):
    """The four collision solvers of sph.h:514-681 applied to one point."""
    lib = _load(ORACLE_SO if kind == "oracle" else REF_SO)
    pre = "oracle_" if kind == "oracle" else "ref_"

    def mk(name, args):
        fn = getattr(lib, pre + name)
        fn.restype = None
        fn.argtypes = [vp] + args
        return fn

    plane = mk("solve_plane", [f32, f32, f32])
    circle = mk("solve_circle", [f32, f32, f32])
    segment = mk("solve_segment", [f32, f32, f32, f32])
    polygon = mk("solve_polygon", [C.c_int, vp])

    def run(fn, p, *args):
        buf = np.array(p, np.float32)
        fn(buf.ctypes.data, *args)
        return buf

    def poly(p, verts):
        v = np.ascontiguousarray(verts, np.float32).reshape(-1)
        buf = np.array(p, np.float32)
        polygon(buf.ctypes.data, len(v) // 2, v.ctypes.data)
        return buf

    return {
        "plane": lambda p, nx, ny, d: run(plane, p, nx, ny, d),
        "circle": lambda p, cx, cy, r: run(circle, p, cx, cy, r),
        "segment": lambda p, ax, ay, bx, by: run(segment, p, ax, ay, bx, by),
        "polygon": poly,
    }

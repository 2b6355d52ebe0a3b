"""ctypes wrapper of oracle/liboracle.so (the CPU restatement, oracle/dxmc_oracle.cpp). TEST INFRASTRUCTURE:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg only — never by dxmclib_b200/.

The restatement takes the same plain-data structs as the CUDA runtime (include/dxmcb200.h); the ctypes struct
definitions are shared with dxmclib_b200/cabi.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from dxmclib_b200 import cabi

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "liboracle.so")

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i64p = C.POINTER(C.c_int64)

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            raise FileNotFoundError(f"{ORACLE_LIB} is not built (make -C oracle oracle)")
        _lib = C.CDLL(ORACLE_LIB)
        _lib.dxmc_oracle_create.restype = C.c_void_p
        _lib.dxmc_oracle_destroy.argtypes = [C.c_void_p]
        _lib.dxmc_oracle_destroy.restype = None
        _lib.dxmc_oracle_history_stream.restype = None
    return _lib


def history_stream(seed, exposure, history):
    out = (C.c_uint64 * 2)()
    lib().dxmc_oracle_history_stream(C.c_uint64(seed), C.c_uint64(exposure), C.c_uint64(history), out)
    return int(out[0]), int(out[1])


def host_log10f(x):
    """std::log10(float) of the host libm, elementwise."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(x)
    lib().dxmc_oracle_log10f(C.c_uint64(x.size), x.ctypes.data_as(_f32p), out.ctypes.data_as(_f32p))
    return out


class Oracle:
    def __init__(self):
        self.l = lib()
        self.h = C.c_void_p(self.l.dxmc_oracle_create())
        self.n_voxels = 0

    def close(self):
        if self.h:
            self.l.dxmc_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _chk(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed with status {rc}")

    def load(self, flat: dict):
        """flat: the dictionary tests/support.flatten_scene builds (world, luts, beam tables)."""
        w = cabi.World()
        w.dim[:] = [int(x) for x in flat["dim"]]
        w.spacing[:] = [float(x) for x in flat["spacing"]]
        w.extent_safe[:] = [float(x) for x in flat["extent_safe"]]
        w.density = flat["density"].ctypes.data_as(_f32p)
        w.material = flat["material"].ctypes.data_as(_u8p)
        if flat.get("measurement") is not None:
            w.measurement = flat["measurement"].ctypes.data_as(_u8p)
        self._chk(self.l.dxmc_oracle_set_world(self.h, C.byref(w)), "dxmc_oracle_set_world")
        self.n_voxels = int(np.prod(flat["dim"]))
        l = cabi.Luts()
        lt = flat["luts"]
        l.n_materials, l.n_segments, l.linear_index = lt["n_materials"], lt["n_segments"], lt["linear_index"]
        l.linear_step, l.linear_energy = lt["linear_step"], lt["linear_energy"]
        for k in ("knots", "coefficients", "max_coefficients", "rita", "spline", "shells"):
            setattr(l, k, lt[k].ctypes.data_as(_f32p))
        self._chk(self.l.dxmc_oracle_set_luts(self.h, C.byref(l)), "dxmc_oracle_set_luts")
        sp, he, bo = flat["spectra"], flat["heels"], flat["bowties"]
        S = (cabi.Spectrum * max(len(sp), 1))()
        for i, (p, a, e) in enumerate(sp):
            S[i] = cabi.Spectrum(p.size, p.ctypes.data_as(_f32p), a.ctypes.data_as(_u32p), e.ctypes.data_as(_f32p))
        H = (cabi.Heel * max(len(he), 1))()
        for i, (e0, de, ne, a0, da, na, wts) in enumerate(he):
            H[i] = cabi.Heel(e0, de, ne, a0, da, na, wts.ctypes.data_as(_f32p))
        B = (cabi.Bowtie * max(len(bo), 1))()
        for i, (a, wts) in enumerate(bo):
            B[i] = cabi.Bowtie(a.size, a.ctypes.data_as(_f32p), wts.ctypes.data_as(_f32p))
        self._chk(self.l.dxmc_oracle_set_beam_tables(self.h, len(sp), S, len(he), H, len(bo), B), "dxmc_oracle_set_beam_tables")
        self._flat = flat  # keep the arrays alive

    def clear(self):
        self._chk(self.l.dxmc_oracle_clear(self.h), "dxmc_oracle_clear")

    def run(self, exposures, begin, end, model=1, seed=1, per_history_streams=True):
        arr = (cabi.Exposure * len(exposures))(*exposures)
        self._chk(self.l.dxmc_oracle_run(self.h, arr, C.c_uint64(begin), C.c_uint64(end), int(model), C.c_uint64(seed),
                                         int(per_history_streams)), "dxmc_oracle_run")

    def get_raw(self):
        n = self.n_voxels
        dose, ev, var = np.zeros(n, np.float32), np.zeros(n, np.uint32), np.zeros(n, np.float32)
        self._chk(self.l.dxmc_oracle_get_raw(self.h, dose.ctypes.data_as(_f32p), ev.ctypes.data_as(_u32p), var.ctypes.data_as(_f32p)),
                  "dxmc_oracle_get_raw")
        return dose, ev, var

    def set_fixed_point(self, energy_bits, energy_sq_bits):
        self._chk(self.l.dxmc_oracle_set_fixed_point(self.h, int(energy_bits), int(energy_sq_bits)), "dxmc_oracle_set_fixed_point")

    def get_fixed(self):
        n = self.n_voxels
        e, e2 = np.zeros(n, np.int64), np.zeros(n, np.uint64)
        self._chk(self.l.dxmc_oracle_get_fixed(self.h, e.ctypes.data_as(_i64p), e2.ctypes.data_as(C.POINTER(C.c_uint64))), "dxmc_oracle_get_fixed")
        return e, e2

    def set_tracking(self, tracking: int, brick_mm: float = 16.0):
        """0: the reference's Woodcock loop; 1: Woodcock + empty-space traversal through air bricks (call after load)."""
        self._chk(self.l.dxmc_oracle_set_tracking(self.h, int(tracking), C.c_float(brick_mm)), "dxmc_oracle_set_tracking")

    def bricks(self) -> dict:
        shift, nb, f_air = (C.c_uint32 * 3)(), (C.c_uint32 * 3)(), C.c_float()
        self._chk(self.l.dxmc_oracle_get_bricks(self.h, shift, nb, C.byref(f_air), None, None, None), "dxmc_oracle_get_bricks")
        n = int(nb[0]) * int(nb[1]) * int(nb[2])
        n_mat = int(self._flat["luts"]["n_materials"])
        ratio, bmax, air = np.zeros(n_mat, np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        self._chk(self.l.dxmc_oracle_get_bricks(self.h, shift, nb, C.byref(f_air), ratio.ctypes.data_as(_f32p), bmax.ctypes.data_as(_f32p),
                                                air.ctypes.data_as(_u8p)), "dxmc_oracle_get_bricks")
        dist = np.zeros(8 * n, np.uint8)
        self._chk(self.l.dxmc_oracle_get_brick_distance(self.h, dist.ctypes.data_as(_u8p)), "dxmc_oracle_get_brick_distance")
        return {"shift": list(shift), "nb": list(nb), "f_air": float(f_air.value), "ratio": ratio, "brick_max": bmax, "air": air, "distance": dist}

    def walk_stats(self):
        out = (C.c_uint64 * 3)()
        self._chk(self.l.dxmc_oracle_get_walk_stats(self.h, out), "dxmc_oracle_get_walk_stats")
        return [int(x) for x in out]

    def stats(self) -> dict:
        s = cabi.Stats()
        self._chk(self.l.dxmc_oracle_get_stats(self.h, C.byref(s)), "dxmc_oracle_get_stats")
        return {k: getattr(s, k) for k, _ in cabi.Stats._fields_}

    def eval_attenuation(self, material, energy):
        m = np.ascontiguousarray(material, np.uint8)
        e = np.ascontiguousarray(energy, np.float32)
        out = np.zeros((e.size, 3), np.float32)
        mx = np.zeros(e.size, np.float32)
        self._chk(self.l.dxmc_oracle_eval_attenuation(self.h, C.c_uint64(e.size), m.ctypes.data_as(_u8p), e.ctypes.data_as(_f32p),
                                                      out.ctypes.data_as(_f32p), mx.ctypes.data_as(_f32p)), "dxmc_oracle_eval_attenuation")
        return out, mx

    def trace_indices(self, pos, direction, steps):
        p = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(direction, np.float32).reshape(-1, 3)
        s = np.ascontiguousarray(steps, np.float32)
        idx = np.zeros((p.shape[0], s.size + 1), np.int64)
        entry = np.zeros((p.shape[0], 3), np.float32)
        self._chk(self.l.dxmc_oracle_trace_indices(self.h, C.c_uint64(p.shape[0]), p.ctypes.data_as(_f32p), d.ctypes.data_as(_f32p), s.size,
                                                   s.ctypes.data_as(_f32p), idx.ctypes.data_as(_i64p), entry.ctypes.data_as(_f32p)),
                  "dxmc_oracle_trace_indices")
        return idx, entry

    def trace_air_runs(self, pos, direction):
        """(length, cubes crossed, exits the grid, starts in an air brick, reaches the world, end point) of the air run of fixed rays"""
        p = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(direction, np.float32).reshape(-1, 3)
        length, info, end = np.zeros(p.shape[0], np.float32), np.zeros(p.shape[0], np.uint32), np.zeros((p.shape[0], 3), np.float32)
        self._chk(self.l.dxmc_oracle_trace_air_runs(self.h, C.c_uint64(p.shape[0]), p.ctypes.data_as(_f32p), d.ctypes.data_as(_f32p), length.ctypes.data_as(_f32p),
                  info.ctypes.data_as(C.POINTER(C.c_uint32)), end.ctypes.data_as(_f32p)), "dxmc_oracle_trace_air_runs")
        return length, info & 0xffff, (info >> 16) & 1, (info >> 17) & 1, (info >> 18) & 1, end

    def sample_particles(self, exposure, exposure_index, seed, n):
        out = np.zeros((n, 8), np.float32)
        self._chk(self.l.dxmc_oracle_sample_particles(self.h, C.byref(exposure), C.c_uint64(exposure_index), C.c_uint64(seed), C.c_uint64(n),
                                                      out.ctypes.data_as(_f32p)), "dxmc_oracle_sample_particles")
        return out

    def sample_interaction(self, kind, model, material, energy, seed, n):
        out = np.zeros((n, 5), np.float32)
        self._chk(self.l.dxmc_oracle_sample_interaction(self.h, int(kind), int(model), C.c_uint8(material), C.c_float(energy), C.c_uint64(seed),
                                                        C.c_uint64(n), out.ctypes.data_as(_f32p)), "dxmc_oracle_sample_interaction")
        return out

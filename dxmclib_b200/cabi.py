"""ctypes bindings of include/dxmcb200.h — the device C ABI of libdxmcb200.so.

Used by the parity tests and bench.py to call the CUDA path directly with plain arrays. Loading fails
loudly when the library is not built; creating a context fails loudly when there is no CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import scene as _scene

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)

# every symbol include/dxmcb200.h declares
CABI_SYMBOLS = [
    "dxmcb200_physics_backend", "dxmcb200_device_count", "dxmcb200_create", "dxmcb200_destroy", "dxmcb200_last_error", "dxmcb200_set_world", "dxmcb200_material_max_density", "dxmcb200_set_pool_limit", "dxmcb200_trim_pool",
    "dxmcb200_set_luts", "dxmcb200_set_beam_tables", "dxmcb200_suggest_fixed_point", "dxmcb200_set_fixed_point", "dxmcb200_set_tracking", "dxmcb200_get_bricks", "dxmcb200_get_brick_distance", "dxmcb200_get_grid_form", "dxmcb200_trace_air_runs", "dxmcb200_tube_bremsstrahlung",
    "dxmcb200_clear", "dxmcb200_history_stream", "dxmcb200_run", "dxmcb200_upload_exposures", "dxmcb200_exposure_table", "dxmcb200_generate_exposures", "dxmcb200_run_range", "dxmcb200_run_resident", "dxmcb200_run_strided", "dxmcb200_run_strided_monitored",
    "dxmcb200_last_run_ms", "dxmcb200_get_result", "dxmcb200_get_raw", "dxmcb200_accumulators", "dxmcb200_reduce", "dxmcb200_comm_create", "dxmcb200_comm_destroy", "dxmcb200_reduce_collect",
    "dxmcb200_get_stats", "dxmcb200_get_kernel_times", "dxmcb200_enable_stats", "dxmcb200_eval_attenuation", "dxmcb200_trace_indices", "dxmcb200_sample_particles",
    "dxmcb200_sample_interaction",
]


class World(C.Structure):
    _fields_ = [("dim", C.c_uint64 * 3), ("spacing", C.c_float * 3), ("extent_safe", C.c_float * 6), ("density", _f32p),
                ("material", _u8p), ("measurement", _u8p)]


class Luts(C.Structure):
    _fields_ = [("n_materials", C.c_uint32), ("n_segments", C.c_uint32), ("linear_index", C.c_uint32), ("linear_step", C.c_float),
                ("linear_energy", C.c_float), ("knots", _f32p), ("coefficients", _f32p), ("max_coefficients", _f32p),
                ("rita", _f32p), ("spline", _f32p), ("shells", _f32p)]


class Spectrum(C.Structure):
    _fields_ = [("n", C.c_uint32), ("probs", _f32p), ("alias", _u32p), ("energies", _f32p)]


class Heel(C.Structure):
    _fields_ = [("energy_start", C.c_float), ("energy_step", C.c_float), ("energy_size", C.c_uint32), ("angle_start", C.c_float),
                ("angle_step", C.c_float), ("angle_size", C.c_uint32), ("weights", _f32p)]


class Bowtie(C.Structure):
    _fields_ = [("n", C.c_uint32), ("angles", _f32p), ("weights", _f32p)]


class Exposure(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("cosines", C.c_float * 6), ("beam_direction", C.c_float * 3),
                ("collimation", C.c_float * 4), ("weight", C.c_float), ("mono_energy", C.c_float), ("spectrum", C.c_int32),
                ("heel", C.c_int32), ("bowtie", C.c_int32), ("reserved", C.c_uint32), ("histories", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("histories", C.c_uint64), ("histories_in_world", C.c_uint64), ("steps", C.c_uint64), ("lookups", C.c_uint64),
                ("interactions", C.c_uint64), ("score_events", C.c_uint64), ("kernel_launches", C.c_uint64), ("kernel_ms", C.c_double),
                ("air_walks", C.c_uint64), ("bricks_crossed", C.c_uint64)]


def lib() -> C.CDLL:
    l = _scene.product_lib()
    if not getattr(l, "_cabi_ready", False):
        l.dxmcb200_last_error.restype = C.c_char_p
        l.dxmcb200_last_error.argtypes = [C.c_void_p]
        l.dxmcb200_destroy.restype = None
        l.dxmcb200_destroy.argtypes = [C.c_void_p]
        l.dxmcb200_history_stream.restype = None
        l.dxmcb200_physics_backend.restype = C.c_char_p
        for name in CABI_SYMBOLS:
            if name not in ("dxmcb200_last_error", "dxmcb200_destroy", "dxmcb200_history_stream", "dxmcb200_physics_backend"):
                getattr(l, name).restype = C.c_int
        l._cabi_ready = True
    return l


class CabiError(RuntimeError):
    pass


def history_stream(seed: int, exposure: int, history: int):
    out = (C.c_uint64 * 2)()
    lib().dxmcb200_history_stream(C.c_uint64(seed), C.c_uint64(exposure), C.c_uint64(history), out)
    return int(out[0]), int(out[1])


def device_count() -> int:
    n = C.c_int(0)
    lib().dxmcb200_device_count(C.byref(n))
    return int(n.value)


def physics_backend():
    """(name, approximate) of the element data the library's host-side table builders use."""
    flag = C.c_int()
    name = lib().dxmcb200_physics_backend(C.byref(flag))
    return name.decode(), bool(flag.value)


def suggest_fixed_point(total_histories: int, max_energy_weight: float):
    """(energy_bits, energy_sq_bits) that cannot overflow for a job of this size (dxmcb200_suggest_fixed_point)."""
    a, b = C.c_int(), C.c_int()
    rc = lib().dxmcb200_suggest_fixed_point(C.c_uint64(int(total_histories)), C.c_double(float(max_energy_weight)), C.byref(a), C.byref(b))
    if rc != 0:
        raise CabiError(f"dxmcb200_suggest_fixed_point failed with status {rc}")
    return int(a.value), int(b.value)


class Context:
    """Thin owner of a dxmcb200_ctx; `handle` may also wrap a ctx borrowed from Scene.b200_context()."""

    def __init__(self, device: int = 0, handle: C.c_void_p | None = None):
        self.l = lib()
        self.owned = handle is None
        if handle is None:
            h = C.c_void_p()
            rc = self.l.dxmcb200_create(int(device), C.byref(h))
            if rc != 0:
                raise CabiError(f"dxmcb200_create(device={device}) failed with status {rc}: no CUDA device and no CPU fallback")
            handle = h
        self.h = handle
        self._keep = []

    def close(self):
        if self.owned and self.h:
            self.l.dxmcb200_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc != 0:
            raise CabiError(f"{what} failed with status {rc}: {self.l.dxmcb200_last_error(self.h).decode()}")

    def set_world(self, dim, spacing, extent_safe, density, material, measurement=None):
        w = World()
        w.dim[:] = [int(x) for x in dim]
        w.spacing[:] = [float(x) for x in spacing]
        w.extent_safe[:] = [float(x) for x in extent_safe]
        d = np.ascontiguousarray(density, np.float32)
        m = np.ascontiguousarray(material, np.uint8)
        w.density = d.ctypes.data_as(_f32p)
        w.material = m.ctypes.data_as(_u8p)
        if measurement is not None:
            me = np.ascontiguousarray(measurement, np.uint8)
            w.measurement = me.ctypes.data_as(_u8p)
        self._chk(self.l.dxmcb200_set_world(self.h, C.byref(w)), "dxmcb200_set_world")
        self.n_voxels = int(np.prod([int(x) for x in dim]))

    def material_max_density(self, n_materials: int) -> np.ndarray:
        """Per-material maximum density of the uploaded grid (device pass; input of the Woodcock majorant)."""
        out = np.zeros(int(n_materials), np.float32)
        self._chk(self.l.dxmcb200_material_max_density(self.h, C.c_uint32(int(n_materials)), out.ctypes.data_as(_f32p)),
                  "dxmcb200_material_max_density")
        return out

    def set_luts_from_scene(self, sc: "_scene.Scene"):
        """Flatten the LUT tables of a scene (after lut_generate) with the layout of dxmcb200_luts."""
        knots = sc.lut_table(0)
        coeff = sc.lut_table(1)
        maxc = sc.lut_table(2)
        lin = sc.lut_table(3)
        rita = sc.lut_table(4)
        spl = sc.lut_table(5)
        n_seg = knots.size
        n_mat = coeff.size // (n_seg * 6)
        spline = np.zeros((n_mat, 63), np.float32)
        raw = spl.reshape(n_mat, 79)  # 60 coefficients, 16 knots, step, start, stop
        spline[:, :60] = raw[:, :60]
        spline[:, 60] = raw[:, 77]  # start
        spline[:, 61] = raw[:, 76]  # step
        spline[:, 62] = raw[:, 78]  # stop
        shells = np.zeros((n_mat, 12, 11), np.float32)
        for i in range(n_mat):
            shells[i] = sc.material_shells(i)[:, :11].astype(np.float32)
        l = Luts()
        l.n_materials, l.n_segments, l.linear_index = n_mat, n_seg, int(lin[0])
        l.linear_step, l.linear_energy = float(lin[1]), float(lin[2])
        keep = [knots, coeff, maxc, rita, np.ascontiguousarray(spline), np.ascontiguousarray(shells)]
        l.knots, l.coefficients, l.max_coefficients, l.rita, l.spline, l.shells = [a.ctypes.data_as(_f32p) for a in keep]
        self._chk(self.l.dxmcb200_set_luts(self.h, C.byref(l)), "dxmcb200_set_luts")
        self.n_materials = n_mat

    def set_beam_tables(self, spectra=(), heels=(), bowties=()):
        """spectra: [(probs, alias, energies)], heels: [(e0, de, ne, a0, da, na, weights)], bowties: [(angles, weights)]"""
        keep = []
        S = (Spectrum * max(len(spectra), 1))()
        for i, (p, a, e) in enumerate(spectra):
            p, a, e = np.ascontiguousarray(p, np.float32), np.ascontiguousarray(a, np.uint32), np.ascontiguousarray(e, np.float32)
            keep += [p, a, e]
            S[i] = Spectrum(p.size, p.ctypes.data_as(_f32p), a.ctypes.data_as(_u32p), e.ctypes.data_as(_f32p))
        H = (Heel * max(len(heels), 1))()
        for i, (e0, de, ne, a0, da, na, w) in enumerate(heels):
            w = np.ascontiguousarray(w, np.float32)
            keep.append(w)
            H[i] = Heel(e0, de, ne, a0, da, na, w.ctypes.data_as(_f32p))
        B = (Bowtie * max(len(bowties), 1))()
        for i, (a, w) in enumerate(bowties):
            a, w = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(w, np.float32)
            keep += [a, w]
            B[i] = Bowtie(a.size, a.ctypes.data_as(_f32p), w.ctypes.data_as(_f32p))
        self._chk(self.l.dxmcb200_set_beam_tables(self.h, len(spectra), S, len(heels), H, len(bowties), B), "dxmcb200_set_beam_tables")

    def set_fixed_point(self, energy_bits, energy_sq_bits):
        self._chk(self.l.dxmcb200_set_fixed_point(self.h, energy_bits, energy_sq_bits), "dxmcb200_set_fixed_point")

    def clear(self):
        self._chk(self.l.dxmcb200_clear(self.h), "dxmcb200_clear")

    def run(self, exposures, begin, end, model=1, seed=1):
        arr = (Exposure * len(exposures))(*exposures)
        self._chk(self.l.dxmcb200_run(self.h, arr, C.c_uint64(begin), C.c_uint64(end), int(model), C.c_uint64(seed), None, None, None),
                  "dxmcb200_run")

    def resident_exposures(self, n: int) -> np.ndarray:
        """The exposure table resident on the device (uploaded or generated there), as a structured array."""
        import torch

        dt = np.dtype([("position", "<f4", 3), ("cosines", "<f4", 6), ("beam_direction", "<f4", 3), ("collimation", "<f4", 4), ("weight", "<f4"),
                       ("mono_energy", "<f4"), ("spectrum", "<i4"), ("heel", "<i4"), ("bowtie", "<i4"), ("reserved", "<u4"), ("histories", "<u8")])
        assert dt.itemsize == C.sizeof(Exposure)
        out = np.zeros(int(n), dt)
        ptr = C.c_void_p()
        self._chk(self.l.dxmcb200_exposure_table(self.h, C.byref(ptr), None), "dxmcb200_exposure_table")

        class _Block:
            __cuda_array_interface__ = {"shape": (int(n) * dt.itemsize,), "typestr": "|u1", "data": (ptr.value, False), "version": 2}

        raw = torch.as_tensor(_Block(), device="cuda").cpu().numpy()
        return raw.view(dt).copy()

    def upload_exposures(self, exposures):
        arr = (Exposure * len(exposures))(*exposures)
        self._chk(self.l.dxmcb200_upload_exposures(self.h, arr, C.c_uint64(len(exposures))), "dxmcb200_upload_exposures")

    def run_resident(self, begin, end, model=1, seed=1):
        self._chk(self.l.dxmcb200_run_resident(self.h, C.c_uint64(begin), C.c_uint64(end), int(model), C.c_uint64(seed)), "dxmcb200_run_resident")

    def run_strided(self, first, stride, count, model=1, seed=1):
        self._chk(self.l.dxmcb200_run_strided(self.h, C.c_uint64(first), C.c_uint64(stride), C.c_uint64(count), int(model), C.c_uint64(seed)),
                  "dxmcb200_run_strided")

    def last_run_ms(self) -> float:
        ms = C.c_double(0)
        self._chk(self.l.dxmcb200_last_run_ms(self.h, C.byref(ms)), "dxmcb200_last_run_ms")
        return float(ms.value)

    def get_result(self, mode, histories, calibration=1.0, n_voxels=None):
        n = n_voxels or self.n_voxels
        dose, ev, var = np.zeros(n, np.float32), np.zeros(n, np.uint32), np.zeros(n, np.float32)
        self._chk(self.l.dxmcb200_get_result(self.h, int(mode), C.c_uint64(histories), C.c_float(calibration), dose.ctypes.data_as(_f32p),
                                             ev.ctypes.data_as(_u32p), var.ctypes.data_as(_f32p)), "dxmcb200_get_result")
        return dose, ev, var

    def get_raw(self, n_voxels=None):
        n = n_voxels or self.n_voxels
        e, e2, ev = np.zeros(n, np.int64), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        self._chk(self.l.dxmcb200_get_raw(self.h, e.ctypes.data_as(_i64p), e2.ctypes.data_as(_u64p), ev.ctypes.data_as(_u64p)), "dxmcb200_get_raw")
        return e, e2, ev

    def accumulators(self):
        ptr = C.c_void_p()
        n = C.c_uint64(0)
        self._chk(self.l.dxmcb200_accumulators(self.h, C.byref(ptr), C.byref(n)), "dxmcb200_accumulators")
        return int(ptr.value), int(n.value)

    def stats(self) -> dict:
        s = Stats()
        self._chk(self.l.dxmcb200_get_stats(self.h, C.byref(s)), "dxmcb200_get_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def kernel_times(self) -> dict:
        """Device milliseconds and launch counts of generate / transport / air walk / interact kernels since clear()."""
        ms = (C.c_double * 4)()
        n = (C.c_uint64 * 4)()
        self._chk(self.l.dxmcb200_get_kernel_times(self.h, ms, n), "dxmcb200_get_kernel_times")
        names = ("generate", "transport", "airwalk", "interact")
        return {k: {"ms": float(ms[i]), "launches": int(n[i])} for i, k in enumerate(names)}

    def set_tracking(self, tracking: int, brick_mm: float = 0.0):
        """0: the reference's Woodcock loop everywhere; 1: + empty-space traversal through air bricks (default)."""
        self._chk(self.l.dxmcb200_set_tracking(self.h, int(tracking), C.c_float(brick_mm)), "dxmcb200_set_tracking")

    def bricks(self, n_materials: int) -> dict:
        shift, nb, f_air = (C.c_uint32 * 3)(), (C.c_uint32 * 3)(), C.c_float()
        self._chk(self.l.dxmcb200_get_bricks(self.h, shift, nb, C.byref(f_air), None, None, None), "dxmcb200_get_bricks")
        n = int(nb[0]) * int(nb[1]) * int(nb[2])
        ratio, bmax, air = np.zeros(n_materials, np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        self._chk(self.l.dxmcb200_get_bricks(self.h, shift, nb, C.byref(f_air), ratio.ctypes.data_as(_f32p), bmax.ctypes.data_as(_f32p),
                                             air.ctypes.data_as(_u8p)), "dxmcb200_get_bricks")
        dist = np.zeros(8 * n, np.uint8)
        self._chk(self.l.dxmcb200_get_brick_distance(self.h, dist.ctypes.data_as(_u8p)), "dxmcb200_get_brick_distance")
        return {"shift": list(shift), "nb": list(nb), "f_air": float(f_air.value), "ratio": ratio, "brick_max": bmax, "air": air, "distance": dist}

    def grid_form(self):
        """(bits per voxel the kernels read: 64 = records, 8 / 4 = palette indices; distinct records of a palette grid)"""
        bits, distinct = C.c_int(), C.c_uint32()
        self._chk(self.l.dxmcb200_get_grid_form(self.h, C.byref(bits), C.byref(distinct)), "dxmcb200_get_grid_form")
        return int(bits.value), int(distinct.value)

    def enable_stats(self, on=True):
        self._chk(self.l.dxmcb200_enable_stats(self.h, int(on)), "dxmcb200_enable_stats")

    def eval_attenuation(self, material, energy):
        m = np.ascontiguousarray(material, np.uint8)
        e = np.ascontiguousarray(energy, np.float32)
        out = np.zeros((e.size, 3), np.float32)
        mx = np.zeros(e.size, np.float32)
        self._chk(self.l.dxmcb200_eval_attenuation(self.h, C.c_uint64(e.size), m.ctypes.data_as(_u8p), e.ctypes.data_as(_f32p),
                                                   out.ctypes.data_as(_f32p), mx.ctypes.data_as(_f32p)), "dxmcb200_eval_attenuation")
        return out, mx

    def trace_indices(self, pos, direction, steps):
        p = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(direction, np.float32).reshape(-1, 3)
        s = np.ascontiguousarray(steps, np.float32)
        idx = np.zeros((p.shape[0], s.size + 1), np.int64)
        entry = np.zeros((p.shape[0], 3), np.float32)
        self._chk(self.l.dxmcb200_trace_indices(self.h, C.c_uint64(p.shape[0]), p.ctypes.data_as(_f32p), d.ctypes.data_as(_f32p), s.size,
                                                s.ctypes.data_as(_f32p), idx.ctypes.data_as(_i64p), entry.ctypes.data_as(_f32p)),
                  "dxmcb200_trace_indices")
        return idx, entry

    def trace_air_runs(self, pos, direction):
        """(length, cubes crossed, exits the grid, starts in an air brick, reaches the world, end point) of the air run of fixed rays"""
        p = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(direction, np.float32).reshape(-1, 3)
        length, info, end = np.zeros(p.shape[0], np.float32), np.zeros(p.shape[0], np.uint32), np.zeros((p.shape[0], 3), np.float32)
        self._chk(self.l.dxmcb200_trace_air_runs(self.h, C.c_uint64(p.shape[0]), p.ctypes.data_as(_f32p), d.ctypes.data_as(_f32p), length.ctypes.data_as(_f32p),
                  info.ctypes.data_as(C.POINTER(C.c_uint32)), end.ctypes.data_as(_f32p)), "dxmcb200_trace_air_runs")
        return length, info & 0xffff, (info >> 16) & 1, (info >> 17) & 1, (info >> 18) & 1, end

    def sample_particles(self, exposure: Exposure, exposure_index, seed, n):
        out = np.zeros((n, 8), np.float32)
        self._chk(self.l.dxmcb200_sample_particles(self.h, C.byref(exposure), C.c_uint64(exposure_index), C.c_uint64(seed), C.c_uint64(n),
                                                   out.ctypes.data_as(_f32p)), "dxmcb200_sample_particles")
        return out

    def sample_interaction(self, kind, model, material, energy, seed, n):
        out = np.zeros((n, 5), np.float32)
        self._chk(self.l.dxmcb200_sample_interaction(self.h, int(kind), int(model), C.c_uint8(material), C.c_float(energy), C.c_uint64(seed),
                                                     C.c_uint64(n), out.ctypes.data_as(_f32p)), "dxmcb200_sample_interaction")
        return out

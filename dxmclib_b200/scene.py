"""ctypes view of include/dxmcb200_scene.h.

`Scene(lib)` wraps one `dxs_scene`. The same class drives
  * the product: dxmclib_b200/libdxmcb200.so  (B200 kernels behind the reference C++ API), and
  * the checker: oracle/_ref/libdxmc_ref.so   (the unmodified reference; tests and bench baseline only).
Both export identical symbols, see the header for the reference members each call forwards to.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
PRODUCT_LIB = os.path.join(_HERE, "libdxmcb200.so")
REFERENCE_LIB = os.path.join(ROOT, "oracle", "_ref", "libdxmc_ref.so")
# the same reference sources built -O3 -march=x86-64-v3 (its own Release flags, portable level): timing baseline only
REFERENCE_TIMING_LIB = os.path.join(ROOT, "oracle", "_ref", "libdxmc_ref_o3.so")

MODEL_NONE, MODEL_LIVERMORE, MODEL_IA = 0, 1, 2
OUT_EV_PER_HISTORY, OUT_DOSE = 0, 1
WORKERS_COUNTER_STREAMS = -1  # reference harness only: one RandomState per history, keyed like the kernels' streams

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class Tube(C.Structure):
    _fields_ = [("voltage", C.c_float), ("anode_angle_deg", C.c_float), ("al_mm", C.c_float), ("cu_mm", C.c_float),
                ("sn_mm", C.c_float), ("energy_resolution", C.c_float)]


class DXParams(C.Structure):
    _fields_ = [("tube", Tube), ("position", C.c_float * 3), ("sdd", C.c_float), ("field_size", C.c_float * 2),
                ("source_angles_deg", C.c_float * 2), ("tube_rotation_deg", C.c_float), ("dap", C.c_float),
                ("model_heel", C.c_int), ("histories", C.c_uint64), ("exposures", C.c_uint64)]


class CTParams(C.Structure):
    _fields_ = [("tube", Tube), ("spiral", C.c_int), ("position", C.c_float * 3), ("cosines", C.c_float * 6),
                ("sdd", C.c_float), ("collimation", C.c_float), ("fov", C.c_float), ("start_angle_deg", C.c_float),
                ("exposure_step_deg", C.c_float), ("scan_length", C.c_float), ("pitch", C.c_float), ("step", C.c_float),
                ("gantry_tilt_deg", C.c_float), ("ctdi_vol", C.c_float), ("ctdi_phantom_diameter", C.c_uint64),
                ("model_heel", C.c_int), ("use_xcare", C.c_int), ("xcare_filter_angle_deg", C.c_float),
                ("xcare_span_deg", C.c_float), ("xcare_ramp_deg", C.c_float), ("xcare_low_weight", C.c_float),
                ("histories", C.c_uint64)]


class CTDualParams(C.Structure):
    _fields_ = [("a", CTParams), ("tube_b", Tube), ("sdd_b", C.c_float), ("fov_b", C.c_float), ("start_angle_b_deg", C.c_float),
                ("mas_a", C.c_float), ("mas_b", C.c_float)]


class CBCTParams(C.Structure):
    _fields_ = [("dx", DXParams), ("span_deg", C.c_float), ("step_deg", C.c_float)]


class Exposure(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("cosines", C.c_float * 6), ("beam_direction", C.c_float * 3),
                ("collimation", C.c_float * 4), ("weight", C.c_float), ("mono_energy", C.c_float), ("has_spectrum", C.c_int32),
                ("has_heel", C.c_int32), ("has_bowtie", C.c_int32), ("histories", C.c_uint64)]


class ResultInfo(C.Structure):
    _fields_ = [("histories", C.c_uint64), ("seconds", C.c_double), ("units", C.c_char * 16)]


class ProgressReport(C.Structure):
    _fields_ = [("percent_seen", C.c_double), ("percent_final", C.c_double), ("images_polled", C.c_uint32), ("images_nonzero", C.c_uint32),
                ("image_width", C.c_uint32), ("image_height", C.c_uint32), ("cancelled", C.c_int32), ("eta", C.c_char * 96)]


# every symbol include/dxmcb200_scene.h declares (the CPU test suite checks the export list against this)
SCENE_SYMBOLS = [
    "dxs_backend", "dxs_last_error", "dxs_create", "dxs_destroy", "dxs_world_geometry", "dxs_world_add_material",
    "dxs_world_add_element", "dxs_world_arrays", "dxs_world_ctdi_phantom", "dxs_world_validate",
    "dxs_world_dimensions", "dxs_world_get_arrays", "dxs_world_ctdi_holes", "dxs_trace_indices", "dxs_material_attenuation",
    "dxs_material_form_factor_sq", "dxs_material_scatter_factor", "dxs_material_binding_energies",
    "dxs_material_shells", "dxs_material_density", "dxs_lut_generate", "dxs_lut_attenuation",
    "dxs_lut_max_inverse", "dxs_lut_scatter_factor", "dxs_lut_sample_form_factor", "dxs_lut_table",
    "dxs_source_pencil", "dxs_source_isotropic", "dxs_source_dx", "dxs_source_ct", "dxs_source_ct_dual", "dxs_source_topogram", "dxs_source_cbct", "dxs_source_bowtie",
    "dxs_source_aec", "dxs_source_total_exposures", "dxs_source_max_energy", "dxs_source_exposure",
    "dxs_source_table", "dxs_source_spectrum", "dxs_source_calibration", "dxs_transport", "dxs_transport_monitored",
    "dxs_b200_prepare", "dxs_b200_run", "dxs_b200_run_strided", "dxs_b200_collect", "dxs_b200_context", "dxs_b200_release", "dxs_b200_set_devices",
]

_libs: dict[str, C.CDLL] = {}


def load(path: str) -> C.CDLL:
    """Load a library implementing the scene API. Fails loudly: there is no fallback."""
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)")
    lib = C.CDLL(path)  # RTLD_LOCAL: product and reference libraries export the same C++ names
    lib.dxs_backend.restype = C.c_char_p
    lib.dxs_last_error.restype = C.c_char_p
    lib.dxs_create.restype = C.c_void_p
    lib.dxs_destroy.argtypes = [C.c_void_p]
    lib.dxs_destroy.restype = None
    for name in SCENE_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("dxs_backend", "dxs_last_error", "dxs_create", "dxs_destroy"):
            fn.restype = C.c_int
    _libs[path] = lib
    return lib


def product_lib() -> C.CDLL:
    return load(PRODUCT_LIB)


def reference_lib() -> C.CDLL:
    return load(REFERENCE_LIB)


def reference_timing_lib() -> tuple[C.CDLL, str]:
    """The reference library bench.py times as the CPU baseline, and a word on how it was built: the -O3 AVX2 build when it
    exists and this host can execute it (probed in a child process, an illegal instruction must not take the caller down),
    otherwise the exact-arithmetic build the parity tests use."""
    import subprocess
    import sys

    if os.path.exists(REFERENCE_TIMING_LIB):
        probe = ("import ctypes as C, sys; l = C.CDLL(sys.argv[1]); l.dxs_create.restype = C.c_void_p; s = C.c_void_p(l.dxs_create()); "
                 "assert l.dxs_world_add_material(s, b'Water, Liquid', C.c_double(1.0)) == 0; "
                 "out = (C.c_double * 4)(); assert l.dxs_material_attenuation(s, 0, C.c_double(60.0), out) == 0 and out[3] > 0")
        try:
            ok = subprocess.run([sys.executable, "-c", probe, REFERENCE_TIMING_LIB], capture_output=True, timeout=120).returncode == 0
        except Exception:
            ok = False
        if ok:
            return load(REFERENCE_TIMING_LIB), "g++ -O3 -march=x86-64-v3"
    return load(REFERENCE_LIB), "g++ -O2 -ffp-contract=off"


class SceneError(RuntimeError):
    pass


def _chk(code: int, what: str):
    if code != 0:
        msg = ""
        for lib in _libs.values():
            m = lib.dxs_last_error().decode()
            if m:
                msg = ": " + m
        raise SceneError(f"{what} failed with status {code}{msg}")


def _f32(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} floats, got {a.size}")
    return a


@dataclass
class Result:
    dose: np.ndarray
    n_events: np.ndarray
    variance: np.ndarray
    histories: int
    seconds: float
    units: str


class Scene:
    def __init__(self, lib: C.CDLL):
        self.lib = lib
        self.h = C.c_void_p(lib.dxs_create())
        if not self.h:
            raise MemoryError("dxs_create")
        self.dim = None

    def close(self):
        if self.h:
            self.lib.dxs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def backend(self) -> str:
        return self.lib.dxs_backend().decode()

    # ---- world
    def world(self, dim, spacing, origin=(0, 0, 0), cosines=(1, 0, 0, 0, 1, 0)):
        d = (C.c_uint64 * 3)(*[int(x) for x in dim])
        sp, o, c = _f32(spacing, 3), _f32(origin, 3), _f32(cosines, 6)
        _chk(self.lib.dxs_world_geometry(self.h, d, sp.ctypes.data_as(_f32p), o.ctypes.data_as(_f32p),
                                         c.ctypes.data_as(_f32p)), "dxs_world_geometry")
        self.dim = tuple(int(x) for x in dim)
        return self

    def add_material(self, name: str, density: float = -1.0):
        _chk(self.lib.dxs_world_add_material(self.h, name.encode(), C.c_double(density)), f"add_material({name})")
        return self

    def add_element(self, Z: int):
        _chk(self.lib.dxs_world_add_element(self.h, int(Z)), f"add_element({Z})")
        return self

    def arrays(self, density, material, measurement=None):
        n = int(np.prod(self.dim))
        dens = _f32(density, n)
        mat = np.ascontiguousarray(material, dtype=np.uint8)
        meas = None if measurement is None else np.ascontiguousarray(measurement, dtype=np.uint8)
        assert mat.size == n and (meas is None or meas.size == n)
        _chk(self.lib.dxs_world_arrays(self.h, dens.ctypes.data_as(_f32p), mat.ctypes.data_as(_u8p),
                                       None if meas is None else meas.ctypes.data_as(_u8p)), "dxs_world_arrays")
        return self

    def ctdi_phantom(self, diameter=320):
        _chk(self.lib.dxs_world_ctdi_phantom(self.h, C.c_uint64(diameter)), "dxs_world_ctdi_phantom")
        dim, _, _ = self.dimensions()
        self.dim = dim
        return self

    def validate(self) -> bool:
        v = C.c_int(0)
        _chk(self.lib.dxs_world_validate(self.h, C.byref(v)), "dxs_world_validate")
        return bool(v.value)

    def dimensions(self):
        d = (C.c_uint64 * 3)()
        sp = (C.c_float * 3)()
        ext = (C.c_float * 6)()
        _chk(self.lib.dxs_world_dimensions(self.h, d, sp, ext), "dxs_world_dimensions")
        return tuple(int(x) for x in d), np.array(sp[:], dtype=np.float32), np.array(ext[:], dtype=np.float32)

    def get_arrays(self):
        n = int(np.prod(self.dim))
        dens = np.zeros(n, np.float32)
        mat = np.zeros(n, np.uint8)
        meas = np.zeros(n, np.uint8)
        _chk(self.lib.dxs_world_get_arrays(self.h, dens.ctypes.data_as(_f32p), mat.ctypes.data_as(_u8p),
                                           meas.ctypes.data_as(_u8p)), "dxs_world_get_arrays")
        return dens, mat, meas

    def ctdi_holes(self, position: int):
        cnt = C.c_uint64(0)
        _chk(self.lib.dxs_world_ctdi_holes(self.h, position, None, C.byref(cnt)), "dxs_world_ctdi_holes")
        out = np.zeros(cnt.value, np.uint64)
        _chk(self.lib.dxs_world_ctdi_holes(self.h, position, out.ctypes.data_as(_u64p), C.byref(cnt)), "dxs_world_ctdi_holes")
        return out

    def trace_indices(self, pos, direction, steps):
        """Voxel-index sequences of fixed rays (see dxs_trace_indices); returns (indices [n, n_steps+1], entry [n, 3])."""
        p = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(direction, np.float32).reshape(-1, 3)
        s = np.ascontiguousarray(steps, np.float32)
        idx = np.zeros((p.shape[0], s.size + 1), np.int64)
        entry = np.zeros((p.shape[0], 3), np.float32)
        _chk(self.lib.dxs_trace_indices(self.h, C.c_uint64(p.shape[0]), p.ctypes.data_as(_f32p), d.ctypes.data_as(_f32p), int(s.size),
                                        s.ctypes.data_as(_f32p), idx.ctypes.data_as(C.POINTER(C.c_int64)), entry.ctypes.data_as(_f32p)),
             "dxs_trace_indices")
        return idx, entry

    # ---- material
    def material_attenuation(self, idx, energy):
        out = (C.c_double * 4)()
        _chk(self.lib.dxs_material_attenuation(self.h, idx, C.c_double(energy), out), "dxs_material_attenuation")
        return np.array(out[:])

    def material_form_factor_sq(self, idx, q):
        out = C.c_double(0)
        _chk(self.lib.dxs_material_form_factor_sq(self.h, idx, C.c_double(q), C.byref(out)), "form_factor")
        return out.value

    def material_scatter_factor(self, idx, q):
        out = C.c_double(0)
        _chk(self.lib.dxs_material_scatter_factor(self.h, idx, C.c_double(q), C.byref(out)), "scatter_factor")
        return out.value

    def material_binding_energies(self, idx, min_value=1.0):
        n = C.c_int(0)
        _chk(self.lib.dxs_material_binding_energies(self.h, idx, C.c_double(min_value), None, C.byref(n)), "binding")
        out = np.zeros(n.value, np.float64)
        _chk(self.lib.dxs_material_binding_energies(self.h, idx, C.c_double(min_value), out.ctypes.data_as(_f64p), C.byref(n)), "binding")
        return out

    def material_shells(self, idx):
        out = np.zeros(12 * 13, np.float64)
        _chk(self.lib.dxs_material_shells(self.h, idx, out.ctypes.data_as(_f64p)), "dxs_material_shells")
        return out.reshape(12, 13)

    def material_density(self, idx):
        out = C.c_double(0)
        _chk(self.lib.dxs_material_density(self.h, idx, C.byref(out)), "dxs_material_density")
        return out.value

    # ---- LUT
    def lut_generate(self, max_energy):
        _chk(self.lib.dxs_lut_generate(self.h, C.c_float(max_energy)), "dxs_lut_generate")
        return self

    def lut_attenuation(self, material, energy):
        out = (C.c_float * 3)()
        _chk(self.lib.dxs_lut_attenuation(self.h, material, C.c_float(energy), out), "dxs_lut_attenuation")
        return np.array(out[:], dtype=np.float32)

    def lut_max_inverse(self, energy):
        out = C.c_float(0)
        _chk(self.lib.dxs_lut_max_inverse(self.h, C.c_float(energy), C.byref(out)), "dxs_lut_max_inverse")
        return np.float32(out.value)

    def lut_scatter_factor(self, material, q):
        out = C.c_float(0)
        _chk(self.lib.dxs_lut_scatter_factor(self.h, material, C.c_float(q), C.byref(out)), "dxs_lut_scatter_factor")
        return np.float32(out.value)

    def lut_sample_form_factor(self, material, qmax_squared, seed, n):
        s = (C.c_uint64 * 2)(int(seed[0]), int(seed[1]))
        out = np.zeros(n, np.float32)
        _chk(self.lib.dxs_lut_sample_form_factor(self.h, material, C.c_float(qmax_squared), s, n,
                                                 out.ctypes.data_as(_f32p)), "dxs_lut_sample_form_factor")
        return out

    def lut_table(self, what):
        n = C.c_uint64(0)
        _chk(self.lib.dxs_lut_table(self.h, what, None, C.byref(n)), "dxs_lut_table")
        out = np.zeros(n.value, np.float32)
        _chk(self.lib.dxs_lut_table(self.h, what, out.ctypes.data_as(_f32p), C.byref(n)), "dxs_lut_table")
        return out

    # ---- sources
    def source_pencil(self, pos, cosines, energy, histories, exposures):
        p, c = _f32(pos, 3), _f32(cosines, 6)
        _chk(self.lib.dxs_source_pencil(self.h, p.ctypes.data_as(_f32p), c.ctypes.data_as(_f32p), C.c_float(energy),
                                        C.c_uint64(histories), C.c_uint64(exposures)), "dxs_source_pencil")
        return self

    def source_isotropic(self, pos, cosines, collimation, weights, energies, histories, exposures, ct=False):
        p, c, col = _f32(pos, 3), _f32(cosines, 6), _f32(collimation, 4)
        w, e = _f32(weights), _f32(energies)
        assert w.size == e.size
        _chk(self.lib.dxs_source_isotropic(self.h, int(ct), p.ctypes.data_as(_f32p), c.ctypes.data_as(_f32p),
                                           col.ctypes.data_as(_f32p), int(w.size), w.ctypes.data_as(_f32p),
                                           e.ctypes.data_as(_f32p), C.c_uint64(histories), C.c_uint64(exposures)),
             "dxs_source_isotropic")
        return self

    def source_dx(self, **kw):
        p = DXParams()
        p.tube = Tube(kw.get("voltage", 0), kw.get("anode_angle_deg", 0), kw.get("al_mm", 0), kw.get("cu_mm", 0),
                      kw.get("sn_mm", 0), kw.get("energy_resolution", 0))
        p.position[:] = kw.get("position", (0, 0, 0))
        p.sdd = kw.get("sdd", 0)
        p.field_size[:] = kw.get("field_size", (0, 0))
        p.source_angles_deg[:] = kw.get("source_angles_deg", (0, 0))
        p.tube_rotation_deg = kw.get("tube_rotation_deg", 0)
        p.dap = kw.get("dap", 0)
        p.model_heel = int(kw.get("model_heel", True))
        p.histories = kw.get("histories", 1000000)
        p.exposures = kw.get("exposures", 10)
        _chk(self.lib.dxs_source_dx(self.h, C.byref(p)), "dxs_source_dx")
        return self

    def source_ct(self, spiral=True, **kw):
        p = CTParams()
        p.tube = Tube(kw.get("voltage", 0), kw.get("anode_angle_deg", 0), kw.get("al_mm", 0), kw.get("cu_mm", 0),
                      kw.get("sn_mm", 0), kw.get("energy_resolution", 0))
        p.spiral = int(spiral)
        p.position[:] = kw.get("position", (0, 0, 0))
        p.cosines[:] = kw.get("cosines", (0, 0, 0, 0, 0, 0))
        for k in ("sdd", "collimation", "fov", "start_angle_deg", "exposure_step_deg", "scan_length", "pitch", "step",
                  "gantry_tilt_deg", "ctdi_vol", "xcare_filter_angle_deg", "xcare_span_deg", "xcare_ramp_deg",
                  "xcare_low_weight"):
            setattr(p, k, kw.get(k, 0))
        p.ctdi_phantom_diameter = kw.get("ctdi_phantom_diameter", 0)
        p.model_heel = int(kw.get("model_heel", True))
        p.use_xcare = int(kw.get("use_xcare", False))
        p.histories = kw.get("histories", 1000000)
        _chk(self.lib.dxs_source_ct(self.h, C.byref(p)), "dxs_source_ct")
        return self

    def _ct_params(self, spiral, kw):
        p = CTParams()
        p.tube = Tube(kw.get("voltage", 0), kw.get("anode_angle_deg", 0), kw.get("al_mm", 0), kw.get("cu_mm", 0),
                      kw.get("sn_mm", 0), kw.get("energy_resolution", 0))
        p.spiral = int(spiral)
        p.position[:] = kw.get("position", (0, 0, 0))
        p.cosines[:] = kw.get("cosines", (0, 0, 0, 0, 0, 0))
        for k in ("sdd", "collimation", "fov", "start_angle_deg", "exposure_step_deg", "scan_length", "pitch", "step",
                  "gantry_tilt_deg", "ctdi_vol", "xcare_filter_angle_deg", "xcare_span_deg", "xcare_ramp_deg",
                  "xcare_low_weight"):
            setattr(p, k, kw.get(k, 0))
        p.ctdi_phantom_diameter = kw.get("ctdi_phantom_diameter", 0)
        p.model_heel = int(kw.get("model_heel", True))
        p.use_xcare = int(kw.get("use_xcare", False))
        p.histories = kw.get("histories", 1000000)
        return p

    def source_ct_dual(self, spiral=True, **kw):
        """CTSpiralDualSource / CTAxialDualSource; tube B: voltage_b, al_mm_b, sdd_b, fov_b, start_angle_b_deg, mas_a, mas_b."""
        d = CTDualParams()
        d.a = self._ct_params(spiral, kw)
        d.tube_b = Tube(kw.get("voltage_b", 0), kw.get("anode_angle_deg_b", 0), kw.get("al_mm_b", 0), kw.get("cu_mm_b", 0), kw.get("sn_mm_b", 0), 0)
        for k in ("sdd_b", "fov_b", "start_angle_b_deg", "mas_a", "mas_b"):
            setattr(d, k, kw.get(k, 0))
        _chk(self.lib.dxs_source_ct_dual(self.h, C.byref(d)), "dxs_source_ct_dual")
        return self

    def source_topogram(self, **kw):
        p = self._ct_params(False, kw)
        _chk(self.lib.dxs_source_topogram(self.h, C.byref(p)), "dxs_source_topogram")
        return self

    def source_cbct(self, **kw):
        c = CBCTParams()
        p = c.dx
        p.tube = Tube(kw.get("voltage", 0), kw.get("anode_angle_deg", 0), kw.get("al_mm", 0), kw.get("cu_mm", 0), kw.get("sn_mm", 0), kw.get("energy_resolution", 0))
        p.position[:] = kw.get("position", (0, 0, 0))
        p.sdd = kw.get("sdd", 0)
        p.field_size[:] = kw.get("field_size", (0, 0))
        p.source_angles_deg[:] = kw.get("source_angles_deg", (0, 0))
        p.tube_rotation_deg = kw.get("tube_rotation_deg", 0)
        p.dap = kw.get("dap", 0)
        p.model_heel = int(kw.get("model_heel", True))
        p.histories = kw.get("histories", 1000000)
        c.span_deg = kw.get("span_deg", 0)
        c.step_deg = kw.get("step_deg", 0)
        _chk(self.lib.dxs_source_cbct(self.h, C.byref(c)), "dxs_source_cbct")
        return self

    def source_bowtie(self, angles, weights):
        a, w = _f32(angles), _f32(weights)
        _chk(self.lib.dxs_source_bowtie(self.h, int(a.size), a.ctypes.data_as(_f32p), w.ctypes.data_as(_f32p)),
             "dxs_source_bowtie")
        return self

    def source_aec(self, profile):
        a = _f32(profile)
        _chk(self.lib.dxs_source_aec(self.h, int(a.size), a.ctypes.data_as(_f32p)), "dxs_source_aec")
        return self

    def total_exposures(self) -> int:
        n = C.c_uint64(0)
        _chk(self.lib.dxs_source_total_exposures(self.h, C.byref(n)), "dxs_source_total_exposures")
        return int(n.value)

    def max_energy(self) -> float:
        e = C.c_float(0)
        _chk(self.lib.dxs_source_max_energy(self.h, C.byref(e)), "dxs_source_max_energy")
        return float(e.value)

    def exposure(self, i) -> dict:
        e = Exposure()
        _chk(self.lib.dxs_source_exposure(self.h, C.c_uint64(i), C.byref(e)), "dxs_source_exposure")
        return {"position": np.array(e.position[:], np.float32), "cosines": np.array(e.cosines[:], np.float32),
                "beam_direction": np.array(e.beam_direction[:], np.float32),
                "collimation": np.array(e.collimation[:], np.float32), "weight": np.float32(e.weight),
                "mono_energy": np.float32(e.mono_energy), "has_spectrum": bool(e.has_spectrum), "has_heel": bool(e.has_heel),
                "has_bowtie": bool(e.has_bowtie), "histories": int(e.histories)}

    def source_table(self, what) -> np.ndarray:
        n = C.c_uint64(0)
        _chk(self.lib.dxs_source_table(self.h, int(what), None, C.byref(n)), "dxs_source_table")
        out = np.zeros(n.value, np.float32)
        if n.value:
            _chk(self.lib.dxs_source_table(self.h, int(what), out.ctypes.data_as(_f32p), C.byref(n)), "dxs_source_table")
        return out

    def spectrum(self):
        n = C.c_int(0)
        _chk(self.lib.dxs_source_spectrum(self.h, None, None, C.byref(n)), "dxs_source_spectrum")
        e = np.zeros(n.value, np.float32)
        w = np.zeros(n.value, np.float32)
        _chk(self.lib.dxs_source_spectrum(self.h, e.ctypes.data_as(_f32p), w.ctypes.data_as(_f32p), C.byref(n)),
             "dxs_source_spectrum")
        return e, w

    def calibration(self, model=MODEL_LIVERMORE) -> float:
        v = C.c_float(0)
        _chk(self.lib.dxs_source_calibration(self.h, model, C.byref(v)), "dxs_source_calibration")
        return float(v.value)

    # ---- transport
    def transport(self, model=MODEL_LIVERMORE, output=OUT_EV_PER_HISTORY, use_calibration=False, seed=0, workers=0,
                  want_events=True, want_variance=True) -> Result:
        n = int(np.prod(self.dim))
        dose = np.zeros(n, np.float32)
        ev = np.zeros(n, np.uint32) if want_events else None
        var = np.zeros(n, np.float32) if want_variance else None
        info = ResultInfo()
        _chk(self.lib.dxs_transport(self.h, model, output, int(use_calibration), C.c_uint64(seed), int(workers),
                                    dose.ctypes.data_as(_f32p), None if ev is None else ev.ctypes.data_as(_u32p),
                                    None if var is None else var.ctypes.data_as(_f32p), C.byref(info)), "dxs_transport")
        return Result(dose, ev, var, int(info.histories), float(info.seconds), info.units.decode())

    def transport_monitored(self, model=MODEL_LIVERMORE, output=OUT_EV_PER_HISTORY, use_calibration=False, seed=0, workers=0,
                            cancel_at_percent=0.0):
        """Transport::operator() with a ProgressBar polled (and optionally cancelled) from a second thread; returns
        (Result, report dict)."""
        n = int(np.prod(self.dim))
        dose, ev, var = np.zeros(n, np.float32), np.zeros(n, np.uint32), np.zeros(n, np.float32)
        info, rep = ResultInfo(), ProgressReport()
        _chk(self.lib.dxs_transport_monitored(self.h, model, output, int(use_calibration), C.c_uint64(seed), int(workers),
                                              C.c_double(cancel_at_percent), dose.ctypes.data_as(_f32p), ev.ctypes.data_as(_u32p),
                                              var.ctypes.data_as(_f32p), C.byref(info), C.byref(rep)), "dxs_transport_monitored")
        report = {k: getattr(rep, k) for k, _ in ProgressReport._fields_}
        report["eta"] = rep.eta.decode(errors="replace")
        return Result(dose, ev, var, int(info.histories), float(info.seconds), info.units.decode()), report

    # ---- B200 extensions (product library only)
    def b200_set_devices(self, devices):
        """Spread subsequent transport() calls over these GPUs of the machine (Transport::setDevices); [] = one GPU."""
        arr = (C.c_int * max(len(devices), 1))(*devices)
        _chk(self.lib.dxs_b200_set_devices(self.h, len(devices), arr), "dxs_b200_set_devices")
        return self

    def b200_prepare(self, device=0, model=MODEL_LIVERMORE, seed=0, total_histories_all_ranks=0):
        _chk(self.lib.dxs_b200_prepare(self.h, int(device), int(model), C.c_uint64(seed), C.c_uint64(total_histories_all_ranks)),
             "dxs_b200_prepare")
        return self

    def b200_run(self, exp_begin, exp_end) -> float:
        """Transport exposures [exp_begin, exp_end); returns the CUDA-event time of the kernels in ms."""
        ms = C.c_double(0)
        _chk(self.lib.dxs_b200_run(self.h, C.c_uint64(exp_begin), C.c_uint64(exp_end), C.byref(ms)), "dxs_b200_run")
        return float(ms.value)

    def b200_run_strided(self, exp_first, exp_stride, exp_count) -> float:
        """Transport exposures exp_first + k * exp_stride, k < exp_count (the interleaved multi-GPU partition)."""
        ms = C.c_double(0)
        _chk(self.lib.dxs_b200_run_strided(self.h, C.c_uint64(exp_first), C.c_uint64(exp_stride), C.c_uint64(exp_count), C.byref(ms)),
             "dxs_b200_run_strided")
        return float(ms.value)

    def b200_collect(self, output=OUT_EV_PER_HISTORY, use_calibration=False, histories=0, want_events=True,
                     want_variance=True) -> Result:
        n = int(np.prod(self.dim))
        dose = np.empty(n, np.float32)  # every element is written by the download
        ev = np.empty(n, np.uint32) if want_events else None
        var = np.empty(n, np.float32) if want_variance else None
        info = ResultInfo()
        _chk(self.lib.dxs_b200_collect(self.h, output, int(use_calibration), C.c_uint64(histories), dose.ctypes.data_as(_f32p),
                                       None if ev is None else ev.ctypes.data_as(_u32p),
                                       None if var is None else var.ctypes.data_as(_f32p), C.byref(info)), "dxs_b200_collect")
        return Result(dose, ev, var, int(info.histories), 0.0, info.units.decode())

    def b200_context(self) -> C.c_void_p:
        ctx = C.c_void_p()
        _chk(self.lib.dxs_b200_context(self.h, C.byref(ctx)), "dxs_b200_context")
        return ctx

    def b200_release(self):
        _chk(self.lib.dxs_b200_release(self.h), "dxs_b200_release")

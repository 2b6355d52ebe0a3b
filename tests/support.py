"""Shared helpers of the test suite: scene builders (the same calls go to the product and to the reference
library), flattening of a scene into the plain-data inputs of include/dxmcb200.h, and statistics helpers."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dxmclib_b200 import cabi  # noqa: E402
from dxmclib_b200 import scene as S  # noqa: E402

SEED = 20261017

# TG-195 style tissue formulas the reference's validation uses (validation/validation.cpp:320-322, 1321-1340)
AIR = "C0.0150228136551869N78.439632744437O21.0780510531616Ar0.467293388746132"
SOFT = "H62.9539171935344C12.9077870263354N1.16702581276482O22.7840718642933Na0.026328553360443P0.0390933975009805S0.0566470278101205Cl0.0341543557411274K0.0309747686593447"
BONE = "H39.229963C15.009010N3.487490O31.621690Na0.050590Mg0.095705P3.867606S0.108832Ca6.529115"
THYROID = "H63.845575C6.130922N1.060311O28.814466Na0.053834P0.019979S0.019302Cl0.034909K0.015827I0.004876"

# a 120 kVp-like spectrum (bin lower edges keV, relative weights) for IsotropicSource::setSpecter
SPECTRUM_E = np.arange(15.0, 121.0, 5.0, dtype=np.float32)
SPECTRUM_W = np.array([0.2, 1.5, 4.5, 7.8, 10.0, 11.2, 11.4, 10.9, 14.5, 12.8, 8.2, 7.0, 5.9, 4.9, 4.0, 3.2, 2.4, 1.7, 1.0, 0.5, 0.2, 0.05],
                      dtype=np.float32)


def have_reference() -> bool:
    return os.path.exists(S.REFERENCE_LIB)


def have_gpu() -> bool:
    try:
        return cabi.device_count() > 0
    except Exception:
        return False


# ---------------------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------------------
def layered_box(lib, n=32):
    """The reference example's geometry (examples/pencilbeam/pencilbeam.cpp:21-58): air / water / aluminium thirds."""
    sc = S.Scene(lib)
    sc.world((n, n, n), (1, 1, 1))
    sc.add_material("Air, Dry (near sea level)").add_material("Water, Liquid").add_element(13)
    d = [sc.material_density(i) for i in range(3)]
    mat = np.zeros((n, n, n), np.uint8)
    mat[n // 3: 2 * n // 3] = 1
    mat[2 * n // 3:] = 2
    sc.arrays(np.array(d, np.float32)[mat], mat)
    assert sc.validate()
    return sc


def pencil_scene(lib, n=32, histories=20000, exposures=4, energy=60.0):
    sc = layered_box(lib, n)
    sc.source_pencil((0.3, 0.2, -float(n)), (1, 0, 0, 0, 1, 0), energy, histories, exposures)
    return sc


def tissue_block(lib, dim=(24, 20, 28), spacing=(4.0, 5.0, 3.0), origin=(3.0, -2.0, 10.0), forced=False, cosines=(1, 0, 0, 0, 1, 0)):
    """Soft tissue block with a bone insert, an iodine-loaded insert and varying density; anisotropic voxels,
    off-centre origin. With forced=True a slab of voxels is flagged in the measurement map."""
    sc = S.Scene(lib)
    sc.world(dim, spacing, origin, cosines)
    sc.add_material(AIR, 0.001205).add_material(SOFT, 1.03).add_material(BONE, 1.92).add_material(THYROID, 1.05)
    nx, ny, nz = dim
    mat = np.ones((nz, ny, nx), np.uint8)
    mat[:2] = 0
    mat[nz // 3: nz // 2, ny // 4: ny // 2, nx // 4: nx // 2] = 2
    mat[nz // 2: 3 * nz // 4, ny // 2: 3 * ny // 4, nx // 2: 3 * nx // 4] = 3
    dens = np.array([0.001205, 1.03, 1.92, 1.05], np.float32)[mat]
    zz = np.arange(nz, dtype=np.float32)[:, None, None]
    dens = (dens * (1.0 + 0.1 * np.sin(zz / 3.0))).astype(np.float32)  # smoothly varying density
    meas = None
    if forced:
        meas = np.zeros((nz, ny, nx), np.uint8)
        meas[nz // 2 - 2: nz // 2 + 2, :, nx // 3: 2 * nx // 3] = 1
    sc.arrays(dens, mat, meas)
    assert sc.validate()
    return sc


def dx_slab_scene(lib, histories=500000, exposures=8):
    """BASELINE config #2 in small: a 120 kVp tungsten-anode DX source (the built-in tube model with heel effect, 2.5 mm Al)
    onto a voxelised soft-tissue slab with a bone and a water insert, 10 mm voxels, tube 1 m above the isocentre (beam
    along -z)."""
    dim, sp = (24, 24, 20), (10.0, 10.0, 10.0)
    sc = S.Scene(lib)
    sc.world(dim, sp)
    sc.add_material(AIR, 0.001205).add_material(SOFT, 1.03).add_material(BONE, 1.92).add_material("Water, Liquid", 1.0)
    nx, ny, nz = dim
    mat = np.zeros((nz, ny, nx), np.uint8)
    mat[4:16, 2:22, 2:22] = 1  # 120 mm thick slab
    mat[7:11, 6:12, 6:18] = 2  # bone insert
    mat[7:13, 14:20, 6:18] = 3  # water insert
    sc.arrays(np.array([0.001205, 1.03, 1.92, 1.0], np.float32)[mat], mat)
    assert sc.validate()
    sc.source_dx(voltage=120.0, al_mm=2.5, sdd=1000.0, field_size=(180.0, 180.0), source_angles_deg=(0.0, 90.0), tube_rotation_deg=0.0,
                 dap=1.0, histories=histories, exposures=exposures, position=(0.0, 0.0, 0.0), model_heel=True)
    return sc


def isotropic_scene(lib, histories=20000, exposures=3, forced=False, ct=False, mono=None):
    sc = tissue_block(lib, forced=forced)
    if mono is not None:
        w, e = np.array([1.0], np.float32), np.array([mono], np.float32)
    else:
        w, e = SPECTRUM_W, SPECTRUM_E
    sc.source_isotropic((2.0, 1.0, -300.0) if not ct else (0.0, -300.0, 10.0), (1, 0, 0, 0, 1, 0) if not ct else (-1, 0, 0, 0, 0, 1),
                        (-0.12, 0.14, -0.10, 0.11), w, e, histories, exposures, ct=ct)
    return sc


def ct_scene(lib, spiral=True, histories=5000, aec=True, xcare=True, tilt=5.0, dim=(40, 40, 30), spacing=(10.0, 10.0, 8.0)):
    """Small CT configuration exercising tube spectrum, heel effect, bow-tie, AEC, XCare and gantry tilt."""
    from dxmclib_b200 import phantoms

    sc = S.Scene(lib)
    sc.world(dim, spacing)
    for name, dens in phantoms.ANTHROPOMORPHIC_MATERIALS:
        sc.add_material(name, dens)
    mat, dens = phantoms.anthropomorphic(dim, spacing)
    sc.arrays(dens, mat)
    assert sc.validate()
    scan = dim[2] * spacing[2]
    sc.source_ct(spiral=spiral, voltage=110.0, al_mm=6.0, cu_mm=0.1, sdd=1100.0, collimation=38.4, fov=480.0, pitch=0.9, step=38.4,
                 scan_length=scan, position=(2.0, -3.0, -scan / 2 if spiral else -scan / 2 + 19.2), exposure_step_deg=12.0,
                 start_angle_deg=20.0, gantry_tilt_deg=tilt, histories=histories, model_heel=True, ctdi_vol=12.0, use_xcare=xcare,
                 xcare_filter_angle_deg=180.0, xcare_span_deg=110.0, xcare_ramp_deg=15.0, xcare_low_weight=0.55)
    a, w = phantoms.bowtie_profile()
    sc.source_bowtie(a, w)
    if aec:
        z = np.arange(dim[2], dtype=np.float32)
        sc.source_aec(1.0 + 0.5 * np.sin(2 * np.pi * z / dim[2]))
    return sc


def air_gap_scene(lib, histories=20000, exposures=4, forced=False, dim=(40, 36, 44), spacing=(2.0, 2.5, 2.0)):
    """A tissue body with a bone core, a lung-like insert and an internal air cavity, surrounded by a wide margin of air, lit by
    a wide isotropic spectrum source from outside the world: most bricks of the empty-space traversal are air bricks, photons
    are born into them, leave through them and cross the cavity. With forced=True a slab is flagged in the measurement map."""
    sc = S.Scene(lib)
    sc.world(dim, spacing)
    sc.add_material(AIR, 0.001205).add_material(SOFT, 1.03).add_material(BONE, 1.92).add_material(SOFT, 0.26)
    nx, ny, nz = dim
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    r2 = ((x - nx / 2 + 0.5) / (nx * 0.28)) ** 2 + ((y - ny / 2 + 0.5) / (ny * 0.30)) ** 2
    body = (r2 <= 1.0) & (z >= nz // 6) & (z < nz - nz // 6)
    mat = np.zeros((nz, ny, nx), np.uint8)
    mat[body] = 1
    mat[body & (r2 <= 0.08)] = 2
    mat[body & (np.abs(x - nx * 0.62) < nx * 0.06) & (np.abs(y - ny / 2) < ny * 0.12)] = 3
    mat[body & (np.abs(x - nx * 0.36) < nx * 0.05) & (np.abs(y - ny / 2) < ny * 0.08) & (np.abs(z - nz / 2) < nz * 0.2)] = 0  # cavity
    dens = np.array([0.001205, 1.03, 1.92, 0.26], np.float32)[mat]
    meas = None
    if forced:
        meas = np.zeros((nz, ny, nx), np.uint8)
        meas[nz // 2 - 1: nz // 2 + 1, ny // 3: 2 * ny // 3, nx // 3: 2 * nx // 3] = 1
    sc.arrays(dens, mat, meas)
    assert sc.validate()
    sc.source_isotropic((1.0, -260.0, 2.0), (-1, 0, 0, 0, 0, 1), (-0.16, 0.16, -0.17, 0.17), SPECTRUM_W, SPECTRUM_E, histories, exposures, ct=True)
    return sc


def ct_dual_scene(lib, spiral=True, histories=100):
    sc = tissue_block(lib)
    sc.source_ct_dual(spiral=spiral, voltage=120, al_mm=6, voltage_b=80, al_mm_b=4, sdd=1100, sdd_b=1000, fov=480, fov_b=300, collimation=38.4,
                      pitch=0.8, step=30.0, scan_length=90, position=(1, 2, 3), exposure_step_deg=10, start_angle_deg=15, start_angle_b_deg=110,
                      gantry_tilt_deg=7, mas_a=80, mas_b=120, histories=histories, use_xcare=True, xcare_filter_angle_deg=90, xcare_span_deg=100,
                      xcare_ramp_deg=20, xcare_low_weight=0.6)
    return sc


def topogram_scene(lib, histories=50):
    sc = tissue_block(lib)
    sc.source_topogram(voltage=100, al_mm=5, sdd=1000, fov=400, collimation=20, scan_length=57.5, position=(1, -2, 3), start_angle_deg=33,
                       gantry_tilt_deg=-4, histories=histories)
    return sc


def cbct_scene(lib, histories=70):
    sc = tissue_block(lib)
    sc.source_cbct(voltage=90, al_mm=3, sdd=700, field_size=(200, 150), source_angles_deg=(25, -12), tube_rotation_deg=10, dap=1.5,
                   histories=histories, position=(2, 3, -4), span_deg=200, step_deg=7)
    return sc


def ctdi_scene(lib, histories=4000, diameter=160):
    sc = S.Scene(lib)
    sc.ctdi_phantom(diameter)
    sc.source_ct(spiral=False, voltage=120.0, al_mm=7.0, collimation=40.0, scan_length=40.0, position=(0, 0, 0), exposure_step_deg=10.0,
                 histories=histories, model_heel=True, ctdi_phantom_diameter=diameter)
    return sc


# ---------------------------------------------------------------------------------------------------------
# scene -> plain data (include/dxmcb200.h layout)
# ---------------------------------------------------------------------------------------------------------
def flatten_scene(sc: S.Scene, max_energy=None) -> dict:
    """Everything dxmcb200_set_world / set_luts / set_beam_tables / run need, read back through the scene API
    (works for the product and for the reference library)."""
    dim, spacing, ext = sc.dimensions()
    dens, mat, meas = sc.get_arrays()
    sc.lut_generate(float(max_energy if max_energy is not None else sc.max_energy()))
    knots, coeff, maxc, lin, rita, spl = (sc.lut_table(i) for i in range(6))
    n_seg = knots.size
    n_mat = coeff.size // (n_seg * 6)
    raw = spl.reshape(n_mat, 79)  # 60 coefficients, 16 knots, step, start, stop
    spline = np.zeros((n_mat, 63), np.float32)
    spline[:, :60] = raw[:, :60]
    spline[:, 60], spline[:, 61], spline[:, 62] = raw[:, 77], raw[:, 76], raw[:, 78]
    shells = np.stack([sc.material_shells(i)[:, :11].astype(np.float32) for i in range(n_mat)])
    flat = {
        "dim": dim, "spacing": spacing, "extent_safe": ext, "density": dens, "material": mat,
        "measurement": meas if meas.any() else None,
        "luts": {"n_materials": n_mat, "n_segments": n_seg, "linear_index": int(lin[0]), "linear_step": float(lin[1]),
                 "linear_energy": float(lin[2]), "knots": knots, "coefficients": coeff, "max_coefficients": maxc, "rita": rita,
                 "spline": np.ascontiguousarray(spline), "shells": np.ascontiguousarray(shells)},
        "spectra": [], "heels": [], "bowties": [],
    }
    try:
        t = [sc.source_table(i) for i in range(7)]
    except S.SceneError:  # scene without a source: world + LUTs only
        t = [np.zeros(0, np.float32)] * 7
    if t[0].size:
        flat["spectra"].append((t[0], t[1].astype(np.uint32), t[2]))
    if t[3].size:
        d = t[3]
        flat["heels"].append((float(d[0]), float(d[1]), int(d[2]), float(d[3]), float(d[4]), int(d[5]), t[4]))
    if t[5].size:
        flat["bowties"].append((t[5], t[6]))
    return flat


def exposures_of(sc: S.Scene, n=None):
    out = []
    total = sc.total_exposures()
    for i in range(total if n is None else min(n, total)):
        e = sc.exposure(i)
        x = cabi.Exposure()
        x.position[:] = e["position"].tolist()
        x.cosines[:] = e["cosines"].tolist()
        x.beam_direction[:] = e["beam_direction"].tolist()
        x.collimation[:] = e["collimation"].tolist()
        x.weight = float(e["weight"])
        x.mono_energy = float(e["mono_energy"])
        x.spectrum = 0 if e["has_spectrum"] else -1
        x.heel = 0 if e["has_heel"] else -1
        x.bowtie = 0 if e["has_bowtie"] else -1
        x.histories = e["histories"]
        out.append(x)
    return out


def load_context(ctx: cabi.Context, flat: dict):
    """Upload a flattened scene into a dxmcb200_ctx through the C ABI."""
    import ctypes as C

    f32p = C.POINTER(C.c_float)
    ctx.set_world(flat["dim"], flat["spacing"], flat["extent_safe"], flat["density"], flat["material"], flat["measurement"])
    lt = flat["luts"]
    l = cabi.Luts()
    l.n_materials, l.n_segments, l.linear_index = lt["n_materials"], lt["n_segments"], lt["linear_index"]
    l.linear_step, l.linear_energy = lt["linear_step"], lt["linear_energy"]
    for k in ("knots", "coefficients", "max_coefficients", "rita", "spline", "shells"):
        setattr(l, k, lt[k].ctypes.data_as(f32p))
    ctx._chk(ctx.l.dxmcb200_set_luts(ctx.h, C.byref(l)), "dxmcb200_set_luts")
    ctx.n_materials = lt["n_materials"]
    ctx.set_beam_tables(flat["spectra"], flat["heels"], flat["bowties"])
    ctx._flat = flat


# ---------------------------------------------------------------------------------------------------------
# statistics
# ---------------------------------------------------------------------------------------------------------
def bit_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def dose_sigma(sum_e, sum_e2, n_events):
    """Standard deviation of a voxel's summed energy from the per-event second moment: var(sum) ~ sum(e^2) (Poisson-like
    event counts), the estimator the reference's variance array feeds."""
    return np.sqrt(np.maximum(sum_e2, 0.0))


def compare_dose(a_sum, a_sum2, b_sum, b_sum2, rel_err_limit=0.02, n_sigma=3.0):
    """north_star criterion: per-voxel agreement within n_sigma of the combined Monte Carlo uncertainty in all voxels whose
    relative error is under rel_err_limit. Returns (fraction outside, number tested, worst z)."""
    sa, sb = dose_sigma(a_sum, a_sum2, None), dose_sigma(b_sum, b_sum2, None)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(b_sum > 0, sb / b_sum, np.inf)
    sel = rel < rel_err_limit
    if not sel.any():
        return 0.0, 0, 0.0
    z = np.abs(a_sum[sel] - b_sum[sel]) / np.sqrt(sa[sel] ** 2 + sb[sel] ** 2)
    return float((z > n_sigma).mean()), int(sel.sum()), float(z.max())


def allowed_outliers(tested, n_sigma_tail=0.0027):
    """Voxels allowed beyond 3 sigma when two correct, statistically independent runs are compared over `tested` voxels: the
    Gaussian tail puts 0.27 % of them there, the count is binomial, the bound is its mean plus three standard deviations (+1)."""
    expected = n_sigma_tail * tested
    return expected + 3.0 * np.sqrt(expected) + 1.0


def air_run_rays(flat, n, seed):
    """Fixed rays for the traversal tests: sources on a sphere around the grid aimed at random points inside it; a tenth of the
    rays run along an axis (two zero direction components)."""
    rng = np.random.default_rng(seed)
    ext = np.array(flat["extent_safe"], np.float64)
    lo, hi = ext[0::2], ext[1::2]
    centre, half = (lo + hi) / 2, (hi - lo) / 2
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    pos = centre + u * 2.5 * np.linalg.norm(half)
    target = lo + rng.uniform(0.02, 0.98, (n, 3)) * (hi - lo)
    d = target - pos
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 10
    axis = rng.integers(0, 3, k)
    pos[:k] = target[:k]
    pos[np.arange(k), axis] = lo[axis] - 50.0
    d[:k] = 0.0
    d[np.arange(k), axis] = 1.0
    return pos.astype(np.float32), d.astype(np.float32)


def assert_air_runs_true(flat, bricks, entry, d32, result, max_cubes=3):
    """Ground truth for trace_air_runs by dense sampling of the brick flags along each ray: every point of a run lies in an air
    brick (or outside the grid), and a run that ended neither at the grid's edge nor at the cube cap stops in front of a non-air
    brick. `entry`: the rays' entry points into the world (trace_indices)."""
    length, crossed, exits, in_air, reaches, _ = result
    dim, sp, ext = np.array(flat["dim"], np.int64), np.array(flat["spacing"], np.float64), np.array(flat["extent_safe"], np.float64)
    lo = ext[0::2]
    shift, nb = np.array(bricks["shift"]), np.array(bricks["nb"])
    air = bricks["air"].reshape(nb[2], nb[1], nb[0]).astype(bool)
    assert reaches.mean() > 0.95 and in_air.mean() > 0.3 and (crossed[in_air == 1] >= 1).all() and crossed.max() <= max_cubes

    def is_air(points):  # [m, 3] -> air brick or outside the grid
        vox = np.floor((points - lo) / sp).astype(np.int64)
        outside = ((vox < 0) | (vox >= dim)).any(axis=1)
        b = np.clip(vox, 0, dim - 1) >> shift
        return outside | air[b[:, 2], b[:, 1], b[:, 0]]

    sel = np.flatnonzero((in_air == 1) & np.isfinite(length))
    e64, d64, len64 = entry[sel].astype(np.float64), d32[sel].astype(np.float64), length[sel].astype(np.float64)
    for f in np.linspace(0.001, 0.999, 97):  # points strictly inside the run
        assert is_air(e64 + d64 * (len64 * f)[:, None]).all(), f"a run crosses a non-air brick at {f:.3f} of its length"
    stopped = (exits[sel] == 0) & (crossed[sel] < max_cubes)
    assert stopped.sum() > 200
    beyond = e64[stopped] + d64[stopped] * (len64[stopped] + 0.01 * sp.min())[:, None]
    assert (~is_air(beyond)).mean() > 0.999, "runs that stopped inside the grid must stop in front of a non-air brick"

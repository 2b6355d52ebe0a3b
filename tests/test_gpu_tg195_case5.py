"""AAPM TG-195 Case 5 (CT of the voxelised anthropomorphic phantom, 500 x 320 x 260 voxels of 1 mm, 20 materials) and the
polychromatic / 15 degree / other-model variants of Case 2, as the reference's validation program sets them up
(validation/validation.cpp:1308-1573 Case 5; :420-470, :496-512 Case 2 variants; :1575-1627 the model loop).

The phantom and the tabulated TG-195 spectra are committed fixtures (tests/golden/case5world.tar.gz, tg195_spectra.npz, written
by tests/golden/make_tg195_fixtures.py; the GPU box has no /root/reference).

Checks: (1) product (default tracking: Woodcock + empty-space traversal) vs the unmodified reference on identical inputs:
deposited energy in the body within 0.5 %, every organ / VOI within 3.5 sigma of the combined Monte Carlo uncertainty (sigma
measured from independent product replicas); (2) published TG-195 values as an informational bound (both sides run on the
approximate xrl_lite cross sections)."""
import io
import math
import os
import tarfile

import numpy as np
import pytest

import support as T
from dxmclib_b200 import scene as S
from test_gpu_tg195 import case2_scene, voi_sums

pytestmark = pytest.mark.gpu
G = os.path.join(T.ROOT, "tests", "golden")

# validation.cpp:1321-1340 (number-density formulas, name, density)
CASE5_MATERIALS = [
    ("C0.015019N78.443071O21.074800Ar0.467110", 0.001205),  # 0 air
    ("H51.869709C36.108118N4.019974O8.002200", 0.075),  # 1 cushion foam
    ("C100.000000", 1.2),  # 2 carbon fibre
    ("H63.000070C12.890598N1.165843O22.756479Na0.026307P0.039052S0.056594Cl0.034118K0.030937", 1.03),  # 3 soft tissue
    ("H63.688796C7.143744N1.278063O27.701991Na0.026851P0.039859S0.038509Cl0.034823K0.047365", 1.05),  # 4 heart
    ("H63.731478C5.452396N1.380394O29.198156Na0.054259P0.040273S0.058363Cl0.052777K0.031904", 0.26),  # 5 lung
    ("H63.217465C7.229913N1.338082O27.958043Na0.054349P0.060510S0.058460Cl0.035243K0.047936", 1.06),  # 6 liver
    ("H63.000070C12.890598N1.165843O22.756479Na0.026307P0.039052S0.056594Cl0.034118K0.030937", 1.03),  # 7 gallbladder
    ("H63.655092C5.860784N1.423215O28.851671Na0.027097P0.060337S0.038862Cl0.035143K0.047799", 1.06),  # 8 spleen
    ("H64.343953C5.858428N0.961057O28.720940Na0.026615P0.019755S0.019085Cl0.034518K0.015650", 1.03),  # 9 stomach
    ("H64.343953C5.858428N0.961057O28.720940Na0.026615P0.019755S0.019085Cl0.034518K0.015650", 1.03),  # 10 large intestine
    ("H63.939187C8.555183N0.955012O26.374094Na0.052895P0.039261S0.018965Cl0.034300K0.031102", 1.04),  # 11 pancreas
    ("H63.000070C12.890598N1.165843O22.756479Na0.026307P0.039052S0.056594Cl0.034118K0.030937", 1.03),  # 12 adrenal
    ("H63.845575C6.130922N1.060311O28.814466Na0.053834P0.019979S0.019302Cl0.034909K0.015827I0.004876", 1.05),  # 13 thyroid
    ("H63.000070C12.890598N1.165843O22.756479Na0.026307P0.039052S0.056594Cl0.034118K0.030937", 1.03),  # 14 thymus
    ("H64.343953C5.858428N0.961057O28.720940Na0.026615P0.019755S0.019085Cl0.034518K0.015650", 1.03),  # 15 small intestine
    ("H64.343953C5.858428N0.961057O28.720940Na0.026615P0.019755S0.019085Cl0.034518K0.015650", 1.03),  # 16 esophagus
    ("H62.083429C10.628873N1.876505O25.228547Na0.054442P0.020204S0.039039Cl0.052955K0.016006", 1.09),  # 17 skin
    ("H61.873627C28.698524N0.675867O8.736110P0.004495S0.004342K0.003561Ca0.003473", 0.93),  # 18 breast
    ("H39.229963C15.009010N3.487490O31.621690Na0.050590Mg0.095705P3.867606S0.108832Ca6.529115", 1.92),  # 19 cortical bone
]
ORGANS = list(range(3, 20))  # validation.cpp:1436
# published organ energies [eV/history], validation.cpp:1441-1459 (discrete angles 0, 90) and :1526-1530 (continuous)
PUBLISHED = {
    ("mono", 0): [11574.28, 3086.42, 1301.17, 679.47, 6.37, 17.57, 134.40, 16.73, 8.79, 0.15, 1.74, 44.22, 10.72, 36.90, 456.36, 21.68, 8761.23],
    ("mono", 90): [9975.20, 786.73, 679.42, 495.34, 3.94, 6.27, 36.21, 3.15, 3.15, 0.10, 1.06, 16.90, 3.38, 23.03, 285.87, 6.80, 5611.52],
    ("120kv", 0): [12374.98, 2917.75, 1275.86, 612.31, 5.78, 16.68, 121.04, 15.16, 8.17, 0.15, 1.65, 40.66, 9.78, 33.37, 559.77, 21.49, 7727.77],
    ("mono", "ct"): [10410.69, 1670.94, 889.97, 438.66, 3.57, 22.80, 103.46, 11.89, 6.71, 0.14, 1.40, 21.02, 6.75, 29.55, 305.22, 9.88, 7854.65],
    ("120kv", "ct"): [11090.33, 1567.72, 852.32, 401.38, 3.39, 21.10, 94.86, 10.96, 6.33, 0.14, 1.34, 19.45, 6.43, 27.27, 370.97, 9.85, 6840.76],
}


def case5_phantom():
    with tarfile.open(os.path.join(G, "case5world.tar.gz")) as t:
        raw = t.extractfile(t.getmembers()[0]).read()
    mat = np.frombuffer(raw, np.uint8)
    assert mat.size == 500 * 320 * 260 and mat.max() == 19
    return mat


def spectrum(kind):
    if kind == "mono":
        return np.array([1.0], np.float32), np.array([56.4], np.float32)
    s = np.load(os.path.join(G, "tg195_spectra.npz"))
    return s["120kv_weight"], s["120kv_energy"]


def case5_scene(lib, mat, kind, angle, histories, exposures):
    """validation.cpp:1310-1347 (world), :1385-1417 and :1462-1480 / :1516-1521 (sources)."""
    sc = S.Scene(lib)
    sc.world((500, 320, 260), (1.0, 1.0, 1.0))
    for formula, density in CASE5_MATERIALS:
        sc.add_material(formula, density)
    dens = np.array([d for _, d in CASE5_MATERIALS], np.float32)[mat]
    sc.arrays(dens, mat)
    assert sc.validate()
    w, e = spectrum(kind)
    coll = (-math.atan(250.0 / 600.0), math.atan(250.0 / 600.0), -math.atan(5.0 / 600.0), math.atan(5.0 / 600.0))
    if angle == "ct":
        sc.source_isotropic((0.0, -600.0, 0.0), (-1, 0, 0, 0, 0, 1), coll, w, e, histories, exposures, ct=True)
    else:
        rad = -math.radians(angle)  # vectormath::rotate about z by -angle (validation.cpp:1470-1474)
        c, s_ = math.cos(rad), math.sin(rad)

        def rot(v):
            return (c * v[0] - s_ * v[1], s_ * v[0] + c * v[1], v[2])

        pos = rot((0.0, -600.0, 0.0))
        cos = rot((-1.0, 0.0, 0.0)) + rot((0.0, 0.0, 1.0))
        sc.source_isotropic(pos, cos, coll, w, e, histories, exposures)
    return sc


def organ_sums(result, mat):
    d = result.dose.astype(np.float64)  # eV / history
    return np.bincount(mat, weights=d, minlength=20)[ORGANS], float(d[mat >= 3].sum())


@pytest.mark.parametrize("kind,angle,model", [
    ("mono", 0, S.MODEL_LIVERMORE), ("mono", 90, S.MODEL_NONE), ("120kv", 0, S.MODEL_IA), ("mono", "ct", S.MODEL_LIVERMORE),
    ("120kv", "ct", S.MODEL_LIVERMORE),
])
def test_tg195_case5_against_reference_and_published(gpu, product, reference, kind, angle, model):
    mat = case5_phantom()
    replicas, exposures = 8, 36 if angle == "ct" else 4
    per_exposure = 1_000_000 if angle != "ct" else 120_000
    organs, body = [], []
    for r in range(replicas):
        sc = case5_scene(product, mat, kind, angle, per_exposure, exposures)
        res = sc.transport(model=model, output=S.OUT_EV_PER_HISTORY, seed=T.SEED + 31 * r)
        assert res.histories == per_exposure * exposures and res.units == "eV/history"
        o, b = organ_sums(res, mat)
        organs.append(o)
        body.append(b)
        sc.close()
    organs, body = np.array(organs), np.array(body)
    mean_a, sigma_rep = organs.mean(axis=0), organs.std(axis=0, ddof=1)
    # the unmodified reference on the same inputs, one run of a replica's size on the host cores
    sb = case5_scene(reference, mat, kind, angle, per_exposure, exposures)
    rb = sb.transport(model=model, output=S.OUT_EV_PER_HISTORY, seed=T.SEED + 5, workers=S.WORKERS_COUNTER_STREAMS)
    mean_b, body_b = organ_sums(rb, mat)
    sb.close()
    sigma_body = body.std(ddof=1) * math.sqrt(1.0 + 1.0 / replicas)
    assert abs(body.mean() - body_b) / body_b < 5e-3, (body.mean(), body_b)
    assert abs(body.mean() - body_b) < 4.0 * sigma_body
    sigma = sigma_rep * math.sqrt(1.0 + 1.0 / replicas)
    z = np.abs(mean_a - mean_b) / np.maximum(sigma, 1e-12)
    assert np.all(z < 3.5), (mean_a, mean_b, sigma, z)
    pub = np.array(PUBLISHED[(kind, angle)])
    big = pub > 100.0  # the small organs are a few eV per history: Monte Carlo noise at this history count, not physics
    print(f"TG-195 case 5 {kind} {angle}: organs product/published " + ", ".join(f"{a:.1f}/{p:.1f}" for a, p in zip(mean_a[big], pub[big]))
          + f"; worst z vs reference {z.max():.2f}; body {body.mean():.1f} vs reference {body_b:.1f} eV/history")
    assert np.all(np.abs(mean_a[big] - pub[big]) / pub[big] < 0.15)


# ---- Case 2 variants: 120 kVp spectrum, 15 degree incidence, the other low-energy models ---------------------------------
CASE2_PUBLISHED = {  # validation.cpp:496-512: total, VOI 1..9
    ("120kv", 0): (33125.98, [24.97, 24.95, 33.52, 24.96, 24.97, 72.70, 49.99, 21.73, 13.48]),
    ("120kv", 15): (30923.13, [30.35, 23.52, 31.64, 23.52, 8.90, 70.53, 47.74, 20.31, 12.51]),
    ("mono", 15): (30883.83, [33.0807985, 25.475272, 34.62570725, 25.50542125, 9.79069025, 70.80499875, 51.0616275, 22.2764985, 13.54431025]),
    ("mono", 0): (33171.4, [27.01, 27.00, 36.67, 27.01, 27.01, 72.86, 53.35, 23.83, 14.60]),
}


def case2_variant(lib, kind, angle, histories, exposures):
    """validation.cpp:429-470: spectrum choice and the 15 degree geometry (source lifted by h = 1800 tan 15 deg, beam tilted back
    onto the slab, asymmetric collimation so that the field still covers 390 x 390 mm at the slab)."""
    sc, mat = case2_scene(lib, histories, exposures)
    w, e = spectrum(kind)
    if angle == 0:
        half = math.atan(195.0 / 1800.0)
        sc.source_isotropic((0.0, 0.0, 0.0), (1, 0, 0, 0, 1, 0), (-half, half, -half, half), w, e, histories, exposures)
        return sc, mat
    a = math.radians(15.0)
    d = 1800.0
    h = d * math.tan(a)
    s = 195.0
    a2 = math.acos((h * (h - s) + d * d) / math.sqrt((h * h + d * d) * ((h - s) * (h - s) + d * d)))
    a1 = math.acos((h * (h + s) + d * d) / math.sqrt((h * h + d * d) * ((h + s) * (h + s) + d * d)))
    ang_x = math.atan(s / d)
    # vectormath::rotate of both cosine vectors about (-1, 0, 0) by 15 degrees (Rodrigues, validation.cpp:455-457)
    c, sn = math.cos(a), math.sin(a)

    def rot(v):  # axis k = (-1, 0, 0): k x v = (0, v[2], -v[1])
        kv = (0.0, v[2], -v[1])
        kdot = -v[0]
        k = (-1.0, 0.0, 0.0)
        return tuple(c * v[i] + (1 - c) * kdot * k[i] + sn * kv[i] for i in range(3))

    cos = rot((1.0, 0.0, 0.0)) + rot((0.0, 1.0, 0.0))
    sc.source_isotropic((0.0, -h, 0.0), cos, (-ang_x, ang_x, -a1, a2), w, e, histories, exposures)
    return sc, mat


@pytest.mark.parametrize("kind,angle,model", [
    ("120kv", 0, S.MODEL_LIVERMORE), ("120kv", 15, S.MODEL_LIVERMORE), ("mono", 15, S.MODEL_IA), ("mono", 0, S.MODEL_NONE),
])
def test_tg195_case2_variants(gpu, product, reference, kind, angle, model):
    replicas, per_replica = 8, 8 * 600_000
    voi, tissue = [], []
    mat = None
    for r in range(replicas):
        sc, mat = case2_variant(product, kind, angle, 600_000, 8)
        res = sc.transport(model=model, output=S.OUT_EV_PER_HISTORY, seed=T.SEED + 13 * r)
        assert res.histories == per_replica
        e, _, t = voi_sums(res, mat)
        voi.append(e / per_replica)
        tissue.append(t / per_replica)
        sc.close()
    voi, tissue = np.array(voi), np.array(tissue)
    sb, _ = case2_variant(reference, kind, angle, 600_000, 8)
    b = sb.transport(model=model, output=S.OUT_EV_PER_HISTORY, seed=T.SEED + 2, workers=S.WORKERS_COUNTER_STREAMS)
    eb, _, tb = voi_sums(b, mat)
    mean_a, mean_b = voi.mean(axis=0), eb / b.histories
    assert abs(tissue.mean() - tb / b.histories) / (tb / b.histories) < 5e-3
    sigma = voi.std(axis=0, ddof=1) * math.sqrt(1.0 + 1.0 / replicas)
    z = np.abs(mean_a - mean_b) / sigma
    assert np.all(z < 3.5), (mean_a * 1e3, mean_b * 1e3, sigma * 1e3, z)
    total_pub, voi_pub = CASE2_PUBLISHED[(kind, angle)]
    print(f"TG-195 case 2 {kind} {angle} deg model {model}: total {tissue.mean() * 1e3:.1f} (published {total_pub}); worst z vs reference {z.max():.2f}")
    assert abs(tissue.mean() * 1e3 - total_pub) / total_pub < 0.10
    assert np.all(np.abs(mean_a * 1e3 - np.array(voi_pub)) / np.array(voi_pub) < 0.20)

"""Deterministic synthetic phantoms for the BASELINE configs (no RNG, no data files).

`anthropomorphic(dim)` is the "synthetic 512x512x400 anthropomorphic segmentation (~10 materials)" of
BASELINE config #4: an analytic thorax/abdomen — elliptical torso with skin and fat layers, two lungs,
heart with blood pool, liver, spine and rib rings, lying on a carbon-fibre table — voxelised at 1 mm.
Materials use the compositions the reference's TG-195 case 5 uses for the same tissues
(validation/validation.cpp:1321-1340) or NIST names.
"""
from __future__ import annotations

import numpy as np

# (material string for Material(name, "", density), density g/cm3)
ANTHROPOMORPHIC_MATERIALS = [
    ("C0.015019N78.443071O21.074800Ar0.467110", 0.001205),  # 0 air
    ("H63.000070C12.890598N1.165843O22.756479Na0.026307P0.039052S0.056594Cl0.034118K0.030937", 1.03),  # 1 soft tissue
    ("Adipose Tissue (ICRP)", 0.92),  # 2 fat
    ("H62.083429C10.628873N1.876505O25.228547Na0.054442P0.020204S0.039039Cl0.052955K0.016006", 1.09),  # 3 skin
    ("H63.731478C5.452396N1.380394O29.198156Na0.054259P0.040273S0.058363Cl0.052777K0.031904", 0.26),  # 4 lung
    ("H63.688796C7.143744N1.278063O27.701991Na0.026851P0.039859S0.038509Cl0.034823K0.047365", 1.05),  # 5 heart muscle
    ("Blood (ICRP)", 1.06),  # 6 blood pool
    ("H39.229963C15.009010N3.487490O31.621690Na0.050590Mg0.095705P3.867606S0.108832Ca6.529115", 1.92),  # 7 cortical bone
    ("H63.217465C7.229913N1.338082O27.958043Na0.054349P0.060510S0.058460Cl0.035243K0.047936", 1.06),  # 8 liver
    ("C", 1.2),  # 9 carbon-fibre table
]


def anthropomorphic(dim=(512, 512, 400), spacing=(1.0, 1.0, 1.0)):
    """Returns (material u8 [nz,ny,nx] flattened x-fastest, density f32) for a world centred on the origin."""
    nx, ny, nz = (int(d) for d in dim)
    sx, sy, sz = (float(s) for s in spacing)
    x = ((np.arange(nx, dtype=np.float32) + 0.5) - nx / 2) * sx
    y = ((np.arange(ny, dtype=np.float32) + 0.5) - ny / 2) * sy
    z = ((np.arange(nz, dtype=np.float32) + 0.5) - nz / 2) * sz
    # geometry is defined for a 512 x 512 x 400 mm field and scales with the actual extent
    fx, fy, fz = nx * sx / 512.0, ny * sy / 512.0, nz * sz / 400.0
    X = x[None, None, :] / fx
    Y = y[None, :, None] / fy
    mat = np.zeros((nz, ny, nx), np.uint8)
    chunk = max(1, 8_000_000 // (nx * ny))
    for z0 in range(0, nz, chunk):
        Z = z[z0:z0 + chunk, None, None] / fz
        m = np.zeros((Z.shape[0], ny, nx), np.uint8)
        body = (X / 170.0) ** 2 + (Y / 115.0) ** 2
        m[np.broadcast_to(body <= 1.0, m.shape)] = 3  # skin
        m[np.broadcast_to((X / 168.0) ** 2 + (Y / 113.0) ** 2 <= 1.0, m.shape)] = 2  # subcutaneous fat
        m[np.broadcast_to((X / 158.0) ** 2 + (Y / 103.0) ** 2 <= 1.0, m.shape)] = 1  # soft tissue
        # rib rings: 10 mm thick every 25 mm in the upper 60 % of the volume
        ring = ((X / 150.0) ** 2 + (Y / 97.0) ** 2 <= 1.0) & ((X / 141.0) ** 2 + (Y / 88.0) ** 2 > 1.0)
        ribz = (np.mod(Z + 200.0, 25.0) < 10.0) & (Z > -60.0)
        m[ring & ribz] = 7
        # lungs
        for cx in (-75.0, 75.0):
            lung = ((X - cx) / 55.0) ** 2 + ((Y + 5.0) / 72.0) ** 2 + ((Z - 50.0) / 115.0) ** 2 <= 1.0
            m[lung] = 4
        # heart and blood pool
        m[((X - 15.0) / 45.0) ** 2 + ((Y + 25.0) / 38.0) ** 2 + ((Z - 30.0) / 50.0) ** 2 <= 1.0] = 5
        m[((X - 15.0) / 30.0) ** 2 + ((Y + 25.0) / 24.0) ** 2 + ((Z - 30.0) / 35.0) ** 2 <= 1.0] = 6
        # liver
        m[((X + 60.0) / 80.0) ** 2 + ((Y + 5.0) / 60.0) ** 2 + ((Z + 115.0) / 60.0) ** 2 <= 1.0] = 8
        # spine
        m[np.broadcast_to(X ** 2 + (Y - 70.0) ** 2 <= 20.0 ** 2, m.shape)] = 7
        # table
        m[np.broadcast_to((np.abs(X) <= 220.0) & (Y >= 122.0) & (Y <= 132.0), m.shape)] = 9
        mat[z0:z0 + chunk] = m
    dens_table = np.array([d for _, d in ANTHROPOMORPHIC_MATERIALS], np.float32)
    density = dens_table[mat]
    return mat.reshape(-1), density.reshape(-1)


def bowtie_profile(n=25, max_angle=0.45):
    """A generic body bow-tie: relative fluence vs |fan angle| (rad)."""
    a = np.linspace(0.0, max_angle, n, dtype=np.float32)
    w = np.exp(-3.0 * (a / 0.40) ** 2).astype(np.float32) + np.float32(0.05)
    return a, w


def water_cube(n=64, density=1.0):
    mat = np.zeros(n ** 3, np.uint8)
    dens = np.full(n ** 3, density, np.float32)
    return mat, dens

#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libdxmc_ref.so, built from
/root/reference by oracle/Makefile) in this container. The vectors travel with the repo so the GPU box — where
/root/reference does not exist — can check the CUDA path and the restatement against reference outputs.

    python tests/golden/make_golden.py

Contents
  traces.npz    voxel-index sequences of fixed rays through awkward worlds (reference transportParticleToWorld /
                particleInsideWorld / indexFromPosition, transport.hpp:485-521, 702-728)
  lut.npz       photoComptRayAttenuation / maxTotalAttenuationInverse / comptonScatterFactor on an energy grid
                (attenuationinterpolator.hpp:207-248, interpolation.hpp:169-176)
  transport_*.npz  seeded single-worker reference runs (sequential RandomState): depth profiles, totals, event counts
  streams_*.npz    reference runs with the product's per-history stream keying (counter streams): full grids of small scenes
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import support as T  # noqa: E402
from dxmclib_b200 import scene as S  # noqa: E402

ref = S.reference_lib()


def trace_worlds():
    return {
        "unit": dict(dim=(64, 48, 40), spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)),
        "aniso": dict(dim=(37, 51, 29), spacing=(0.7, 1.3, 2.5), origin=(12.5, -7.25, 100.0)),
        "fine": dict(dim=(200, 10, 10), spacing=(0.1, 3.0, 3.0), origin=(-3.0, 0.5, 0.25)),
    }


def rays_for(dim, spacing, origin, n=400, seed=1):
    rng = np.random.default_rng(seed)
    ext = np.array(dim) * np.array(spacing) / 2
    pos = (np.array(origin) + rng.uniform(-2.5, 2.5, (n, 3)) * ext).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    aim = (np.array(origin) + rng.uniform(-0.9, 0.9, (n, 3)) * ext) - pos  # half of the rays are aimed into the volume
    aim /= np.linalg.norm(aim, axis=1, keepdims=True)
    d[n // 2:] = aim[n // 2:]
    d[: n // 10] = np.eye(3)[rng.integers(0, 3, n // 10)] * rng.choice([-1.0, 1.0], (n // 10, 1))  # axis-aligned rays
    # rays starting exactly on voxel boundaries inside the world
    k = n // 10
    pos[n // 10: n // 10 + k] = (np.array(origin) + (rng.integers(-3, 4, (k, 3)) * np.array(spacing))).astype(np.float32)
    steps = np.concatenate([rng.exponential(3.0, 40), [0.0, 1e-6, 0.5, 1.0, 2.5]]).astype(np.float32)
    return pos, d.astype(np.float32), steps


def world_scene(lib, dim, spacing, origin):
    sc = S.Scene(lib)
    sc.world(dim, spacing, origin)
    sc.add_material("Water, Liquid")
    n = int(np.prod(dim))
    sc.arrays(np.ones(n, np.float32), np.zeros(n, np.uint8))
    assert sc.validate()
    return sc


def main():
    out = {}
    for name, w in trace_worlds().items():
        sc = world_scene(ref, **w)
        pos, d, steps = rays_for(**w)
        idx, entry = sc.trace_indices(pos, d, steps)
        out[f"{name}_pos"], out[f"{name}_dir"], out[f"{name}_steps"] = pos, d, steps
        out[f"{name}_idx"], out[f"{name}_entry"] = idx, entry
        print(name, "rays hitting:", int((idx[:, 0] >= 0).sum()), "of", len(pos))
    np.savez_compressed(os.path.join(HERE, "traces.npz"), **out)

    sc = T.tissue_block(ref)
    sc.lut_generate(140.0)
    e = np.unique(np.concatenate([np.geomspace(1.0, 140.0, 160), [4.0385, 4.0386, 33.1694, 33.17, 60.0, 100.0]])).astype(np.float32)
    att = np.stack([[sc.lut_attenuation(m, x) for x in e] for m in range(4)])
    mx = np.array([sc.lut_max_inverse(x) for x in e], np.float32)
    q = np.linspace(0.0, 12.0, 97).astype(np.float32)
    sf = np.stack([[sc.lut_scatter_factor(m, x) for x in q] for m in range(4)])
    np.savez_compressed(os.path.join(HERE, "lut.npz"), energy=e, attenuation=att, max_inverse=mx, q=q, scatter_factor=sf)

    def summarize(sc, r):
        nx, ny, nz = sc.dim
        d = r.dose.reshape(nz, ny, nx).astype(np.float64)
        return dict(total=d.sum(), profile_z=d.sum(axis=(1, 2)), profile_x=d.sum(axis=(0, 1)), events=int(r.n_events.sum()),
                    events_z=r.n_events.reshape(nz, ny, nx).sum(axis=(1, 2)), histories=r.histories)

    for name, build in (("pencil", lambda: T.pencil_scene(ref, histories=40000, exposures=4)),
                        ("isotropic_forced", lambda: T.isotropic_scene(ref, histories=30000, forced=True)),
                        ("ct_spiral", lambda: T.ct_scene(ref, histories=1500))):
        for model in (0, 1, 2):
            sc = build()
            r = sc.transport(model=model, output=S.OUT_EV_PER_HISTORY, seed=T.SEED)
            np.savez_compressed(os.path.join(HERE, f"transport_{name}_m{model}.npz"), **summarize(sc, r))
            sc = build()
            r = sc.transport(model=model, output=S.OUT_EV_PER_HISTORY, seed=T.SEED, workers=S.WORKERS_COUNTER_STREAMS)
            s = summarize(sc, r)
            s["n_events"] = r.n_events
            np.savez_compressed(os.path.join(HERE, f"streams_{name}_m{model}.npz"), **s)
            print(name, model, s["total"], s["events"])


if __name__ == "__main__":
    main()

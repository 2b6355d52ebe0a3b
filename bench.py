#!/usr/bin/env python3
"""bench.py — photon histories/s of the transport hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                      # this repo (CUDA kernels)
    python bench.py --impl reference --gpus N --steps K ...            # the unmodified reference on the host CPU
    python bench.py --config {1,2,3,4,5} [--scaling weak|strong] ...   # another BASELINE config / strong scaling

Workloads (config.workload names the one that ran; BASELINE.json configs[0..4] = --config 1..5):
  1  60 keV pencil beam into a 64^3 water cube, 10 exposures x 1e6 histories
  2  120 kVp tungsten-anode DX source (built-in tube model, heel effect) onto a voxelised tissue slab with bone and water
     inserts (TG-195 case 2 geometry), 100 exposures x 1e7
  3  CTDI 32 cm PMMA phantom, CT axial source 120 kVp, 40 mm collimation, 360 exposures x 2 777 778 = 1e9 histories,
     forced interactions in the five chamber bores; centre / periphery / weighted dose are reported
  4  (default at N=1, the configuration the metric is quoted on) CT spiral source, pitch 1.0, 40 mm collimation, 120 kV,
     bow-tie + heel effect, Livermore model, over the synthetic 512x512x400 anthropomorphic phantom (10 materials),
     3600 exposures x 2 777 778 histories = 1e10 histories
  5  (default at N>1) the same scan over the 512^3 phantom with tube-current modulation (AECFilter, I(z) = 1 + 0.5 sin(2 pi z / 512)),
     4608 exposures; weak scaling: 1.25e10 histories per GPU, i.e. BASELINE config #5 as written at N=8
     (4608 exposures x 21 701 389 histories = 1e11)
One step = one full pass of the run. With N GPUs rank r transports exposures r, r+N, ... and the fixed-point dose grids are
summed with one NCCL all-reduce inside the timed region. --scaling weak (default): per-GPU work fixed (config 4: the angular
step shrinks to 1/N degree; config 5: histories per exposure grow with N); --scaling strong: the N=1 job split over N GPUs.

Prints ONE JSON line (rank 0). `value` is measured with everything resident on the GPU; `e2e` goes through the
reference-facing call (dxs_transport = Transport::operator()) with host arrays in and host arrays out.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# Transport calls run back to back here: let destroyed contexts park their big device blocks for the next call
# (off by default for library users, dxmcb200_set_pool_limit; reported in config.device_pool)
os.environ.setdefault("DXMCB200_POOL_GB", "96")

from dxmclib_b200 import phantoms, sharding  # noqa: E402
from dxmclib_b200 import scene as S  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of one full-size launch of the dominant kernel, from the committed
# `ncu --set full` capture under profiles/ (bytes); None until a capture of the current kernels is committed
TRAFFIC_PER_LAUNCH = {"transportKernel": 9_741_176_000, "interactKernel": 8_909_366_000}
TRAFFIC_SOURCE = ("profiles/r2_v11_{transport,interact}Kernel_ncu_summary.csv: one steady-state launch (a wave of 2^26 photon segments). The captured "
                  "transportKernel launch made 2.3e8 look-ups = 1.4 GB of algorithmic bytes; the rest of its traffic is the record hand-over "
                  "(64 B read per segment, 64 B written per event, 64 B per air-walk photon) and the sectors of missed palette look-ups")

SEED = 0xD1C02026
MODEL = S.MODEL_LIVERMORE
TISSUE = "H62.9539171935344C12.9077870263354N1.16702581276482O22.7840718642933Na0.026328553360443P0.0390933975009805S0.0566470278101205Cl0.0341543557411274K0.0309747686593447"
AIR = "C0.0150228136551869N78.439632744437O21.0780510531616Ar0.467293388746132"
BONE = "H39.229963C15.009010N3.487490O31.621690Na0.050590Mg0.095705P3.867606S0.108832Ca6.529115"


class Workload:
    """One BASELINE config as concrete inputs: `build(lib)` makes the scene (world + source) through the scene API."""

    def __init__(self, config, n_ranks, scaling, histories=None):
        self.config, self.n, self.scaling = config, n_ranks, scaling
        grow = n_ranks if scaling == "weak" else 1
        self.step_deg = 1.0
        if config == 1:
            self.dim, self.spacing, self.exposures, self.hist = (64, 64, 64), (1.0, 1.0, 1.0), 10 * grow, 1_000_000
            self.name = "60 keV pencil beam into a 64^3 water cube @1 mm"
        elif config == 2:
            self.dim, self.spacing, self.exposures, self.hist = (80, 200, 360), (5.0, 5.0, 5.0), 100 * grow, 10_000_000
            self.name = "120 kVp W-anode DX source (tube model, heel effect) onto a 390x390x200 mm tissue slab with bone and water inserts, 80x200x360 @5 mm"
        elif config == 3:
            self.dim, self.spacing, self.exposures, self.hist = None, None, 360 * grow, 2_777_778
            self.step_deg = 1.0 / grow
            self.name = "CTDI 32 cm PMMA phantom, CT axial 120 kVp, 40 mm collimation, forced interactions in the chamber bores"
        elif config == 4:
            self.dim, self.spacing, self.exposures, self.hist = (512, 512, 400), (1.0, 1.0, 1.0), 3600 * grow, 2_777_778
            self.step_deg = 1.0 / grow
            self.name = "CT spiral pitch 1.0, 120 kV, 40 mm collimation, bow-tie + heel, Livermore, synthetic anthropomorphic 512x512x400 @1 mm, 10 materials"
        elif config == 5:
            self.dim, self.spacing, self.exposures, self.hist = (512, 512, 512), (1.0, 1.0, 1.0), 4608, 2_712_674 * grow
            if scaling == "strong":
                self.hist = 21_701_389  # config #5 as written: 1e11 histories, split over however many GPUs there are
            self.name = ("CT spiral pitch 1.0, 120 kV, 40 mm collimation, bow-tie + heel, AEC tube-current modulation 1+0.5sin(2 pi z/512), "
                         "Livermore, synthetic anthropomorphic 512^3 @1 mm, 10 materials")
        else:
            raise SystemExit(f"unknown --config {config}")
        if histories:
            self.hist = int(histories)
        self._phantom = None

    @property
    def total_histories(self):
        return self.exposures * self.hist

    def phantom(self):
        if self._phantom is None and self.config in (4, 5):
            self._phantom = phantoms.anthropomorphic(self.dim, self.spacing)
        return self._phantom

    def build(self, lib, hist=None):
        hist = self.hist if hist is None else int(hist)
        sc = S.Scene(lib)
        if self.config == 1:
            sc.world(self.dim, self.spacing)
            sc.add_material("Water, Liquid", 1.0)
            mat, dens = phantoms.water_cube(self.dim[0])
            sc.arrays(dens, mat)
            assert sc.validate()
            sc.source_pencil((0.0, 0.0, -float(self.dim[2])), (1, 0, 0, 0, 1, 0), 60.0, hist, self.exposures)
        elif self.config == 2:
            nx, ny, nz = self.dim
            sp = self.spacing[0]
            x = (np.arange(nx) + 0.5) * sp - nx * sp / 2
            y = (np.arange(ny) + 0.5) * sp - ny * sp / 2
            z = (np.arange(nz) + 0.5) * sp
            Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
            mat = np.zeros((nz, ny, nx), np.uint8)
            mat[(np.abs(X) < 195) & (np.abs(Y) < 195) & (Z > 1550) & (Z < 1750)] = 1
            mat[(np.abs(X) < 60) & (np.abs(Y + 100) < 60) & (Z > 1600) & (Z < 1660)] = 2  # bone
            mat[(np.abs(X) < 60) & (np.abs(Y - 100) < 60) & (Z > 1600) & (Z < 1700)] = 3  # water
            sc.world(self.dim, self.spacing, (0.0, 0.0, 900.0))
            sc.add_material(AIR, 0.001205).add_material(TISSUE, 1.03).add_material(BONE, 1.92).add_material("Water, Liquid", 1.0)
            sc.arrays(np.array([0.001205, 1.03, 1.92, 1.0], np.float32)[mat], mat)
            assert sc.validate()
            sc.source_dx(voltage=120.0, al_mm=2.5, sdd=1800.0, field_size=(390.0, 390.0), source_angles_deg=(0.0, -90.0), tube_rotation_deg=0.0,  # tube at the origin, beam along +z
                         dap=1.0, histories=hist, exposures=self.exposures, position=(0.0, 0.0, 1800.0), model_heel=True)
        elif self.config == 3:
            sc.ctdi_phantom(320)
            self.dim, self.spacing, _ = sc.dimensions()
            self.dim, self.spacing = tuple(int(d) for d in self.dim), tuple(float(s) for s in self.spacing)
            sc.source_ct(spiral=False, voltage=120.0, al_mm=7.0, sdd=1190.0, collimation=40.0, fov=500.0, scan_length=40.0, step=40.0,
                         position=(0.0, 0.0, 0.0), exposure_step_deg=self.step_deg, histories=hist, model_heel=True, ctdi_phantom_diameter=320)
        else:
            sc.world(self.dim, self.spacing)
            for name, dens in phantoms.ANTHROPOMORPHIC_MATERIALS:
                sc.add_material(name, dens)
            mat, dens = self.phantom()
            sc.arrays(dens, mat)
            if not sc.validate():
                raise RuntimeError("phantom world did not validate")
            scan = self.dim[2] * self.spacing[2]
            sc.source_ct(spiral=True, voltage=120.0, al_mm=7.0, sdd=1190.0, collimation=40.0, fov=500.0, pitch=1.0, scan_length=scan,
                         position=(0.0, 0.0, -scan / 2), exposure_step_deg=self.step_deg, histories=hist, model_heel=True, ctdi_vol=10.0)
            a, w = phantoms.bowtie_profile()
            sc.source_bowtie(a, w)
            if self.config == 5:
                zz = np.arange(self.dim[2], dtype=np.float32)
                sc.source_aec(1.0 + 0.5 * np.sin(2 * np.pi * zz / self.dim[2]))
        n_exp = sc.total_exposures()
        assert n_exp == self.exposures, (n_exp, self.exposures)
        return sc

    def describe(self, hist=None):
        hist = self.hist if hist is None else int(hist)
        tracking = os.environ.get("DXMCB200_TRACKING", "1")
        return {
            "workload": f"BASELINE config #{self.config}: {self.name}, {self.exposures} exposures x {hist} histories",
            "baseline_config": self.config, "voxels": int(np.prod(self.dim)) if self.dim else None, "exposures": self.exposures,
            "histories_per_exposure": hist, "histories_per_step": self.exposures * hist,
            "sharding": f"exposures interleaved over {self.n} GPU(s) (rank r: r, r+N, ...), one all-reduce of the fixed-point grids",
            "tracking": "Woodcock + empty-space traversal through air bricks (default)" if tracking != "0" else "Woodcock, global majorant (reference algorithm)",
            "device_pool": f"DXMCB200_POOL_GB={os.environ.get('DXMCB200_POOL_GB')} (big device blocks parked between Transport calls; library default is off)",
            "l2_note": "accumulators (32 B/voxel) + photon/event record streams (tens of GB per wave pair) exceed the 126 MB L2; accumulators are cleared every step",
        }


# kept for tests/test_gpu_fullsize.py and tools/: the N=1 bench workload
DIM = (512, 512, 400)
SPACING = (1.0, 1.0, 1.0)
EXPOSURES = 3600
HIST_PER_EXPOSURE = 2_777_778


def build_scene(lib, histories_per_exposure, n_ranks=1, dim=DIM, phantom=None):
    w = Workload(4, n_ranks, "weak", histories_per_exposure)
    w.dim = tuple(dim)
    w._phantom = phantom
    return w.build(lib)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = max((float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.samples)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ctdi_report(sc, result):
    """centre / periphery / weighted bore dose of the CTDI phantom (reference source.hpp:972-986)."""
    holes = [float(result.dose[sc.ctdi_holes(p).astype(np.int64)].astype(np.float64).mean()) for p in range(5)]
    centre, periphery = holes[0], float(np.mean(holes[1:]))
    return {"centre": centre, "periphery_mean": periphery, "weighted": centre / 3.0 + 2.0 * periphery / 3.0, "units": result.units}


def run_reference(args, rank, emit):
    """The reference's own multithreaded CPU Transport (oracle/_ref, built from the unmodified reference) on a bounded
    sample of the same workload: same world, same source, fewer histories per exposure."""
    if rank != 0:
        return
    lib, how = S.reference_timing_lib()
    cores = os.cpu_count() or 1
    w = Workload(args.config, 1, "weak")
    hist = max(1, args.ref_histories // w.exposures)
    t0 = time.time()
    sc = w.build(lib, hist)
    setup = time.time() - t0
    times = []
    for i in range(args.warmup + args.steps):
        r = sc.transport(model=MODEL, output=S.OUT_EV_PER_HISTORY, seed=0, workers=0, want_events=False, want_variance=False)
        if i >= args.warmup:
            times.append(r.seconds)
    per_step = float(np.mean(times))
    value = hist * w.exposures / per_step
    sample = f"{w.exposures} exposures x {hist} histories per step (stock Transport::operator(), Result::simulationTime; reference built with {how})"
    line = {
        "impl": "reference", "metric": "photon histories/s", "value": value, "unit": "histories/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": w.describe(hist) | {"setup_s": round(setup, 1)},
        "cpu_baseline": {"value": value, "unit": "histories/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    # stdout carries exactly ONE JSON line: anything libraries print (NCCL banners, warnings) is sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=0, help="BASELINE config 1..5 (default: 4 on one GPU, 5 on several)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--histories", type=int, default=0, help="histories per exposure (default: the config's own)")
    ap.add_argument("--ref-histories", type=int, default=30_000_000, help="histories per step of the CPU reference sample")
    ap.add_argument("--cpu-baseline-histories", type=int, default=150_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least three warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config == 0:
        args.config = 4 if max(world, args.gpus) == 1 or args.impl == "reference" else 5

    if args.impl == "reference":
        run_reference(args, rank, emit)
        return

    import torch
    import torch.distributed as dist

    from dxmclib_b200 import cabi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = world
    lib = S.product_lib()

    w = Workload(args.config, n, args.scaling, args.histories)
    hist = w.hist
    sc = w.build(lib)
    n_exp = w.exposures
    total_hist_all = w.total_histories
    nvox = int(np.prod(w.dim))
    t0 = time.time()
    sc.b200_prepare(device=local, model=MODEL, seed=SEED, total_histories_all_ranks=total_hist_all)
    prepare_s = time.time() - t0
    ctx = cabi.Context(handle=sc.b200_context())
    ctx.n_voxels = nvox
    # interleaved partition: rank r transports exposures r, r + N, ... so that every GPU sees the same mix of scan positions
    first, stride, count = sharding.exposure_stride(n_exp, rank, n)
    hist_rank = count * hist

    acc_tensor = sharding.accumulator_tensor(ctx, torch.device("cuda", local)) if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_ms = []
    reduce_ms = []
    launches = 0
    per_kernel = {"generate": [0.0, 0], "transport": [0.0, 0], "airwalk": [0.0, 0], "interact": [0.0, 0]}

    def step(record):
        nonlocal launches
        ctx.clear()
        ms = sc.b200_run_strided(first, stride, count)
        if acc_tensor is not None:
            t1 = time.perf_counter()
            dist.all_reduce(acc_tensor, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if record:
                reduce_ms.append((time.perf_counter() - t1) * 1e3)
        if record:
            kernel_ms.append(ms)
            launches += ctx.stats()["kernel_launches"]  # generate / transport / air walk / interact launches (cursor resets and NCCL not counted)
            for k, v in ctx.kernel_times().items():
                per_kernel[k][0] += v["ms"]
                per_kernel[k][1] += v["launches"]

    for _ in range(args.warmup):
        step(False)
    barrier()
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(True)
        barrier()
        elapsed = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([elapsed, float(np.sum(kernel_ms))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, kernel_total_ms = float(t[0]), float(t[1])
    else:
        kernel_total_ms = float(np.sum(kernel_ms))
    ms_per_step = elapsed / args.steps * 1e3
    value = total_hist_all * args.steps / elapsed

    # ---- roofline: algorithmic bytes per history from the kernels' own work counters (short counted run)
    ctx.enable_stats(True)
    ctx.clear()
    sample_hist = 0
    for k in range(0, count, max(1, count // 20)):  # ~20 exposures spread over the rank's share: the counters must sample the whole scan
        sc.b200_run_strided(first + k * stride, 1, 1)
    st = ctx.stats()
    sample_hist = max(st["histories"], 1)
    L = st["lookups"] / sample_hist
    Sev = st["score_events"] / sample_hist
    # the same sample with the reference's own tracking (Woodcock against the global majorant everywhere): the look-ups per history
    # of the REFERENCE ALGORITHM, which is what SURVEY 8d's per-history figure B_alg = 6 L + 24 S was defined on (and round 1 reported)
    L_ref = L
    if os.environ.get("DXMCB200_TRACKING", "1") != "0":
        ctx.set_tracking(0)
        ctx.clear()
        for k in range(0, count, max(1, count // 20)):
            sc.b200_run_strided(first + k * stride, 1, 1)
        ref_st = ctx.stats()
        L_ref = ref_st["lookups"] / max(ref_st["histories"], 1)
        ctx.set_tracking(1)
    ctx.enable_stats(False)
    peak, peak_src = measured_peak()
    # Algorithmic bytes: 6 B per voxel look-up (u8 material + f32 density + u8 measurement of the reference layout),
    # all issued by transportKernel (the air walk's few look-ups included in L); 24 B per scoring event (read + write of f32
    # dose, u32 events, f32 variance), all issued by interactKernel. The roofline object describes the kernel with the largest
    # device time; the other one and the whole pipeline are listed beside it.
    def kernel_line(name, bytes_per_history):
        ms, cnt = per_kernel[name]
        launch_ms = ms / max(cnt, 1)
        per_launch = bytes_per_history * hist_rank * args.steps / max(cnt, 1)
        ach = per_launch / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
        return {"achieved": ach, "frac": ach / peak, "algorithmic_bytes_per_launch": per_launch, "launch_ms": launch_ms, "launches": cnt}

    lines = {"transportKernel": kernel_line("transport", 6.0 * L), "interactKernel": kernel_line("interact", 24.0 * Sev)}
    dominant = "transportKernel" if per_kernel["transport"][0] >= per_kernel["interact"][0] else "interactKernel"
    b_alg = 6.0 * L + 24.0 * Sev
    pipeline_s = kernel_total_ms / args.steps / 1e3
    roofline = {"bound": "hbm", "achieved": lines[dominant]["achieved"], "peak": peak, "unit": "GB/s", "frac": lines[dominant]["frac"],
                "traffic": TRAFFIC_PER_LAUNCH[dominant], "traffic_source": TRAFFIC_SOURCE,
                "kernel": dominant + (" (Woodcock stepping; one launch = one wave of <= 2^26 photon segments)" if dominant == "transportKernel"
                                      else " (interaction sampling + fixed-point scoring; one launch = the events of one wave)"),
                "algorithmic_bytes_per_launch": lines[dominant]["algorithmic_bytes_per_launch"], "launch_ms": lines[dominant]["launch_ms"],
                "launches": lines[dominant]["launches"], "kernels": lines,
                "kernel_share_of_step": {k: v[0] / max(sum(x[0] for x in per_kernel.values()), 1e-9) for k, v in per_kernel.items()},
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in per_kernel.items()},
                "pipelines": 2, "note": "two wave pipelines overlap on the GPU, so per-kernel device times add up to more than the step. The algorithmic "
                                        "bytes are 6 B x look-ups: the empty-space traversal removes two thirds of the reference algorithm's look-ups "
                                        "(32 -> 11.5 per history), so `achieved`/`frac` fall while histories/s rise; the kernels are bound by instruction "
                                        "issue and latency, not by HBM (profiles/README.md)",
                "whole_pipeline": {"bytes_per_history": b_alg, "achieved": hist_rank * b_alg / pipeline_s / 1e9,
                                   "frac": hist_rank * b_alg / pipeline_s / 1e9 / peak},
                "lookups_per_history": L, "score_events_per_history": Sev, "steps_per_history": st["steps"] / sample_hist,
                "interactions_per_history": st["interactions"] / sample_hist, "air_walks_per_history": st["air_walks"] / sample_hist,
                "bricks_crossed_per_history": st["bricks_crossed"] / sample_hist,
                "pipeline_ms_per_step": kernel_total_ms / args.steps, "peak_source": peak_src}
    # the same two figures per unit of the reference algorithm's work (a history costs 6 L_ref + 24 S algorithmic bytes in the
    # reference; the product does the history with fewer look-ups): comparable with round 1 and across tracking modes
    b_ref = 6.0 * L_ref + 24.0 * Sev
    t_ms, t_cnt = per_kernel["transport"]
    ref_transport = 6.0 * L_ref * hist_rank * args.steps / max(t_ms * 1e-3, 1e-12) / 1e9
    roofline["reference_algorithm"] = {
        "lookups_per_history": L_ref, "bytes_per_history": b_ref,
        "transportKernel": {"achieved": ref_transport, "frac": ref_transport / peak},
        "whole_pipeline": {"achieved": hist_rank * b_ref / pipeline_s / 1e9, "frac": hist_rank * b_ref / pipeline_s / 1e9 / peak},
        "note": "algorithmic bytes of the work as the reference does it (look-ups counted with DXMCB200_TRACKING=0 on the same sample) over the measured "
                "device times; `achieved` / `frac` above count only the look-ups the product actually makes"}
    if reduce_ms:
        roofline["all_reduce_ms_per_step"] = float(np.mean(reduce_ms))
        roofline["all_reduce_share_of_step"] = float(np.mean(reduce_ms)) / ms_per_step

    # ---- config #3: the dose figures the configuration is about
    extra = {}
    if args.config == 3 and rank == 0:
        r = sc.b200_collect(output=S.OUT_DOSE, use_calibration=False, histories=total_hist_all)
        extra["ctdi_bore_dose_keV_per_kg"] = ctdi_report(sc, r)

    # ---- e2e: Transport::operator() with host arrays in and out (N=1), or prepare/run/all-reduce/collect (N>1)
    e2e = None
    if not args.no_e2e:
        sc.b200_release()
        has_meas = args.config == 3
        h2d = nvox * (4 + 1 + (1 if has_meas else 0)) + n_exp * 96
        d2h = nvox * (4 + 4 + 4)

        def e2e_step(output=S.OUT_EV_PER_HISTORY, calibrate=False):
            if world == 1:
                return sc.transport(model=MODEL, output=output, use_calibration=calibrate, seed=SEED)
            sc.b200_prepare(device=local, model=MODEL, seed=SEED, total_histories_all_ranks=total_hist_all)
            ctx2 = cabi.Context(handle=sc.b200_context())
            sc.b200_run_strided(first, stride, count)
            dist.all_reduce(sharding.accumulator_tensor(ctx2, torch.device("cuda", local)), op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if rank == 0:
                sc.b200_collect(output=output, histories=total_hist_all)
            sc.b200_release()
            return None

        # every e2e step is one complete call: host arrays in (world, tables, exposures), host arrays out (dose, events,
        # variance); K calls are timed like the device-resident steps, the per-call times are listed
        calls = []
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            t1 = time.perf_counter()
            e2e_step()
            calls.append(time.perf_counter() - t1)
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        if world > 1:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t[0])
        e2e = {"value": total_hist_all / e2e_s, "unit": "histories/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "seconds": e2e_s, "calls": args.steps, "seconds_per_call_rank0": [round(x, 3) for x in calls],
               "path": "dxs_transport (Transport::operator(): LUT build + upload + transport + download), output EV_PER_HISTORY"
               if world == 1 else "prepare + run + NCCL all-reduce + collect"}
        if world == 1 and args.config in (3, 4, 5):
            # the reference's DEFAULT call: OUTPUTMODE::DOSE with the source's dose calibration (transport.hpp:838-839), which for
            # a CT source runs a second Transport over the CTDI phantom (source.hpp:925-988) inside the call
            dose_calls = []
            for _ in range(2):
                t1 = time.perf_counter()
                r = e2e_step(S.OUT_DOSE, True)
                dose_calls.append(time.perf_counter() - t1)
            dose_s = float(np.mean(dose_calls))
            e2e["dose_mode_with_calibration"] = {"value": total_hist_all / dose_s, "unit": "histories/s", "seconds": dose_s,
                                                 "seconds_per_call": [round(x, 3) for x in dose_calls], "dose_units": r.units,
                                                 "note": "histories of the main run / wall time of the whole call incl. the calibration run (mean of two calls)"}

    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # N=1 only: at N>1 the host cores are busy feeding the other ranks
        try:
            ref, how = S.reference_timing_lib()
            h = max(1, args.cpu_baseline_histories // n_exp)
            rs = w.build(ref, h)
            r = rs.transport(model=MODEL, output=S.OUT_EV_PER_HISTORY, seed=0, workers=0, want_events=False, want_variance=False)
            cpu = {"value": r.histories / r.seconds, "unit": "histories/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"{n_exp} exposures x {h} histories, stock multithreaded Transport::operator(), Result::simulationTime; "
                             f"reference built with {how}",
                   "seconds": r.seconds}
            rs.close()
        except Exception as e:  # the checker is optional for the measurement, never for the product
            cpu = {"value": None, "unit": "histories/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}

    if rank == 0:
        line = {
            "metric": "photon histories/s", "value": value, "unit": "histories/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": w.describe() | {"prepare_s": round(prepare_s, 2)} | extra,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

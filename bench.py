#!/usr/bin/env python3
"""bench.py — photon histories/s of the transport hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference on the host CPU

Workload (config.workload): BASELINE config #4 — CT spiral source, pitch 1.0, 40 mm collimation, 120 kV,
bow-tie + heel effect, Livermore model, over the synthetic 512x512x400 anthropomorphic phantom
(dxmclib_b200/phantoms.py, 10 materials), 3600 exposures x 2 777 778 histories = 1e10 histories.
One step = one full pass of that run. With N GPUs the exposure angle step is 1/N degree, so every rank
transports its own interleaved set of 3600 exposures (rank r: r, r+N, ...; weak scaling: N x 1e10 histories) and the fixed-point dose grids
are summed with one NCCL all-reduce inside the timed region.

Prints ONE JSON line (rank 0). `value` is measured with everything resident on the GPU; `e2e` goes through the
reference-facing call (dxs_transport = Transport::operator()) with host arrays in and host arrays out.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from dxmclib_b200 import phantoms, sharding  # noqa: E402
from dxmclib_b200 import scene as S  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of one full-size transportKernel launch, from the committed
# `ncu --set full` capture (profiles/r1_v7_*), bytes
TRAFFIC_PER_LAUNCH = 9.441e9  # profiles/r1_v7_transportKernel_ncu_summary.csv: 5.713 GB read + 3.728 GB written, 2^26-record wave

DIM = (512, 512, 400)
SPACING = (1.0, 1.0, 1.0)
EXPOSURES = 3600
HIST_PER_EXPOSURE = 2_777_778
SEED = 0xD1C02026
MODEL = S.MODEL_LIVERMORE


def build_scene(lib, histories_per_exposure, n_ranks=1, dim=DIM, phantom=None):
    """CT spiral over the anthropomorphic phantom; with n_ranks>1 the angular step shrinks so that the scan holds
    n_ranks x 3600 exposures."""
    sc = S.Scene(lib)
    sc.world(dim, SPACING)
    for name, dens in phantoms.ANTHROPOMORPHIC_MATERIALS:
        sc.add_material(name, dens)
    mat, dens = phantom if phantom is not None else phantoms.anthropomorphic(dim, SPACING)
    sc.arrays(dens, mat)
    if not sc.validate():
        raise RuntimeError("phantom world did not validate")
    scan = dim[2] * SPACING[2]
    sc.source_ct(spiral=True, voltage=120.0, al_mm=7.0, sdd=1190.0, collimation=40.0, fov=500.0, pitch=1.0,
                 scan_length=scan, position=(0.0, 0.0, -scan / 2), exposure_step_deg=1.0 / n_ranks,
                 histories=histories_per_exposure, model_heel=True, ctdi_vol=10.0)
    a, w = phantoms.bowtie_profile()
    sc.source_bowtie(a, w)
    return sc


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


def run_reference(args, rank, emit):
    """The reference's own multithreaded CPU Transport (oracle/_ref, built from the unmodified reference) on a bounded
    sample of the same workload: same world, same source, fewer histories per exposure."""
    if rank != 0:
        return
    lib, how = S.reference_timing_lib()
    cores = os.cpu_count() or 1
    hist = max(1, args.ref_histories // EXPOSURES)
    t0 = time.time()
    sc = build_scene(lib, hist)
    setup = time.time() - t0
    times = []
    for i in range(args.warmup + args.steps):
        r = sc.transport(model=MODEL, output=S.OUT_EV_PER_HISTORY, seed=0, workers=0, want_events=False, want_variance=False)
        if i >= args.warmup:
            times.append(r.seconds)
    per_step = float(np.mean(times))
    value = hist * EXPOSURES / per_step
    sample = f"{EXPOSURES} exposures x {hist} histories per step (stock Transport::operator(), Result::simulationTime; reference built with {how})"
    line = {
        "impl": "reference", "metric": "photon histories/s", "value": value, "unit": "histories/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, hist) | {"setup_s": round(setup, 1)},
        "cpu_baseline": {"value": value, "unit": "histories/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n_ranks, hist):
    return {
        "workload": f"CT spiral pitch 1.0, 120 kV, 40 mm collimation, bow-tie + heel, Livermore, synthetic anthropomorphic "
                    f"{DIM[0]}x{DIM[1]}x{DIM[2]} @1 mm, 10 materials, {EXPOSURES * n_ranks} exposures x {hist} histories",
        "voxels": int(np.prod(DIM)), "exposures": EXPOSURES * n_ranks, "histories_per_exposure": hist,
        "histories_per_step": EXPOSURES * n_ranks * hist, "sharding": f"exposures interleaved over {n_ranks} GPU(s) (rank r: r, r+N, ...), one all-reduce of the fixed-point grids",
        "l2_note": "accumulators 3.4 GB + photon/event record streams (>20 GB per wave pair) exceed the 126 MB L2; the 4-bit palette voxel "
                   "grid is 52 MB; accumulators are cleared every step",
    }


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
    ap.add_argument("--histories", type=int, default=HIST_PER_EXPOSURE, help="histories per exposure (default: the 1e10 run)")
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

    hist = args.histories
    total_hist_all = EXPOSURES * n * hist
    phantom = phantoms.anthropomorphic(DIM, SPACING)
    sc = build_scene(lib, hist, n_ranks=n, phantom=phantom)
    n_exp = sc.total_exposures()
    assert n_exp == EXPOSURES * n, (n_exp, EXPOSURES * n)
    t0 = time.time()
    sc.b200_prepare(device=local, model=MODEL, seed=SEED, total_histories_all_ranks=total_hist_all)
    prepare_s = time.time() - t0
    ctx = cabi.Context(handle=sc.b200_context())
    ctx.n_voxels = int(np.prod(DIM))
    # interleaved partition: rank r transports exposures r, r + N, ... so that every GPU sees the same mix of scan positions
    first, stride, count = sharding.exposure_stride(n_exp, rank, n)
    assert count == EXPOSURES

    acc_tensor = None
    if world > 1:
        ptr, n_u64 = ctx.accumulators()

        class _Acc:  # expose the accumulator block to torch without a copy
            __cuda_array_interface__ = {"shape": (n_u64,), "typestr": "<i8", "data": (ptr, False), "version": 2}

        acc_tensor = torch.as_tensor(_Acc(), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_ms = []
    launches = 0
    per_kernel = {"generate": [0.0, 0], "transport": [0.0, 0], "airwalk": [0.0, 0], "interact": [0.0, 0]}

    def step(record):
        nonlocal launches
        ctx.clear()
        ms = sc.b200_run_strided(first, stride, count)
        if acc_tensor is not None:
            dist.all_reduce(acc_tensor, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
        if record:
            kernel_ms.append(ms)
            launches += ctx.stats()["kernel_launches"]  # generate / transport / interact launches (cursor resets and NCCL not counted)
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

    # ---- roofline: algorithmic bytes per history from the kernel's own work counters (short counted run)
    ctx.enable_stats(True)
    ctx.clear()
    for k in range(0, EXPOSURES, 180):  # every 180th exposure of the rank: the counters must sample the whole scan, not one end of it
        sc.b200_run_strided(first + k * stride, 1, 1)
    st = ctx.stats()
    ctx.enable_stats(False)
    L = st["lookups"] / max(st["histories"], 1)
    Sev = st["score_events"] / max(st["histories"], 1)
    hist_rank = EXPOSURES * hist
    peak, peak_src = measured_peak()
    # Dominant kernel = transportKernel (Woodcock stepping): its algorithmic bytes are the voxel look-ups, 6 B each in the
    # reference layout (u8 material + f32 density + u8 measurement); the scoring bytes (24 B per event) belong
    # to interactKernel. achieved = bytes per launch / average launch duration, both from the timed steps.
    t_ms, t_n = per_kernel["transport"]
    t_launch_ms = t_ms / max(t_n, 1)
    lookups_per_launch = L * hist_rank * args.steps / max(t_n, 1)
    achieved = 6.0 * lookups_per_launch / (t_launch_ms * 1e-3) / 1e9 if t_launch_ms > 0 else 0.0
    b_alg = 6.0 * L + 24.0 * Sev
    pipeline_s = kernel_total_ms / args.steps / 1e3
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": TRAFFIC_PER_LAUNCH,
                "kernel": "transportKernel<false> (Woodcock stepping; one launch = one wave of <= 2^26 photon segments)",
                "algorithmic_bytes_per_launch": 6.0 * lookups_per_launch, "launch_ms": t_launch_ms, "launches": t_n,
                "kernel_share_of_step": {k: v[0] / max(kernel_total_ms, 1e-9) for k, v in per_kernel.items()},
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in per_kernel.items()},
                "pipelines": 2, "note": "two wave pipelines overlap on the GPU, so per-kernel device times add up to more than the step",
                "whole_pipeline": {"bytes_per_history": b_alg, "achieved": hist_rank * b_alg / pipeline_s / 1e9,
                                   "frac": hist_rank * b_alg / pipeline_s / 1e9 / peak},
                "lookups_per_history": L, "score_events_per_history": Sev, "steps_per_history": st["steps"] / max(st["histories"], 1),
                "interactions_per_history": st["interactions"] / max(st["histories"], 1),
                "pipeline_ms_per_step": kernel_total_ms / args.steps, "peak_source": peak_src}

    # ---- e2e: Transport::operator() with host arrays in and out (N=1), or prepare/run/all-reduce/collect (N>1)
    e2e = None
    if not args.no_e2e:
        sc.b200_release()
        nvox = int(np.prod(DIM))
        h2d = nvox * (4 + 1 + 1) + n_exp * 96
        d2h = nvox * (4 + 4 + 4)
        def e2e_step():
            if world == 1:
                sc.transport(model=MODEL, output=S.OUT_EV_PER_HISTORY, seed=SEED)
                return
            sc.b200_prepare(device=local, model=MODEL, seed=SEED, total_histories_all_ranks=total_hist_all)
            ctx2 = cabi.Context(handle=sc.b200_context())
            sc.b200_run_strided(first, stride, count)
            ptr, n_u64 = ctx2.accumulators()

            class _Acc2:
                __cuda_array_interface__ = {"shape": (n_u64,), "typestr": "<i8", "data": (ptr, False), "version": 2}

            dist.all_reduce(torch.as_tensor(_Acc2(), device=torch.device("cuda", local)), op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if rank == 0:
                sc.b200_collect(output=S.OUT_EV_PER_HISTORY, histories=total_hist_all)
            sc.b200_release()

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
               "path": "dxs_transport (Transport::operator(): LUT build + upload + transport + download)"
               if world == 1 else "prepare + run + NCCL all-reduce + collect"}

    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # N=1 only: at N>1 the host cores are busy feeding the other ranks
        try:
            ref, how = S.reference_timing_lib()
            h = max(1, args.cpu_baseline_histories // EXPOSURES)
            rs = build_scene(ref, h, phantom=phantom)
            r = rs.transport(model=MODEL, output=S.OUT_EV_PER_HISTORY, seed=0, workers=0, want_events=False, want_variance=False)
            cpu = {"value": r.histories / r.seconds, "unit": "histories/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"{EXPOSURES} exposures x {h} histories, stock multithreaded Transport::operator(), Result::simulationTime; "
                             f"reference built with {how}",
                   "seconds": r.seconds}
            rs.close()
        except Exception as e:  # the checker is optional for the measurement, never for the product
            cpu = {"value": None, "unit": "histories/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}

    if rank == 0:
        line = {
            "metric": "photon histories/s", "value": value, "unit": "histories/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(n, hist) | {"prepare_s": round(prepare_s, 2)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

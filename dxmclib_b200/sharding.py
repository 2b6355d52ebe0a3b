"""Multi-GPU partitioning of a run: exposures are independent units, every GPU transports a contiguous
block of exposure indices with world, tables and beam data replicated, and the 64-bit fixed-point accumulator blocks are
summed with ONE all-reduce. Integer sums are associative, so the result is bit-identical for any number of GPUs."""
from __future__ import annotations

import ctypes as C

from . import cabi


def exposure_block(n_exposures: int, rank: int, world_size: int) -> tuple[int, int]:
    """[begin, end) of the exposures rank `rank` transports; block sizes differ by at most one."""
    base, extra = divmod(int(n_exposures), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def exposure_stride(n_exposures: int, rank: int, world_size: int) -> tuple[int, int, int]:
    """(first, stride, count) of the interleaved partition: rank r transports exposures r, r + N, r + 2N, ... Every GPU
    then sees the same mix of scan positions (a CT spiral's cost per history varies along the patient), so the slowest
    rank is no slower than the average one; the summed grids are the same bits as with exposure_block."""
    n, r, w = int(n_exposures), int(rank), int(world_size)
    return r, w, (n - r + w - 1) // w if r < n else 0


def fixed_point_bits(total_histories_all_ranks: int, max_energy_weight: float) -> tuple[int, int]:
    """The fixed-point scales every rank must agree on (dxmcb200_suggest_fixed_point over the WHOLE job)."""
    e, e2 = C.c_int(0), C.c_int(0)
    rc = cabi.lib().dxmcb200_suggest_fixed_point(C.c_uint64(total_histories_all_ranks), C.c_double(max_energy_weight), C.byref(e), C.byref(e2))
    if rc != 0:
        raise cabi.CabiError(f"dxmcb200_suggest_fixed_point failed with status {rc}")
    return int(e.value), int(e2.value)


def all_reduce_sum(tensor):
    """Sum an int64 accumulator tensor over all ranks in place (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def accumulator_tensor(ctx: "cabi.Context", device):
    """Zero-copy torch int64 view of a context's device accumulator block (n_voxels x {energy, energy^2, events, 0})."""
    import torch

    ptr, n_u64 = ctx.accumulators()

    class _Block:
        __cuda_array_interface__ = {"shape": (n_u64,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    return torch.as_tensor(_Block(), device=device)

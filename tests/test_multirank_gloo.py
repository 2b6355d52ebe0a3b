"""Multi-GPU host logic on CPU: two gloo ranks each transport their block of exposures (with the restatement oracle in
place of the kernels), sum the fixed-point grids with one all-reduce, and must reproduce the single-rank grids bit for
bit — the property the NCCL path relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import support as T
from dxmclib_b200 import scene as S
from dxmclib_b200 import sharding
from oracle import pyoracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, strided=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = T.isotropic_scene(S.product_lib(), histories=1200, exposures=7)
    flat = T.flatten_scene(sc)
    exps = T.exposures_of(sc)
    b, e = sharding.exposure_block(len(exps), rank, world)
    o = pyoracle.Oracle()
    o.load(flat)
    o.set_fixed_point(*sharding.fixed_point_bits(sum(x.histories for x in exps), 140.0))
    if strided:  # the interleaved partition bench.py uses: rank r transports exposures r, r + N, ...
        first, stride, count = sharding.exposure_stride(len(exps), rank, world)
        for k in range(count):
            o.run(exps, first + k * stride, first + k * stride + 1, model=1, seed=9, per_history_streams=True)
    else:
        o.run(exps, b, e, model=1, seed=9, per_history_streams=True)
    energy, energy_sq = o.get_fixed()
    _, events, _ = o.get_raw()
    block = torch.from_numpy(np.stack([energy, energy_sq.view(np.int64), events.astype(np.int64)]))
    sharding.all_reduce_sum(block)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), block.numpy())
    dist.destroy_process_group()


def test_exposure_blocks_partition_the_range():
    for n in (1, 7, 3600, 28800):
        for world in (1, 2, 3, 8):
            blocks = [sharding.exposure_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_exposure_strides_partition_the_range():
    for n in (1, 7, 3600, 28800):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, stride, count = sharding.exposure_stride(n, r, world)
                seen += [first + k * stride for k in range(count)]
            assert sorted(seen) == list(range(n))


@pytest.mark.parametrize("strided", [False, True])
def test_two_rank_sum_equals_single_rank_bit_for_bit(tmp_path, strided):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), strided), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    sc = T.isotropic_scene(S.product_lib(), histories=1200, exposures=7)
    flat = T.flatten_scene(sc)
    exps = T.exposures_of(sc)
    o = pyoracle.Oracle()
    o.load(flat)
    o.set_fixed_point(*sharding.fixed_point_bits(sum(x.histories for x in exps), 140.0))
    o.run(exps, 0, len(exps), model=1, seed=9, per_history_streams=True)
    energy, energy_sq = o.get_fixed()
    _, events, _ = o.get_raw()
    assert energy.any()
    assert np.array_equal(reduced[0], energy)
    assert np.array_equal(reduced[1], energy_sq.view(np.int64))
    assert np.array_equal(reduced[2], events.astype(np.int64))

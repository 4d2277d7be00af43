"""CPU suite, part 4: the multi-GPU plumbing of bench.py with world_size 2 over gloo -- slice ranges partition the slices,
the zero / non-zero column flags combine with MAX (= logical OR), and the per-rank framebuffers (each rank owns the
pixels of its slices, zero elsewhere) reduce to the full image.  The per-rank work is done by the CPU oracle here (the
oracle honours set_slice_range like the CUDA library does); on the GPU box test_multi_handle_slice_ranges_compose covers
the same composition with two CUDA handles."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, small_case, setup


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import orc
    pkg = orc._pkg
    scene, vrls, params = small_case(pkg, "C1", 40, 40, 60, seed=2, targetNumSlices=9)
    o = setup(orc.Oracle(threads=2, **params), scene, vrls)
    o.build_slices()
    S, _ = o.num_slices()
    b, e = pkg.sharding.balanced_ranges(pkg.sharding.slice_sizes(o.pixel_to_slice(), S), world)[rank]
    o.set_slice_range(b, e)
    o.sample_slice_mapping()
    o.build_R()
    flags = torch.from_numpy(o.column_nonzero())                                 # local rows only (the others are zero)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    o.set_column_nonzero(flags.numpy())
    o.build_clusters()
    fb = torch.from_numpy(o.render())
    dist.reduce(fb, dst=0, op=dist.ReduceOp.SUM)
    ranges = [None] * world
    dist.all_gather_object(ranges, (b, e))
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), image=fb.numpy(), flags=flags.numpy(), ranges=np.array(ranges), S=S)
    dist.barrier()
    dist.destroy_process_group()


def test_slice_sharding_world_size_2(tmp_path, pkg, orc):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(tmp_path, "out.npz"))
    ranges, S = got["ranges"], int(got["S"])
    assert ranges[0][0] == 0 and ranges[-1][1] == S and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    # single-process reference of the same job
    scene, vrls, params = small_case(pkg, "C1", 40, 40, 60, seed=2, targetNumSlices=9)
    o = setup(orc.Oracle(threads=2, **params), scene, vrls)
    o.build_slices(); o.prepass()
    np.testing.assert_array_equal(got["image"], o.render())          # slices are independent: bit-identical composition
    np.testing.assert_array_equal(got["flags"], o.column_nonzero())


def test_balanced_ranges_cover_and_balance(pkg):
    rng = np.random.default_rng(5)
    for world in (1, 2, 4, 8):
        for S in (1, 3, 8, 100):
            sizes = rng.integers(1, 4000, S)
            r = pkg.sharding.balanced_ranges(sizes, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == S
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1)) and all(b <= e for b, e in r)
            if S >= 4 * world:                                   # every rank within one largest slice of its fair share
                loads = np.array([sizes[b:e].sum() for b, e in r])
                assert np.abs(loads - sizes.sum() / world).max() <= sizes.max()
    p2s = np.array([0, 0, 1, 0xFFFFFFFF, 2, 2, 2], dtype=np.uint32)
    assert list(pkg.sharding.slice_sizes(p2s, 3)) == [2, 1, 3]

"""GPU suite, part 4: the library's multi-GPU entry point (alvrl_group_*, csrc/group.cu).  A group of one rank must give the
single-handle image; with two GPUs on the box (gpurun --gpus 2) the slice-sharded frame -- column-flag all-reduce and
framebuffer reduce over NCCL inside the library -- must equal the single-GPU frame bit for bit (every pixel is rendered by
exactly one rank; the reduce adds zeros)."""
import numpy as np
import pytest
import torch

from conftest import small_case, setup

pytestmark = pytest.mark.gpu


def _single(pkg, scene, vrls, params):
    g = setup(pkg.integrator(0, **params), scene, vrls)
    g.build_slices(); g.prepass()
    return g, g.render()


def test_group_of_one_rank_equals_single_handle(pkg):
    scene, vrls, params = small_case(pkg, "C1", 64, 64, 150, seed=3, targetNumSlices=12)
    g, img = _single(pkg, scene, vrls, params)
    grp = pkg.binding.Group.local(pkg.api(), [0], **params)
    assert grp.comm_size() == 1
    for m in grp.members:
        setup(m, scene, vrls)
    out = grp.frame()
    assert np.array_equal(out, img)
    assert grp.slice_range(0) == (0, 12)
    grp.close()
    # the rank form with one rank needs no NCCL id
    g2 = setup(pkg.integrator(0, **params), scene, vrls)
    r = pkg.binding.Group.rank(g2, 0, 1)
    assert np.array_equal(r.frame(), img)
    r.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_group_local_multi_gpu_equals_single_gpu(pkg, ndev):
    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    scene, vrls, params = small_case(pkg, "C1", 96, 96, 300, seed=5, targetNumSlices=24)
    g, img = _single(pkg, scene, vrls, params)
    grp = pkg.binding.Group.local(pkg.api(), list(range(ndev)), **params)
    assert grp.comm_size() == ndev
    for m in grp.members:
        setup(m, scene, vrls)
    out = grp.frame()
    ranges = [grp.slice_range(i) for i in range(ndev)]
    assert ranges[0][0] == 0 and ranges[-1][1] == 24 and all(ranges[i][1] == ranges[i + 1][0] for i in range(ndev - 1))
    assert np.array_equal(out, img)
    for _ in range(3):                                   # and again: buffers are reused, and the ranges are recut from the ranks'
        out2 = grp.frame()                               # measured times (sharding.h) -- the image does not depend on the cut
        assert np.array_equal(out2, img)
        ranges = [grp.slice_range(i) for i in range(ndev)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 24 and all(ranges[i][1] == ranges[i + 1][0] for i in range(ndev - 1))
    grp.close()


def test_sharded_handle_refuses_global_clustering(pkg):
    """a handle that owns a slice range holds zeros in the other rows of R, so the global Clustering object
    (globalCluster, fallback) must be refused instead of silently clustering zeros; changing the range invalidates R."""
    scene, vrls, params = small_case(pkg, "C1", 48, 48, 80, seed=2, targetNumSlices=10, globalCluster=1, globalUndersampling=10.0)
    g = setup(pkg.integrator(0, **params), scene, vrls)
    g.build_slices(); g.set_slice_range(0, 4); g.sample_slice_mapping(); g.build_R()
    with pytest.raises(pkg.binding.AlvrlError) as e:
        g.build_clusters()
    assert e.value.code == -5
    g.set_slice_range(2, 6)
    with pytest.raises(pkg.binding.AlvrlError) as e:
        g.build_clusters()                                # R of the old range is gone
    assert e.value.code == -2

"""The multi-GPU result gather (sloam_b200/csrc/comm.cu) on ONE device: a communicator of size 1
goes through the same entry points -- unique id, ncclCommInitRank, the all-gather on the side
stream, the wait that orders the next run behind it -- so the path is exercised by the
single-GPU test run too.  The N > 1 behaviour is checked by bench.py itself (every gathered
slice against a checksum of the producing rank's buffers, `run.gather_check`) and, for the
host-side sharding logic, by the gloo tests in tests/test_distributed.py."""
import numpy as np
import pytest

from sloam_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from sloam_b200 import capi as c
    return c


def _batch(capi, K):
    from sloam_b200 import configs
    p, cfg = configs.make(capi, "vlp-16", max_trees=64, max_map_models=64)
    pts, mask = capi.synth_generate_host(cfg, 0, K)
    scene = capi.synth_scene(cfg)
    PP, M = p.max_prev_planes, p.max_map_models
    pose = np.array([capi.synth_pose(cfg, k)[1] for k in range(K)])
    first = np.zeros(K, np.uint8)
    first[0] = 1
    maps = np.zeros((K, M), abi.CYLINDER)
    maps[:, :len(scene)] = scene
    inp = dict(points=pts, mask=mask, pose_est=pose, first_scan=first, map_models=maps,
               n_map_models=np.full(K, len(scene), np.int32), prev_planes=np.zeros((K, PP), abi.PLANE),
               n_prev_planes=np.zeros(K, np.int32))
    return p, {k: capi.to_dev(v) for k, v in inp.items()}


def test_gather_results_on_a_communicator_of_one(capi):
    K = 24
    p, d_in = _batch(capi, K)
    T = p.max_trees
    ctx = capi.Context(p, K)
    assert capi.lib().sloam_b200_comm_size(ctx.h) == 1 and capi.lib().sloam_b200_comm_rank(ctx.h) == 0
    ctx.comm_init(0, 1, capi.comm_unique_id())
    assert capi.lib().sloam_b200_comm_size(ctx.h) == 1
    out = ctx.alloc_outputs_dev(K)
    sizes = dict(results=K * abi.KF_RESULT.itemsize, matches=K * T * 4, tm=K * T * abi.CYLINDER.itemsize, tm_id=K * T * 4)
    gathered = {k: capi.dev_empty(n) for k, n in sizes.items()}
    for v in list(out.values()) + list(gathered.values()):
        if v is not None:
            v.zero_()
    ctx.sync()
    # two batches back to back: the second run's output kernels have to wait for the first gather
    for _ in range(2):
        ctx.run_keyframes_dev(K, d_in, out)
        ctx.gather_results(K, out, gathered)
    ctx.comm_wait()
    ctx.sync()
    res = capi.to_host(out["results"], abi.KF_RESULT, (K,))
    assert (res["success"] == 1).sum() >= K - 1 and (res["n_landmarks"] > 0).any()
    for k, n in sizes.items():
        assert np.array_equal(capi.to_host(gathered[k], np.uint8)[:n], capi.to_host(out[k], np.uint8)[:n]), k
    ctx.close()


def test_gather_without_a_communicator_is_an_error(capi):
    p, d_in = _batch(capi, 2)
    ctx = capi.Context(p, 2)
    out = ctx.alloc_outputs_dev(2)
    with pytest.raises(Exception):
        ctx.gather_results(2, out, out)
    ctx.comm_wait()  # a no-op without a pending gather
    ctx.close()

"""Ouster destagger (Segmentation::_destaggerCloud, sloam/src/segmentation/inference.cpp:200-228,
applied to the dense tree cloud by maskCloud :255-259).

CPU: the oracle against a line-by-line Python transcription of the reference loop (including
the `im_col > W` bound, SURVEY B-12) and against the closed form the GPU kernels use.
GPU: the stage entries and the fused path with do_destagger = 1 against the oracle."""
import numpy as np
import pytest

from sloam_b200 import abi


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from sloam_b200 import capi as c
    c.lib()
    return c


def reference_destagger(temp, H, W):
    """inference.cpp:200-228 / :255 transcribed statement by statement (out = copy of temp first)."""
    out = temp.copy()
    col_valid = True
    for irow in range(H):
        for icol in range(W):
            im_col = icol
            if irow % 2 == 0:
                im_col += 32
                if im_col < 0 or im_col > W:
                    col_valid = False
                    im_col = im_col % W
            if col_valid:
                dst = irow * W + im_col
                if dst < len(out):  # dst == H*W only for an even last row (UB in the reference)
                    out[dst]["x"], out[dst]["y"], out[dst]["z"] = temp[irow * W + icol][["x", "y", "z"]]
            col_valid = True
    return out


def closed_form(temp, H, W):
    out = temp.copy().reshape(H, W)
    t = temp.reshape(H, W)
    if W > 32:
        for f in ("x", "y", "z"):
            out[f][0::2, 32:] = t[f][0::2, :W - 32]
    return out.reshape(-1)


@pytest.mark.parametrize("H,W", [(6, 80), (5, 64), (4, 32), (3, 20), (8, 100)])
def test_oracle_destagger_matches_reference_loop(oracle, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    N = H * W
    pts = np.zeros(N, abi.POINT)
    for f in ("x", "y", "z", "intensity"):
        pts[f] = rng.normal(0, 5, N).astype(np.float32)
    mask = rng.choice(np.array([0, 1, 255], np.uint8), N)
    pix = rng.permutation(N).astype(np.int32)  # any pixel assignment: the destagger only sees the masked cloud
    p0 = oracle.default_params(img_h=H, img_w=W, do_destagger=0)
    p1 = oracle.default_params(img_h=H, img_w=W, do_destagger=1)
    temp, g0 = oracle.mask_cloud(p0, pts, pix, mask)
    got, g1 = oracle.mask_cloud(p1, pts, pix, mask)
    assert g0.tobytes() == g1.tobytes()                      # the sparse (ground) cloud is never destaggered
    assert got.tobytes() == reference_destagger(temp, H, W).tobytes()
    assert got.tobytes() == closed_form(temp, H, W).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,n_trees", [(64, 2048, 30), (16, 1800, 12), (64, 1024, 20)])
def test_gpu_destagger_stage_entries(capi, oracle, H, W, n_trees):
    K = 2
    cfg = capi.synth_config(H, W, n_trees)
    pts, mask = capi.synth_generate_host(cfg, 0, K)
    p = capi.default_params(img_h=H, img_w=W, do_destagger=1)
    ctx = capi.Context(p, K)
    N = H * W
    d_pts, d_mask = capi.to_dev(pts), capi.to_dev(mask)
    pix, rng, tree, ground, cnt = ctx.project_split(d_pts, d_mask, K)
    tree2, ground2, cnt2 = ctx.mask_cloud(d_pts, pix, d_mask, K)
    ctx.sync()
    tree = capi.to_host(tree, abi.POINT, (K, N))
    assert capi.to_host(tree2, abi.POINT, (K, N)).tobytes() == tree.tobytes()
    moved = 0
    for k in range(K):
        o_pix, _ = oracle.project(p, pts[k], want_range=False)
        o_tree, o_ground = oracle.mask_cloud(p, pts[k], o_pix, mask[k])
        assert tree[k].tobytes() == o_tree.tobytes()
        p0 = oracle.default_params(img_h=H, img_w=W, do_destagger=0)
        moved += int((oracle.mask_cloud(p0, pts[k], o_pix, mask[k])[0].tobytes() != o_tree.tobytes()))
    assert moved == K  # the destagger pass changed something
    ctx.close()


@pytest.mark.gpu
def test_gpu_destagger_fused_path(capi, oracle):
    """64 x 2048 scans (the reference's Ouster configuration, sloam.yaml:31-33 + ouster.yaml:5)
    through the fused path: results, landmarks and the dense tree cloud of the intermediates."""
    from test_gpu_parity import run_sequence, compare_keyframe, compare_graph
    H, W, K = 64, 2048, 3
    p = capi.default_params(img_h=H, img_w=W, do_destagger=1)
    cfg = capi.synth_config(H, W, 30)
    # an Ouster-style staggered scan: even rows arrive 32 columns early, which the destagger
    # pass undoes (the first 32 columns of those rows stay sheared, as with the real sensor)
    gen = capi.synth_generate_host

    def staggered(cfg_, k0, K_):
        pts, mask = gen(cfg_, k0, K_)
        grid = pts.reshape(K_, H, W)
        grid[:, 0::2, :] = np.roll(grid[:, 0::2, :], -32, axis=2)
        return grid.reshape(K_, H * W), mask
    capi.synth_generate_host = staggered
    try:
        inp, exp = run_sequence(capi, oracle, p, cfg, K, True)
    finally:
        capi.synth_generate_host = gen
    T, PP, N = p.max_trees, p.max_prev_planes, H * W
    ctx = capi.Context(p, K)
    out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32),
               tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32),
               planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32), range_image=None)
    d_in = {k: capi.to_dev(v) for k, v in inp.items()}
    d_out = ctx.alloc_outputs_dev(K)
    ctx.run_keyframes_dev(K, d_in, d_out)
    ctx.sync()
    res = capi.to_host(d_out["results"], abi.KF_RESULT, (K,))
    matches = capi.to_host(d_out["matches"], np.int32, (K, T))
    tm = capi.to_host(d_out["tm"], abi.CYLINDER, (K, T))
    tm_id = capi.to_host(d_out["tm_id"], np.int32, (K, T))
    planes = capi.to_host(d_out["planes"], abi.PLANE, (K, PP))
    npl = capi.to_host(d_out["n_planes"], np.int32, (K,))
    for k in range(K):
        compare_keyframe(res[k], matches[k], tm[k], tm_id[k], planes[k], npl[k], exp[k])
    assert res["n_trees"].min() > 3
    it = ctx.intermediates()
    tree = capi.read_dev(it.tree, K * N * abi.POINT.itemsize, ctx.device).view(np.uint32).reshape(K, N, 4)
    for k in range(K):
        e_tree, _ = oracle.mask_cloud(p, inp["points"][k], exp[k].pix, inp["mask"][k])
        assert np.array_equal(tree[k], e_tree.view(np.uint32).reshape(N, 4))
    ctx.close()

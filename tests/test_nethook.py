"""the tensor producer and the mask decoder either side of the segmentation
network (Segmentation::_makeTensor / _mask, inference.cpp:167-198, 275-300)."""
import numpy as np
import pytest

from sloam_b200 import abi


def test_oracle_make_tensor_known_values(oracle):
    r = np.array([0.0, 0.5, 0.999, 1.0, 12.97, 25.32, -0.5, 100.0], np.float32)
    t, inv, idx = oracle.make_tensor(r)
    # the reference tests the value converted to int (:183): |v| < 1 is "zero"
    assert list(inv) == [1, 1, 1, 0, 0, 0, 1, 0] and list(idx) == [0, 1, 2, 6]
    assert np.array_equal(t[inv == 1], r[inv == 1])          # invalid pixels pass through
    exp = (r[inv == 0] - np.float32(12.97)) / np.float32(12.35)
    assert np.array_equal(t[inv == 0], exp.astype(np.float32))
    assert t[4] == 0.0


def test_oracle_mask_first_maximum_wins(oracle):
    logits = np.array([[1, 0, 0, 2, 2, np.nan],      # class 0
                       [0, 1, 0, 2, 1, 1],           # class 1
                       [0, 0, 1, 2, 2, 5]], np.float32)  # class 2 -> 255
    m = oracle.mask_from_logits(logits)
    # ties keep the earlier class; NaN compares false, so class 0 = NaN is never replaced
    assert list(m) == [0, 1, 255, 0, 0, 0]
    m2 = oracle.mask_from_logits(logits, invalid=np.array([0, 1, 1, 0, 0, 0], np.uint8))
    assert list(m2) == [0, 0, 0, 0, 0, 0]


@pytest.mark.gpu
def test_gpu_nethook_bit_exact(oracle):
    from sloam_b200 import capi, configs
    K = 3
    p, cfg = configs.make(capi, "vlp-16")   # N = 28800, not a multiple of the block size
    N = p.img_h * p.img_w
    ctx = capi.Context(p, K)
    pts, mask = capi.synth_generate_host(cfg, 0, K)
    d_pts = capi.to_dev(pts)
    rng_img = capi.dev_empty(K * N * 4, ctx.device)
    pix = capi.dev_empty(K * N * 4, ctx.device)
    ctx.check(capi.lib().sloam_b200_project_dev(ctx.h, K, capi.dptr(d_pts), capi.dptr(pix), capi.dptr(rng_img)))
    tensor, invalid, n_inv = ctx.make_tensor(K, rng_img)
    ctx.sync()
    r = capi.to_host(rng_img, np.float32, (K, N))
    t = capi.to_host(tensor, np.float32, (K, N))
    inv = capi.to_host(invalid, np.uint8, (K, N))
    ninv = capi.to_host(n_inv, np.int32, (K,))
    assert (r > 0).sum() > K * N // 4
    for k in range(K):
        et, einv, eidx = oracle.make_tensor(r[k])
        assert np.array_equal(t[k].view(np.uint32), et.view(np.uint32))
        assert np.array_equal(inv[k], einv) and ninv[k] == len(eidx)
    # decoder: random scores with exact ties and NaNs
    g = np.random.default_rng(5)
    logits = g.integers(-3, 4, size=(K, 3, N)).astype(np.float32)
    logits[g.random((K, 3, N)) < 0.01] = np.nan
    got = capi.to_host(ctx.mask_from_logits(K, capi.to_dev(logits), invalid), np.uint8, (K, N))
    got_noinv = capi.to_host(ctx.mask_from_logits(K, capi.to_dev(logits)), np.uint8, (K, N))
    for k in range(K):
        assert np.array_equal(got[k], oracle.mask_from_logits(logits[k], inv[k]))
        assert np.array_equal(got_noinv[k], oracle.mask_from_logits(logits[k]))
    assert set(np.unique(got)) <= {0, 1, 255}
    ctx.close()

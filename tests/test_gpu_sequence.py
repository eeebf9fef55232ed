"""SURVEY 8(f)-1/2 on the GPU: the device-resident semantic map (MapManager::getSubmap /
updateMap, sloam/src/core/mapManager.cpp) and the sequential SLOAMNode::run call sequence
(sloam/src/core/sloamNode.cpp:186-282) against the oracle, keyframe by keyframe."""
import ctypes as C

import numpy as np
import pytest

from sloam_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available()
    from sloam_b200 import capi as c
    c.lib()
    return c


def quat_angle(qa, qb):
    return 2.0 * np.arccos(min(1.0, abs(float(np.dot(qa, qb)))))


@pytest.mark.parametrize("two_step", [True, False])
def test_sequence_matches_oracle(capi, oracle, two_step):
    from sloam_b200 import configs
    K = 12
    p, cfg = configs.make(capi, "os1-64", twoStepOptim=int(two_step))
    N, T, M = p.img_h * p.img_w, p.max_trees, p.max_map_models
    pts, mask = capi.synth_generate_host(cfg, 0, K)
    ctx = capi.Context(p, 1)
    assert capi.lib().sloam_b200_map_init(ctx.h, 4096) == 0
    omap = oracle.OracleMap()
    o_first, o_prev = True, np.zeros(0, abi.PLANE)
    n_opt = 0
    for k in range(K):
        pose = np.array([capi.synth_pose(cfg, k)[1]])     # poseEstimate = prevKeyPose * initialGuess
        # ---- oracle: getSubmap -> RunSloam -> updateMap
        sub, sub_idx = omap.get_submap(pose, M)
        e = oracle.run_keyframe(p, pts[k], mask[k], pose, o_first, sub, o_prev)
        ran = e.result["status"] in (abi.KF_OK, abi.KF_NOT_CONVERGED)
        n = int(e.result["n_landmarks"]) if ran else 0
        omap.update(e.tm[:n], e.tm_id[:n], e.matches[:n])
        o_prev = e.planes[:e.n_planes].copy()
        o_first = False
        # ---- GPU: one call
        res = np.zeros(1, abi.KF_RESULT)
        matches = np.zeros(T, np.int32); tm = np.zeros(T, abi.CYLINDER); tm_id = np.zeros(T, np.int32)
        rc = capi.lib().sloam_b200_sequence_step_host(ctx.h, abi.ptr(pts[k]), abi.ptr(mask[k]), abi.ptr(pose),
                                                      abi.ptr(res), abi.ptr(matches), abi.ptr(tm), abi.ptr(tm_id))
        assert rc == 0, capi.lib().sloam_b200_last_error(ctx.h)
        r, er = res[0], e.result
        for f in ("status", "success", "n_ground", "n_planes", "n_trees", "n_landmarks", "n_tree_matches",
                  "n_plane_matches"):
            assert r[f] == er[f], (k, f, r[f], er[f])
        assert np.array_equal(r["lm_termination"], er["lm_termination"]), k
        assert np.array_equal(matches[:n], e.matches[:n]), k              # association: bit-exact
        assert np.max(np.abs(r["T_Map_Curr"]["t"] - er["T_Map_Curr"]["t"])) <= 1e-5
        assert quat_angle(r["T_Map_Curr"]["q"], er["T_Map_Curr"]["q"]) <= 1e-5
        n_opt += int(er["lm_termination"][0] == 0)
        # ---- the maps stay identical
        gm, gh = np.zeros(4096, abi.CYLINDER), np.zeros(4096, np.int32)
        gn = capi.lib().sloam_b200_map_dump_host(ctx.h, abi.ptr(gm), abi.ptr(gh), 4096)
        om, oh = omap.dump(4096)
        assert gn == len(om), k
        assert np.array_equal(gh[:gn], oh)
        assert np.allclose(gm[:gn]["root"], om["root"], atol=1e-4) and np.allclose(gm[:gn]["radius"], om["radius"], atol=1e-4)
    assert n_opt >= K // 2 and omap.size() >= 10
    om, oh = omap.dump(4096)
    assert (oh > 2).sum() >= 5          # getMap() would publish these (hits > 2)
    ctx.close()


def test_submap_knn_and_recent_filter(capi, oracle):
    """getSubmap on a large synthetic map: kNN(100) by (distance, index), then only the last
    199 landmarks survive (mapManager.cpp:57-66); exact duplicates exercise the tie order."""
    rng = np.random.default_rng(9)
    p = capi.default_params(img_h=16, img_w=64, max_map_models=128, max_trees=512)
    ctx = capi.Context(p, 1)
    assert capi.lib().sloam_b200_map_init(ctx.h, 8192) == 0
    omap = oracle.OracleMap()
    n_total = 0
    for batch in range(12):
        n = 400
        tm = np.zeros(n, abi.CYLINDER)
        tm["root"][:, :2] = rng.uniform(-30, 30, (n, 2))
        tm["root"][:, 2] = rng.normal(1.0, 0.3, n)
        tm["root"][5] = tm["root"][4]                    # duplicate root: distance tie
        tm["ray"][:, 2] = 1.0
        tm["radius"] = 0.2
        ids = np.arange(n_total, n_total + n, dtype=np.int32)
        matches = np.full(n, -1, np.int32)
        res = np.zeros(1, abi.KF_RESULT); res["n_landmarks"] = n
        d_res, d_tm, d_ids, d_matches = capi.to_dev(res), capi.to_dev(tm), capi.to_dev(ids), capi.to_dev(matches)
        ctx.check(capi.lib().sloam_b200_map_update_dev(ctx.h, capi.dptr(d_res), capi.dptr(d_tm), capi.dptr(d_ids),
                                                        capi.dptr(d_matches)))
        ctx.sync()   # the device buffers must outlive the asynchronous kernel
        omap.update(tm, ids, matches)
        n_total += n
        for q in range(3):
            pose = np.zeros(1, abi.POSE); pose["q"][:, 3] = 1
            # queries near the most recent landmarks so that some of the 100 neighbours pass the filter
            pose["t"][0, :2] = tm["root"][rng.integers(n - 150, n), :2] + rng.normal(0, 0.5, 2)
            d_sub = capi.dev_empty(128 * abi.CYLINDER.itemsize); d_n = capi.dev_empty(4); d_pose = capi.to_dev(pose)
            ctx.check(capi.lib().sloam_b200_map_get_submap_dev(ctx.h, capi.dptr(d_pose), capi.dptr(d_sub),
                                                                capi.dptr(d_n)))
            ctx.sync()
            gn = int(capi.to_host(d_n, np.int32, (1,))[0])
            gsub = capi.to_host(d_sub, abi.CYLINDER, (128,))[:gn]
            esub, eidx = omap.get_submap(pose, 128)
            omap.update(np.zeros(0, abi.CYLINDER), np.zeros(0, np.int32), np.zeros(0, np.int32))  # clears matchesMap
            assert gn == len(esub)
            assert gsub.tobytes() == esub.tobytes()
    assert omap.size() == n_total
    ctx.close()


def test_capacities_are_reported_not_silent(capi, oracle):
    """more clusters than max_trees, a full semantic map, and counts beyond the
    buffer capacities must be visible to the caller (status flags / SLOAM_E_INVALID); the
    trees that are kept are the first max_trees big clusters in PCL label order."""
    from sloam_b200 import configs
    p, cfg = configs.make(capi, "os1-64")
    full = oracle.compute_graph(p, oracle.mask_cloud(p, *_scan(capi, oracle, p, cfg))[0])[0]
    assert len(full) >= 8
    small = p.copy(); small.max_trees = 4; small.max_map_models = 8
    pts, mask = capi.synth_generate_host(cfg, 0, 1)
    ctx = capi.Context(small, 1)
    assert capi.lib().sloam_b200_map_init(ctx.h, 1) == 0          # room for one landmark only
    T = small.max_trees
    flags, n_lm = [], []
    for rep in range(2):                                           # twice: identical (deterministic) truncation
        res = np.zeros(1, abi.KF_RESULT)
        matches = np.zeros(T, np.int32); tm = np.zeros(T, abi.CYLINDER); tm_id = np.zeros(T, np.int32)
        pose = np.array([capi.synth_pose(cfg, 0)[1]])
        rc = capi.lib().sloam_b200_sequence_step_host(ctx.h, abi.ptr(pts[0]), abi.ptr(mask[0]), abi.ptr(pose),
                                                      abi.ptr(res), abi.ptr(matches), abi.ptr(tm), abi.ptr(tm_id))
        assert rc == 0
        flags.append(int(res[0]["status"]))
        n_lm.append(int(res[0]["n_landmarks"]))
        it = ctx.intermediates()
        trees = capi.read_dev(it.trees, T * abi.TREE.itemsize, ctx.device).view(abi.TREE)
        ntr = int(capi.read_dev(it.n_trees, 4, ctx.device).view(np.int32)[0])
        # kept trees = oracle trees of the first clusters, same ids, in order
        want = [t for t in full["tree_id"]][:ntr]
        assert ntr >= 1 and list(trees[:ntr]["tree_id"]) == want
    assert flags[0] & 0x100 and flags[1] & 0x100                   # SLOAM_KF_FLAG_TREE_CAPACITY
    assert n_lm[0] >= 2 and flags[0] & 0x200                        # SLOAM_KF_FLAG_MAP_CAPACITY: 2+ new landmarks, room for 1
    assert (flags[0] & 0xFF) == abi.KF_OK
    ctx.close()
    # counts beyond the capacities: rejected by the host entry before any copy
    K = 1
    ctx = capi.Context(small, K)
    PP, M = small.max_prev_planes, small.max_map_models
    inp = dict(points=pts, mask=mask, pose_est=np.array([capi.synth_pose(cfg, 0)[1]]), first_scan=np.zeros(1, np.uint8),
               map_models=np.zeros((K, M), abi.CYLINDER), n_map_models=np.array([M + 1], np.int32),
               prev_planes=np.zeros((K, PP), abi.PLANE), n_prev_planes=np.zeros(K, np.int32))
    out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32),
               tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32),
               planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32), range_image=None)
    with pytest.raises(RuntimeError, match="exceed"):
        ctx.run_keyframes_host(K, inp, out)
    ctx.close()


def _scan(capi, oracle, p, cfg):
    pts, mask = capi.synth_generate_host(cfg, 0, 1)
    pix, _ = oracle.project(p, pts[0], want_range=False)
    return pts[0], pix, mask[0]

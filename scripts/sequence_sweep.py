"""Long sequential runs (GPU box): sloam_b200_sequence_step_host (device MapManager, device
firstScan_/prevGPlanes_) against the oracle's getSubmap -> RunSloam -> updateMap loop, keyframe
by keyframe, including the map contents.  Long enough for the map to exceed 200 landmarks (the
"last 200" filter) and 100 neighbours (kNN cut).
usage: python scripts/sequence_sweep.py [keyframes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from sloam_b200 import abi, capi, configs

K = int(sys.argv[1]) if len(sys.argv) > 1 else 150
CAP = 16384
BIG = len(sys.argv) > 2 and sys.argv[2] == "big"   # a dense, long drive: the map outgrows 200 landmarks
CASES = [("vlp-16", True, 7, 0.6)] if BIG else [("os1-64", True, 3, 0.35), ("vlp-16", True, 4, 0.5), ("os1-64", False, 5, 0.8),
                                                 ("vlp-16", False, 6, 1.2)]
for preset, two_step, seed, step in CASES:
    p, cfg = configs.make(capi, preset, twoStepOptim=int(two_step))
    if preset == "vlp-16" and not two_step:
        p.minGroundModels = 10
    cfg.seed, cfg.step_per_keyframe, cfg.n_trees = seed, step, (2000 if BIG else 400)   # a long drive through a big forest
    cfg.tree_r_max = 100.0 if BIG else 60.0
    T, M = p.max_trees, p.max_map_models
    pts, mask = capi.synth_generate_host(cfg, 0, K)
    ctx = capi.Context(p, 1)
    assert capi.lib().sloam_b200_map_init(ctx.h, CAP) == 0
    omap = orc.OracleMap()
    o_first, o_prev = True, np.zeros(0, abi.PLANE)
    bad = 0
    for k in range(K):
        pose = np.array([capi.synth_pose(cfg, k)[1]])
        sub, _ = omap.get_submap(pose, M)
        e = orc.run_keyframe(p, pts[k], mask[k], pose, o_first, sub, o_prev)
        ran = e.result["status"] in (abi.KF_OK, abi.KF_NOT_CONVERGED)
        n = int(e.result["n_landmarks"]) if ran else 0
        omap.update(e.tm[:n], e.tm_id[:n], e.matches[:n])
        o_prev = e.planes[:e.n_planes].copy()
        o_first = False
        res = np.zeros(1, abi.KF_RESULT)
        matches = np.zeros(T, np.int32); tm = np.zeros(T, abi.CYLINDER); tm_id = np.zeros(T, np.int32)
        rc = capi.lib().sloam_b200_sequence_step_host(ctx.h, abi.ptr(pts[k]), abi.ptr(mask[k]), abi.ptr(pose), abi.ptr(res),
                                                      abi.ptr(matches), abi.ptr(tm), abi.ptr(tm_id))
        assert rc == 0, capi.lib().sloam_b200_last_error(ctx.h)
        r, er = res[0], e.result
        ok = all(r[f] == er[f] for f in ("status", "success", "n_ground", "n_planes", "n_trees", "n_landmarks",
                                         "n_tree_matches", "n_plane_matches"))
        ok = ok and np.array_equal(r["lm_termination"], er["lm_termination"]) and np.array_equal(matches[:n], e.matches[:n])
        ok = ok and np.max(np.abs(r["T_Map_Curr"]["t"] - er["T_Map_Curr"]["t"])) <= 1e-5
        if ok and k % 10 == 9:
            gm, gh = np.zeros(CAP, abi.CYLINDER), np.zeros(CAP, np.int32)
            gn = capi.lib().sloam_b200_map_dump_host(ctx.h, abi.ptr(gm), abi.ptr(gh), CAP)
            om, oh = omap.dump(CAP)
            ok = gn == len(om) and np.array_equal(gh[:gn], oh) and np.allclose(gm[:gn]["root"], om["root"], atol=1e-4)
        if not ok:
            bad += 1
            if bad <= 3:
                print(f"  MISMATCH {preset} two_step={two_step} kf={k}: gpu", [int(r[f]) for f in ("status", "n_landmarks", "n_tree_matches")],
                      "oracle", [int(er[f]) for f in ("status", "n_landmarks", "n_tree_matches")], "map", omap.size())
    print(f"{preset} two_step={two_step}: {K} keyframes, {bad} mismatching, final map {omap.size()} landmarks, submap {len(sub)}", flush=True)
    ctx.close()

"""The reference's `moving` fixture pair + poses.txt (sloam/src/tests/aux/moving_{ground,tree}_t{0,1}.pcd,
moving_landmarks_t{0,1}, poses.txt:11-14,46-49): real-sensor data the reference ships but its
own gtests never run (core_test.cpp loads only the `still` pair).  The assertions of
core_test.cpp:148-194 (RunSloam true, pose, >= 1 association) are applied to it with the pose
guess = the relative odometry of poses.txt, as SLOAMNode::run would pass it
(sloamNode.cpp:192), for the oracle (CPU) and for the CUDA path (GPU == oracle on every field).
Golden vectors: tests/golden/moving_*.npz, moving_poses.json (scripts/make_golden.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import golden_io
from sloam_b200 import abi

EMPTY_MAP = np.zeros(0, abi.CYLINDER)
EMPTY_PLANES = np.zeros(0, abi.PLANE)


def quat_mul(a, b):
    x1, y1, z1, w1 = a
    x2, y2, z2, w2 = b
    return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                     w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])


def quat_rot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R @ np.asarray(v)


def odometry_step():
    """T_t0^-1 * T_t1 of poses.txt: the initialGuess of SLOAMNode::run for keyframe t1."""
    with open(os.path.join(golden_io.GOLDEN, "moving_poses.json")) as f:
        poses = json.load(f)
    q0, t0 = np.array(poses["t0"]["q"]), np.array(poses["t0"]["t"])
    q1, t1 = np.array(poses["t1"]["q"]), np.array(poses["t1"]["t"])
    q0i = q0 * np.array([-1.0, -1.0, -1.0, 1.0])
    est = np.zeros(1, abi.POSE)
    est["t"][0] = quat_rot(q0i, t1 - t0)
    est["q"][0] = quat_mul(q0i, q1)
    return est


def moving_input(stamp):
    trees, verts, vpts = golden_io.Trellis("moving", stamp).as_flat()
    return golden_io.ground("moving", stamp), trees, verts, vpts


# shipped parameters (sloam/params/sloam.yaml) and a relaxed acceptance under which enough of the
# fixture's trunks become landmarks for the tree residuals to take part in the optimisation
CASES = {
    "yaml": dict(),
    "relaxed": dict(maxLidarDist=40.0, maxTreeRadius=0.6, maxAxisTheta=25.0, minTreeModels=2, treeMatchThresh=1.0,
                    featuresPerTree=4, minGroundModels=10),
}


def params_for(mod, case, two_step):
    return mod.default_params(img_h=64, img_w=2048, twoStepOptim=int(two_step), **CASES[case])


def oracle_pair(oracle, p):
    g0, tr0, ve0, vp0 = moving_input("t0")
    g1, tr1, ve1, vp1 = moving_input("t1")
    o0 = oracle.run_sloam(p, g0, tr0, ve0, vp0, oracle.identity_pose(), True, EMPTY_MAP, EMPTY_PLANES)
    n0 = int(o0.result["n_landmarks"])
    o1 = oracle.run_sloam(p, g1, tr1, ve1, vp1, odometry_step(), False, o0.tm[:n0], o0.planes[:o0.n_planes])
    return o0, o1


def check_pair(o0, o1, case, two_step):
    step = float(np.linalg.norm(odometry_step()["t"][0]))
    assert abs(step - 0.627) < 1e-3                          # poses.txt: the two odometry poses are 0.63 m apart
    assert o0.result["success"] == 1 and o0.result["n_landmarks"] > 0 and o0.n_planes > 0   # FirstScan
    r = o1.result
    assert r["success"] == 1                                 # SLOAMSucess
    n1 = int(r["n_landmarks"])
    assert np.any(o1.matches[:n1] != -1)                     # ObjectAssociation
    t = np.linalg.norm(r["T_Map_Curr"]["t"])
    if case == "yaml":
        assert abs(t - step) < 0.1                           # PoseOptimization, around the odometry step
        if two_step:
            assert r["lm_termination"][1] == 0               # the ground problem ran and converged
    else:
        assert r["lm_termination"][0] == 0 and r["lm_iterations"][0] >= 2   # the tree residuals were optimised
        assert int((o1.matches[:n1] != -1).sum()) >= 5
        assert abs(t - step) < 0.2                           # SLOAM's estimate vs. the odometry prior
    assert abs(np.linalg.norm(r["T_Map_Curr"]["q"]) - 1) < 1e-9


@pytest.mark.parametrize("two_step", [False, True])
@pytest.mark.parametrize("case", ["yaml", "relaxed"])
def test_oracle_moving_pair(oracle, case, two_step):
    o0, o1 = oracle_pair(oracle, params_for(oracle, case, two_step))
    check_pair(o0, o1, case, two_step)


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from sloam_b200 import capi as c
    c.lib()
    return c


class GpuOut:
    pass


def gpu_run_sloam(capi, ctx, p, ground, trees, verts, vpts, pose_est, first, map_models, prev_planes):
    """sloam_b200_run_sloam_dev for one keyframe (K = 1)."""
    N, T, M, PP = p.img_h * p.img_w, p.max_trees, p.max_map_models, p.max_prev_planes
    dev = ctx.device

    def padded(a, cap, dtype):
        out = np.zeros(cap, dtype)
        out[:len(a)] = a
        return out
    assert len(ground) <= N and len(trees) <= T and len(vpts) <= N
    d = dict(
        ground=capi.to_dev(padded(ground, N, abi.POINT), dev), gcount=capi.to_dev(np.array([len(ground)], np.int32), dev),
        trees=capi.to_dev(padded(trees, T, abi.TREE), dev), ntrees=capi.to_dev(np.array([len(trees)], np.int32), dev),
        verts=capi.to_dev(padded(verts, max(len(verts), 1), abi.VERTEX), dev),
        vpts=capi.to_dev(padded(vpts, N, abi.POINT), dev),
        pose=capi.to_dev(pose_est, dev), first=capi.to_dev(np.array([1 if first else 0], np.uint8), dev),
        map=capi.to_dev(padded(map_models, M, abi.CYLINDER), dev), nmap=capi.to_dev(np.array([len(map_models)], np.int32), dev),
        prev=capi.to_dev(padded(prev_planes, PP, abi.PLANE), dev), nprev=capi.to_dev(np.array([len(prev_planes)], np.int32), dev))
    out = ctx.alloc_outputs_dev(1)
    bi = abi.BatchIn(None, None, capi.dptr(d["pose"]), capi.dptr(d["first"]), capi.dptr(d["map"]), capi.dptr(d["nmap"]), 0,
                     capi.dptr(d["prev"]), capi.dptr(d["nprev"]))
    bo = abi.BatchOut(capi.dptr(out["results"]), capi.dptr(out["matches"]), capi.dptr(out["tm"]), capi.dptr(out["tm_id"]),
                      capi.dptr(out["planes"]), capi.dptr(out["n_planes"]), None)
    ctx.check(capi.lib().sloam_b200_run_sloam_dev(ctx.h, 1, capi.dptr(d["ground"]), capi.dptr(d["gcount"]), N,
                                                  capi.dptr(d["trees"]), capi.dptr(d["ntrees"]), capi.dptr(d["verts"]),
                                                  max(len(verts), 1), capi.dptr(d["vpts"]), N, C.byref(bi), C.byref(bo)))
    ctx.sync()
    o = GpuOut()
    o.result = capi.to_host(out["results"], abi.KF_RESULT, (1,))[0]
    o.matches = capi.to_host(out["matches"], np.int32, (T,))
    o.tm = capi.to_host(out["tm"], abi.CYLINDER, (T,))
    o.tm_id = capi.to_host(out["tm_id"], np.int32, (T,))
    o.planes = capi.to_host(out["planes"], abi.PLANE, (PP,))
    o.n_planes = int(capi.to_host(out["n_planes"], np.int32, (1,))[0])
    return o


def same_as_oracle(g, e):
    for f in ("status", "success", "n_planes", "n_landmarks", "n_tree_matches", "n_plane_matches"):
        assert g.result[f] == e.result[f], f
    assert np.array_equal(g.result["lm_termination"], e.result["lm_termination"])
    assert np.array_equal(g.result["lm_iterations"], e.result["lm_iterations"])
    n = int(e.result["n_landmarks"])
    assert np.array_equal(g.matches[:n], e.matches[:n])                     # association indices: bit-exact
    assert np.array_equal(g.tm_id[:n], e.tm_id[:n])
    assert np.max(np.abs(g.result["T_Map_Curr"]["t"] - e.result["T_Map_Curr"]["t"])) <= 1e-5
    assert np.max(np.abs(g.result["T_Map_Curr"]["q"] - e.result["T_Map_Curr"]["q"])) <= 1e-5
    assert np.allclose(g.tm[:n]["root"], e.tm[:n]["root"], atol=1e-4) and np.allclose(g.tm[:n]["radius"], e.tm[:n]["radius"], atol=1e-4)
    assert g.n_planes == e.n_planes and np.allclose(g.planes[:g.n_planes]["plane"], e.planes[:e.n_planes]["plane"], atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("two_step", [False, True])
@pytest.mark.parametrize("case", ["yaml", "relaxed"])
def test_gpu_moving_pair(capi, oracle, case, two_step):
    p = params_for(capi, case, two_step)
    e0, e1 = oracle_pair(oracle, p)
    ctx = capi.Context(p, 1)
    g0, tr0, ve0, vp0 = moving_input("t0")
    g1, tr1, ve1, vp1 = moving_input("t1")
    o0 = gpu_run_sloam(capi, ctx, p, g0, tr0, ve0, vp0, oracle.identity_pose(), True, EMPTY_MAP, EMPTY_PLANES)
    same_as_oracle(o0, e0)
    n0 = int(o0.result["n_landmarks"])
    o1 = gpu_run_sloam(capi, ctx, p, g1, tr1, ve1, vp1, odometry_step(), False, o0.tm[:n0], o0.planes[:o0.n_planes])
    same_as_oracle(o1, e1)
    check_pair(o0, o1, case, two_step)      # the reference's assertions hold for the CUDA path itself
    ctx.close()

"""The reference's 12 gtest assertions (sloam/src/tests/{plane,cylinder,core}_test.cpp)
restated against the CPU oracle on the reference's own `still` fixtures
(committed as tests/golden/*.npz).  The three parameters the reference tests
leave uninitialised (core_test.cpp:94-119: minTreeModels, minGroundModels,
twoStepOptim; SURVEY B-16) are set explicitly and both optimiser variants run."""
import numpy as np
import pytest

import golden_io
from sloam_b200 import abi


def pts(xyz):
    a = np.zeros(len(xyz), abi.POINT)
    a["x"], a["y"], a["z"] = np.asarray(xyz, np.float32).T
    return a


# ------------------------------------------------------------------ plane_test.cpp
def test_plane_initializes(oracle):  # plane_test.cpp:57-66
    pl = oracle.plane_fit(pts([[0, 0, 0], [0, 1, 0], [1, 0, 0]]), 1)
    assert pl["is_valid"] == 1


def test_plane_distance_to_feature(oracle):  # plane_test.cpp:68-80
    p3 = pts([[0, 0, 0], [0, 1, 0], [1, 0, 0]])
    pl = oracle.plane_fit(p3, 1)
    m = np.array([pl["model"]])
    d = oracle.lib().orc_plane_distance_point(abi.ptr(m), abi.ptr(p3[:1]))
    assert abs(d) < 0.1
    # stronger than the reference: the fitted normal is +-z and d = 0
    assert np.allclose(np.abs(pl["model"]["plane"][:3]), [0, 0, 1], atol=1e-12)
    assert abs(pl["model"]["plane"][3]) < 1e-7


def test_plane_translate_model(oracle):  # plane_test.cpp:82-97
    pl = oracle.plane_fit(pts([[0, 0, 0], [0, 1, 0], [1, 0, 0]]), 1)
    m = np.array([pl["model"]])
    cz = m[0]["centroid"][2]
    tf = oracle.identity_pose()
    tf["t"][0, 2] = 1
    oracle.lib().orc_plane_project(abi.ptr(m), abi.ptr(tf))
    assert m[0]["centroid"][2] == cz + 1


# --------------------------------------------------------------- cylinder_test.cpp
def cylinder_test_setup(oracle):
    """cylinder_test.cpp:20-62: a 4-point z=0 plane and still_landmarks_t0."""
    p = oracle.default_params(
        maxLidarDist=30.0, maxGroundLidarDist=30.0, minGroundLidarDist=0.0, groundRadiiBins=1,
        groundThetaBins=1, groundRetainThresh=0.1, maxTreeRadius=0.3, maxAxisTheta=10,
        treeMatchThresh=1.0, AddNewTreeThreshDist=2.0, featuresPerTree=2, numGroundFeatures=3,
        defaultTreeRadius=0.1, img_h=64, img_w=2048)
    plane = oracle.plane_fit(pts([[0, 0, 0], [0, 1, 0], [1, 0, 0], [1, 1, 0]]), 3)
    assert plane["is_valid"] == 1
    cells = np.zeros(1, abi.CELL_PLANE)
    cells[0] = plane
    cells[0]["accepted"] = 1
    trees, verts, vpts = golden_io.Trellis("still", "t0").as_flat()
    models, feats = oracle.cylinders(p, trees, verts, vpts, cells)
    return p, models, feats


def test_cylinder_initializes(oracle):  # cylinder_test.cpp:69-77
    p, models, feats = cylinder_test_setup(oracle)
    assert len(models) == 14


def test_cylinder_distance_to_model(oracle):  # cylinder_test.cpp:79-93
    p, models, feats = cylinder_test_setup(oracle)
    valid = models[models["is_valid"] == 1]
    assert len(valid) > 0
    m = np.array([valid[0]["model"]])
    assert oracle.lib().orc_cylinder_distance_model(abi.ptr(m), abi.ptr(m)) == 0.0


def test_cylinder_distance_to_feature(oracle):  # cylinder_test.cpp:95-108
    p, models, feats = cylinder_test_setup(oracle)
    idx = np.flatnonzero(models["is_valid"] == 1)
    m = np.array([models[idx[0]]["model"]])
    f0 = np.ascontiguousarray(feats[idx[0]][:1])
    d = oracle.lib().orc_cylinder_distance_point(abi.ptr(m), abi.ptr(f0))
    assert abs(d) < 0.1
    # features carry the tree id as intensity (cylinder.cpp:89)
    assert f0["intensity"][0] == float(models[idx[0]]["id"])


def test_cylinder_translate_model(oracle):  # cylinder_test.cpp:110-127
    p, models, feats = cylinder_test_setup(oracle)
    valid = models[models["is_valid"] == 1]
    m = np.array([valid[0]["model"]])
    rx = m[0]["root"][0]
    tf = oracle.identity_pose()
    tf["t"][0, 0] = 1
    oracle.lib().orc_cylinder_project(abi.ptr(m), abi.ptr(tf))
    assert m[0]["root"][0] == rx + 1


# ------------------------------------------------------------------- core_test.cpp
def core_params(oracle, two_step):
    """core_test.cpp:94-119 plus the YAML values for what it leaves unset."""
    return oracle.default_params(
        maxLidarDist=15.0, maxGroundLidarDist=30.0, minGroundLidarDist=0.0, groundRadiiBins=1,
        groundThetaBins=18, groundRetainThresh=0.05, maxTreeRadius=0.3, maxAxisTheta=10,
        treeMatchThresh=1.0, AddNewTreeThreshDist=2.0, featuresPerTree=2, numGroundFeatures=60,
        defaultTreeRadius=0.1, minTreeModels=5, minGroundModels=36, twoStepOptim=int(two_step),
        img_h=64, img_w=2048)


def core_input(stamp):
    """core_test.cpp:72-92: landmarks with every vertex's points resized to 5."""
    trees, verts, vpts = golden_io.Trellis("still", stamp).as_flat()
    nv = int(trees["n_vertices"].sum())
    new_pts = np.zeros(nv * 5, abi.POINT)  # resize(5): truncate or zero-pad
    for k in range(nv):
        b, n = verts["point_begin"][k], min(verts["n_points"][k], 5)
        new_pts[k * 5:k * 5 + n] = vpts[b:b + n]
        verts["point_begin"][k] = k * 5
        verts["n_points"][k] = 5
    trees["n_points"] = trees["n_vertices"] * 5
    return golden_io.ground("still", stamp), trees, verts, new_pts


EMPTY_MAP = np.zeros(0, abi.CYLINDER)
EMPTY_PLANES = np.zeros(0, abi.PLANE)


@pytest.mark.parametrize("stamp", ["t0", "t1"])
def test_first_scan(oracle, stamp):  # core_test.cpp:128-146 (FirstScan, SecondScan)
    p = core_params(oracle, False)
    g, trees, verts, vpts = core_input(stamp)
    out = oracle.run_sloam(p, g, trees, verts, vpts, oracle.identity_pose(), True, EMPTY_MAP,
                           EMPTY_PLANES)
    assert out.result["n_landmarks"] > 0   # out.tm.size() > 0
    assert out.n_planes > 0                # getPrevGroundModel().size() > 0
    assert out.result["success"] == 1
    assert np.all(out.matches[:out.result["n_landmarks"]] == -1)


@pytest.mark.parametrize("two_step", [False, True])
def test_sloam_success_pose_and_association(oracle, two_step):
    """core_test.cpp:148-194: SLOAMSucess, PoseOptimization, ObjectAssociation."""
    p = core_params(oracle, two_step)
    g0, tr0, ve0, vp0 = core_input("t0")
    g1, tr1, ve1, vp1 = core_input("t1")
    ident = oracle.identity_pose()
    out0 = oracle.run_sloam(p, g0, tr0, ve0, vp0, ident, True, EMPTY_MAP, EMPTY_PLANES)
    n0 = out0.result["n_landmarks"]
    map_models = out0.tm[:n0]                      # t1Input.mapModels = out.tm
    prev = out0.planes[:out0.n_planes]
    out1 = oracle.run_sloam(p, g1, tr1, ve1, vp1, ident, False, map_models, prev)
    assert out1.result["success"] == 1             # SLOAMSucess
    t = out1.result["T_Delta"]["t"]
    # PoseOptimization asserts |T_Delta.t| < 0.1.  In two-step mode T_Delta stays
    # identity (SURVEY B-17) so the assertion is trivially true -- which is also what
    # the reference test most likely exercises (twoStepOptim is uninitialised there,
    # core_test.cpp:94-119).  In joint mode only 18 tree residuals (9 trees x 2 lowest
    # trunk points) constrain x/y and the LM optimum sits 0.15 m away; the LM itself is
    # cross-checked against scipy in test_oracle_lm.py.
    assert np.linalg.norm(t) < (0.1 if two_step else 0.2)
    n1 = out1.result["n_landmarks"]
    assert np.any(out1.matches[:n1] != -1)         # ObjectAssociation
    # stronger than the reference: the `still` pair barely moves
    tm = out1.result["T_Map_Curr"]
    assert np.linalg.norm(tm["t"]) < 0.2
    assert abs(np.linalg.norm(tm["q"]) - 1) < 1e-12 and abs(tm["q"][3]) > 0.999

"""Host build of the device pose-optimiser code (sloam_b200/csrc/dev_lm.h, compiled by
tests/hd_check.cpp with a serial evaluator) against the oracle and finite differences.

Runs without a GPU: it checks the scalar code lm_kernel executes (residuals, analytic
tangent-space Jacobians, the trust-region state machine), not the kernel itself -- the
kernel is covered by tests/test_gpu_parity.py and tests/test_gpu_sequence.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import synth_matches as sm
from sloam_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hd(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hd") / "hd_check.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "hd_check.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.hd_residual_row.restype = C.c_double
    return lib


def _x_joint(pose):
    return np.array([*pose["q"], *pose["t"]], np.float64)


def _solve(hd, mode, x, pb, use_t=True, use_p=True, max_it=50):
    tf = np.ascontiguousarray(pb["tree_feat"], np.float64)
    pf = np.ascontiguousarray(pb["plane_feat"], np.float64)
    to = np.ascontiguousarray(pb["tree_obj"])
    po = np.ascontiguousarray(pb["plane_obj"])
    it = C.c_int()
    costs = np.zeros(2)
    term = hd.hd_lm_solve(mode, abi.ptr(x), abi.ptr(tf), abi.ptr(to), len(to) if use_t else 0,
                          abi.ptr(pf), abi.ptr(po), len(po) if use_p else 0, C.c_double(0.1), max_it,
                          C.byref(it), abi.ptr(costs))
    return term, it.value, costs


@pytest.mark.parametrize("seed", [0, 1, 2, 5])
def test_joint_state_machine_matches_the_oracle(hd, oracle, seed):
    pb = sm.make_problem(seed)
    p = oracle.default_params()
    out, it, term = oracle.optimize_pose(p, 0, pb["guess"], pb["tree_feat"], pb["tree_obj"],
                                         pb["plane_feat"], pb["plane_obj"])
    x = _x_joint(pb["guess"][0])
    t, n_it, costs = _solve(hd, 0, x, pb)
    # same termination and iteration count: the accept/reject sequence is identical
    assert (t, n_it) == (int(term[0]), int(it[0]))
    assert costs[1] < costs[0]
    # tolerance of north_star: pose 1e-5 m / 1e-5 rad (normal equations vs the oracle's QR)
    assert np.allclose(x[4:], out["t"], atol=1e-5)
    q = x[:4] / np.linalg.norm(x[:4])
    assert min(np.abs(q - out["q"]).max(), np.abs(q + out["q"]).max()) < 1e-5


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_analytic_jacobian_against_central_differences(hd, mode):
    pb = sm.make_problem(7)
    g = pb["guess"][0]
    if mode == 0:
        x = _x_joint(g)
        n = 6
    else:
        from scipy.spatial.transform import Rotation as R
        x = np.array([*g["t"], *R.from_quat(g["q"]).as_rotvec(), 0.0])
        n = 3
    rows = []
    if mode != 2:
        rows += [(pb["tree_feat"][i], pb["tree_obj"][i:i + 1], None) for i in range(0, 40, 7)]
    if mode != 1:
        rows += [(pb["plane_feat"][i], None, pb["plane_obj"][i:i + 1]) for i in range(0, 40, 7)]
    h = 1e-6
    for feat, cyl, pl in rows:
        feat = np.ascontiguousarray(feat, np.float64)
        cp = abi.ptr(np.ascontiguousarray(cyl)) if cyl is not None else None
        pp = abi.ptr(np.ascontiguousarray(pl)) if pl is not None else None
        J = np.zeros(6)
        hd.hd_residual_row(mode, abi.ptr(x), abi.ptr(feat), cp, pp, abi.ptr(J))
        for c in range(n):
            d = np.zeros(6)
            xp, xm = np.zeros(7), np.zeros(7)
            d[c] = h
            hd.hd_plus(mode, abi.ptr(x), abi.ptr(d), abi.ptr(xp))
            d[c] = -h
            hd.hd_plus(mode, abi.ptr(x), abi.ptr(d), abi.ptr(xm))
            rp = hd.hd_residual_row(mode, abi.ptr(xp), abi.ptr(feat), cp, pp, None)
            rm = hd.hd_residual_row(mode, abi.ptr(xm), abi.ptr(feat), cp, pp, None)
            assert abs((rp - rm) / (2 * h) - J[c]) < 1e-5 * max(1.0, abs(J[c]))


def test_rejected_steps_reuse_the_same_normal_matrix(hd, oracle):
    """A start far from the optimum makes the first trust-region steps fail; the state
    machine must shrink the radius and retry on the SAME linearisation (this is the path
    that a compiler-introduced aliasing of the matrix once broke on the device)."""
    pb = sm.make_problem(11)
    p = oracle.default_params()
    g = pb["guess"].copy()
    g["t"][0] += np.array([1.5, -1.2, 0.4])
    out, it, term = oracle.optimize_pose(p, 0, g, pb["tree_feat"], pb["tree_obj"],
                                         pb["plane_feat"], pb["plane_obj"])
    x = _x_joint(g[0])
    t, n_it, costs = _solve(hd, 0, x, pb)
    assert (t, n_it) == (int(term[0]), int(it[0]))
    assert np.allclose(x[4:], out["t"], atol=1e-5)


def test_empty_problem_is_convergence(hd):
    pb = sm.make_problem(4)
    x = _x_joint(pb["guess"][0])
    x0 = x.copy()
    t, n_it, _ = _solve(hd, 0, x, pb, use_t=False, use_p=False)
    assert (t, n_it) == (0, 0) and np.array_equal(x, x0)


@pytest.mark.parametrize("seed", [0, 3, 8])
def test_two_step_state_machines_match_the_oracle(hd, oracle, seed):
    """TwoStepOptimizePose (sloam.cpp:48-152): XYYaw over the tree residuals and ZRollPitch over
    the plane residuals, both started at the pose estimate, composed like pose_pack_kernel."""
    from scipy.spatial.transform import Rotation as R
    pb = sm.make_problem(seed)
    p = oracle.default_params()
    out, it, term = oracle.optimize_pose(p, 1, pb["guess"], pb["tree_feat"], pb["tree_obj"],
                                         pb["plane_feat"], pb["plane_obj"])
    g = pb["guess"][0]
    x0 = np.array([*g["t"], *R.from_quat(g["q"]).as_rotvec(), 0.0])
    x1, x2 = x0.copy(), x0.copy()
    t1, n1, c1 = _solve(hd, 1, x1, pb)
    t2, n2, c2 = _solve(hd, 2, x2, pb)
    assert (t1, n1) == (int(term[0]), int(it[0]))
    assert (t2, n2) == (int(term[1]), int(it[1]))
    assert c1[1] < c1[0] and c2[1] < c2[0]
    # XYYaw moves only x, y, yaw; ZRollPitch only z, roll, pitch (SubsetParameterization)
    assert np.array_equal(x1[[2, 3, 4]], x0[[2, 3, 4]]) and np.array_equal(x2[[0, 1, 5]], x0[[0, 1, 5]])
    t = np.array([x1[0], x1[1], x2[2]])
    q = R.from_rotvec([x2[3], x2[4], x1[5]]).as_quat()
    assert np.allclose(t, out["t"], atol=1e-5)
    assert 2.0 * min(np.linalg.norm(q - out["q"]), np.linalg.norm(q + out["q"])) < 1e-5


@pytest.mark.parametrize("kind", sm.DEGENERATE)
@pytest.mark.parametrize("mode", [0, 1])
def test_ill_conditioned_problems_follow_the_oracle(hd, oracle, kind, mode):
    """Rank-deficient / badly conditioned geometry (collinear trunks, one or two planes,
    zero-padded features, a single trunk, fewer rows than unknowns): the device solver
    (square-root-free LDL^T on the damped normal equations) must make the same accept / reject
    decisions as the oracle's Ceres restatement (QR on the augmented Jacobian): identical
    termination type and iteration count, pose within the north-star tolerance."""
    from scipy.spatial.transform import Rotation as R
    pb = sm.make_degenerate(kind, seed=3)
    p = oracle.default_params()
    out, it, term = oracle.optimize_pose(p, mode, pb["guess"], pb["tree_feat"], pb["tree_obj"],
                                         pb["plane_feat"], pb["plane_obj"])
    g = pb["guess"][0]
    if mode == 0:
        x = _x_joint(g)
        t, n_it, _ = _solve(hd, 0, x, pb)
        assert (t, n_it) == (int(term[0]), int(it[0])), (kind, t, n_it, term, it)
        if t == 0:
            assert np.allclose(x[4:], out["t"], atol=1e-5)
            q = x[:4] / np.linalg.norm(x[:4])
            assert min(np.abs(q - out["q"]).max(), np.abs(q + out["q"]).max()) < 1e-5
    else:
        x0 = np.array([*g["t"], *R.from_quat(g["q"]).as_rotvec(), 0.0])
        x1, x2 = x0.copy(), x0.copy()
        t1, n1, _ = _solve(hd, 1, x1, pb)
        t2, n2, _ = _solve(hd, 2, x2, pb)
        assert (t1, n1) == (int(term[0]), int(it[0])), (kind, "xyyaw")
        assert (t2, n2) == (int(term[1]), int(it[1])), (kind, "zrollpitch")
        # TwoStepOptimizePose keeps the estimate of a block that did not converge (sloam.cpp:43-50)
        t = np.array([x1[0] if t1 == 0 else x0[0], x1[1] if t1 == 0 else x0[1], x2[2] if t2 == 0 else x0[2]])
        q = R.from_rotvec([x2[3] if t2 == 0 else x0[3], x2[4] if t2 == 0 else x0[4], x1[5] if t1 == 0 else x0[5]]).as_quat()
        assert np.allclose(t, out["t"], atol=1e-5)
        assert 2.0 * min(np.linalg.norm(q - out["q"]), np.linalg.norm(q + out["q"])) < 1e-5

"""Cross-check the oracle's Ceres-style LM (oracle/orc_ceres.cpp, restating
sloam.cpp:33-255 + SURVEY A.7) against an independent solver (scipy
least_squares with the same Huber loss) on synthetic matched features."""
import numpy as np
import pytest
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation as R

import synth_matches as sm


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_joint_lm_reaches_the_huber_optimum(oracle, seed):
    pb = sm.make_problem(seed)
    p = oracle.default_params()
    out, it, term = oracle.optimize_pose(p, 0, pb["guess"], pb["tree_feat"], pb["tree_obj"],
                                         pb["plane_feat"], pb["plane_obj"])
    assert term[0] == 0 and 1 <= it[0] <= 50
    x = np.concatenate([out["t"], R.from_quat(out["q"]).as_rotvec()])
    x0 = np.concatenate([pb["guess"]["t"][0], R.from_quat(pb["guess"]["q"][0]).as_rotvec()])
    sol = least_squares(sm.residuals, x0, args=(pb,), loss="huber", f_scale=0.1, xtol=1e-14,
                        ftol=1e-14, gtol=1e-14)
    c_or, c_sp = sm.huber_cost(sm.residuals(x, pb)), sm.huber_cost(sm.residuals(sol.x, pb))
    assert c_or < sm.huber_cost(sm.residuals(x0, pb))
    # Ceres stops at |dcost| <= 1e-6 cost: within 1e-5 relative of the true optimum
    assert c_or <= c_sp * (1 + 1e-5)
    assert np.allclose(x, sol.x, atol=2e-4)
    assert np.linalg.norm(out["t"] - pb["true_t"]) < 0.02


def test_two_step_freezes_the_right_coordinates(oracle):
    pb = sm.make_problem(3)
    p = oracle.default_params()
    g = pb["guess"]
    # only the tree problem runs: z and roll/pitch keep the guess values
    out, it, term = oracle.optimize_pose(p, 1, g, pb["tree_feat"], pb["tree_obj"], pb["plane_feat"],
                                         pb["plane_obj"], optim_trees=True, optim_ground=False)
    rv, rv_g = R.from_quat(out["q"]).as_rotvec(), R.from_quat(g["q"][0]).as_rotvec()
    assert term[0] == 0 and term[1] == -1
    assert out["t"][2] == g["t"][0][2]
    assert np.allclose(rv[:2], rv_g[:2], atol=1e-12)
    assert abs(out["t"][0] - pb["true_t"][0]) < abs(g["t"][0][0] - pb["true_t"][0])
    # neither runs: pose guess is returned (through angle-axis and back)
    out2, _, term2 = oracle.optimize_pose(p, 1, g, pb["tree_feat"], pb["tree_obj"], pb["plane_feat"],
                                          pb["plane_obj"], optim_trees=False, optim_ground=False)
    assert np.allclose(out2["t"], g["t"][0]) and np.allclose(out2["q"], g["q"][0], atol=1e-12)
    assert list(term2) == [-1, -1]


def test_no_residuals_is_convergence_with_unchanged_pose(oracle):
    pb = sm.make_problem(4)
    p = oracle.default_params()
    e3, ec, ep = np.zeros((0, 3)), pb["tree_obj"][:0], pb["plane_obj"][:0]
    out, it, term = oracle.optimize_pose(p, 0, pb["guess"], e3, ec, e3, ep)
    assert term[0] == 0 and it[0] == 0
    assert np.allclose(out["t"], pb["guess"]["t"][0])

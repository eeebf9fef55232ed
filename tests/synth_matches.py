"""Synthetic matched residual sets for the pose optimiser tests (host-side helper)."""
import numpy as np
from scipy.spatial.transform import Rotation as R

from sloam_b200 import abi


def make_problem(seed=0, ncyl=12, npl=10, feats_per_tree=20, feats_per_plane=5, outliers=True,
                 rotvec=(0.02, -0.03, 0.4), trans=(1.0, -2.0, 3.4)):
    rng = np.random.default_rng(seed)
    cyl = np.zeros(ncyl, abi.CYLINDER)
    cyl["root"][:, 0] = rng.uniform(-10, 10, ncyl)
    cyl["root"][:, 1] = rng.uniform(-10, 10, ncyl)
    ray = np.stack([rng.normal(0, 0.05, ncyl), rng.normal(0, 0.05, ncyl), np.ones(ncyl)], 1)
    cyl["ray"] = ray / np.linalg.norm(ray, axis=1)[:, None]
    cyl["radius"] = rng.uniform(0.1, 0.28, ncyl)
    pl = np.zeros(npl, abi.PLANE)
    n = np.stack([rng.normal(0, 0.03, npl), rng.normal(0, 0.03, npl), np.ones(npl)], 1)
    pl["plane"][:, :3] = n / np.linalg.norm(n, axis=1)[:, None]
    pl["plane"][:, 3] = rng.normal(0, 0.05, npl)
    Rt, tt = R.from_rotvec(rotvec), np.asarray(trans, float)
    tf, to, pf, po = [], [], [], []
    for c in cyl:
        a = c["ray"]
        u = np.cross(a, [1, 0, 0]); u /= np.linalg.norm(u)
        v = np.cross(a, u)
        for _ in range(feats_per_tree):
            h, ang = rng.uniform(0.5, 8), rng.uniform(0, 2 * np.pi)
            pw = c["root"] + h * a + c["radius"] * (np.cos(ang) * u + np.sin(ang) * v) + rng.normal(0, 0.01, 3)
            tf.append(Rt.inv().apply(pw - tt)); to.append(c)
    for q in pl:
        for _ in range(feats_per_plane):
            xy = rng.uniform(-15, 15, 2)
            z = -(q["plane"][0] * xy[0] + q["plane"][1] * xy[1] + q["plane"][3]) / q["plane"][2]
            pw = np.array([xy[0], xy[1], z + rng.normal(0, 0.02)])
            pf.append(Rt.inv().apply(pw - tt)); po.append(q)
    tf, pf = np.array(tf), np.array(pf)
    if outliers:
        tf[::17] += rng.normal(0, 0.5, tf[::17].shape)
    guess = np.zeros(1, abi.POSE)
    Rg = R.from_rotvec(np.asarray(rotvec) + [0.01, -0.008, 0.012])
    guess["t"][0] = tt + [0.05, -0.04, 0.03]
    guess["q"][0] = Rg.as_quat()
    return dict(tree_feat=tf, tree_obj=np.array(to, abi.CYLINDER), plane_feat=pf,
                plane_obj=np.array(po, abi.PLANE), guess=guess, true_t=tt, true_rotvec=np.asarray(rotvec))


def residuals(x, pb):
    """x = [t(3), rotvec(3)] -> raw residuals (cylinder.h:118-124, plane.h:99-110)."""
    Rm, t = R.from_rotvec(x[3:6]), x[:3]
    to, po = pb["tree_obj"], pb["plane_obj"]
    lp = Rm.apply(pb["tree_feat"]) + t
    d = lp - to["root"]
    proj = to["root"] + (np.sum(d * to["ray"], 1) / np.sum(to["ray"] ** 2, 1))[:, None] * to["ray"]
    rc = np.linalg.norm(proj - lp, axis=1) - to["radius"]
    lq = Rm.apply(pb["plane_feat"]) + t
    rp = np.abs(np.sum(po["plane"][:, :3] * lq, 1) + po["plane"][:, 3]) / np.linalg.norm(po["plane"][:, :3], axis=1)
    return np.concatenate([rc, rp])


def huber_cost(r, a=0.1):
    s = r * r
    return 0.5 * np.sum(np.where(s > a * a, 2 * a * np.sqrt(s) - a * a, s))


def make_degenerate(kind, seed=0):
    """Ill-conditioned variants of make_problem: geometry where the normal
    equations lose rank or conditioning, so that a solver difference (LDL^T on J^T J vs the
    oracle's QR on the augmented Jacobian) would show up as a different accept/reject sequence.
      collinear   all trunks on one line with parallel axes: the translation along the line is
                  constrained only through the trunk radii (weak curvature), yaw about the line poorly
      one_plane   a single ground plane patch: z / roll / pitch from one small patch
      two_planes  two nearly parallel planes
      zero_pad    trees whose feature lists were zero-padded by features.resize() (SURVEY B-6):
                  half of the tree rows are the point (0, 0, 0)
      one_tree    a single trunk: yaw and the translation along its axis unobservable from the trees
      tiny        3 tree rows + 3 plane rows: fewer residuals than the 6 tangent dimensions spanned"""
    rng = np.random.default_rng(1000 + seed)
    if kind == "collinear":
        pb = make_problem(seed, ncyl=8, npl=10, outliers=False)
        x = np.linspace(-9, 9, 8)
        roots = np.stack([x, np.full(8, 2.0), np.zeros(8)], 1)
        return _rebuild(pb, roots, np.tile([0.0, 0.0, 1.0], (8, 1)), seed)
    if kind == "one_plane":
        return make_problem(seed, ncyl=8, npl=1, feats_per_plane=5, outliers=False)
    if kind == "two_planes":
        return make_problem(seed, ncyl=8, npl=2, feats_per_plane=5, outliers=False)
    if kind == "zero_pad":
        pb = make_problem(seed, ncyl=8, npl=8, outliers=False)
        tf = pb["tree_feat"].copy()
        tf.reshape(8, 20, 3)[:, 10:, :] = 0.0        # resize(20) of a tree with 10 points
        pb["tree_feat"] = tf
        return pb
    if kind == "one_tree":
        return make_problem(seed, ncyl=1, npl=10, outliers=False)
    if kind == "tiny":
        pb = make_problem(seed, ncyl=3, npl=3, feats_per_tree=1, feats_per_plane=1, outliers=False)
        return pb
    raise ValueError(kind)


def _rebuild(pb, roots, rays, seed):
    """Re-sample the tree rows of pb on trunks with the given roots / axes (same pose, same planes)."""
    rng = np.random.default_rng(2000 + seed)
    Rt, tt = R.from_rotvec(pb["true_rotvec"]), pb["true_t"]
    n = len(roots)
    cyl = np.zeros(n, abi.CYLINDER)
    cyl["root"], cyl["ray"], cyl["radius"] = roots, rays, rng.uniform(0.1, 0.28, n)
    tf, to = [], []
    for c in cyl:
        a = c["ray"]
        u = np.cross(a, [1, 0, 0]); u /= np.linalg.norm(u)
        v = np.cross(a, u)
        for _ in range(20):
            h, ang = rng.uniform(0.5, 8), rng.uniform(0, 2 * np.pi)
            pw = c["root"] + h * a + c["radius"] * (np.cos(ang) * u + np.sin(ang) * v) + rng.normal(0, 0.01, 3)
            tf.append(Rt.inv().apply(pw - tt)); to.append(c)
    out = dict(pb)
    out["tree_feat"], out["tree_obj"] = np.array(tf), np.array(to, abi.CYLINDER)
    return out


DEGENERATE = ["collinear", "one_plane", "two_planes", "zero_pad", "one_tree", "tiny"]

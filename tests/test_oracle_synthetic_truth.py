"""The oracle against the GROUND TRUTH of the synthetic scenes (CPU only).

The third-party arithmetic the oracle restates from published algorithms (PCL line RANSAC,
Eigen JacobiSVD, "parity unpinned" in oracle/orc.h) has no reference fixture; this pins it at
least physically: on a generated forest with known trunk axes and a known ground plane the
fitted models must be the scene's.  The trunk models are lines through the VISIBLE side of the
trunk (cylinder.cpp fits the vertex medians), so the root is up to one radius off the axis.
"""
import numpy as np
import pytest

from sloam_b200 import abi, capi, configs


@pytest.mark.parametrize("preset,min_landmarks", [("os1-64", 8), ("vlp-16", 20)])
def test_models_of_the_first_scan_are_the_scene(oracle, preset, min_landmarks):
    p, cfg = configs.make(capi, preset)
    scene = capi.synth_scene(cfg)
    pts, mask = capi.synth_generate_host(cfg, 0, 1)
    gt, _ = capi.synth_pose(cfg, 0)
    e = oracle.run_keyframe(p, pts[0], mask[0], np.array([gt]), True, np.zeros(0, abi.CYLINDER),
                            np.zeros(0, abi.PLANE))
    assert e.result["status"] == abi.KF_OK
    n = int(e.result["n_landmarks"])
    assert n >= min_landmarks
    used = set()
    for t in e.tm[:n]:                                   # first scan: tm is in the map (= scene) frame
        tr = t["ray"] / np.linalg.norm(t["ray"])
        best = None
        for i, s in enumerate(scene):
            ray = s["ray"] / np.linalg.norm(s["ray"])
            at = s["root"] + (t["root"][2] - s["root"][2]) / ray[2] * ray      # true axis at the root's height
            d = np.linalg.norm(at[:2] - t["root"][:2])
            if best is None or d < best[0]:
                best = (d, np.degrees(np.arccos(min(1.0, abs(float(tr @ ray))))), abs(t["radius"] - s["radius"]), i,
                        s["radius"])
        d, ang, dr, i, r_true = best
        assert d < r_true + 0.12, (d, r_true)            # on the trunk's surface
        assert ang < 2.0                                 # axis direction [deg]
        assert dr < 0.12                                 # radius [m]
        used.add(i)
    assert len(used) >= n - 2                            # (a partly occluded trunk can split in two clusters)
    # accepted ground planes: normal along the scene's up axis, sensor height above the plane
    planes = e.planes[:e.n_planes]
    assert e.n_planes >= 10
    assert np.abs(planes["plane"][:, 2]).min() > 0.99
    cen = planes["centroid"]
    assert np.all(cen[:, 2] < gt["t"][2])                # heightCheck (sloam.cpp:404)


def test_plane_fit_is_the_total_least_squares_plane(oracle):
    """Plane::computeModel (plane.cpp:96-128) through the restated Eigen JacobiSVD against
    numpy's LAPACK SVD: same normal up to sign, plane through the float32 centroid."""
    rng = np.random.default_rng(3)
    for trial in range(50):
        n = int(rng.integers(5, 400))
        nrm = rng.normal(size=3)
        nrm /= np.linalg.norm(nrm)
        basis = np.linalg.svd(nrm[None, :])[2][1:]                       # two in-plane directions
        xyz = (rng.uniform(-10, 10, (n, 2)) @ basis + rng.normal(0, 0.03, (n, 1)) * nrm + rng.normal(0, 5, 3)).astype(np.float32)
        pts = np.zeros(n, abi.POINT)
        pts["x"], pts["y"], pts["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        fit = oracle.plane_fit(pts, 5)
        assert fit["is_valid"]
        got = np.array(fit["model"]["plane"][:3])
        cen = xyz.astype(np.float64).mean(0)
        want = np.linalg.svd((xyz.astype(np.float64) - cen).T)[0][:, 2]
        assert abs(abs(float(got @ want)) - 1.0) < 1e-9, trial
        assert abs(np.linalg.norm(got) - 1.0) < 1e-12
        assert np.allclose(fit["model"]["centroid"], cen, atol=1e-4)
        # plane[3] = -n . centroid (plane.cpp:124-126)
        assert abs(fit["model"]["plane"][3] + float(got @ np.array(fit["model"]["centroid"]))) < 1e-9

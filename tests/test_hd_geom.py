"""Host build of the kernels' bit-exact per-point arithmetic (sloam_b200/csrc/proj_math.h,
compiled by tests/hd_geom_check.cpp) against the oracle: pixel indices and polar ground cells
must be identical on random and on adversarial (boundary) inputs.  Runs without a GPU; the
kernels themselves (fp32 estimate + this exact fallback) are covered by tests/test_gpu_parity.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from sloam_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hg(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hg") / "hd_geom_check.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "hd_geom_check.cpp"), "-o", so])
    return C.CDLL(so)


def _points(xyz):
    pts = np.zeros(len(xyz), abi.POINT)
    pts["x"], pts["y"], pts["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    return pts


def _cloud(rng, n, H, W, fov_up, fov_down):
    """Random returns plus the cases the projection is sensitive to: pixel-boundary azimuths and
    elevations, the +-x axis, zeros, NaN no-returns, beams outside the field of view."""
    r = rng.uniform(0.5, 40.0, n)
    yaw = rng.uniform(-np.pi, np.pi, n)
    pitch = np.radians(rng.uniform(fov_down - 3.0, fov_up + 3.0, n))
    k = n // 4
    yaw[:k] = -np.pi + 2.0 * np.pi * rng.integers(0, W + 1, k) / W          # column boundaries
    fov = np.radians(fov_up - fov_down)
    pitch[k:2 * k] = np.radians(fov_up) - fov * rng.integers(0, H + 1, k) / H   # row boundaries
    xyz = np.stack([r * np.cos(pitch) * np.cos(yaw), r * np.cos(pitch) * np.sin(yaw), r * np.sin(pitch)], 1)
    special = np.array([[0, 0, 0], [-0.0, 0.0, 1.0], [1, 0, 0], [-1, 0, 0], [-1, -0.0, 0], [0, 1, 0], [0, -1, 0],
                        [np.nan, 1, 1], [1, np.nan, 1], [1, 1, np.nan], [np.nan] * 3, [np.inf, 1, 1],
                        [1e-30, 1e-30, 1e-30], [1e18, -1e18, 1e18], [0, 0, 5], [0, 0, -5]], np.float64)
    return _points(np.concatenate([xyz, special]).astype(np.float32))


@pytest.mark.parametrize("H,W,fov_up,fov_down", [(64, 1024, 22.5, -22.5), (16, 1800, 15.0, -15.0),
                                                 (128, 2048, 22.5, -22.5), (32, 1000, 10.0, -30.0)])
def test_pixel_indices_match_the_oracle(hg, oracle, H, W, fov_up, fov_down):
    rng = np.random.default_rng(H * W)
    p = oracle.default_params(img_h=H, img_w=W, fov_up_deg=fov_up, fov_down_deg=fov_down)
    pts = _cloud(rng, 200_000, H, W, fov_up, fov_down)
    pix = np.zeros(len(pts), np.int32)
    rng_out = np.zeros(len(pts), np.float32)
    hg.hd_project(C.byref(p), abi.ptr(pts), len(pts), abi.ptr(pix), abi.ptr(rng_out))
    want, _ = oracle.project(p, pts, want_range=False)
    assert np.array_equal(pix, want)                      # bit-exact (north_star: indices)
    assert pix.min() >= 0 and pix.max() < H * W


@pytest.mark.parametrize("RB,TB,dmin,dmax", [(2, 18, 5.0, 25.0), (1, 18, 0.0, 30.0), (3, 7, 2.5, 40.0), (5, 36, 1.0, 12.0)])
def test_ground_cells_match_the_oracle(hg, oracle, RB, TB, dmin, dmax):
    rng = np.random.default_rng(RB * 100 + TB)
    p = oracle.default_params(groundRadiiBins=RB, groundThetaBins=TB, minGroundLidarDist=dmin,
                              maxGroundLidarDist=dmax, max_prev_planes=max(64, RB * TB))
    n = 200_000
    r = rng.uniform(0.0, dmax * 1.2, n)
    th = rng.uniform(-np.pi, np.pi, n)
    k = n // 4
    th[:k] = -3.14159265 + (2 * 3.14159265 / TB) * rng.integers(0, TB + 1, k)     # theta-bin edges
    r[k:2 * k] = (dmax / RB) * rng.integers(0, RB + 1, k)                           # radial-bin edges
    r[2 * k:2 * k + 50] = dmin
    xyz = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(-1.5, 0.1, n)], 1).astype(np.float32)
    pts = _points(xyz)
    cell = np.zeros(n, np.int32)
    hg.hd_ground_cells(C.byref(p), abi.ptr(pts), n, abi.ptr(cell))
    assert cell.max() < RB * TB
    total = np.zeros(RB * TB, np.int64)
    for lo in range(0, n, 2000):                          # small chunks: a swap between two cells cannot hide
        c = cell[lo:lo + 2000]
        hist = np.bincount(c[c >= 0], minlength=RB * TB)
        cells, _, _, _ = oracle.ground_planes(p, pts[lo:lo + 2000], oracle.identity_pose())
        assert np.array_equal(hist, cells["n_cell"]), lo  # bit-exact (north_star: indices)
        total += hist
    assert total.sum() < n and total.min() > 0


@pytest.mark.parametrize("RB,TB,dmin,dmax", [(2, 18, 5.0, 25.0), (1, 18, 0.0, 30.0), (3, 7, 2.5, 40.0), (4, 36, 0.7, 12.3),
                                              (4, 9, 1.0, 1e-3 + 33.333), (5, 18, 1.0, 20.0)])
def test_radius_thresholds_equal_the_exact_radial_decisions(hg, oracle, RB, TB, dmin, dmax):
    """project_split_kernel decides the radius tests and the radial bin by comparing
    r2 = x*x + y*y with precomputed float thresholds (proj_math.h: ground_geom_thresholds).
    They must give the decisions of the exact ground_cell_of() for every point -- here on
    random radii and on the floats around every threshold radius (min, max, bin edges)."""
    rng = np.random.default_rng(RB * 1000 + TB)
    p = oracle.default_params(groundRadiiBins=RB, groundThetaBins=TB, minGroundLidarDist=dmin,
                              maxGroundLidarDist=dmax, max_prev_planes=max(64, RB * TB))
    edges = np.array([dmin, dmax] + [dmax / RB * j for j in range(1, RB)], np.float64)
    near = []
    for e in edges:                                   # +-64 floats around each edge radius
        f = np.float32(e)
        lo = f
        for _ in range(64):
            lo = np.nextafter(lo, np.float32(-np.inf))
        xs = [lo]
        for _ in range(128):
            xs.append(np.nextafter(xs[-1], np.float32(np.inf)))
        near.append(np.array(xs, np.float32))
    near = np.concatenate(near)
    r = np.concatenate([near, rng.uniform(0.0, dmax * 1.2, 100_000).astype(np.float32)])
    # three families with different rounding of x*x + y*y: on an axis, on the diagonal, random
    th = np.concatenate([np.zeros(near.size), rng.uniform(-np.pi, np.pi, 100_000)])
    xyz = [np.stack([r * np.cos(th), r * np.sin(th), np.zeros(r.size)], 1),
           np.stack([near / np.sqrt(2.0), near / np.sqrt(2.0), np.zeros(near.size)], 1),
           np.stack([np.zeros(near.size), -near, np.zeros(near.size)], 1)]
    pts = _points(np.concatenate(xyz).astype(np.float32))
    n = pts.shape[0]
    cell = np.zeros(n, np.int32)
    rb = np.zeros(n, np.int32)
    hg.hd_ground_cells(C.byref(p), abi.ptr(pts), n, abi.ptr(cell))
    hg.hd_ground_radial_by_threshold(C.byref(p), abi.ptr(pts), n, abi.ptr(rb))
    want = np.where(cell < 0, -1, cell // TB)
    if RB > 4:                                        # no thresholds: the kernel uses estimate + exact fallback
        assert (rb == -2).all()
        return
    assert np.array_equal(rb, want)
    assert (want == -1).any() and all((want == j).any() for j in range(RB))


def test_plane_acceptance_matches_the_oracle(hg, oracle):
    """The acceptance test (FromTwoVectors -> eulerAngles(0,1,2) -> tolerance, and the height
    check) decides which cells become planes; tilted grounds around the 0.1 rad tolerance and
    poses with large yaw exercise both outcomes and both branches of the Euler extraction."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(5)
    p = oracle.default_params(minGroundLidarDist=1.0, numGroundFeatures=5)
    n_valid = n_acc = 0
    for trial in range(40):
        tilt = np.radians(rng.uniform(0.0, 11.0))
        az = rng.uniform(-np.pi, np.pi)
        nrm = np.array([np.sin(tilt) * np.cos(az), np.sin(tilt) * np.sin(az), np.cos(tilt)])
        h = rng.uniform(0.8, 2.5) * (1 if trial % 7 else -1)      # every 7th ground is ABOVE the sensor
        r = rng.uniform(1.5, 24.0, 6000)
        th = rng.uniform(-np.pi, np.pi, 6000)
        x, y = r * np.cos(th), r * np.sin(th)
        z = (-h - nrm[0] * x - nrm[1] * y) / nrm[2] + rng.normal(0, 0.02, 6000)
        pts = _points(np.stack([x, y, z], 1).astype(np.float32))
        pose = oracle.identity_pose()
        pose["q"][0] = R.from_euler("zyx", [rng.uniform(-np.pi, np.pi), rng.normal(0, 0.03), rng.normal(0, 0.03)]).as_quat()
        pose["t"][0] = rng.normal(0, 5.0, 3)
        cells, _, _, _ = oracle.ground_planes(p, pts, pose)
        for c in cells:
            if not c["is_valid"]:
                continue
            plane = np.ascontiguousarray(c["model"]["plane"], np.float64)
            cen = np.ascontiguousarray(c["model"]["centroid"], np.float64)
            got = hg.hd_plane_accept(abi.ptr(pose), abi.ptr(plane), abi.ptr(cen), C.c_double(p.ground_angle_tol))
            assert got == int(c["accepted"]), (trial, plane, cen)
            n_valid += 1
            n_acc += got
    assert n_valid > 1000 and 0.2 * n_valid < n_acc < 0.8 * n_valid   # both outcomes are well represented

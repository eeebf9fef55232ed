"""The C++ mirror of the reference API (sloam_b200/host/sloam_host.h) running the
reference's own gtest cases on the GPU through the C ABI (tests/host_api_test.cpp)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import golden_io
from sloam_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "sloam_b200", "lib")


def write_fixture(f, stamp):
    g = golden_io.ground("still", stamp)
    trees, verts, pts = golden_io.Trellis("still", stamp).as_flat()
    f.write(struct.pack("<i", len(g)))
    f.write(g.tobytes())
    f.write(struct.pack("<i", len(trees)))
    for t in trees:
        f.write(struct.pack("<i", t["n_vertices"]))
        for v in verts[t["vertex_begin"]:t["vertex_begin"] + t["n_vertices"]]:
            f.write(struct.pack("<iffffi", int(t["tree_id"]), float(v["radius"]), float(v["cx"]), float(v["cy"]),
                                float(v["cz"]), int(v["n_points"])))
            f.write(pts[v["point_begin"]:v["point_begin"] + v["n_points"]].tobytes())


def build(tmp):
    exe = os.path.join(tmp, "host_api_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "host_api_test.cpp"),
                           f"-L{LIBDIR}", "-lsloam_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe])
    return exe


def test_host_mirror_compiles_against_the_c_abi(tmp_path):
    """CPU-only: the header-only mirror builds with plain g++ and links the library."""
    assert os.path.exists(build(str(tmp_path)))


@pytest.mark.gpu
def test_reference_gtests_through_the_cpp_mirror(tmp_path, oracle):
    exe = build(str(tmp_path))
    blob = os.path.join(str(tmp_path), "fixtures.bin")
    with open(blob, "wb") as f:
        write_fixture(f, "t0")
        write_fixture(f, "t1")
    K = 6
    r = subprocess.run([exe, blob, str(K)], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout
    # SLOAMNodeCore::run (8(f)-2) over a synthetic sequence against the oracle's
    # getSubmap -> RunSloam -> updateMap loop: status, landmark count and map size exact,
    # pose within north_star's 1e-5 m / 1e-5 rad
    seq = [l.split()[1:] for l in r.stdout.splitlines() if l.startswith("SEQ ")]
    assert len(seq) == K
    from sloam_b200 import capi, configs
    p, cfg = configs.make(capi, "os1-64")
    pts, mask = capi.synth_generate_host(cfg, 0, K)
    omap = oracle.OracleMap()
    first, prev = True, np.zeros(0, abi.PLANE)
    for k in range(K):
        pose = np.array([capi.synth_pose(cfg, k)[1]])
        sub, _ = omap.get_submap(pose, p.max_map_models)
        e = oracle.run_keyframe(p, pts[k], mask[k], pose, first, sub, prev)
        ran = e.result["status"] in (abi.KF_OK, abi.KF_NOT_CONVERGED)
        n = int(e.result["n_landmarks"]) if ran else 0
        omap.update(e.tm[:n], e.tm_id[:n], e.matches[:n])
        prev, first = e.planes[:e.n_planes].copy(), False
        g = seq[k]
        assert [int(v) for v in g[:5]] == [k, int(e.result["status"]), int(e.result["success"]),
                                           int(e.result["n_landmarks"]), omap.size()], (k, g)
        if e.result["success"]:
            t, q = np.array(g[5:8], float), np.array(g[8:12], float)
            assert np.max(np.abs(t - e.result["T_Map_Curr"]["t"])) <= 1e-5
            qe = e.result["T_Map_Curr"]["q"]     # small-angle form: arccos(dot) loses half the digits near 1
            assert 2.0 * min(np.linalg.norm(q - qe), np.linalg.norm(q + qe)) <= 1e-5

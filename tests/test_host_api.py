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
def test_reference_gtests_through_the_cpp_mirror(tmp_path):
    exe = build(str(tmp_path))
    blob = os.path.join(str(tmp_path), "fixtures.bin")
    with open(blob, "wb") as f:
        write_fixture(f, "t0")
        write_fixture(f, "t1")
    r = subprocess.run([exe, blob], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout

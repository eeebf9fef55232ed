"""PCD v0.7 (ascii / binary), the Boost text archive v17 of the landmark
fixtures and the ROS1 wire encoding of ROSCylinder (sloam_b200/host/formats.h).

tests/formats_test.cpp does the checks; with the reference tree present (this container, not
the GPU box) it also reads the reference's eight .pcd and four landmark fixtures and requires
that writing them back reproduces the files byte for byte."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AUX = "/root/reference/sloam/src/tests/aux"


def test_formats_round_trips_and_reference_fixtures(tmp_path):
    exe = str(tmp_path / "formats_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "formats_test.cpp"), "-o", exe])
    args = [exe] + ([AUX] if os.path.isdir(AUX) else [])
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK")
    if os.path.isdir(AUX):
        assert "8 pcd + 4 archives" in out.stdout

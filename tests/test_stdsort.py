"""sloam_b200/csrc/dev_stdsort.h replays libstdc++'s std::sort (introsort + final insertion
sort) so that exact ties come out in the order the reference's unstable sorts leave them
(SURVEY B-3).  tests/stdsort_test.cpp compares it with std::sort on ~60k arrays full of ties,
including the reference's three consecutive sorts by x, y, z."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_std_sort_replay_matches_libstdcxx(tmp_path):
    exe = str(tmp_path / "stdsort_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "stdsort_test.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr

"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/sloam_b200.h declares, agrees with the Python/oracle record layouts, and refuses to
run without a GPU (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import torch

from sloam_b200 import abi, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sloam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sloam_(?:b200|synth)_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(capi.EXPORTS) <= set(names)
    assert lib.sloam_b200_version().startswith(b"sloam_b200")


def test_record_layouts_agree(oracle):
    sizes = (C.c_int32 * 8)()
    n = oracle.lib().orc_abi_sizes(sizes)
    got = list(sizes)[:n]
    assert got == [C.sizeof(abi.Params), abi.KF_RESULT.itemsize, abi.CELL_PLANE.itemsize,
                   abi.TREE_MODEL.itemsize, abi.VERTEX.itemsize, abi.TREE.itemsize]
    # the library and the oracle fill the same defaults (sloam/params/sloam.yaml over sloamNode.cpp:57-128)
    a, b = capi.default_params(), oracle.default_params()
    assert bytes(a) == bytes(b)
    assert (a.groundRadiiBins, a.groundThetaBins, a.featuresPerTree, a.numGroundFeatures) == (2, 18, 20, 5)
    assert a.twoStepOptim == 1 and a.treeMatchThresh == 0.5 and a.AddNewTreeThreshDist == 1.5


def test_no_cpu_fallback():
    """Without a usable sm_100 device create() must fail with SLOAM_E_NODEVICE; bad
    parameters are rejected before any device work."""
    p = capi.default_params()
    h = C.c_void_p()
    rc = capi.lib().sloam_b200_create(C.byref(p), 0, 1, C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        capi.lib().sloam_b200_destroy(h)
    else:
        assert rc == -4 and not h.value
    assert capi.lib().sloam_b200_create(None, 0, 1, C.byref(h)) == -1
    # nothing in the product package refers to the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sloam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liborc" not in text and "oracle/" not in text.replace("oracle/)", ""), f


def test_synthetic_generator_is_seeded_and_labelled():
    cfg = capi.synth_config(16, 256, 6)
    a = capi.synth_generate_host(cfg, 3, 2)
    b = capi.synth_generate_host(cfg, 3, 2)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()
    assert set(np.unique(a[1])) <= {0, 1, 255}
    gt, guess = capi.synth_pose(cfg, 3)
    assert abs(np.linalg.norm(gt["q"]) - 1) < 1e-12 and abs(np.linalg.norm(guess["q"]) - 1) < 1e-9
    assert 0 < np.linalg.norm(np.array(gt["t"]) - np.array(guess["t"])) < 0.5

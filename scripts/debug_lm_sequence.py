"""Debug helper (GPU box): joint-mode LM on the real match lists of a short sequence,
GPU stage entry vs oracle."""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from sloam_b200 import abi, capi, configs

p, cfg = configs.make(capi, "os1-64", twoStepOptim=0)
K, M = 6, p.max_map_models
pts, mask = capi.synth_generate_host(cfg, 0, K)
ctx = capi.Context(p, 1)
omap = orc.OracleMap(); first = True; prev = np.zeros(0, abi.PLANE)
for k in range(K):
    pose = np.array([capi.synth_pose(cfg, k)[1]])
    sub, _ = omap.get_submap(pose, M)
    e = orc.run_keyframe(p, pts[k], mask[k], pose, first, sub, prev)
    r = e.result
    if not first:
        tf = np.zeros((4096, 3)); to = np.zeros(4096, abi.CYLINDER); pf = np.zeros((512, 3)); po = np.zeros(512, abi.PLANE)
        npl = C.c_int32()
        nt = orc.lib().orc_last_matches(abi.ptr(tf), abi.ptr(to), 4096, abi.ptr(pf), abi.ptr(po), 512, C.byref(npl))
        d = [capi.to_dev(a) for a in (pose, tf, to, np.array([nt], np.int32), pf, po, np.array([npl.value], np.int32),
                                      np.ones(1, np.uint8), np.ones(1, np.uint8))]
        out, it, term = ctx.optimize_pose(0, d[0], d[1], d[2], d[3], 4096, d[4], d[5], d[6], 512, d[7], d[8], 1)
        ctx.sync()
        out = capi.to_host(out, abi.POSE, (1,)); it = capi.to_host(it, np.int32, (1, 2)); term = capi.to_host(term, np.int32, (1, 2))
        print(k, "oracle it", r["lm_iterations"], "term", r["lm_termination"], "t", r["T_Map_Curr"]["t"],
              "| gpu it", it[0], "term", term[0], "t", out[0]["t"], "nt", nt, "np", npl.value)
    n = int(r["n_landmarks"]) if r["status"] in (0, 3) else 0
    omap.update(e.tm[:n], e.tm_id[:n], e.matches[:n]); prev = e.planes[:e.n_planes].copy(); first = False

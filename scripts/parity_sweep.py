"""Randomised parity sweep (GPU box): fused path vs the CPU oracle on many keyframes of several
presets / seeds, full comparison of every result field (tests/test_gpu_parity.compare_keyframe).
usage: [SLOAM_SWEEP_SEED_OFFSET=n] python scripts/parity_sweep.py [keyframes_per_case]"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from sloam_b200 import abi, capi, configs
import test_gpu_parity as tg

K = int(sys.argv[1]) if len(sys.argv) > 1 else 60
only = sys.argv[2] if len(sys.argv) > 2 else None
cases = [("vlp-16", True, 11), ("vlp-16", True, 12), ("vlp-16", False, 13), ("os1-64", True, 21), ("os1-64", False, 22),
         ("os1-128", True, 31), ("os1-64-dense", True, 41), ("vlp-16", True, 51), ("os1-64", True, 52), ("vlp-16", True, 53),
         ("os1-64", False, 54)]
# seeds >= 50 also vary the sensor model and the parameters
VARIANTS = {
    51: (dict(nan_no_return=0, range_noise=0.03), dict(groundRetainThresh=0.2, featuresPerTree=12)),
    52: (dict(nan_no_return=0, ground_noise=0.05, max_tilt_deg=4.0), dict(numGroundFeatures=8, groundRetainThresh=0.05)),
    53: (dict(step_per_keyframe=0.6, guess_sigma_t=0.15), dict(ransac_fixed_hypotheses=256)),
    54: (dict(azimuth_offset_cols=17.25, ground_slope_deg=2.5), dict(huber_delta=0.05)),
}
if only:
    cases = [c for c in cases if c[0] == only]
total = bad = 0
for preset, two_step, seed in cases:
    kk = K if preset not in ("os1-128", "os1-64-dense") else max(4, K // 8)
    cfg_kw, p_kw = VARIANTS.get(seed, ({}, {}))
    p, cfg = configs.make(capi, preset, twoStepOptim=int(two_step), **p_kw)
    for name, val in cfg_kw.items():
        setattr(cfg, name, val)
    if preset == "vlp-16" and not two_step:
        p.minGroundModels = 10
    cfg.seed = seed + 1000 * int(os.environ.get("SLOAM_SWEEP_SEED_OFFSET", "0"))  # new scenes, same case list
    cfg.n_trees = cfg.n_trees + seed % 7
    inp, exp = tg.run_sequence(capi, orc, p, cfg, kk, two_step)
    T, PP = p.max_trees, p.max_prev_planes
    ctx = capi.Context(p, kk)
    out = dict(results=np.zeros(kk, abi.KF_RESULT), matches=np.zeros((kk, T), np.int32),
               tm=np.zeros((kk, T), abi.CYLINDER), tm_id=np.zeros((kk, T), np.int32),
               planes=np.zeros((kk, PP), abi.PLANE), n_planes=np.zeros(kk, np.int32), range_image=None)
    ctx.run_keyframes_host(kk, inp, out)
    nb = 0
    for k in range(kk):
        total += 1
        try:
            tg.compare_keyframe(out["results"][k], out["matches"][k], out["tm"][k], out["tm_id"][k], out["planes"][k],
                                out["n_planes"][k], exp[k])
        except AssertionError:
            nb += 1
            if nb <= 3:
                tb = traceback.format_exc().strip().splitlines()
                r, er = out["results"][k], exp[k].result
                print(f"  MISMATCH {preset} two_step={two_step} seed={seed} kf={k}: {tb[-1][:160]}")
                print("    gpu it/term", r["lm_iterations"], r["lm_termination"], "oracle", er["lm_iterations"], er["lm_termination"],
                      "n_lm", r["n_landmarks"], er["n_landmarks"])
    bad += nb
    opt = sum(int(e.result["lm_termination"][0] == 0) for e in exp)
    print(f"{preset} two_step={two_step} seed={seed}: {kk} keyframes, {nb} mismatching, {opt} optimised", flush=True)
    ctx.close()
print(f"TOTAL {total} keyframes, {bad} mismatching")

"""Random image sizes / bin layouts / capacities through the fused path against the oracle (GPU
box).  usage: python scripts/config_fuzz.py [n_configs] [seed]"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from sloam_b200 import abi, capi
import test_gpu_parity as tg

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for it in range(n_cfg):
    H = int(rng.choice([8, 16, 21, 32, 40, 64, 100, 128]))
    W = int(rng.choice([257, 512, 600, 900, 1024, 1031, 1800, 2048]))
    fov = float(rng.choice([15.0, 16.6, 22.5]))
    rb, tb = int(rng.integers(1, 5)), int(rng.integers(4, 40))
    fg = int(rng.integers(3, 9))
    kw = dict(img_h=H, img_w=W, fov_up_deg=fov, fov_down_deg=-fov, groundRadiiBins=rb, groundThetaBins=tb,
              numGroundFeatures=fg, featuresPerTree=int(rng.integers(8, 25)), max_prev_planes=max(64, rb * tb),
              groundRetainThresh=float(rng.choice([0.05, 0.1, 0.25])), max_trees=int(rng.choice([64, 128, 300])),
              min_tree_vertices=int(min(8, max(3, H // 3))), min_cluster_points=int(rng.choice([20, 30, 80])),
              minTreeModels=3, minGroundModels=float(rng.choice([5, 10, 36])), twoStepOptim=int(rng.integers(0, 2)))
    try:
        p = capi.default_params(**kw)
    except Exception as e:
        print(it, "params rejected", kw, e); continue
    cfg = capi.synth_config(H, W, int(rng.integers(10, 60)), fov_up_deg=fov, fov_down_deg=-fov,
                            sensor_height=float(rng.choice([1.5, 3.4])), tree_r_max=float(rng.choice([9.0, 15.0])),
                            max_tilt_deg=float(rng.choice([1.5, 4.0])), nan_no_return=int(rng.integers(0, 2)))
    cfg.seed = int(rng.integers(1, 1 << 30))
    K = 3
    T, PP, N = p.max_trees, p.max_prev_planes, H * W
    tag = f"cfg {it}: {H}x{W} fov {fov} bins {rb}x{tb} Fg {fg} T {T} two_step {kw['twoStepOptim']}"
    try:
        ctx = capi.Context(p, K)
    except Exception as e:
        print(tag, "create rejected:", e); continue
    try:
        inp, exp = tg.run_sequence(capi, orc, p, cfg, K, bool(kw["twoStepOptim"]))
        out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32), tm=np.zeros((K, T), abi.CYLINDER),
                   tm_id=np.zeros((K, T), np.int32), planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32),
                   range_image=np.zeros((K, N), np.float32))
        ctx.run_keyframes_host(K, inp, out)
        for k in range(K):
            tg.compare_keyframe(out["results"][k], out["matches"][k], out["tm"][k], out["tm_id"][k], out["planes"][k],
                                out["n_planes"][k], exp[k])
            assert np.array_equal(out["range_image"][k].view(np.uint32), exp[k].range_image.view(np.uint32)), "range image"
        print(tag, "ok, landmarks", [int(e.result["n_landmarks"]) for e in exp], "planes", [int(e.n_planes) for e in exp], flush=True)
    except Exception:
        bad += 1
        print(tag, "MISMATCH/ERROR:", traceback.format_exc().strip().splitlines()[-1][:200], flush=True)
    ctx.close()
print("configs", n_cfg, "bad", bad)

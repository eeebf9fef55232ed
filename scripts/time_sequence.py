"""Latency of the sequential mode (sloam_b200_sequence_step_host, one keyframe per call):
usage: python scripts/time_sequence.py [preset] [keyframes]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from sloam_b200 import abi, capi, configs

preset = sys.argv[1] if len(sys.argv) > 1 else "os1-64"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
p, cfg = configs.make(capi, preset)
T = p.max_trees
pts, mask = capi.synth_generate_host(cfg, 0, K)
import torch
pts_t = torch.from_numpy(pts.view(np.uint8).reshape(K, -1)).pin_memory()
mask_t = torch.from_numpy(mask.reshape(K, -1)).pin_memory()
ctx = capi.Context(p, 1)
assert capi.lib().sloam_b200_map_init(ctx.h, 8192) == 0
res = np.zeros(1, abi.KF_RESULT)
matches = np.zeros(T, np.int32); tm = np.zeros(T, abi.CYLINDER); tm_id = np.zeros(T, np.int32)
lat = []
for k in range(K):
    pose = np.array([capi.synth_pose(cfg, k)[1]])
    t0 = time.perf_counter()
    rc = capi.lib().sloam_b200_sequence_step_host(ctx.h, C.c_void_p(pts_t[k].data_ptr()), C.c_void_p(mask_t[k].data_ptr()), abi.ptr(pose),
                                                  abi.ptr(res), abi.ptr(matches), abi.ptr(tm), abi.ptr(tm_id))
    lat.append(time.perf_counter() - t0)
    assert rc == 0
lat = np.array(lat[5:]) * 1e3
print(f"{preset}: {K} keyframes, sequence_step_host latency ms: median {np.median(lat):.3f} p90 {np.percentile(lat, 90):.3f} "
      f"max {lat.max():.3f}; landmarks last {int(res[0]['n_landmarks'])}, status {int(res[0]['status'])}")

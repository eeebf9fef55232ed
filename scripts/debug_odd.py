import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, orc
from sloam_b200 import abi, capi
import test_gpu_parity as tg
K, H, W = 4, 21, 1031
p = capi.default_params(img_h=H, img_w=W, fov_up_deg=15.0, fov_down_deg=-15.0, min_tree_vertices=8, min_cluster_points=30, minTreeModels=3)
cfg = capi.synth_config(H, W, 40, fov_up_deg=15.0, fov_down_deg=-15.0, sensor_height=1.5, tree_r_max=9.0, max_tilt_deg=1.5)
N, T, PP = H * W, p.max_trees, p.max_prev_planes
inp, exp = tg.run_sequence(capi, orc, p, cfg, K, True)
ctx = capi.Context(p, K)
out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32), tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32), planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32), range_image=None)
ctx.run_keyframes_host(K, inp, out)
it = ctx.intermediates()
V = p.max_tree_vertices
trees = capi.read_dev(it.trees, K * T * abi.TREE.itemsize, ctx.device).view(abi.TREE).reshape(K, T)
verts = capi.read_dev(it.vertices, K * T * V * abi.VERTEX.itemsize, ctx.device).view(abi.VERTEX).reshape(K, T * V)
models = capi.read_dev(it.tree_models, K * T * abi.TREE_MODEL.itemsize, ctx.device).view(abi.TREE_MODEL).reshape(K, T)
for k in range(K):
    e = exp[k]
    nt = e.n_trees
    for t in range(nt):
        gt, et = trees[k][t], e.trees[t]
        gv = verts[k][gt["vertex_begin"]:gt["vertex_begin"] + gt["n_vertices"]]
        ev = e.vertices[et["vertex_begin"]:et["vertex_begin"] + et["n_vertices"]]
        bad = gt["n_vertices"] != et["n_vertices"] or not np.array_equal(gv["n_points"], ev["n_points"]) or not np.array_equal(gv["radius"].view(np.uint32), ev["radius"].view(np.uint32))
        gm, em = models[k][t], e.tree_models[t]
        if bad or abs(gm["model"]["radius"] - em["model"]["radius"]) > 1e-6:
            print("kf", k, "tree", t, "nv", gt["n_vertices"], et["n_vertices"], "model radius", gm["model"]["radius"], em["model"]["radius"], "valid", gm["is_valid"], em["is_valid"])
            print("  gpu npts", gv["n_points"], "\n  orc npts", ev["n_points"])
            print("  gpu rad", np.round(gv["radius"], 5), "\n  orc rad", np.round(ev["radius"], 5))

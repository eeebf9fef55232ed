"""ctypes binding of the CPU oracle (oracle/build/liborc.so).

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from sloam_b200 import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "oracle", "build", "liborc.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(_ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_plane_distance_point.restype = C.c_double
        _lib.orc_cylinder_distance_model.restype = C.c_double
        _lib.orc_cylinder_distance_point.restype = C.c_double
        _lib.orc_time_keyframes.restype = C.c_double
    return _lib


def default_params(**kw):
    p = abi.Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def project(p, pts, use_libm=False, want_range=True):
    n = len(pts)
    pix = np.empty(n, np.int32)
    rng = np.empty(p.img_h * p.img_w, np.float32) if want_range else None
    lib().orc_project(C.byref(p), int(use_libm), abi.ptr(pts), n, abi.ptr(pix), abi.ptr(rng))
    return pix, rng


def mask_cloud(p, pts, pix, mask):
    n = len(pts)
    tree = np.empty(n, abi.POINT)
    ground = np.empty(n, abi.POINT)
    ng = C.c_int32()
    lib().orc_mask_cloud(C.byref(p), abi.ptr(pts), n, abi.ptr(pix), abi.ptr(mask), abi.ptr(tree),
                         abi.ptr(ground), C.byref(ng))
    return tree, ground[:ng.value].copy()


def ground_planes(p, ground, pose, use_libm=False):
    B, Fg = p.n_cells(), p.numGroundFeatures
    cells = np.zeros(B, abi.CELL_PLANE)
    feats = np.zeros((B, Fg), abi.POINT)
    kept = np.zeros(max(len(ground), 1), abi.POINT)
    offs = np.zeros(B + 1, np.int32)
    pose = np.ascontiguousarray(pose)
    lib().orc_ground_planes(C.byref(p), int(use_libm), abi.ptr(ground), len(ground), abi.ptr(pose),
                            abi.ptr(cells), abi.ptr(feats), abi.ptr(kept), abi.ptr(offs))
    return cells, feats, kept[:offs[B]].copy(), offs


def plane_fit(pts, num_ground_features):
    out = np.zeros(1, abi.CELL_PLANE)
    lib().orc_plane_fit(abi.ptr(pts), len(pts), num_ground_features, abi.ptr(out))
    return out[0]


def find_clusters(p, tree):
    labels = np.empty(p.img_h * p.img_w, np.uint32)
    n = C.c_int32()
    lib().orc_find_clusters(C.byref(p), abi.ptr(tree), abi.ptr(labels), C.byref(n))
    return labels, n.value


def compute_graph(p, tree):
    N = p.img_h * p.img_w
    trees = np.zeros(p.max_trees, abi.TREE)
    verts = np.zeros(p.max_trees * p.max_tree_vertices, abi.VERTEX)
    vpts = np.zeros(N, abi.POINT)
    n = C.c_int32()
    lib().orc_compute_graph(C.byref(p), abi.ptr(tree), abi.ptr(trees), C.byref(n), abi.ptr(verts),
                            abi.ptr(vpts))
    return trees[:n.value].copy(), verts, vpts


def cylinders(p, trees, verts, vpts, cells):
    T, Ft = len(trees), p.featuresPerTree
    models = np.zeros(max(T, 1), abi.TREE_MODEL)
    feats = np.zeros((max(T, 1), Ft), abi.POINT)
    trees = np.ascontiguousarray(trees)
    lib().orc_cylinders(C.byref(p), abi.ptr(trees), T, abi.ptr(verts), abi.ptr(vpts),
                        abi.ptr(cells), abi.ptr(models), abi.ptr(feats))
    return models[:T], feats[:T]


def ransac_draw_table(n, n_draws):
    pairs = np.empty((n_draws, 2), np.int32)
    lib().orc_ransac_draw_table(n, n_draws, abi.ptr(pairs))
    return pairs


def associate(det, tf, mp):
    det = np.ascontiguousarray(det)
    mp = np.ascontiguousarray(mp)
    bi = np.empty(len(det), np.int32)
    bd = np.empty(len(det), np.float64)
    tfp = abi.ptr(np.ascontiguousarray(tf)) if tf is not None else None
    lib().orc_associate(abi.ptr(det), len(det), tfp, abi.ptr(mp), len(mp), abi.ptr(bi), abi.ptr(bd))
    return bi, bd


def optimize_pose(p, mode, pose_est, tree_feat, tree_obj, plane_feat, plane_obj, optim_trees=True,
                  optim_ground=True):
    out = np.zeros(1, abi.POSE)
    it = np.zeros(2, np.int32)
    term = np.zeros(2, np.int32)
    tree_feat = np.ascontiguousarray(tree_feat, np.float64)
    plane_feat = np.ascontiguousarray(plane_feat, np.float64)
    tree_obj = np.ascontiguousarray(tree_obj)
    plane_obj = np.ascontiguousarray(plane_obj)
    pose_est = np.ascontiguousarray(pose_est)
    lib().orc_optimize_pose(C.byref(p), mode, abi.ptr(pose_est), abi.ptr(tree_feat), abi.ptr(tree_obj),
                            len(tree_obj), abi.ptr(plane_feat), abi.ptr(plane_obj), len(plane_obj),
                            int(optim_trees), int(optim_ground), abi.ptr(out), abi.ptr(it),
                            abi.ptr(term))
    return out[0], it, term


class KeyframeOut:
    pass


def run_keyframe(p, points, mask, pose_est, first_scan, map_models, prev_planes, use_libm=False,
                 intermediates=False):
    N = p.img_h * p.img_w
    B = p.n_cells()
    o = KeyframeOut()
    res = np.zeros(1, abi.KF_RESULT)
    o.matches = np.full(p.max_trees, -1, np.int32)
    o.tm = np.zeros(p.max_trees, abi.CYLINDER)
    o.tm_id = np.zeros(p.max_trees, np.int32)
    o.planes = np.zeros(p.max_prev_planes, abi.PLANE)
    npl = C.c_int32()
    map_models = np.ascontiguousarray(map_models)
    prev_planes = np.ascontiguousarray(prev_planes)
    pose_est = np.ascontiguousarray(pose_est)
    if intermediates:
        o.pix = np.empty(N, np.int32)
        o.range_image = np.empty(N, np.float32)
        o.cells = np.zeros(B, abi.CELL_PLANE)
        o.trees = np.zeros(p.max_trees, abi.TREE)
        o.vertices = np.zeros(p.max_trees * p.max_tree_vertices, abi.VERTEX)
        o.vertex_points = np.zeros(N, abi.POINT)
        o.tree_models = np.zeros(p.max_trees, abi.TREE_MODEL)
        ntr = C.c_int32()
        extra = [abi.ptr(o.pix), abi.ptr(o.range_image), abi.ptr(o.cells), abi.ptr(o.trees),
                 C.byref(ntr), abi.ptr(o.vertices), abi.ptr(o.vertex_points), abi.ptr(o.tree_models)]
    else:
        extra = [None] * 8
    lib().orc_run_keyframe(C.byref(p), int(use_libm), abi.ptr(points), abi.ptr(mask),
                           abi.ptr(pose_est), int(first_scan), abi.ptr(map_models), len(map_models),
                           abi.ptr(prev_planes), len(prev_planes), abi.ptr(res), abi.ptr(o.matches),
                           abi.ptr(o.tm), abi.ptr(o.tm_id), abi.ptr(o.planes), C.byref(npl), *extra)
    o.result = res[0]
    o.n_planes = npl.value
    if intermediates:
        o.n_trees = ntr.value
    return o


def run_sloam(p, ground, trees, verts, vpts, pose_est, first_scan, map_models, prev_planes,
              use_libm=False):
    o = KeyframeOut()
    res = np.zeros(1, abi.KF_RESULT)
    o.matches = np.full(p.max_trees, -1, np.int32)
    o.tm = np.zeros(p.max_trees, abi.CYLINDER)
    o.tm_id = np.zeros(p.max_trees, np.int32)
    o.planes = np.zeros(p.max_prev_planes, abi.PLANE)
    o.tree_models = np.zeros(p.max_trees, abi.TREE_MODEL)
    npl = C.c_int32()
    trees = np.ascontiguousarray(trees)
    map_models = np.ascontiguousarray(map_models)
    prev_planes = np.ascontiguousarray(prev_planes)
    pose_est = np.ascontiguousarray(pose_est)
    lib().orc_run_sloam(C.byref(p), int(use_libm), abi.ptr(ground), len(ground), abi.ptr(trees),
                        len(trees), abi.ptr(verts), abi.ptr(vpts), abi.ptr(pose_est),
                        int(first_scan), abi.ptr(map_models), len(map_models), abi.ptr(prev_planes),
                        len(prev_planes), abi.ptr(res), abi.ptr(o.matches), abi.ptr(o.tm),
                        abi.ptr(o.tm_id), abi.ptr(o.planes), C.byref(npl), abi.ptr(o.tree_models))
    o.result = res[0]
    o.n_planes = npl.value
    return o


def identity_pose(n=1):
    p = np.zeros(n, abi.POSE)
    p["q"][:, 3] = 1.0
    return p


class OracleMap:
    """MapManager restatement (oracle/orc_map.cpp)."""

    def __init__(self):
        lib().orc_map_create.restype = C.c_void_p
        self.h = C.c_void_p(lib().orc_map_create())

    def __del__(self):
        try:
            lib().orc_map_destroy(self.h)
        except Exception:
            pass

    def size(self):
        return lib().orc_map_size(self.h)

    def get_submap(self, pose, cap):
        sub = np.zeros(cap, abi.CYLINDER)
        idx = np.zeros(cap, np.int32)
        pose = np.ascontiguousarray(pose)
        n = lib().orc_map_get_submap(self.h, abi.ptr(pose), abi.ptr(sub), abi.ptr(idx), cap)
        return sub[:n].copy(), idx[:n].copy()

    def update(self, tm, ids, matches):
        tm = np.ascontiguousarray(tm)
        ids = np.ascontiguousarray(ids, np.int32)
        matches = np.ascontiguousarray(matches, np.int32)
        lib().orc_map_update(self.h, abi.ptr(tm), abi.ptr(ids), abi.ptr(matches), len(tm))

    def dump(self, cap=1 << 16):
        models = np.zeros(cap, abi.CYLINDER)
        hits = np.zeros(cap, np.int32)
        n = lib().orc_map_dump(self.h, abi.ptr(models), abi.ptr(hits), cap)
        return models[:n].copy(), hits[:n].copy()


def make_tensor(range_image, mean=12.97, std=12.35):
    """Segmentation::_makeTensor restatement -> (tensor, invalid flags, ordered invalid indices)"""
    r = np.ascontiguousarray(range_image, np.float32).reshape(-1)
    tensor = np.empty_like(r)
    invalid = np.empty(r.size, np.uint8)
    idx = np.empty(r.size, np.int32)
    n = lib().orc_make_tensor(abi.ptr(r), r.size, C.c_float(mean), C.c_float(std), abi.ptr(tensor), abi.ptr(invalid),
                              abi.ptr(idx))
    return tensor, invalid, idx[:n].copy()


def mask_from_logits(logits, invalid=None):
    """Segmentation::_mask restatement; logits [3][n] channel-major"""
    o = np.ascontiguousarray(logits, np.float32)
    n = o.size // 3
    mask = np.empty(n, np.uint8)
    inv = np.ascontiguousarray(invalid, np.uint8) if invalid is not None else None
    lib().orc_mask_from_logits(abi.ptr(o), n, abi.ptr(inv) if inv is not None else None, abi.ptr(mask))
    return mask

"""ctypes / numpy mirrors of the records in include/sloam_b200.h.

Host-side plumbing only: the product is the CUDA library behind the C ABI
(sloam_b200/csrc); these definitions let tests and bench.py pass buffers to
it (and to the CPU oracle, which uses the same flattened records).
"""
import ctypes as C

import numpy as np

POINT = np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("intensity", "f4")])
POSE = np.dtype([("t", "f8", 3), ("q", "f8", 4)])  # q = x, y, z, w
CYLINDER = np.dtype([("root", "f8", 3), ("ray", "f8", 3), ("radius", "f8")])
PLANE = np.dtype([("plane", "f8", 4), ("centroid", "f8", 3)])
VERTEX = np.dtype([("cx", "f4"), ("cy", "f4"), ("cz", "f4"), ("radius", "f4"),
                   ("n_points", "i4"), ("point_begin", "i4"), ("row", "i4"),
                   ("is_valid", "i4")])
TREE = np.dtype([("tree_id", "i4"), ("n_vertices", "i4"), ("vertex_begin", "i4"),
                 ("n_points", "i4")])
CELL_PLANE = np.dtype([("model", PLANE), ("n_cell", "i4"), ("n_kept", "i4"),
                       ("is_valid", "i4"), ("accepted", "i4")])
TREE_MODEL = np.dtype([("model", CYLINDER), ("id", "i4"), ("is_valid", "i4"),
                       ("plane_index", "i4"), ("n_inliers", "i4"),
                       ("best_hypothesis", "i4"), ("n_hypotheses", "i4"),
                       ("n_refit_inliers", "i4"), ("reserved", "i4")])
KF_RESULT = np.dtype([("status", "i4"), ("success", "i4"), ("n_ground", "i4"),
                      ("n_planes", "i4"), ("n_trees", "i4"), ("n_landmarks", "i4"),
                      ("n_tree_matches", "i4"), ("n_plane_matches", "i4"),
                      ("lm_iterations", "i4", 2), ("lm_termination", "i4", 2),
                      ("T_Map_Curr", POSE), ("T_Delta", POSE)])

assert POINT.itemsize == 16 and POSE.itemsize == 56 and CYLINDER.itemsize == 56
assert PLANE.itemsize == 56 and VERTEX.itemsize == 32 and TREE.itemsize == 16
assert CELL_PLANE.itemsize == 72 and TREE_MODEL.itemsize == 88
assert KF_RESULT.itemsize == 160

KF_OK, KF_EMPTY_MAP, KF_NO_MODELS, KF_NOT_CONVERGED = 0, 1, 2, 3


class Params(C.Structure):
    """sloam_params (include/sloam_b200.h)."""
    _fields_ = [
        ("img_h", C.c_int32), ("img_w", C.c_int32),
        ("fov_up_deg", C.c_float), ("fov_down_deg", C.c_float),
        ("do_destagger", C.c_int32),
        ("scansPerSweep", C.c_int32),
        ("minTreeModels", C.c_double), ("minGroundModels", C.c_double),
        ("maxLidarDist", C.c_double), ("maxGroundLidarDist", C.c_double),
        ("minGroundLidarDist", C.c_double),
        ("twoStepOptim", C.c_int32),
        ("groundRadiiBins", C.c_int32), ("groundThetaBins", C.c_int32),
        ("groundRetainThresh", C.c_double),
        ("groundMatchThresh", C.c_double), ("roughTreeMatchThresh", C.c_double),
        ("treeMatchThresh", C.c_double),
        ("maxTreeRadius", C.c_double), ("maxAxisTheta", C.c_double),
        ("maxFocusOutlierDistance", C.c_double),
        ("AddNewTreeThreshDist", C.c_double),
        ("featuresPerTree", C.c_int32), ("numGroundFeatures", C.c_int32),
        ("defaultTreeRadius", C.c_double),
        ("max_dist_to_centroid", C.c_float), ("cluster_dist_thresh", C.c_float),
        ("min_cluster_points", C.c_int32), ("min_vertex_points", C.c_int32),
        ("min_tree_vertices", C.c_int32), ("max_tree_vertices", C.c_int32),
        ("ransac_threshold", C.c_double), ("ransac_max_iterations", C.c_int32),
        ("ransac_probability", C.c_double), ("ransac_fixed_hypotheses", C.c_int32),
        ("min_tree_height_sq", C.c_double), ("root_plane_max_dist", C.c_double),
        ("plane_match_thresh", C.c_double), ("ground_angle_tol", C.c_double),
        ("huber_delta", C.c_double), ("lm_max_iterations", C.c_int32),
        ("max_trees", C.c_int32), ("max_map_models", C.c_int32),
        ("max_prev_planes", C.c_int32),
    ]

    def copy(self):
        p = Params()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(Params))
        return p

    def n_cells(self):
        return self.groundRadiiBins * self.groundThetaBins


class SynthConfig(C.Structure):
    """sloam_synth_config (include/sloam_b200.h)."""
    _fields_ = [
        ("img_h", C.c_int32), ("img_w", C.c_int32),
        ("fov_up_deg", C.c_float), ("fov_down_deg", C.c_float),
        ("n_trees", C.c_int32),
        ("tree_r_min", C.c_float), ("tree_r_max", C.c_float),
        ("trunk_radius_min", C.c_float), ("trunk_radius_max", C.c_float),
        ("max_tilt_deg", C.c_float), ("sensor_height", C.c_float),
        ("ground_slope_deg", C.c_float),
        ("ground_noise", C.c_float), ("range_noise", C.c_float),
        ("max_range", C.c_float), ("step_per_keyframe", C.c_float),
        ("azimuth_offset_cols", C.c_float),
        ("guess_sigma_t", C.c_float), ("guess_sigma_r", C.c_float),
        ("nan_no_return", C.c_int32),
        ("seed", C.c_uint64),
    ]


class ProfKernel(C.Structure):
    """sloam_prof_kernel"""
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_double), ("launches", C.c_int32), ("reserved", C.c_int32)]


class BatchIn(C.Structure):
    """sloam_batch_in: raw pointers (host or device, depending on the entry)."""
    _fields_ = [
        ("points", C.c_void_p), ("mask", C.c_void_p), ("pose_est", C.c_void_p),
        ("first_scan", C.c_void_p), ("map_models", C.c_void_p),
        ("n_map_models", C.c_void_p), ("map_shared", C.c_int32),
        ("prev_planes", C.c_void_p), ("n_prev_planes", C.c_void_p),
    ]


class BatchOut(C.Structure):
    _fields_ = [
        ("results", C.c_void_p), ("matches", C.c_void_p), ("tm", C.c_void_p),
        ("tm_id", C.c_void_p), ("planes", C.c_void_p), ("n_planes", C.c_void_p),
        ("range_image", C.c_void_p),
    ]


class Intermediates(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "pix", "tree", "ground", "ground_count", "cells", "cell_features",
        "trees", "n_trees", "vertices", "vertex_points", "tree_models",
        "tree_features")]


def ptr(a):
    """void* of a numpy array (must stay alive while the callee runs)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)

"""Load the committed golden vectors (tests/golden/*.npz, made by scripts/make_golden.py
from the reference's sloam/src/tests/aux fixtures)."""
import os

import numpy as np

from sloam_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAIRS = [("still", "t0"), ("still", "t1"), ("moving", "t0"), ("moving", "t1")]


class Trellis:
    """One *_tree_*.pcd -> *_landmarks_* pair."""

    def __init__(self, prefix, stamp):
        z = np.load(os.path.join(GOLDEN, f"{prefix}_{stamp}_trellis.npz"))
        self.W, self.H = int(z["width"]), int(z["height"])
        tree = np.zeros(self.W * self.H, abi.POINT)
        tree["x"] = tree["y"] = tree["z"] = np.nan
        idx = z["finite_index"]
        f = z["finite_xyzi"]
        tree["x"][idx], tree["y"][idx], tree["z"][idx], tree["intensity"][idx] = f.T
        self.tree_cloud = tree
        self.tree_nvertices = z["tree_nvertices"]
        meta = z["vertex_meta"]
        self.v_tree_id, self.v_valid, self.v_npoints = meta[:, 0], meta[:, 3], meta[:, 4]
        assert np.all(meta[:, 1] == 0) and np.all(meta[:, 2] == 0)  # beam, prevVertexSize
        self.v_radius = z["vertex_radius"]
        self.v_coords = z["vertex_coords"]
        self.v_points = z["vertex_points"]

    def as_flat(self):
        """The stored landmarks in the flattened ABI layout (trees, vertices, points).
        radius is stored as the fixture has it (0.5 x what trellis.cpp:98 computes now)."""
        nt = len(self.tree_nvertices)
        trees = np.zeros(nt, abi.TREE)
        verts = np.zeros(len(self.v_radius), abi.VERTEX)
        vb = np.concatenate([[0], np.cumsum(self.tree_nvertices)])
        pb = np.concatenate([[0], np.cumsum(self.v_npoints)])
        for i in range(nt):
            trees[i] = (self.v_tree_id[vb[i]], self.tree_nvertices[i], vb[i],
                        pb[vb[i + 1]] - pb[vb[i]])
        verts["cx"], verts["cy"], verts["cz"] = self.v_coords.T.astype(np.float32)
        verts["radius"] = (2.0 * self.v_radius).astype(np.float32)
        verts["n_points"] = self.v_npoints
        verts["point_begin"] = pb[:-1]
        verts["row"] = -1
        verts["is_valid"] = self.v_valid
        pts = np.zeros(len(self.v_points), abi.POINT)
        pts["x"], pts["y"], pts["z"] = self.v_points.T.astype(np.float32)
        return trees, verts, pts


def ground(prefix, stamp):
    z = np.load(os.path.join(GOLDEN, f"{prefix}_{stamp}_ground.npz"))["xyzi"]
    g = np.zeros(len(z), abi.POINT)
    g["x"], g["y"], g["z"], g["intensity"] = z.T
    return g

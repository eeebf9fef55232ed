#!/usr/bin/env python
"""Convert the reference's own test fixtures into compact golden vectors.

Reads /root/reference/sloam/src/tests/aux/{still,moving}_{tree,ground}_{t0,t1}.pcd
(PCD v0.7 ASCII) and {still,moving}_landmarks_{t0,t1} (Boost text archive v17 of
std::vector<std::vector<TreeVertex>>, written by the dump code at
sloam/src/segmentation/inferenceNode.cpp:128-135 through
sloam/include/helpers/serialization.h:13-31) and writes tests/golden/*.npz.

/root/reference does not exist on the GPU box, so the tests read only the
committed .npz files.  Run once in the build container:
    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

REF = "/root/reference/sloam/src/tests/aux"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def read_pcd_ascii(path):
    """-> (width, height, float32 [n,4])"""
    with open(path) as f:
        width = height = None
        while True:
            line = f.readline()
            if line.startswith("WIDTH"):
                width = int(line.split()[1])
            elif line.startswith("HEIGHT"):
                height = int(line.split()[1])
            elif line.startswith("FIELDS"):
                assert line.split()[1:] == ["x", "y", "z", "intensity"], line
            elif line.startswith("DATA"):
                assert line.split()[1] == "ascii"
                break
        data = np.loadtxt(f, dtype=np.float32, ndmin=2)
    assert data.shape == (width * height, 4), (data.shape, width, height)
    return width, height, data


def read_landmarks(path):
    """Boost text archive v17 of vector<vector<TreeVertex>> (SURVEY appendix C grammar).

    -> list of trees, each a list of dicts(treeId, beam, prevVertexSize, radius,
       isValid, coords[3], points[n,3])
    """
    tok = open(path).read().split()
    pos = 0

    def take(n=1):
        nonlocal pos
        out = tok[pos:pos + n]
        pos += n
        return out

    assert take(3) == ["22", "serialization::archive", "17"]
    take(2)                      # outer vector: tracking / version
    ntrees = int(take()[0])
    take(1)                      # item_version
    trees = []
    first_inner = first_vertex = first_coords = first_points = True
    for _ in range(ntrees):
        if first_inner:
            take(2)
            first_inner = False
        nvtx = int(take()[0])
        take(1)                  # item_version
        tree = []
        for _ in range(nvtx):
            if first_vertex:
                take(2)
                first_vertex = False
            treeId, beam, prev = (int(x) for x in take(3))
            radius = float(take()[0])
            valid = int(take()[0])
            if first_coords:
                take(2)
                first_coords = False
            coords = [float(x) for x in take(3)]
            if first_points:
                take(2)
                first_points = False
            npts = int(take()[0])
            take(1)              # item_version
            pts = np.array([float(x) for x in take(3 * npts)], np.float64).reshape(npts, 3)
            tree.append(dict(treeId=treeId, beam=beam, prevVertexSize=prev, radius=radius,
                             isValid=valid, coords=coords, points=pts))
        trees.append(tree)
    assert pos == len(tok), (pos, len(tok))
    return trees


def main():
    if not os.path.isdir(REF):
        sys.exit("reference fixtures not found at " + REF)
    os.makedirs(OUT, exist_ok=True)
    for prefix in ("still", "moving"):
        for stamp in ("t0", "t1"):
            w, h, tree = read_pcd_ascii(f"{REF}/{prefix}_tree_{stamp}.pcd")
            finite = np.flatnonzero(np.isfinite(tree[:, 0]))
            # non-finite pixels are (nan, nan, nan, *): their intensity is never read
            assert np.all(np.isnan(tree[~np.isfinite(tree[:, 0]), :3]))
            lm = read_landmarks(f"{REF}/{prefix}_landmarks_{stamp}")
            t_nv = np.array([len(t) for t in lm], np.int32)
            verts = [v for t in lm for v in t]
            v_meta = np.array([(v["treeId"], v["beam"], v["prevVertexSize"], v["isValid"],
                                len(v["points"])) for v in verts], np.int32)
            v_radius = np.array([v["radius"] for v in verts], np.float64)
            v_coords = np.array([v["coords"] for v in verts], np.float64)
            v_points = np.concatenate([v["points"] for v in verts]).astype(np.float64)
            np.savez_compressed(
                f"{OUT}/{prefix}_{stamp}_trellis.npz", width=w, height=h,
                finite_index=finite.astype(np.int32), finite_xyzi=tree[finite],
                tree_nvertices=t_nv, vertex_meta=v_meta, vertex_radius=v_radius,
                vertex_coords=v_coords, vertex_points=v_points)
            print(prefix, stamp, "trees", len(lm), "vertices", len(verts), "finite", len(finite))
        for stamp in ("t0", "t1"):
            # the reference's tests only load the `still` ground clouds (core_test.cpp:30-70); the
            # `moving` pair + poses.txt pin the whole a3..a19 path on data it never exercised
            w, h, ground = read_pcd_ascii(f"{REF}/{prefix}_ground_{stamp}.pcd")
            np.savez_compressed(f"{OUT}/{prefix}_{stamp}_ground.npz", xyzi=ground)
            print(prefix, stamp, "ground", ground.shape)
    poses = read_poses(f"{REF}/poses.txt")
    import json
    with open(f"{OUT}/moving_poses.json", "w") as f:
        json.dump(poses, f, indent=1)
    print("poses", poses)


def read_poses(path):
    """poses.txt: two nav_msgs/Odometry dumps (t0, t1) of the `moving` pair -> {stamp: {t, q}}."""
    out, stamp, sect = {}, None, None
    for line in open(path):
        s = line.strip()
        if s in ("t0", "t1"):
            stamp = s
            out[stamp] = {"t": [None] * 3, "q": [None] * 4}
        elif s in ("position:", "orientation:"):
            sect = s[:-1]
        elif stamp and sect and len(s) > 2 and s[0] in "xyzw" and s[1] == ":":
            v = float(s[2:])
            if sect == "position":
                out[stamp]["t"]["xyz".index(s[0])] = v
            else:
                out[stamp]["q"]["xyzw".index(s[0])] = v
            if (sect == "position" and s[0] == "z") or (sect == "orientation" and s[0] == "w"):
                sect = None
    return out


if __name__ == "__main__":
    main()

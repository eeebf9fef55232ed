"""GPU parity tests: every stage of the CUDA path, called through the C ABI,
against the CPU oracle on identical inputs (and against the reference's golden
fixtures for the tree detector).

Bars (BASELINE.json north_star): range-image indices, labels, cluster ids,
inlier counts and association indices bit-exact; cylinder axis <= 1e-4 rad,
radius <= 1e-4 m; pose <= 1e-5 m / 1e-5 rad.
"""
import numpy as np
import pytest

import golden_io
import synth_matches as sm
from sloam_b200 import abi

pytestmark = pytest.mark.gpu

AXIS_TOL = 1e-4      # rad
RADIUS_TOL = 1e-4    # m
POSE_T_TOL = 1e-5    # m
POSE_R_TOL = 1e-5    # rad


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from sloam_b200 import capi as c
    c.lib()
    return c


def quat_angle(qa, qb):
    d = abs(float(np.dot(qa, qb)))
    return 2.0 * np.arccos(min(1.0, d))


def axis_angle(a, b):
    a, b = np.asarray(a), np.asarray(b)
    c = np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))
    return float(np.arccos(np.clip(c, -1, 1)))


def make_scene(capi, H, W, n_trees, K, k0=0, **kw):
    cfg = capi.synth_config(H, W, n_trees, **kw)
    pts, mask = capi.synth_generate_host(cfg, k0, K)
    return cfg, pts, mask


# ------------------------------------------------------------------ stage a1 + a2
@pytest.mark.parametrize("H,W,n_trees", [(64, 1024, 20), (16, 1800, 12), (128, 2048, 30)])
def test_project_split_bit_exact(capi, oracle, H, W, n_trees):
    K = 3
    cfg, pts, mask = make_scene(capi, H, W, n_trees, K)
    # exercise the NaN / zero / clamp paths as well
    pts[0, 5] = (0.0, 0.0, 0.0, 1.0)
    pts[1, 7] = (np.nan, 1.0, 2.0, 1.0)
    pts[2, 11] = (0.0, 0.0, 5.0, 1.0)      # straight up: clamps to row 0
    pts[2, 12] = (1e-3, 0.0, -7.0, 1.0)    # straight down: clamps to the last row
    p = capi.default_params(img_h=H, img_w=W)
    ctx = capi.Context(p, K)
    d_pts, d_mask = capi.to_dev(pts), capi.to_dev(mask)
    pix, rng, tree, ground, cnt = ctx.project_split(d_pts, d_mask, K)
    ctx.sync()
    N = H * W
    pix = capi.to_host(pix, np.int32, (K, N))
    rng = capi.to_host(rng, np.float32, (K, N))
    tree = capi.to_host(tree, abi.POINT, (K, N))
    ground = capi.to_host(ground, abi.POINT, (K, N))
    cnt = capi.to_host(cnt, np.int32, (K,))
    # the unfused entries must agree with the fused kernel
    pix2, rng2 = ctx.project(d_pts, K)
    tree2, ground2, cnt2 = ctx.mask_cloud(d_pts, pix2, d_mask, K)
    ctx.sync()
    assert np.array_equal(capi.to_host(pix2, np.int32, (K, N)), pix)
    assert np.array_equal(capi.to_host(rng2, np.float32, (K, N)).view(np.uint32), rng.view(np.uint32))
    assert np.array_equal(capi.to_host(cnt2, np.int32, (K,)), cnt)
    assert capi.to_host(tree2, abi.POINT, (K, N)).tobytes() == tree.tobytes()
    for k in range(K):
        o_pix, o_rng = oracle.project(p, pts[k])
        assert np.array_equal(pix[k], o_pix)                                  # indices: bit-exact
        assert np.array_equal(rng[k].view(np.uint32), o_rng.view(np.uint32))   # range image: bit-exact
        o_tree, o_ground = oracle.mask_cloud(p, pts[k], o_pix, mask[k])
        assert cnt[k] == len(o_ground)
        assert ground[k, :cnt[k]].tobytes() == o_ground.tobytes()              # order-preserving split
        assert tree[k].tobytes() == o_tree.tobytes()                            # NaN pattern included
    ctx.close()


def test_projection_det_vs_libm_deviation_is_tiny(capi, oracle):
    """The GPU matches the oracle in `det` mode (include/sloam_b200_detmath.h); this
    measures how far that is from glibc's atan2f/asinf, which the reference calls."""
    cfg, pts, mask = make_scene(capi, 64, 1024, 20, 4)
    p = oracle.default_params(img_h=64, img_w=1024)
    diff = 0
    for k in range(4):
        a, _ = oracle.project(p, pts[k], use_libm=False, want_range=False)
        b, _ = oracle.project(p, pts[k], use_libm=True, want_range=False)
        diff += int((a != b).sum())
    assert diff <= 4 * 64 * 1024 * 1e-4


# -------------------------------------------------------------- stage a3 + a4 + a5
def check_ground(capi, oracle, p, ground_list, poses):
    K = len(ground_list)
    stride = p.img_h * p.img_w
    g = np.zeros((K, stride), abi.POINT)
    cnt = np.zeros(K, np.int32)
    for k, gk in enumerate(ground_list):
        g[k, :len(gk)] = gk
        cnt[k] = len(gk)
    ctx = capi.Context(p, K)
    cells, feats, kept, offs = ctx.ground_planes(capi.to_dev(g), capi.to_dev(cnt), stride, capi.to_dev(poses), K)
    ctx.sync()
    B, Fg = p.n_cells(), p.numGroundFeatures
    cells = capi.to_host(cells, abi.CELL_PLANE, (K, B))
    feats = capi.to_host(feats, abi.POINT, (K, B, Fg))
    kept = capi.to_host(kept, abi.POINT, (K, stride))
    offs = capi.to_host(offs, np.int32, (K, B + 1))
    n_valid = 0
    for k in range(K):
        oc, of, ok, oo = oracle.ground_planes(p, ground_list[k], poses[k:k + 1])
        assert np.array_equal(cells[k]["n_cell"], oc["n_cell"])       # binning: bit-exact
        assert np.array_equal(cells[k]["n_kept"], oc["n_kept"])
        assert np.array_equal(offs[k], oo)
        assert kept[k, :oo[B]].tobytes() == ok.tobytes()               # retained sets, in order
        assert np.array_equal(cells[k]["is_valid"], oc["is_valid"])
        assert np.array_equal(cells[k]["accepted"], oc["accepted"])
        v = oc["is_valid"] == 1
        n_valid += int(v.sum())
        # centroid: same float32 sequential sum -> identical; normal: same sign, ~1e-12
        assert np.array_equal(cells[k]["model"]["centroid"][v], oc["model"]["centroid"][v])
        assert np.allclose(cells[k]["model"]["plane"][v], oc["model"]["plane"][v], rtol=0, atol=1e-9)
        assert feats[k][v].tobytes() == of[v].tobytes()
    ctx.close()
    return n_valid


def test_ground_planes_synthetic(capi, oracle):
    H, W, K = 64, 1024, 3
    cfg, pts, mask = make_scene(capi, H, W, 20, K)
    p = capi.default_params(img_h=H, img_w=W)
    grounds, poses = [], np.zeros(K, abi.POSE)
    for k in range(K):
        pix, _ = oracle.project(p, pts[k], want_range=False)
        _, g = oracle.mask_cloud(p, pts[k], pix, mask[k])
        grounds.append(g)
        poses[k] = capi.synth_pose(cfg, k)[1]
    assert check_ground(capi, oracle, p, grounds, poses) > 20


def test_ground_planes_reference_fixture_and_ties(capi, oracle):
    """The reference's own ground cloud (still_ground_t0.pcd) with the parameters of
    core_test.cpp:94-119 (1 x 18 cells, retain 0.05, 60 features), plus a tie-injected
    copy: the reference's std::sort is not stable, so which of the tied points survive and in
    which order is libstdc++'s introsort artefact (SURVEY B-3) -- the oracle calls std::sort,
    the device replays it (csrc/dev_stdsort.h)."""
    g0 = golden_io.ground("still", "t0")
    p = capi.default_params(img_h=64, img_w=2048, maxGroundLidarDist=30.0, minGroundLidarDist=0.0,
                            groundRadiiBins=1, groundThetaBins=18, groundRetainThresh=0.05,
                            numGroundFeatures=60)
    ties = g0.copy()
    ties["z"] = np.round(ties["z"] * 20) / 20   # 5 cm quantisation: thousands of exact z ties
    poses = np.zeros(2, abi.POSE)
    poses["q"][:, 3] = 1
    assert check_ground(capi, oracle, p, [g0, ties], poses) >= 20


def test_ground_tie_replay_stress(capi, oracle):
    """Exact z ties at every density and cell size (SURVEY B-3): the warp-cooperative replay of
    libstdc++'s std::sort (k2_ground.cu: warp_sort_prefix) against the oracle's real std::sort --
    retained sets in order, centroids bit-equal.  Includes cells that are ONE z value (quicksort's
    worst case: the depth limit and the heapsort fallback), cells larger than the replay's
    shared-memory list, and different retain fractions."""
    rng = np.random.default_rng(77)
    H, W = 64, 1024
    clouds = []
    for n_pts, quant in [(3000, 0.05), (9000, 0.01), (20000, 0.002), (40000, 0.1), (60000, 0.0005), (5000, 1000.0)]:
        ang = rng.uniform(-np.pi, np.pi, n_pts)
        rad = rng.uniform(5.5, 24.0, n_pts)
        g = np.zeros(n_pts, abi.POINT)
        g["x"], g["y"] = (rad * np.cos(ang)).astype(np.float32), (rad * np.sin(ang)).astype(np.float32)
        z = -3.4 + 0.02 * rad * np.cos(ang) + rng.normal(0, 0.03, n_pts)
        g["z"] = (np.round(z / quant) * quant).astype(np.float32)     # quant = 1000: every z is 0
        g["intensity"] = rng.uniform(0, 1, n_pts).astype(np.float32)
        clouds.append(g)
    # one cell of 12 000 points (more than the replay keeps in shared memory) with heavy ties
    big = clouds[3][:12000].copy()
    big["x"], big["y"] = rng.uniform(8, 9, 12000).astype(np.float32), rng.uniform(0.5, 1.0, 12000).astype(np.float32)
    clouds.append(big)
    poses = np.zeros(len(clouds), abi.POSE)
    poses["q"][:, 3] = 1
    poses["t"][:, 2] = 5.0
    for retain, fg in [(0.05, 5), (0.3, 12)]:
        p = capi.default_params(img_h=H, img_w=W, groundRetainThresh=retain, numGroundFeatures=fg)
        assert check_ground(capi, oracle, p, clouds, poses) >= 30


def test_ground_small_cells_and_empty(capi, oracle):
    """Cells below 1/retain points keep input order; empty cloud; fewer points than features."""
    rng = np.random.default_rng(5)
    p = capi.default_params(img_h=16, img_w=64, groundRadiiBins=2, groundThetaBins=4, numGroundFeatures=3,
                            minGroundLidarDist=1.0, maxGroundLidarDist=20.0)
    def cloud(n):
        g = np.zeros(n, abi.POINT)
        r, th = rng.uniform(0.5, 22, n), rng.uniform(-np.pi, np.pi, n)
        g["x"], g["y"] = (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)
        g["z"] = rng.normal(-3.4, 0.05, n).astype(np.float32)
        return g
    poses = np.zeros(4, abi.POSE)
    poses["q"][:, 3] = 1
    poses["t"][:, 2] = 3.4
    check_ground(capi, oracle, p, [cloud(0), cloud(7), cloud(90), cloud(1000)], poses)


# ------------------------------------------------------------------- stage a6 + a7
def compare_graph(capi, got, exp, check_points=True):
    trees, ntr, verts, vpts = got
    e_trees, e_verts, e_vpts = exp
    assert ntr == len(e_trees)
    assert np.array_equal(trees["tree_id"][:ntr], e_trees["tree_id"])       # cluster ids: bit-exact
    assert np.array_equal(trees["n_vertices"][:ntr], e_trees["n_vertices"])
    assert np.array_equal(trees["n_points"][:ntr], e_trees["n_points"])
    for t in range(ntr):
        a = verts[trees["vertex_begin"][t]:trees["vertex_begin"][t] + trees["n_vertices"][t]]
        b = e_verts[e_trees["vertex_begin"][t]:e_trees["vertex_begin"][t] + e_trees["n_vertices"][t]]
        for f in ("cx", "cy", "cz", "radius"):
            assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f
        assert np.array_equal(a["n_points"], b["n_points"])
        assert np.array_equal(a["row"], b["row"])
        assert np.all(a["is_valid"] == 1)
        if check_points:
            for va, vb in zip(a, b):
                pa = vpts[va["point_begin"]:va["point_begin"] + va["n_points"]]
                pb = e_vpts[vb["point_begin"]:vb["point_begin"] + vb["n_points"]]
                assert pa.tobytes() == pb.tobytes()


def run_graph(capi, p, clouds):
    K = len(clouds)
    N = p.img_h * p.img_w
    ctx = capi.Context(p, K)
    d_tree = capi.to_dev(np.stack(clouds))
    labels, ncl = ctx.find_clusters(d_tree, K)
    trees, ntr, verts, vpts = ctx.compute_graph(d_tree, K)
    ctx.sync()
    T, V = p.max_trees, p.max_tree_vertices
    out = dict(labels=capi.to_host(labels, np.uint32, (K, N)), ncl=capi.to_host(ncl, np.int32, (K,)),
               trees=capi.to_host(trees, abi.TREE, (K, T)), ntr=capi.to_host(ntr, np.int32, (K,)),
               verts=capi.to_host(verts, abi.VERTEX, (K, T * V)), vpts=capi.to_host(vpts, abi.POINT, (K, N)))
    ctx.close()
    return out


def test_compute_graph_reference_golden_fixtures(capi, oracle):
    """All four tree.pcd -> landmarks pairs of the reference in one batch."""
    gs = [golden_io.Trellis(*pr) for pr in golden_io.PAIRS]
    p = capi.default_params(img_h=gs[0].H, img_w=gs[0].W)
    out = run_graph(capi, p, [g.tree_cloud for g in gs])
    for k, g in enumerate(gs):
        o_labels, o_ncl = oracle.find_clusters(p, g.tree_cloud)
        assert out["ncl"][k] == o_ncl
        assert np.array_equal(out["labels"][k], o_labels)                  # labels: bit-exact
        exp = oracle.compute_graph(p, g.tree_cloud)
        compare_graph(capi, (out["trees"][k], out["ntr"][k], out["verts"][k], out["vpts"][k]), exp)
        # and directly against the reference's stored landmarks
        assert out["ntr"][k] == len(g.tree_nvertices)
        assert np.array_equal(out["trees"][k]["n_vertices"][:out["ntr"][k]], g.tree_nvertices)
        vb = np.concatenate([[0], np.cumsum(g.tree_nvertices)])
        assert np.array_equal(out["trees"][k]["tree_id"][:out["ntr"][k]], g.v_tree_id[vb[:-1]])


@pytest.mark.parametrize("H,W,n_trees", [(64, 1024, 20), (128, 2048, 30)])
def test_compute_graph_synthetic(capi, oracle, H, W, n_trees):
    K = 2
    cfg, pts, mask = make_scene(capi, H, W, n_trees, K)
    p = capi.default_params(img_h=H, img_w=W)
    clouds = []
    for k in range(K):
        pix, _ = oracle.project(p, pts[k], want_range=False)
        clouds.append(oracle.mask_cloud(p, pts[k], pix, mask[k])[0])
    out = run_graph(capi, p, clouds)
    for k in range(K):
        o_labels, o_ncl = oracle.find_clusters(p, clouds[k])
        assert out["ncl"][k] == o_ncl and np.array_equal(out["labels"][k], o_labels)
        exp = oracle.compute_graph(p, clouds[k])
        assert len(exp[0]) >= n_trees // 2
        compare_graph(capi, (out["trees"][k], out["ntr"][k], out["verts"][k], out["vpts"][k]), exp)


def test_compute_graph_edge_cases(capi, oracle):
    """Empty cloud, one giant blob wider than the warp path (wide kernel), isolated pixels,
    a component that wraps in an S shape, NaN y/z with finite x."""
    H, W = 32, 512
    p = capi.default_params(img_h=H, img_w=W, min_cluster_points=20, min_tree_vertices=4)
    rng = np.random.default_rng(3)
    def empty():
        c = np.zeros(H * W, abi.POINT)
        c["x"] = c["y"] = c["z"] = np.nan
        return c
    def put(c, r, col, x, y, z):
        i = r * W + col
        c["x"][i], c["y"][i], c["z"][i] = x, y, z
    c0 = empty()
    c1 = empty()   # blob: rows 4..27, cols 50..349 on a smooth surface (300 > 128 members per row)
    for r in range(4, 28):
        zperm = rng.permutation(300)
        for col in range(50, 350):
            # every third row: exact z ties in groups of 7 among 300 members (wide path + the
            # std::sort replay, SURVEY B-3); the others have distinct z
            dz = 2e-6 * (zperm[col - 50] // 7) if r % 3 == 0 else 2e-6 * zperm[col - 50]
            put(c1, r, col, 5.0 + 0.001 * col + rng.normal(0, 1e-4), 0.01 * col, 3.0 - 0.05 * r + dz)
    c2 = empty()   # S-shaped component + isolated pixels + a finite-x / NaN-y pixel
    # row 5: 50 members with ONE z value, row 9: z ties in pairs (n > 16: the order is
    # libstdc++'s introsort artefact, SURVEY B-3, replayed on the device), row 13: distinct z
    for col in range(10, 60):
        put(c2, 5, col, 4.0 + 0.003 * ((col * 7) % 11), 0.02 * ((col * 5) % 13), 1.0)
        put(c2, 9, col, 4.0, 0.02 * col, 0.6 + 1e-4 * (col // 2))
    for r in range(5, 10):
        put(c2, r, 59, 4.0, 0.02 * 59, 1.0 - 0.1 * (r - 5))
    for r in range(9, 14):
        put(c2, r, 10, 4.0, 0.02 * 10, 0.6 - 0.1 * (r - 9))
    for col in range(10, 60):
        put(c2, 13, col, 4.0, 0.02 * col, 0.2 + 1e-4 * col)
    put(c2, 20, 300, 9.0, 9.0, 9.0)
    put(c2, 0, 0, 1.0, 1.0, 1.0)
    put(c2, 31, 511, 2.0, 2.0, 2.0)
    c2["x"][25 * W + 100] = 3.0   # y, z stay NaN: valid pixel, never connected
    out = run_graph(capi, p, [c0, c1, c2])
    for k, c in enumerate([c0, c1, c2]):
        o_labels, o_ncl = oracle.find_clusters(p, c)
        assert out["ncl"][k] == o_ncl
        assert np.array_equal(out["labels"][k], o_labels)
        exp = oracle.compute_graph(p, c)
        compare_graph(capi, (out["trees"][k], out["ntr"][k], out["verts"][k], out["vpts"][k]), exp)
    assert out["ntr"][0] == 0 and out["ntr"][1] == 1


def test_vertex_tie_replay_stress(capi, oracle):
    """Rows of 17 .. 128 members whose x, y and z values come from small value sets, so that
    each of the three std::sort calls of computeVertexProperties (trellis.cpp:71-82) meets tie
    groups of random sizes in random positions -- the case where the order of the kept points,
    the radius and which points survive the centroid filter depend on libstdc++'s introsort
    (SURVEY B-3; warp-parallel replay, dev_warpsort.cuh).  Every row of every blob is one
    vertex; a wide max_dist_to_centroid keeps most points so that their order is compared."""
    H, W = 32, 768
    p = capi.default_params(img_h=H, img_w=W, min_cluster_points=20, min_tree_vertices=4,
                            max_dist_to_centroid=0.5)
    rng = np.random.default_rng(20260018)
    clouds = []
    for scene in range(6):
        c = np.zeros(H * W, abi.POINT)
        c["x"] = c["y"] = c["z"] = np.nan
        col0 = 8
        for blob in range(5):
            width = int(rng.integers(17, 129))
            for r in range(3, 3 + int(rng.integers(8, 24))):
                n = int(rng.integers(max(17, width - 12), width + 1))
                # 0: distinct, 1: pairs, 2: a handful of values, 3: one value
                kinds = rng.integers(0, 4, 3)
                axes = []
                for a, kind in enumerate(kinds):
                    base = [4.0, 0.3 * blob, 2.5 - 0.05 * r][a]
                    step = [2e-3, 2e-3, 1e-4][a]
                    if kind == 0:
                        v = base + step * rng.permutation(n)
                    elif kind == 1:
                        v = base + step * (rng.permutation(n) // 2)
                    elif kind == 2:
                        v = base + step * rng.integers(0, int(rng.integers(2, 9)), n)
                    else:
                        v = np.full(n, base)
                    axes.append(v.astype(np.float32))
                i0 = r * W + col0
                c["x"][i0:i0 + n], c["y"][i0:i0 + n], c["z"][i0:i0 + n] = axes
            col0 += width + 8
        clouds.append(c)
    out = run_graph(capi, p, clouds)
    n_vertices = 0
    for k, c in enumerate(clouds):
        exp = oracle.compute_graph(p, c)
        compare_graph(capi, (out["trees"][k], out["ntr"][k], out["verts"][k], out["vpts"][k]), exp)
        n_vertices += int(out["trees"][k]["n_vertices"][:out["ntr"][k]].sum())
    assert n_vertices > 200


# ------------------------------------------------------------------- stage a8..a10
def run_models(capi, oracle, p, pts, mask, poses):
    """Oracle front end (projection, split, graph, planes) -> GPU vs oracle cylinders."""
    K = len(pts)
    N, T, V, B, Ft = p.img_h * p.img_w, p.max_trees, p.max_tree_vertices, p.n_cells(), p.featuresPerTree
    trees = np.zeros((K, T), abi.TREE); ntr = np.zeros(K, np.int32)
    verts = np.zeros((K, T * V), abi.VERTEX); vpts = np.zeros((K, N), abi.POINT)
    cells = np.zeros((K, B), abi.CELL_PLANE)
    exp = []
    for k in range(K):
        pix, _ = oracle.project(p, pts[k], want_range=False)
        tree, ground = oracle.mask_cloud(p, pts[k], pix, mask[k])
        t, v, vp = oracle.compute_graph(p, tree)
        c, _, _, _ = oracle.ground_planes(p, ground, poses[k:k + 1])
        trees[k, :len(t)] = t; ntr[k] = len(t); verts[k] = v; vpts[k] = vp; cells[k] = c
        exp.append(oracle.cylinders(p, t, v, vp, c))
    ctx = capi.Context(p, K)
    models, feats = ctx.cylinders(capi.to_dev(trees), capi.to_dev(ntr), capi.to_dev(verts), capi.to_dev(vpts),
                                  capi.to_dev(cells), K)
    ctx.sync()
    models = capi.to_host(models, abi.TREE_MODEL, (K, T))
    feats = capi.to_host(feats, abi.POINT, (K, T, Ft))
    ctx.close()
    return models, feats, exp, ntr


def compare_models(models, feats, exp_models, exp_feats, n):
    m, e = models[:n], exp_models
    for f in ("id", "is_valid", "plane_index", "n_inliers", "best_hypothesis", "n_hypotheses", "n_refit_inliers"):
        assert np.array_equal(m[f], e[f]), f                                   # inlier counts: bit-exact
    ran = e["n_hypotheses"] > 0
    for i in np.flatnonzero(ran & (e["model"]["radius"] > 0)):
        assert axis_angle(m["model"]["ray"][i], e["model"]["ray"][i]) <= AXIS_TOL
        assert np.dot(m["model"]["ray"][i], e["model"]["ray"][i]) > 0          # same sign (B-7)
        assert abs(m["model"]["radius"][i] - e["model"]["radius"][i]) <= RADIUS_TOL
        assert np.allclose(m["model"]["root"][i], e["model"]["root"][i], atol=1e-4)
    v = e["is_valid"] == 1
    assert feats[:n][v].tobytes() == exp_feats[v].tobytes()


def test_cylinders_reference_faithful_ransac(capi, oracle):
    H, W, K = 64, 1024, 3
    cfg, pts, mask = make_scene(capi, H, W, 20, K)
    p = capi.default_params(img_h=H, img_w=W)
    poses = np.array([capi.synth_pose(cfg, k)[1] for k in range(K)])
    models, feats, exp, ntr = run_models(capi, oracle, p, pts, mask, poses)
    n_valid = 0
    for k in range(K):
        compare_models(models[k], feats[k], exp[k][0], exp[k][1], ntr[k])
        n_valid += int(exp[k][0]["is_valid"].sum())
    assert n_valid >= 15


def test_cylinders_fixed_hypothesis_stress_mode(capi, oracle):
    """BASELINE config 3: fixed-count RANSAC (no adaptive exit), first maximum wins, on
    noisy medians so that hypotheses really differ."""
    H, W, K = 64, 1024, 2
    cfg, pts, mask = make_scene(capi, H, W, 20, K, range_noise=0.08)
    p = capi.default_params(img_h=H, img_w=W, ransac_fixed_hypotheses=256, ransac_threshold=0.05)
    poses = np.array([capi.synth_pose(cfg, k)[1] for k in range(K)])
    models, feats, exp, ntr = run_models(capi, oracle, p, pts, mask, poses)
    seen = 0
    for k in range(K):
        compare_models(models[k], feats[k], exp[k][0], exp[k][1], ntr[k])
        ran = exp[k][0]["n_hypotheses"] > 0
        assert np.all(exp[k][0]["n_hypotheses"][ran] == 256)
        seen += int((exp[k][0]["best_hypothesis"][ran] > 0).sum())
    assert seen > 0   # the winner is not always hypothesis 0, i.e. the scoring matters


def test_ransac_draw_table_matches_pcl_stream(capi, oracle):
    """Sampling stream of SURVEY A.3 (mt19937(12345) >> 1, persistent partial Fisher-Yates)."""
    t = oracle.ransac_draw_table(17, 8)
    assert t.shape == (8, 2) and np.all(t[:, 0] != t[:, 1]) and t.min() >= 0 and t.max() < 17
    # first draw of mt19937(12345): 3992670690, 3823185381
    assert t[0, 0] == (3992670690 >> 1) % 17


# ------------------------------------------------------------------ stage a11..a13
def random_cylinders(rng, n, spread):
    c = np.zeros(n, abi.CYLINDER)
    c["root"][:, :2] = rng.uniform(-spread, spread, (n, 2))
    c["root"][:, 2] = rng.normal(0, 0.2, n)
    ray = np.stack([rng.normal(0, 0.05, n), rng.normal(0, 0.05, n), np.ones(n)], 1)
    c["ray"] = ray / np.linalg.norm(ray, axis=1)[:, None]
    c["radius"] = rng.uniform(0.1, 0.3, n)
    return c


@pytest.mark.parametrize("n_det,n_map,shared", [(37, 50, False), (300, 4000, True), (2000, 20000, True)])
def test_association_indices_bit_exact(capi, oracle, n_det, n_map, shared):
    rng = np.random.default_rng(n_det)
    K = 2
    mp = random_cylinders(rng, n_map, 200.0)
    det = np.zeros((K, n_det), abi.CYLINDER)
    tf = np.zeros(K, abi.POSE)
    for k in range(K):
        pick = rng.integers(0, n_map, n_det)
        det[k] = mp[pick]
        det[k]["root"][:, :2] += rng.normal(0, 0.2, (n_det, 2))
        det[k]["root"][: n_det // 10, :2] += 500.0           # unmatched tail
        tf[k]["t"] = rng.normal(0, 0.3, 3)
        from scipy.spatial.transform import Rotation as R
        tf[k]["q"] = R.from_rotvec(rng.normal(0, 0.02, 3)).as_quat()
    mp[3] = mp[2]   # exact duplicate: the first minimum must win
    p = capi.default_params()
    ctx = capi.Context(p, K)
    nd = np.full(K, n_det, np.int32)
    if shared:
        d_map, nm, stride = capi.to_dev(mp), np.array([n_map], np.int32), n_map
    else:
        d_map, nm, stride = capi.to_dev(np.stack([mp, mp])), np.full(K, n_map, np.int32), n_map
    bi, bd = ctx.associate(capi.to_dev(det), capi.to_dev(nd), n_det, capi.to_dev(tf), d_map, capi.to_dev(nm),
                           stride, shared, K)
    ctx.sync()
    bi = capi.to_host(bi, np.int32, (K, n_det))
    bd = capi.to_host(bd, np.float64, (K, n_det))
    for k in range(K):
        ei, ed = oracle.associate(det[k], tf[k:k + 1], mp)
        assert np.array_equal(bi[k], ei)                                    # indices: bit-exact
        assert np.array_equal(bd[k], ed)                                    # and so are the distances
    # empty map -> -1
    bi0, _ = ctx.associate(capi.to_dev(det), capi.to_dev(nd), n_det, None, d_map, capi.to_dev(np.zeros(K, np.int32)),
                           stride, shared, K)
    ctx.sync()
    assert np.all(capi.to_host(bi0, np.int32, (K, n_det)) == -1)
    ctx.close()


# ------------------------------------------------------------------ stage a14..a17
@pytest.mark.parametrize("mode", [0, 1])
def test_pose_optimisation_matches_oracle(capi, oracle, mode):
    K = 6
    pbs = [sm.make_problem(seed) for seed in range(K)]
    # one problem with a zero pose guess (small-angle branch of AngleAxisRotatePoint)
    pbs[5] = sm.make_problem(5, rotvec=(0.0, 0.0, 0.0), trans=(0.0, 0.0, 0.0), outliers=False)
    pbs[5]["guess"]["t"][0] = 0.0
    pbs[5]["guess"]["q"][0] = (0, 0, 0, 1)
    nt = np.array([len(pb["tree_obj"]) for pb in pbs], np.int32)
    npl = np.array([len(pb["plane_obj"]) for pb in pbs], np.int32)
    ts, ps = int(nt.max()), int(npl.max())
    tf, to = np.zeros((K, ts, 3)), np.zeros((K, ts), abi.CYLINDER)
    pf, po = np.zeros((K, ps, 3)), np.zeros((K, ps), abi.PLANE)
    guess = np.zeros(K, abi.POSE)
    for k, pb in enumerate(pbs):
        tf[k, :nt[k]], to[k, :nt[k]] = pb["tree_feat"], pb["tree_obj"]
        pf[k, :npl[k]], po[k, :npl[k]] = pb["plane_feat"], pb["plane_obj"]
        guess[k] = pb["guess"][0]
    ot = np.ones(K, np.uint8); og = np.ones(K, np.uint8)
    if mode == 1:
        ot[2] = 0      # trees not optimised for keyframe 2
        og[3] = 0      # ground not optimised for keyframe 3
    p = capi.default_params()
    ctx = capi.Context(p, K)
    out, it, term = ctx.optimize_pose(mode, capi.to_dev(guess), capi.to_dev(tf), capi.to_dev(to), capi.to_dev(nt),
                                      ts, capi.to_dev(pf), capi.to_dev(po), capi.to_dev(npl), ps,
                                      capi.to_dev(ot), capi.to_dev(og), K)
    ctx.sync()
    out = capi.to_host(out, abi.POSE, (K,))
    it = capi.to_host(it, np.int32, (K, 2)); term = capi.to_host(term, np.int32, (K, 2))
    for k, pb in enumerate(pbs):
        e_pose, e_it, e_term = oracle.optimize_pose(p, mode, pb["guess"], pb["tree_feat"], pb["tree_obj"],
                                                    pb["plane_feat"], pb["plane_obj"], bool(ot[k]), bool(og[k]))
        assert np.array_equal(term[k], e_term)
        assert np.array_equal(it[k], e_it)
        assert np.max(np.abs(out[k]["t"] - e_pose["t"])) <= POSE_T_TOL
        assert quat_angle(out[k]["q"], e_pose["q"]) <= POSE_R_TOL
    ctx.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_pose_optimisation_ill_conditioned(capi, oracle, mode):
    """lm_kernel on rank-deficient / badly conditioned problems (tests/synth_matches.make_degenerate:
    collinear trunks, one or two planes, zero-padded features, a single trunk, fewer rows than
    unknowns): same termination and iteration count as the oracle's QR-based solver."""
    pbs = [sm.make_degenerate(kind, seed=3) for kind in sm.DEGENERATE]
    K = len(pbs)
    nt = np.array([len(pb["tree_obj"]) for pb in pbs], np.int32)
    npl = np.array([len(pb["plane_obj"]) for pb in pbs], np.int32)
    ts, ps = int(nt.max()), int(npl.max())
    tf, to = np.zeros((K, ts, 3)), np.zeros((K, ts), abi.CYLINDER)
    pf, po = np.zeros((K, ps, 3)), np.zeros((K, ps), abi.PLANE)
    guess = np.zeros(K, abi.POSE)
    for k, pb in enumerate(pbs):
        tf[k, :nt[k]], to[k, :nt[k]] = pb["tree_feat"], pb["tree_obj"]
        pf[k, :npl[k]], po[k, :npl[k]] = pb["plane_feat"], pb["plane_obj"]
        guess[k] = pb["guess"][0]
    ones = np.ones(K, np.uint8)
    p = capi.default_params()
    ctx = capi.Context(p, K)
    out, it, term = ctx.optimize_pose(mode, capi.to_dev(guess), capi.to_dev(tf), capi.to_dev(to), capi.to_dev(nt),
                                      ts, capi.to_dev(pf), capi.to_dev(po), capi.to_dev(npl), ps,
                                      capi.to_dev(ones), capi.to_dev(ones), K)
    ctx.sync()
    out = capi.to_host(out, abi.POSE, (K,))
    it = capi.to_host(it, np.int32, (K, 2)); term = capi.to_host(term, np.int32, (K, 2))
    long_runs = 0
    for k, pb in enumerate(pbs):
        e_pose, e_it, e_term = oracle.optimize_pose(p, mode, pb["guess"], pb["tree_feat"], pb["tree_obj"],
                                                    pb["plane_feat"], pb["plane_obj"], True, True)
        assert np.array_equal(term[k], e_term), sm.DEGENERATE[k]
        assert np.array_equal(it[k], e_it), sm.DEGENERATE[k]
        assert np.max(np.abs(out[k]["t"] - e_pose["t"])) <= POSE_T_TOL, sm.DEGENERATE[k]
        assert quat_angle(out[k]["q"], e_pose["q"]) <= POSE_R_TOL, sm.DEGENERATE[k]
        long_runs += int(e_it.max() >= 15)
    assert long_runs >= 1   # at least one problem wanders along an unobservable direction
    ctx.close()


# ------------------------------------------------------------------ fused path a1..a19
def run_sequence(capi, oracle, p, cfg, K, two_step, first_all=False):
    """K keyframes: keyframe 0 is a first scan, the others use the GT scene as submap and
    the oracle's planes of the previous keyframe as prevGPlanes_."""
    N, T, PP, M = p.img_h * p.img_w, p.max_trees, p.max_prev_planes, p.max_map_models
    pts, mask = capi.synth_generate_host(cfg, 0, K)
    scene = capi.synth_scene(cfg)
    pose = np.array([capi.synth_pose(cfg, k)[1] for k in range(K)])
    first = np.zeros(K, np.uint8); first[0] = 1
    if first_all:
        first[:] = 1
    maps = np.zeros((K, M), abi.CYLINDER); nmap = np.zeros(K, np.int32)
    prev = np.zeros((K, PP), abi.PLANE); nprev = np.zeros(K, np.int32)
    exp = []
    for k in range(K):
        if not first[k]:
            maps[k, :len(scene)] = scene; nmap[k] = len(scene)
            prev[k] = exp[k - 1].planes; nprev[k] = exp[k - 1].n_planes
        exp.append(oracle.run_keyframe(p, pts[k], mask[k], pose[k:k + 1], bool(first[k]), maps[k, :nmap[k]],
                                       prev[k, :nprev[k]], intermediates=True))
    inp = dict(points=pts, mask=mask, pose_est=pose, first_scan=first, map_models=maps, n_map_models=nmap,
               prev_planes=prev, n_prev_planes=nprev)
    return inp, exp


def compare_keyframe(got_res, got_matches, got_tm, got_tm_id, got_planes, got_npl, e, tight=True):
    r, er = got_res, e.result
    for f in ("status", "success", "n_ground", "n_planes", "n_trees", "n_landmarks", "n_tree_matches",
              "n_plane_matches"):
        assert r[f] == er[f], f
    assert np.array_equal(r["lm_termination"], er["lm_termination"])
    assert np.array_equal(r["lm_iterations"], er["lm_iterations"])
    n = er["n_landmarks"]
    if er["status"] in (abi.KF_EMPTY_MAP, abi.KF_NO_MODELS):
        return
    assert np.array_equal(got_matches[:n], e.matches[:n])                      # association: bit-exact
    assert np.array_equal(got_tm_id[:n], e.tm_id[:n])
    for f in ("T_Map_Curr", "T_Delta"):
        assert np.max(np.abs(r[f]["t"] - er[f]["t"])) <= POSE_T_TOL
        assert quat_angle(r[f]["q"], er[f]["q"]) <= POSE_R_TOL
    for i in range(n):
        assert axis_angle(got_tm[i]["ray"], e.tm[i]["ray"]) <= AXIS_TOL
        assert abs(got_tm[i]["radius"] - e.tm[i]["radius"]) <= RADIUS_TOL
        assert np.allclose(got_tm[i]["root"], e.tm[i]["root"], atol=1e-4)
    assert got_npl == e.n_planes
    assert np.allclose(got_planes[:got_npl]["plane"], e.planes[:e.n_planes]["plane"], atol=1e-6)
    assert np.allclose(got_planes[:got_npl]["centroid"], e.planes[:e.n_planes]["centroid"], atol=1e-6)


@pytest.mark.parametrize("preset,two_step", [("os1-64", True), ("os1-64", False), ("vlp-16", True),
                                             ("vlp-16", False)])
def test_run_keyframes_matches_oracle(capi, oracle, preset, two_step):
    from sloam_b200 import configs
    K = 5
    p, cfg = configs.make(capi, preset, twoStepOptim=int(two_step))
    if preset == "vlp-16" and not two_step:
        p.minGroundModels = 10   # joint mode needs treeCheck && groundCheck (sloam.cpp:505)
    H, W = p.img_h, p.img_w
    inp, exp = run_sequence(capi, oracle, p, cfg, K, two_step)
    T, PP, N = p.max_trees, p.max_prev_planes, H * W
    # --- host entry (what the nodelet-side wrapper calls) ---
    ctx = capi.Context(p, K)
    out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32),
               tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32),
               planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32),
               range_image=np.zeros((K, N), np.float32))
    ctx.run_keyframes_host(K, inp, out)
    assert ctx.launches() > 0
    n_opt = 0
    for k in range(K):
        compare_keyframe(out["results"][k], out["matches"][k], out["tm"][k], out["tm_id"][k], out["planes"][k],
                         out["n_planes"][k], exp[k])
        assert np.array_equal(out["range_image"][k].view(np.uint32), exp[k].range_image.view(np.uint32))
        n_opt += int(exp[k].result["lm_termination"][0] == 0)
    assert n_opt >= 2, "the scene must actually exercise the optimiser"
    # --- device entry gives the same bytes ---
    d_in = {k: capi.to_dev(v) for k, v in inp.items()}
    d_out = ctx.alloc_outputs_dev(K)
    ctx.run_keyframes_dev(K, d_in, d_out)
    ctx.sync()
    res2 = capi.to_host(d_out["results"], abi.KF_RESULT, (K,))
    assert res2.tobytes() == out["results"].tobytes()
    ctx.close()


def test_run_keyframes_guards(capi, oracle):
    """RunSloam's early exits (sloam.cpp:476-486): empty map, no models; and all-first-scan."""
    H, W, K = 64, 1024, 3
    p = capi.default_params(img_h=H, img_w=W)
    cfg = capi.synth_config(H, W, 20)
    inp, exp = run_sequence(capi, oracle, p, cfg, K, True)
    inp["n_map_models"][1] = 0                 # keyframe 1: empty map
    inp["mask"][2][:] = 0                      # keyframe 2: nothing labelled -> no models
    T, PP = p.max_trees, p.max_prev_planes
    exp[1] = oracle.run_keyframe(p, inp["points"][1], inp["mask"][1], inp["pose_est"][1:2], False,
                                 inp["map_models"][1, :0], inp["prev_planes"][1, :inp["n_prev_planes"][1]],
                                 intermediates=True)
    exp[2] = oracle.run_keyframe(p, inp["points"][2], inp["mask"][2], inp["pose_est"][2:3], False,
                                 inp["map_models"][2, :inp["n_map_models"][2]],
                                 inp["prev_planes"][2, :inp["n_prev_planes"][2]], intermediates=True)
    assert exp[1].result["status"] == abi.KF_EMPTY_MAP and exp[2].result["status"] == abi.KF_NO_MODELS
    ctx = capi.Context(p, K)
    out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32),
               tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32),
               planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32), range_image=None)
    ctx.run_keyframes_host(K, inp, out)
    for k in range(K):
        compare_keyframe(out["results"][k], out["matches"][k], out["tm"][k], out["tm_id"][k], out["planes"][k],
                         out["n_planes"][k], exp[k])
    assert out["results"][1]["success"] == 0 and out["results"][2]["success"] == 0
    # prevGPlanes_ is untouched by a bailed-out call
    assert out["n_planes"][1] == inp["n_prev_planes"][1]
    ctx.close()


def test_host_entry_with_packed_xyz_cloud(capi, oracle):
    """sloam_b200_run_keyframes_host_xyz: the same keyframes as a 12-byte x, y, z cloud give
    byte-identical results (no output of RunSloam depends on the input intensity)."""
    H, W, K = 64, 1024, 3
    p = capi.default_params(img_h=H, img_w=W)
    cfg = capi.synth_config(H, W, 20)
    inp, exp = run_sequence(capi, oracle, p, cfg, K, True)
    T, PP = p.max_trees, p.max_prev_planes
    ctx = capi.Context(p, K)

    def outputs():
        return dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32),
                    tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32),
                    planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32), range_image=None)
    a, b = outputs(), outputs()
    ctx.run_keyframes_host(K, inp, a)
    xyz = np.ascontiguousarray(np.stack([inp["points"]["x"], inp["points"]["y"], inp["points"]["z"]], axis=-1))
    inp_xyz = dict(inp, points=None, points_xyz=xyz)
    ctx.run_keyframes_host(K, inp_xyz, b)
    for key in ("results", "matches", "tm", "tm_id", "planes", "n_planes"):
        assert a[key].tobytes() == b[key].tobytes(), key
    for k in range(K):
        compare_keyframe(b["results"][k], b["matches"][k], b["tm"][k], b["tm_id"][k], b["planes"][k], b["n_planes"][k], exp[k])
    ctx.close()


def test_c_abi_rejects_bad_arguments(capi):
    p = capi.default_params(img_h=16, img_w=64)
    ctx = capi.Context(p, 2)
    import ctypes as C
    assert capi.lib().sloam_b200_project_dev(ctx.h, 3, None, None, None) == -1      # K > max, null buffers
    assert b"bad arguments" in capi.lib().sloam_b200_last_error(ctx.h)
    bad = capi.default_params(groundRetainThresh=2.0)                                 # SURVEY B-5
    h = C.c_void_p()
    assert capi.lib().sloam_b200_create(C.byref(bad), 0, 1, C.byref(h)) == -1
    ctx.close()


# ------------------------------------------------------------ BASELINE.json configurations
@pytest.mark.parametrize("preset", ["os1-64-dense", "os1-128"])
def test_baseline_configs_full_size(capi, oracle, preset):
    """configs[2] (dense forest, 4096 fixed RANSAC hypotheses per tree) and configs[3]
    (OS1-128, 128 x 2048) through the fused path against the oracle."""
    from sloam_b200 import configs
    K = 3
    p, cfg = configs.make(capi, preset)
    inp, exp = run_sequence(capi, oracle, p, cfg, K, True)
    T, PP = p.max_trees, p.max_prev_planes
    ctx = capi.Context(p, K)
    out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32),
               tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32),
               planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32), range_image=None)
    ctx.run_keyframes_host(K, inp, out)
    for k in range(K):
        compare_keyframe(out["results"][k], out["matches"][k], out["tm"][k], out["tm_id"][k], out["planes"][k],
                         out["n_planes"][k], exp[k])
    # inlier counts / winning hypotheses of every tree, bit-exact
    it = ctx.intermediates()
    models = capi.read_dev(it.tree_models, K * T * abi.TREE_MODEL.itemsize, ctx.device).view(abi.TREE_MODEL).reshape(K, T)
    for k in range(K):
        n = exp[k].n_trees
        for f in ("n_inliers", "best_hypothesis", "n_hypotheses", "n_refit_inliers", "is_valid", "plane_index"):
            assert np.array_equal(models[k][:n][f], exp[k].tree_models[:n][f]), f
    if preset == "os1-64-dense":
        ran = exp[0].tree_models[:exp[0].n_trees]["n_hypotheses"]
        assert np.all(ran[ran > 0] == 4096) and exp[0].n_trees > 50
    ctx.close()


def test_workspace_reuse_and_dense_tree_cloud(capi, oracle):
    """Two DIFFERENT scenes through the same context.  The fused path writes only the
    tree-labelled points of its organized tree cloud (plus a bit mask) and never touches the
    connected-component state of the other pixels, so whatever the first batch left behind must
    not leak into the second; the dense cloud handed out with the intermediates (NaN points
    filled in on demand) must equal the reference's maskCloud output bit for bit."""
    from sloam_b200 import configs
    K = 3
    p, cfg_a = configs.make(capi, "os1-64")
    _, cfg_b = configs.make(capi, "os1-64")
    cfg_b.seed, cfg_b.n_trees, cfg_b.azimuth_offset_cols = 977, 33, 311.5
    T, PP, N = p.max_trees, p.max_prev_planes, p.img_h * p.img_w
    ctx = capi.Context(p, K)

    def run(cfg):
        inp, exp = run_sequence(capi, oracle, p, cfg, K, True)
        out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32),
                   tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32),
                   planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32), range_image=None)
        ctx.run_keyframes_host(K, inp, out)
        for k in range(K):
            compare_keyframe(out["results"][k], out["matches"][k], out["tm"][k], out["tm_id"][k],
                             out["planes"][k], out["n_planes"][k], exp[k])
        return inp, exp

    run(cfg_a)
    inp, exp = run(cfg_b)
    it = ctx.intermediates()
    tree = capi.read_dev(it.tree, K * N * abi.POINT.itemsize, ctx.device).view(np.uint32).reshape(K, N, 4)
    pix = capi.read_dev(it.pix, K * N * 4, ctx.device).view(np.int32).reshape(K, N)
    n_tree_px = 0
    for k in range(K):
        e_tree, _ = oracle.mask_cloud(p, inp["points"][k], exp[k].pix, inp["mask"][k])
        assert np.array_equal(pix[k], exp[k].pix)
        assert np.array_equal(tree[k], e_tree.view(np.uint32).reshape(N, 4))
        n_tree_px += int(np.isfinite(e_tree["x"]).sum())
    assert 0 < n_tree_px < K * N // 4
    # and a third batch (scene A again) after the intermediates were materialised
    run(cfg_a)
    ctx.close()


def test_lanes_give_identical_results(capi):
    """sloam_b200_set_lanes: a fused run cut into concurrent sub-batches (own streams, own
    scratch) returns byte-identical results, including an uneven split."""
    from sloam_b200 import configs
    K = 197
    p, cfg = configs.make(capi, "vlp-16", max_trees=128, max_map_models=64)
    pts, mask = capi.synth_generate_host(cfg, 0, K)
    scene = capi.synth_scene(cfg)
    T, PP, M = p.max_trees, p.max_prev_planes, p.max_map_models
    pose = np.array([capi.synth_pose(cfg, k)[1] for k in range(K)])
    first = np.zeros(K, np.uint8); first[::50] = 1
    maps = np.zeros((K, M), abi.CYLINDER); maps[:, :len(scene)] = scene
    nmap = np.full(K, len(scene), np.int32)
    inp = dict(points=pts, mask=mask, pose_est=pose, first_scan=first, map_models=maps, n_map_models=nmap,
               prev_planes=np.zeros((K, PP), abi.PLANE), n_prev_planes=np.zeros(K, np.int32))
    d_in = {k: capi.to_dev(v) for k, v in inp.items()}
    got = []
    for lanes in (1, 2, 3):
        ctx = capi.Context(p, K)
        ctx.set_lanes(lanes)
        for _ in range(2):  # the second run reuses the lanes' scratch
            d_out = ctx.alloc_outputs_dev(K)
            for v in d_out.values():
                if v is not None:
                    v.zero_()  # slots beyond n_landmarks are never written
            ctx.sync()
            ctx.run_keyframes_dev(K, d_in, d_out)
            ctx.sync()
        got.append({k: capi.to_host(d_out[k], np.uint8, None).copy() for k in ("results", "matches", "tm", "tm_id", "planes", "n_planes")})
        ctx.close()
    res = got[0]["results"].view(abi.KF_RESULT)
    assert (res["n_landmarks"] > 0).sum() > K // 2
    for other in got[1:]:
        for k, v in got[0].items():
            assert np.array_equal(v, other[k]), k


def test_odd_image_size_fused_path(capi, oracle):
    """H x W = 21 x 1031 = 21651 pixels: not a multiple of 32 (bit-mask words), of 256 (CTA) or
    of 1024 (split tiles: the last tile is partial and its TMA copy is short), rows that
    straddle the 32-pixel words everywhere.  Fused path against the oracle, plus the pixel
    indices and the dense tree / ground clouds of stage a1 + a2 bit for bit."""
    K = 4
    H, W = 21, 1031
    p = capi.default_params(img_h=H, img_w=W, fov_up_deg=15.0, fov_down_deg=-15.0, min_tree_vertices=8,
                            min_cluster_points=30, minTreeModels=3)
    cfg = capi.synth_config(H, W, 40, fov_up_deg=15.0, fov_down_deg=-15.0, sensor_height=1.5, tree_r_max=9.0,
                            max_tilt_deg=1.5)
    N, T, PP = H * W, p.max_trees, p.max_prev_planes
    inp, exp = run_sequence(capi, oracle, p, cfg, K, True)
    ctx = capi.Context(p, K)
    out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32),
               tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32),
               planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32),
               range_image=np.zeros((K, N), np.float32))
    ctx.run_keyframes_host(K, inp, out)
    assert sum(int(e.result["n_landmarks"]) for e in exp) > 0
    for k in range(K):
        compare_keyframe(out["results"][k], out["matches"][k], out["tm"][k], out["tm_id"][k], out["planes"][k],
                         out["n_planes"][k], exp[k])
        assert np.array_equal(out["range_image"][k].view(np.uint32), exp[k].range_image.view(np.uint32))
    it = ctx.intermediates()
    tree = capi.read_dev(it.tree, K * N * 16, ctx.device).view(np.uint32).reshape(K, N, 4)
    ground = capi.read_dev(it.ground, K * N * 16, ctx.device).view(np.uint32).reshape(K, N, 4)
    gcount = capi.read_dev(it.ground_count, K * 4, ctx.device).view(np.int32)
    pix = capi.read_dev(it.pix, K * N * 4, ctx.device).view(np.int32).reshape(K, N)
    for k in range(K):
        e_tree, e_ground = oracle.mask_cloud(p, inp["points"][k], exp[k].pix, inp["mask"][k])
        assert np.array_equal(pix[k], exp[k].pix)
        assert np.array_equal(tree[k], e_tree.view(np.uint32).reshape(N, 4))
        assert gcount[k] == len(e_ground)
        assert np.array_equal(ground[k][:gcount[k]], e_ground.view(np.uint32).reshape(-1, 4))
    ctx.close()


def test_contexts_of_different_image_width_in_one_process(capi, oracle):
    """A context for a narrow image followed by one for a wide image in the same process:
    per-function launch attributes (opt-in shared memory of the wide-row vertex kernel) must
    follow the widest image, in either order (a process-wide "set once" flag did not)."""
    from sloam_b200 import configs
    for order in (("vlp-16", "os1-128"), ("os1-128", "vlp-16", "os1-128")):
        for preset in order:
            K = 2
            p, cfg = configs.make(capi, preset)
            inp, exp = run_sequence(capi, oracle, p, cfg, K, True)
            T, PP = p.max_trees, p.max_prev_planes
            ctx = capi.Context(p, K)
            out = dict(results=np.zeros(K, abi.KF_RESULT), matches=np.zeros((K, T), np.int32),
                       tm=np.zeros((K, T), abi.CYLINDER), tm_id=np.zeros((K, T), np.int32),
                       planes=np.zeros((K, PP), abi.PLANE), n_planes=np.zeros(K, np.int32), range_image=None)
            ctx.run_keyframes_host(K, inp, out)
            for k in range(K):
                compare_keyframe(out["results"][k], out["matches"][k], out["tm"][k], out["tm_id"][k],
                                 out["planes"][k], out["n_planes"][k], exp[k])
            ctx.close()


def test_large_map_association_config5(capi, oracle):
    """configs[4]: 100 000 map cylinders, 2 000 detections per keyframe (split-map path)."""
    rng = np.random.default_rng(55)
    n_map, n_det, K = 100000, 2000, 1
    mp_ = random_cylinders(rng, n_map, 1000.0)
    det = np.zeros((K, n_det), abi.CYLINDER)
    pick = rng.integers(0, n_map, n_det)
    det[0] = mp_[pick]
    det[0]["root"][:, :2] += rng.normal(0, 0.2, (n_det, 2))
    det[0]["root"][:200, :2] += 5000.0                     # 10 % unmatched
    tf = np.zeros(K, abi.POSE); tf["q"][:, 3] = 1.0; tf["t"][0] = (0.1, -0.2, 0.05)
    p = capi.default_params()
    ctx = capi.Context(p, K)
    bi, bd = ctx.associate(capi.to_dev(det), capi.to_dev(np.full(K, n_det, np.int32)), n_det, capi.to_dev(tf),
                           capi.to_dev(mp_), capi.to_dev(np.array([n_map], np.int32)), n_map, True, K)
    ctx.sync()
    bi = capi.to_host(bi, np.int32, (K, n_det)); bd = capi.to_host(bd, np.float64, (K, n_det))
    ei, ed = oracle.associate(det[0], tf[0:1], mp_)
    assert np.array_equal(bi[0], ei) and np.array_equal(bd[0], ed)
    # property at full size: a matched detection's best map entry is the one it was sampled from
    near = bd[0] < 1.0
    assert near.sum() >= 1700 and np.mean(bi[0][near] == pick[near]) > 0.99
    ctx.close()

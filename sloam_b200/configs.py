"""The BASELINE.json configurations as (parameter overrides, synthetic-scene overrides).

Shared by tests/ and bench.py so that the workload named in a bench line is the
workload the parity tests cover.
"""

# name -> (H, W, n_trees requested, sloam_params overrides, sloam_synth_config overrides)
PRESETS = {
    # configs[0]: OS1-64 64x1024, 20 trees + ground plane (the reference's CPU-runnable case)
    "os1-64": (64, 1024, 20, {}, {}),
    # configs[1]: synthetic VLP-16 sequence, 50-tree submap.  16 scan lines cannot satisfy the
    # detector's hard-coded "> 16 vertices" (trellis.cpp:124): the literals are parameters
    # here.  +-15 deg fov, sensor 1.5 m above ground.
    "vlp-16": (16, 1800, 50,
               dict(fov_up_deg=15.0, fov_down_deg=-15.0, min_tree_vertices=8, min_cluster_points=30,
                    minTreeModels=3),
               dict(fov_up_deg=15.0, fov_down_deg=-15.0, sensor_height=1.5, tree_r_max=9.0,
                    max_tilt_deg=1.5)),
    # configs[2]: OS1-64 dense forest, 300 trees, 4096 RANSAC hypotheses per tree (fixed-count
    # stress mode; noisy medians so the hypotheses differ)
    "os1-64-dense": (64, 1024, 300,
                     dict(ransac_fixed_hypotheses=4096, max_trees=512),
                     dict(tree_r_max=22.0, max_tilt_deg=1.0, range_noise=0.03)),
    # configs[3]: OS1-128 2048-column sequence
    "os1-128": (128, 2048, 40, {}, {}),
}


def make(capi, name, **param_overrides):
    """-> (sloam_params, sloam_synth_config)"""
    H, W, n_trees, pk, sk = PRESETS[name]
    kw = dict(img_h=H, img_w=W)
    kw.update(pk)
    kw.update(param_overrides)
    p = capi.default_params(**kw)
    cfg = capi.synth_config(H, W, n_trees, **sk)
    return p, cfg

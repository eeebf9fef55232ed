"""ctypes binding of the C ABI (include/sloam_b200.h -> sloam_b200/lib/libsloam_b200.so).

PyTorch is used only for device memory and streams.  There is no CPU fallback:
if the CUDA library is missing or no B200 is visible, this fails loudly.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsloam_b200.so")
_lib = None

EXPORTS = [
    "sloam_b200_default_params", "sloam_b200_create", "sloam_b200_destroy", "sloam_b200_set_params",
    "sloam_b200_get_params", "sloam_b200_set_stream", "sloam_b200_sync", "sloam_b200_last_error",
    "sloam_b200_kernel_launches", "sloam_b200_workspace_bytes", "sloam_b200_profile_enable", "sloam_b200_set_lanes", "sloam_b200_make_tensor_dev",
    "sloam_b200_mask_from_logits_dev",
    "sloam_b200_profile_read", "sloam_b200_profile_read_kernels", "sloam_b200_version",
    "sloam_b200_project_dev", "sloam_b200_mask_cloud_dev", "sloam_b200_project_split_dev",
    "sloam_b200_ground_planes_dev", "sloam_b200_find_clusters_dev", "sloam_b200_compute_graph_dev",
    "sloam_b200_cylinders_dev", "sloam_b200_associate_dev", "sloam_b200_associate_planes_dev", "sloam_b200_optimize_pose_dev",
    "sloam_b200_run_keyframes_dev", "sloam_b200_run_keyframes_host", "sloam_b200_run_keyframes_host_xyz",
    "sloam_b200_get_intermediates",
    "sloam_b200_run_sloam_dev", "sloam_b200_dev_alloc", "sloam_b200_dev_free", "sloam_b200_copy_h2d",
    "sloam_b200_copy_d2h", "sloam_b200_map_init", "sloam_b200_map_free", "sloam_b200_map_get_submap_dev",
    "sloam_b200_map_update_dev", "sloam_b200_map_dump_host", "sloam_b200_sequence_step_host",
    "sloam_b200_comm_unique_id", "sloam_b200_comm_init", "sloam_b200_comm_destroy", "sloam_b200_comm_size",
    "sloam_b200_comm_rank", "sloam_b200_gather_results_dev", "sloam_b200_comm_wait",
    "sloam_synth_default_config", "sloam_synth_scene", "sloam_synth_pose",
    "sloam_synth_generate_host", "sloam_synth_generate_dev",
]


def lib():
    """Load the CUDA library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C sloam_b200/csrc).  sloam_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.sloam_b200_last_error.restype = C.c_char_p
        _lib.sloam_b200_version.restype = C.c_char_p
        _lib.sloam_b200_kernel_launches.restype = C.c_int64
        _lib.sloam_b200_workspace_bytes.restype = C.c_int64
    return _lib


def default_params(**kw):
    p = abi.Params()
    lib().sloam_b200_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def synth_config(img_h, img_w, n_trees, **kw):
    c = abi.SynthConfig()
    lib().sloam_synth_default_config(C.byref(c), img_h, img_w, n_trees)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def synth_scene(cfg):
    out = np.zeros(max(cfg.n_trees, 1), abi.CYLINDER)
    n = lib().sloam_synth_scene(C.byref(cfg), abi.ptr(out))
    return out[:n].copy()


def synth_pose(cfg, k):
    gt, guess = np.zeros(1, abi.POSE), np.zeros(1, abi.POSE)
    lib().sloam_synth_pose(C.byref(cfg), C.c_int64(k), abi.ptr(gt), abi.ptr(guess))
    return gt[0], guess[0]


def synth_generate_host(cfg, k0, K):
    N = cfg.img_h * cfg.img_w
    pts = np.zeros((K, N), abi.POINT)
    mask = np.zeros((K, N), np.uint8)
    rc = lib().sloam_synth_generate_host(C.byref(cfg), C.c_int64(k0), K, abi.ptr(pts), abi.ptr(mask))
    assert rc == 0
    return pts, mask


# ----------------------------------------------------------------- device side
def _torch():
    import torch
    return torch


def to_dev(a, device="cuda:0"):
    """numpy (possibly structured) array -> uint8 cuda tensor holding its bytes."""
    torch = _torch()
    a = np.ascontiguousarray(a)
    flat = a.reshape(-1).view(np.uint8) if a.size else np.zeros(0, np.uint8)
    return torch.from_numpy(flat.copy()).to(device)


def dev_empty(nbytes, device="cuda:0"):
    return _torch().empty(max(int(nbytes), 16), dtype=_torch().uint8, device=device)


def to_host(t, dtype, shape=None):
    a = t.cpu().numpy().view(np.uint8)
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize if shape is not None else (a.size // np.dtype(dtype).itemsize) * np.dtype(dtype).itemsize
    out = a[:n].view(dtype)
    return out.reshape(shape) if shape is not None else out


def dptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Context:
    """sloam_ctx: bound to one GPU; not thread-safe (like the reference core)."""

    def __init__(self, params, max_keyframes, device=0, use_torch_stream=True):
        self.p = params.copy()
        self.h = C.c_void_p()
        rc = lib().sloam_b200_create(C.byref(self.p), device, max_keyframes, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"sloam_b200_create failed with {rc} (-4 = no usable sm_100 GPU, -1 = bad params)")
        self.device = f"cuda:{device}"
        self.max_k = max_keyframes
        if use_torch_stream:
            torch = _torch()
            self.stream = torch.cuda.Stream(device=self.device)
            self.check(lib().sloam_b200_set_stream(self.h, C.c_void_p(self.stream.cuda_stream)))
        else:
            self.stream = None

    def close(self):
        if self.h:
            lib().sloam_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(f"sloam_b200 error {rc}: {lib().sloam_b200_last_error(self.h).decode()}")

    def sync(self):
        self.check(lib().sloam_b200_sync(self.h))

    def launches(self):
        return lib().sloam_b200_kernel_launches(self.h)

    def set_lanes(self, n):
        """Cut fused runs into n concurrent sub-batches (1..4); results do not change."""
        self.check(lib().sloam_b200_set_lanes(self.h, int(n)))

    def make_tensor(self, K, range_image, mean=12.97, std=12.35):
        """Segmentation::_makeTensor on device tensors -> (tensor, invalid flags, n_invalid)"""
        tensor = dev_empty(K * self.N * 4, self.device)
        invalid = dev_empty(K * self.N, self.device)
        n_inv = dev_empty(K * 4, self.device)
        self.check(lib().sloam_b200_make_tensor_dev(self.h, K, dptr(range_image), C.c_float(mean), C.c_float(std),
                                                    dptr(tensor), dptr(invalid), dptr(n_inv)))
        return tensor, invalid, n_inv

    def mask_from_logits(self, K, logits, invalid=None):
        mask = dev_empty(K * self.N, self.device)
        self.check(lib().sloam_b200_mask_from_logits_dev(self.h, K, dptr(logits), dptr(invalid), dptr(mask)))
        return mask

    def profile_enable(self, on=True):
        self.check(lib().sloam_b200_profile_enable(self.h, int(on)))

    def profile_read(self):
        """-> (summed split-kernel milliseconds, launches) since the last enable/read"""
        ms, n = C.c_double(), C.c_int32()
        self.check(lib().sloam_b200_profile_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_read_kernels(self):
        """-> [(kernel group name, summed milliseconds, runs counted)] since the last enable/read"""
        cap = 32
        buf = (abi.ProfKernel * cap)()
        n = C.c_int32()
        self.check(lib().sloam_b200_profile_read_kernels(self.h, buf, cap, C.byref(n)))
        return [(buf[i].name.decode(), buf[i].ms, buf[i].launches) for i in range(n.value)]

    def workspace_bytes(self):
        return lib().sloam_b200_workspace_bytes(self.h)

    # sizes
    @property
    def N(self):
        return self.p.img_h * self.p.img_w

    @property
    def B(self):
        return self.p.n_cells()

    # ---- stage entries (device tensors in, device tensors out) ----
    def project(self, d_points, K, want_range=True):
        pix = dev_empty(K * self.N * 4, self.device)
        rng = dev_empty(K * self.N * 4, self.device) if want_range else None
        self.check(lib().sloam_b200_project_dev(self.h, K, dptr(d_points), dptr(pix), dptr(rng)))
        return pix, rng

    def mask_cloud(self, d_points, d_pix, d_mask, K):
        tree = dev_empty(K * self.N * 16, self.device)
        ground = dev_empty(K * self.N * 16, self.device)
        cnt = dev_empty(K * 4, self.device)
        self.check(lib().sloam_b200_mask_cloud_dev(self.h, K, dptr(d_points), dptr(d_pix), dptr(d_mask),
                                                   dptr(tree), dptr(ground), dptr(cnt)))
        return tree, ground, cnt

    def project_split(self, d_points, d_mask, K, want_range=True):
        pix = dev_empty(K * self.N * 4, self.device)
        rng = dev_empty(K * self.N * 4, self.device) if want_range else None
        tree = dev_empty(K * self.N * 16, self.device)
        ground = dev_empty(K * self.N * 16, self.device)
        cnt = dev_empty(K * 4, self.device)
        self.check(lib().sloam_b200_project_split_dev(self.h, K, dptr(d_points), dptr(d_mask), dptr(pix),
                                                      dptr(rng), dptr(tree), dptr(ground), dptr(cnt)))
        return pix, rng, tree, ground, cnt

    def ground_planes(self, d_ground, d_count, stride, d_pose, K, want_kept=True):
        Fg = self.p.numGroundFeatures
        cells = dev_empty(K * self.B * abi.CELL_PLANE.itemsize, self.device)
        feats = dev_empty(K * self.B * Fg * 16, self.device)
        kept = dev_empty(K * stride * 16, self.device) if want_kept else None
        offs = dev_empty(K * (self.B + 1) * 4, self.device) if want_kept else None
        self.check(lib().sloam_b200_ground_planes_dev(self.h, K, dptr(d_ground), dptr(d_count), stride,
                                                      dptr(d_pose), dptr(cells), dptr(feats), dptr(kept),
                                                      dptr(offs)))
        return cells, feats, kept, offs

    def find_clusters(self, d_tree, K):
        labels = dev_empty(K * self.N * 4, self.device)
        n = dev_empty(K * 4, self.device)
        self.check(lib().sloam_b200_find_clusters_dev(self.h, K, dptr(d_tree), dptr(labels), dptr(n)))
        return labels, n

    def compute_graph(self, d_tree, K):
        T, V = self.p.max_trees, self.p.max_tree_vertices
        trees = dev_empty(K * T * abi.TREE.itemsize, self.device)
        n = dev_empty(K * 4, self.device)
        verts = dev_empty(K * T * V * abi.VERTEX.itemsize, self.device)
        vpts = dev_empty(K * self.N * 16, self.device)
        self.check(lib().sloam_b200_compute_graph_dev(self.h, K, dptr(d_tree), dptr(trees), dptr(n),
                                                      dptr(verts), dptr(vpts)))
        return trees, n, verts, vpts

    def cylinders(self, d_trees, d_ntrees, d_verts, d_vpts, d_cells, K):
        T, Ft = self.p.max_trees, self.p.featuresPerTree
        models = dev_empty(K * T * abi.TREE_MODEL.itemsize, self.device)
        feats = dev_empty(K * T * Ft * 16, self.device)
        self.check(lib().sloam_b200_cylinders_dev(self.h, K, dptr(d_trees), dptr(d_ntrees), dptr(d_verts),
                                                  dptr(d_vpts), dptr(d_cells), dptr(models), dptr(feats)))
        return models, feats

    def associate(self, d_det, d_ndet, det_stride, d_tf, d_map, d_nmap, map_stride, map_shared, K):
        bi = dev_empty(K * det_stride * 4, self.device)
        bd = dev_empty(K * det_stride * 8, self.device)
        self.check(lib().sloam_b200_associate_dev(self.h, K, dptr(d_det), dptr(d_ndet), det_stride,
                                                  dptr(d_tf), dptr(d_map), dptr(d_nmap), map_stride,
                                                  int(map_shared), dptr(bi), dptr(bd)))
        return bi, bd

    def optimize_pose(self, mode, d_pose, d_tf, d_to, d_nt, tf_stride, d_pf, d_po, d_np, pf_stride, d_ot,
                      d_og, K):
        out = dev_empty(K * abi.POSE.itemsize, self.device)
        it = dev_empty(K * 8, self.device)
        term = dev_empty(K * 8, self.device)
        self.check(lib().sloam_b200_optimize_pose_dev(self.h, K, mode, dptr(d_pose), dptr(d_tf), dptr(d_to),
                                                      dptr(d_nt), tf_stride, dptr(d_pf), dptr(d_po),
                                                      dptr(d_np), pf_stride, dptr(d_ot), dptr(d_og),
                                                      dptr(out), dptr(it), dptr(term)))
        return out, it, term

    # ---- fused path ----
    def alloc_outputs_dev(self, K, want_range=False):
        T, PP = self.p.max_trees, self.p.max_prev_planes
        o = dict(results=dev_empty(K * abi.KF_RESULT.itemsize, self.device),
                 matches=dev_empty(K * T * 4, self.device),
                 tm=dev_empty(K * T * abi.CYLINDER.itemsize, self.device),
                 tm_id=dev_empty(K * T * 4, self.device),
                 planes=dev_empty(K * PP * abi.PLANE.itemsize, self.device),
                 n_planes=dev_empty(K * 4, self.device),
                 range_image=dev_empty(K * self.N * 4, self.device) if want_range else None)
        return o

    def run_keyframes_dev(self, K, inp, out, map_shared=False):
        bi = abi.BatchIn(dptr(inp["points"]), dptr(inp["mask"]), dptr(inp["pose_est"]),
                         dptr(inp["first_scan"]), dptr(inp["map_models"]), dptr(inp["n_map_models"]),
                         int(map_shared), dptr(inp["prev_planes"]), dptr(inp["n_prev_planes"]))
        bo = abi.BatchOut(dptr(out["results"]), dptr(out["matches"]), dptr(out["tm"]), dptr(out["tm_id"]),
                          dptr(out["planes"]), dptr(out["n_planes"]), dptr(out.get("range_image")))
        self.check(lib().sloam_b200_run_keyframes_dev(self.h, K, C.byref(bi), C.byref(bo)))

    def run_keyframes_host(self, K, inp, out, map_shared=False):
        """inp/out: dicts of numpy arrays or pinned torch tensors (host memory)."""
        def hp(x):
            if x is None:
                return None
            if isinstance(x, np.ndarray):
                return abi.ptr(x)
            return C.c_void_p(x.data_ptr())
        bi = abi.BatchIn(hp(inp.get("points")), hp(inp["mask"]), hp(inp["pose_est"]), hp(inp["first_scan"]),
                         hp(inp["map_models"]), hp(inp["n_map_models"]), int(map_shared),
                         hp(inp["prev_planes"]), hp(inp["n_prev_planes"]))
        bo = abi.BatchOut(hp(out["results"]), hp(out["matches"]), hp(out["tm"]), hp(out["tm_id"]),
                          hp(out["planes"]), hp(out["n_planes"]), hp(out.get("range_image")))
        if inp.get("points_xyz") is not None:  # packed x, y, z cloud
            self.check(lib().sloam_b200_run_keyframes_host_xyz(self.h, K, hp(inp["points_xyz"]), C.byref(bi), C.byref(bo)))
        else:
            self.check(lib().sloam_b200_run_keyframes_host(self.h, K, C.byref(bi), C.byref(bo)))

    # ---- multi-GPU gather (comm.cu) ----
    def comm_init(self, rank, world, id_bytes):
        buf = (C.c_char * 128).from_buffer_copy(bytes(id_bytes))
        self.check(lib().sloam_b200_comm_init(self.h, int(rank), int(world), buf))

    def gather_results(self, K, local, gathered):
        """all-gather of the result arrays of this rank's batch (device tensors) into `gathered`
        (dict with the same keys, world x larger); asynchronous, see sloam_b200_gather_results_dev"""
        def bo(o):
            return abi.BatchOut(dptr(o.get("results")), dptr(o.get("matches")), dptr(o.get("tm")), dptr(o.get("tm_id")),
                                None, None, None)
        lo, al = bo(local), bo(gathered)
        self.check(lib().sloam_b200_gather_results_dev(self.h, K, C.byref(lo), C.byref(al)))

    def comm_wait(self):
        self.check(lib().sloam_b200_comm_wait(self.h))

    def intermediates(self):
        it = abi.Intermediates()
        self.check(lib().sloam_b200_get_intermediates(self.h, C.byref(it)))
        return it

    def synth_generate_dev(self, cfg, k0, K):
        pts = dev_empty(K * self.N * 16, self.device)
        mask = dev_empty(K * self.N, self.device)
        self.check(lib().sloam_synth_generate_dev(self.h, C.byref(cfg), C.c_int64(k0), K, dptr(pts), dptr(mask)))
        return pts, mask


def comm_unique_id():
    """ncclUniqueId (128 bytes) for sloam_b200_comm_init: rank 0 creates it, everyone gets a copy"""
    buf = (C.c_char * 128)()
    rc = lib().sloam_b200_comm_unique_id(buf)
    if rc != 0:
        raise RuntimeError(f"sloam_b200_comm_unique_id failed with {rc} (-4: libnccl.so.2 not found)")
    return bytes(buf)


def read_dev(ptr_value, nbytes, device="cuda:0"):
    """Copy nbytes from a raw device pointer (context scratch) to a numpy uint8 array."""
    torch = _torch()
    out = np.empty(nbytes, np.uint8)
    rt = _cudart()  # torch cannot wrap a raw pointer portably: plain cudaMemcpy D2H
    torch.cuda.synchronize(device)
    rc = rt.cudaMemcpy(abi.ptr(out), C.c_void_p(ptr_value), C.c_size_t(nbytes), 2)
    if rc != 0:
        raise RuntimeError(f"cudaMemcpy failed: {rc}")
    return out


_rt = None


def _cudart():
    global _rt
    if _rt is None:
        import glob
        import torch
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*"))
        cands += glob.glob("/usr/local/cuda/lib64/libcudart.so*")
        for cnd in cands:
            try:
                _rt = C.CDLL(cnd)
                break
            except OSError:
                continue
        if _rt is None:
            _rt = C.CDLL("libcudart.so")
    return _rt

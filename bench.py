#!/usr/bin/env python
"""bench.py -- keyframes/s of the SLOAM per-keyframe hot path on B200.

One step = one pass of the whole hot path (projection/split -> ground cells + plane fits
-> tree clustering + vertices -> cylinder models -> association -> LM pose -> projection
-> association) over one batch of synthetic keyframes.  Default workload: the one
BASELINE.json's metric is quoted on, synthetic OS1-64 64x1024 forest scans (configs[0]'s
scene: 20 trees + ground plane), 1024 keyframes per step and GPU (the batch size of
configs[1]; inputs 1.14 GB, far larger than the 126 MB L2; 256 / 512 / 2048 keyframes per
step give 0.78 / 0.93 / 1.04 of its throughput).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload os1-64|vlp-16|os1-64-dense|os1-128|assoc-100k] [--keyframes B]

    os1-64        BASELINE metric workload (default)
    vlp-16        configs[1]: VLP-16 sequence, 1000 keyframes per step
    os1-64-dense  configs[2]: 300 trees, 4096 RANSAC hypotheses per tree (fixed-count mode), 256 keyframes per
                  step (64 / 128 per step: 0.56 / 0.77 of its throughput -- the per-keyframe and per-tree
                  kernels need that many to fill the GPU)
    os1-128       configs[3]: OS1-128 2048 columns (the multi-GPU sequence; here per-GPU batches of 512;
                  128 / 256 per step: 0.81 / 0.94)
    assoc-100k    configs[4]: association only, 100k map cylinders x 2k detections per keyframe

Under torchrun (N > 1) every rank processes its own B keyframes (weak scaling, no data-path
collective); the per-keyframe result records are gathered over NCCL inside the timed region;
time is the max over ranks of CUDA-event time.
`--impl reference` times the CPU oracle (the restated reference path, oracle/) with all host
cores on the SAME keyframes per step -- the reference itself cannot be compiled in this image
(DESIGN.md section 3).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# BASELINE.json "metric", verbatim, for the workload it names
METRIC_OS1_64 = "keyframes/sec (OS1-64 synthetic forest) at 1/2/4/8 B200; kernel HBM GB/s vs peak"
WORKLOADS = {
    # name: (preset in sloam_b200/configs.py, default keyframes per step and GPU, metric, BASELINE config)
    "os1-64": ("os1-64", 1024, METRIC_OS1_64, "metric workload: configs[0]'s OS1-64 64x1024 scene (20 trees + ground), batched"),
    "vlp-16": ("vlp-16", 1000, "keyframes/sec (VLP-16 synthetic forest)", "configs[1]: VLP-16 sequence, 1000 keyframes, 50-tree scene"),
    "os1-64-dense": ("os1-64-dense", 256, "keyframes/sec (OS1-64 dense synthetic forest, 4096 RANSAC hypotheses/tree)",
                     "configs[2]: OS1-64 dense forest, 300 trees, 4096 hypotheses per tree"),
    "os1-128": ("os1-128", 512, "keyframes/sec (OS1-128 2048-column synthetic forest)", "configs[3]: OS1-128 2048-col sequence (per-GPU batch)"),
    "assoc-100k": (None, 8, "keyframes/sec (data association only: 100k map cylinders, 2k detections/keyframe)",
                   "configs[4]: 100k cylinder landmarks, 2k detections per keyframe"),
}
FP64_PEAK_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # B200: 64 FP64 FMA lanes per SM and clock


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[2 + j] == "Active" for r in self.rows if len(r) > 2 + j)]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons}


def committed_traffic(kernel, B, workload):
    """DRAM bytes of one launch of `kernel` from the committed `ncu --set full` capture
    (profiles/kernel_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per keyframe of a
    workload), scaled to the batch; None when no capture of this workload is committed."""
    path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f)
        per_kf = rec.get(workload, {}).get(kernel)
        return float(per_kf) * B if per_kf is not None else None
    except (OSError, ValueError):
        return None


def workload_config(args, p, n_scene):
    """The `config` object: identical for both arms (the driver compares them)."""
    N = p.img_h * p.img_w
    B = args.keyframes
    return {"workload": args.workload, "baseline_config": WORKLOADS[args.workload][3],
            "img_h": p.img_h, "img_w": p.img_w, "keyframes_per_step_per_gpu": B,
            "submap_cylinders": int(n_scene), "two_step": bool(p.twoStepOptim),
            "ransac_fixed_hypotheses": int(p.ransac_fixed_hypotheses),
            "l2": f"inputs {B * N * 17 / 1e6:.0f} MB per step > 126 MB L2, no flush needed"
                  if B * N * 17 > 2 * 126e6 else "inputs smaller than L2 (short run)"}


def host_state(capi, abi, p, cfg, k0, B):
    """Per-keyframe state inputs that do not need the GPU: pose guesses and the submap."""
    M = p.max_map_models
    scene = capi.synth_scene(cfg)
    assert len(scene) <= M
    pose = np.array([capi.synth_pose(cfg, k0 + k)[1] for k in range(B)])
    maps = np.zeros((B, M), abi.CYLINDER)
    maps[:, :len(scene)] = scene
    nmap = np.full(B, len(scene), np.int32)
    return scene, pose, maps, nmap


# diagnostic only (the scaling control in profiles/): a sharded run without its result gather
NO_GATHER = os.environ.get("SLOAM_BENCH_NO_GATHER") is not None
GATHER_ONLY = os.environ.get("SLOAM_BENCH_GATHER_ONLY") is not None  # ... and the gather alone


def make_inputs(capi, abi, ctx, p, cfg, B, k0, device):
    """Device-resident inputs of keyframes [k0, k0+B): generated on the GPU; prevGPlanes_ come
    from an untimed first-scan pass over the same keyframes (planes of keyframe k-1)."""
    PP = p.max_prev_planes
    pts, mask = ctx.synth_generate_dev(cfg, k0, B)
    scene, pose, maps, nmap = host_state(capi, abi, p, cfg, k0, B)
    inp = dict(points=pts, mask=mask, pose_est=capi.to_dev(pose, device),
               first_scan=capi.to_dev(np.ones(B, np.uint8), device),
               map_models=capi.to_dev(maps, device), n_map_models=capi.to_dev(nmap, device),
               prev_planes=capi.to_dev(np.zeros((B, PP), abi.PLANE), device),
               n_prev_planes=capi.to_dev(np.zeros(B, np.int32), device))
    out = ctx.alloc_outputs_dev(B, want_range=os.environ.get("SLOAM_BENCH_NO_RANGE") is None)
    ctx.run_keyframes_dev(B, inp, out)       # untimed: every keyframe as a first scan
    ctx.sync()
    planes = capi.to_host(out["planes"], abi.PLANE, (B, PP))
    npl = capi.to_host(out["n_planes"], np.int32, (B,))
    prev = np.zeros((B, PP), abi.PLANE)
    nprev = np.zeros(B, np.int32)
    prev[1:], nprev[1:] = planes[:-1], npl[:-1]
    first = np.zeros(B, np.uint8)
    first[0] = 1
    inp["prev_planes"] = capi.to_dev(prev, device)
    inp["n_prev_planes"] = capi.to_dev(nprev, device)
    inp["first_scan"] = capi.to_dev(first, device)
    host = dict(pose_est=pose, first_scan=first, map_models=maps, n_map_models=nmap, prev_planes=prev,
                n_prev_planes=nprev)
    return inp, out, host, len(scene)


def run_threads(fn, bounds):
    """fn(lo, hi) on one thread per slice (the oracle releases the GIL inside ctypes calls)."""
    errors = []

    def guarded(lo, hi):
        try:
            fn(int(lo), int(hi))
        except Exception as e:  # a failed thread must not look like a fast one
            errors.append(repr(e))
    th = [threading.Thread(target=guarded, args=(bounds[i], bounds[i + 1])) for i in range(len(bounds) - 1)
          if bounds[i + 1] > bounds[i]]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errors:
        raise SystemExit(f"oracle thread failed: {errors[:1]}")
    return time.perf_counter() - t0


def cpu_reference_arm(args, p, cfg, capi, abi):
    """--impl reference: the CPU restatement of the reference path on all host cores, on the
    same B keyframes per step as the GPU arm."""
    import ctypes as C
    import orc
    cores = os.cpu_count() or 1
    B = args.keyframes
    M, PP = p.max_map_models, p.max_prev_planes
    pts, mask = capi.synth_generate_host(cfg, 0, B)
    scene, pose, maps, nmap = host_state(capi, abi, p, cfg, 0, B)
    prev = np.zeros((B, PP), abi.PLANE)
    nprev = np.zeros(B, np.int32)
    res = np.zeros(B, abi.KF_RESULT)
    bounds = np.linspace(0, B, cores + 1).astype(int)

    # prevGPlanes_: planes of the previous keyframe from a first-scan pass (untimed, threaded)
    planes_of = np.zeros((B, PP), abi.PLANE)
    nplanes_of = np.zeros(B, np.int32)

    def first_pass(lo, hi):
        for k in range(lo, hi):
            o = orc.run_keyframe(p, pts[k], mask[k], pose[k:k + 1], True, maps[k, :0], prev[k, :0])
            planes_of[k] = o.planes
            nplanes_of[k] = o.n_planes
    run_threads(first_pass, bounds)
    prev[1:], nprev[1:] = planes_of[:-1], nplanes_of[:-1]
    first = np.zeros(B, np.uint8)
    first[0] = 1

    def run_slice(lo, hi):
        orc.lib().orc_time_keyframes(C.byref(p), 0, hi - lo, abi.ptr(pts[lo:hi]), abi.ptr(mask[lo:hi]),
                                     abi.ptr(pose[lo:hi]), abi.ptr(first[lo:hi]), abi.ptr(maps[lo:hi]),
                                     abi.ptr(nmap[lo:hi]), M, abi.ptr(prev[lo:hi]), abi.ptr(nprev[lo:hi]), PP,
                                     abi.ptr(res[lo:hi]))
    for _ in range(args.warmup):
        run_threads(run_slice, bounds)
    times = [run_threads(run_slice, bounds) for _ in range(args.steps)]
    if int((res["n_trees"] > 0).sum()) == 0:
        raise SystemExit("reference arm failed: no keyframe produced trees")
    total = sum(times)
    value = B * args.steps / total
    line = {"metric": WORKLOADS[args.workload][2], "value": value, "unit": "keyframes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(args, p, len(scene)),
            "note": "CPU oracle (restated reference path, oracle/); the reference cannot be compiled here",
            "cpu_baseline": {"value": value, "unit": "keyframes/s", "cores": cores, "kind": "port",
                             "sample": f"all {B} keyframes of the step, {cores} threads"},
            "e2e": {"value": value, "unit": "keyframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------ per-kernel byte model
def kernel_model(name, c):
    """Algorithmic bytes of one launch of a kernel group from the batch
    counts c = {BN pixels, T tree-labelled points, G ground points, Gc binned ground points,
    V vertex points, cells, words}; -> (bound, bytes or None, formula)."""
    BN, T, G = c["BN"], c["T"], c["G"]
    m = {
        "project_split_kernel": ("hbm", 21 * BN + BN // 8 + 16 * T + 9 * G,
                                 "16N pts + 1N mask + 4N range + N/8 tree bits + 16 T tree points + 9 G ground records and tags"),
        "range_finalize_kernel": ("hbm", 8 * BN, "4N read + 4N write of the range image"),
        "ground_offsets+ground_scatter_kernel": ("hbm", 17 * G, "9 G records and tags in + 8 G member records out"),
        "ground_cells_kernel<0>": ("latency", 8 * G + 8 * (G // 20), "8 G member records in + 8 B per retained record out"),
        "ground_fit": ("latency", 24 * (G // 20), "8 B record + 16 B point per retained point"),
        "tree_words_kernel": ("hbm", BN // 8, "N/8 tree bits"),
        "cc_init_kernel": ("latency", 20 * T, "16 T tree points + 4 T parents"),
        "cc_flatten_kernel": ("latency", 8 * T, "4 T parents read + written"),
        "cc_rows_kernel": ("hbm", 16 * T + 3 * BN // 8, "16 T tree points + 3 N/8 bit planes"),
        "cc_label_kernel": ("latency", 3 * BN // 8, "3 N/8 bit planes"),
        "vertex_kernel<0>": ("latency", 20 * T + 16 * c.get("V", T), "16 T points + 4 T parents + 16 V vertex points"),
        "vertex_kernel": ("latency", 16 * T + 16 * c.get("V", T), "16 T points + 16 V vertex points"),
    }
    for k, v in m.items():
        if name.startswith(k):
            return v
    return ("latency", None, "kilobyte working set per keyframe: not an HBM kernel")


def bench_assoc(args, capi, abi):
    """configs[4]: association only (a11-a13).  FP64-ALU bound: report DP throughput."""
    import ctypes as C
    import torch
    import orc
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = f"cuda:{local_rank}"
    torch.cuda.set_device(local_rank)
    M, D, K = 100000, 2000, args.keyframes
    rng = np.random.default_rng(20260005)
    mp = np.zeros(M, abi.CYLINDER)
    mp["root"][:, :2] = rng.uniform(-1000, 1000, (M, 2))
    mp["root"][:, 2] = rng.normal(0, 0.3, M)
    tilt = rng.normal(0, 0.04, (M, 2))
    mp["ray"] = np.stack([tilt[:, 0], tilt[:, 1], np.ones(M)], 1)
    mp["ray"] /= np.linalg.norm(mp["ray"], axis=1, keepdims=True)
    mp["radius"] = rng.uniform(0.1, 0.28, M)
    det = np.zeros((K, D), abi.CYLINDER)
    for k in range(K):
        sel = rng.choice(M, D, replace=False)
        det[k] = mp[sel]
        det[k]["root"] += rng.normal(0, 0.2, (D, 3))
        far = rng.random(D) < 0.1  # 10 % unmatched
        det[k]["root"][far, :2] += 500.0
    p = capi.default_params(max_trees=D, max_map_models=M)
    ctx = capi.Context(p, K, device=local_rank)
    d_det, d_map = capi.to_dev(det, device), capi.to_dev(mp, device)
    d_nd, d_nm = capi.to_dev(np.full(K, D, np.int32), device), capi.to_dev(np.array([M], np.int32), device)

    def step():
        return ctx.associate(d_det, d_nd, D, None, d_map, d_nm, M, True, K)
    with torch.cuda.stream(ctx.stream):
        for _ in range(max(args.warmup, 3)):
            bi, bd = step()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = ctx.launches()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ctx.stream)
        for _ in range(args.steps):
            bi, bd = step()
        ev1.record(ctx.stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        launches = ctx.launches() - l0
        sampler.stop_flag = True
        sampler.join(timeout=2)
        # e2e: host detections in, host indices out (the map stays resident: it is the semantic map)
        h_det = torch.from_numpy(det.reshape(-1).view(np.uint8)).pin_memory()
        h_idx = torch.empty(K * D * 4, dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(e2e_steps):
            d_det.copy_(h_det, non_blocking=True)
            bi, bd = step()
            h_idx.copy_(bi, non_blocking=True)
            torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    gi = capi.to_host(bi, np.int32, (K, D))
    # CPU baseline: the oracle on a bounded sample of detections of keyframe 0
    S = 64
    t0 = time.perf_counter()
    ci, cd = orc.associate(det[0, :S], None, mp)
    cpu_s = time.perf_counter() - t0
    agree = int((ci == gi[0, :S]).sum())
    pairs = float(K) * D * M
    flop = pairs * 48.0  # per pair: 3 x (3 sub + 3 mul + 2 add + sqrt) + 2 add + div  (cylinder.cpp:175-194)
    peak, peak_src = load_peaks()
    sec = ms * 1e-3 / args.steps
    bytes_alg = 56.0 * (K * D + M) + 12.0 * K * D
    line = {"metric": WORKLOADS[args.workload][2], "value": K / sec, "unit": "keyframes/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "baseline_config": WORKLOADS[args.workload][3], "map_cylinders": M,
                       "detections_per_keyframe": D, "keyframes_per_step_per_gpu": K,
                       "l2": "map 5.6 MB is L2-resident by design (SURVEY 8d); compute-bound"},
            "clocks": sampler.summary(),
            "e2e": {"value": K * e2e_steps / e2e_s, "unit": "keyframes/s", "h2d_bytes_per_step": int(h_det.numel()),
                    "d2h_bytes_per_step": int(h_idx.numel())},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "assoc_kernel", "achieved": bytes_alg / sec / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": bytes_alg / sec / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "note": "FP64-ALU bound, not HBM (56 (T + M) bytes in): see fp64"},
            "fp64": {"pair_distances_per_step": pairs, "flop_per_pair": 48, "achieved_tflops": flop / sec / 1e12,
                     "peak_tflops": FP64_PEAK_TFLOPS, "frac": flop / sec / 1e12 / FP64_PEAK_TFLOPS,
                     "note": "3 DSQRT per pair are counted as 1 flop each but cost ~10 issue slots"},
            "cpu_baseline": {"value": (S / D) / cpu_s, "unit": "keyframes/s", "cores": 1, "kind": "port",
                             "sample": f"{S} of the {D} detections of keyframe 0 against the 100k map, oracle single thread",
                             "agree_with_gpu": f"{agree}/{S} association indices"}}
    if rank == 0:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="os1-64", choices=sorted(WORKLOADS))
    ap.add_argument("--keyframes", type=int, default=0, help="keyframes per step per GPU (0 = the workload's default)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU time budget of the single-thread oracle baseline")
    ap.add_argument("--lanes", type=int, default=2, help="concurrent sub-batches of a fused run (1..4)")
    ap.add_argument("--no-kernel-profile", action="store_true", help="no per-kernel event pairs in the timed steps")
    ap.add_argument("--total-keyframes", type=int, default=0,
                    help="strong scaling: a fixed sequence of this many keyframes is sharded over the ranks "
                         "(BASELINE configs[3]: --workload os1-128 --total-keyframes 10000); a step processes "
                         "the rank's whole share in batches of --keyframes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.keyframes <= 0:
        args.keyframes = WORKLOADS[args.workload][1]

    # NCCL prints its version banner to stdout at VERSION level: keep stdout to the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # (the banner is printed at WARN too)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from sloam_b200 import abi, capi, configs
    if args.workload == "assoc-100k":
        if args.impl == "reference":
            raise SystemExit("assoc-100k: the CPU oracle is timed inside the default arm (cpu_baseline)")
        if rank == 0:
            bench_assoc(args, capi, abi)
        return
    p, cfg = configs.make(capi, WORKLOADS[args.workload][0], max_trees=128, max_map_models=64)
    if args.workload == "os1-64-dense":
        p.max_trees, p.max_map_models = 512, 512

    if args.impl == "reference":
        if rank == 0:
            cpu_reference_arm(args, p, cfg, capi, abi)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sloam_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    B = args.keyframes
    ctx = capi.Context(p, B, device=local_rank)
    N, T, M, PP = p.img_h * p.img_w, p.max_trees, p.max_map_models, p.max_prev_planes
    # strong scaling: the rank's share of a fixed sequence, resident in HBM, as batches of B keyframes
    share = (args.total_keyframes // world) if args.total_keyframes > 0 else B
    n_batches = max(1, share // B)
    share = n_batches * B
    batches = []
    with torch.cuda.stream(ctx.stream):
        for b in range(n_batches):
            batches.append(make_inputs(capi, abi, ctx, p, cfg, B, rank * share + b * B, device))
    inp, out, host, n_scene = batches[0]
    ctx.sync()
    # the only exchange of the sharded path: all-gather of the per-keyframe results, inside the
    # library (comm.cu: ncclAllGather on a side stream that overlaps the next batch)
    gathered = None
    if world > 1:
        ident = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        ctx.comm_init(rank, world, ident[0])
        gathered = dict(results=capi.dev_empty(world * B * abi.KF_RESULT.itemsize, device),
                        matches=capi.dev_empty(world * B * T * 4, device),
                        tm=capi.dev_empty(world * B * T * abi.CYLINDER.itemsize, device),
                        tm_id=capi.dev_empty(world * B * T * 4, device))

    def step():
        for b_inp, b_out, _, _ in batches:
            if not GATHER_ONLY:
                ctx.run_keyframes_dev(B, b_inp, b_out)
            if world > 1 and not NO_GATHER:
                ctx.gather_results(B, b_out, gathered)

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record(ctx.stream)
        for _ in range(steps):
            fn()
        if world > 1:
            ctx.comm_wait()  # the last gather is part of the step
        ev1.record(ctx.stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.cuda.stream(ctx.stream):
        # label statistics of the batch (for the algorithmic bytes of the kernels): one
        # untimed un-split run, pixel indices from the intermediates
        ctx.set_lanes(1)
        ctx.run_keyframes_dev(B, inp, out)
        ctx.sync()
        it = ctx.intermediates()
        pix = capi.read_dev(it.pix, B * N * 4, device).view(np.int32).reshape(B, N)
        mk_h = capi.to_host(inp["mask"], np.uint8, (B, N))
        lab = np.take_along_axis(mk_h, pix, axis=1)
        n_tree_pts, n_ground_pts = int((lab == 255).sum()), int((lab == 1).sum())
        trees_h = capi.read_dev(it.trees, B * T * abi.TREE.itemsize, device).view(abi.TREE).reshape(B, T)
        ntr_h = capi.read_dev(it.n_trees, B * 4, device).view(np.int32)
        n_vertex_pts = int(sum(int(trees_h[k, :ntr_h[k]]["n_points"].sum()) for k in range(B)))
        del pix, lab
        ctx.set_lanes(args.lanes)
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = ctx.launches()
        ms = timed(step, args.steps)
        launches = ctx.launches() - l0
        # second timed region, same steps, with CUDA event pairs around every kernel group.  The
        # library then runs a step on one lane and one stream so that each pair times its kernels
        # alone (in the production step above, kernels of two lanes x two streams overlap).
        kernels, ms_serial = [], None
        if not args.no_kernel_profile:
            ctx.profile_enable(True)
            step()
            ctx.profile_read_kernels()  # warm-up of the serial schedule, discarded
            ms_serial = timed(step, args.steps)
            kernels = ctx.profile_read_kernels()
            ctx.profile_enable(False)
        sampler.stop_flag = True
        sampler.join(timeout=2)

        # ---- end to end through the host-buffer C-ABI entry (pinned host memory) ----
        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a).reshape(-1).view(np.uint8)).pin_memory()
            return t
        h_in = dict(points=pin(capi.to_host(inp["points"], abi.POINT, (B, N))), mask=pin(mk_h))
        for kname, v in host.items():
            h_in[kname] = pin(v)
        h_out = dict(results=pin(np.zeros(B, abi.KF_RESULT)), matches=pin(np.zeros((B, T), np.int32)),
                     tm=pin(np.zeros((B, T), abi.CYLINDER)), tm_id=pin(np.zeros((B, T), np.int32)),
                     planes=pin(np.zeros((B, PP), abi.PLANE)), n_planes=pin(np.zeros(B, np.int32)),
                     range_image=None)
        h2d = sum(h_in[kname].numel() for kname in h_in)
        d2h = sum(v.numel() for v in h_out.values() if v is not None)

        # the same cloud packed as x, y, z (12 B per point): what a binding hands over when it
        # repacks PCL's 32-byte points anyway (sloam_b200_run_keyframes_host_xyz)
        pts_h = capi.to_host(inp["points"], abi.POINT, (B, N))
        h_xyz = dict(h_in, points=None,
                     points_xyz=pin(np.stack([pts_h["x"], pts_h["y"], pts_h["z"]], axis=-1)))
        h2d_xyz = h2d - h_in["points"].numel() + h_xyz["points_xyz"].numel()
        e2e_steps = max(2, min(args.steps, 5))

        def e2e_run(h_inputs):
            ctx.run_keyframes_host(B, h_inputs, h_out)  # warm-up (staging buffers)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                ctx.run_keyframes_host(B, h_inputs, h_out)
            torch.cuda.synchronize()
            sec = torch.tensor([time.perf_counter() - t0], device=device)
            if world > 1:
                dist.all_reduce(sec, op=dist.ReduceOp.MAX)
            return float(sec.item())
        e2e_xyzi_s = e2e_run(h_in)
        e2e_s = e2e_run(h_xyz)

    # the gather, checked: one more batch, then every slice of the gathered arrays on every rank
    # against a checksum of the buffers of the rank that produced it
    gather_check = None
    if world > 1 and not NO_GATHER:
        import zlib
        with torch.cuda.stream(ctx.stream):
            ctx.run_keyframes_dev(B, inp, out)
            ctx.gather_results(B, out, gathered)
            ctx.comm_wait()
        ctx.sync()
        names = ("results", "matches", "tm", "tm_id")
        mine = {k: zlib.crc32(capi.to_host(out[k], np.uint8).tobytes()) for k in names}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        good = 0
        for k in names:
            buf = capi.to_host(gathered[k], np.uint8).reshape(world, -1)
            good += sum(int(zlib.crc32(buf[r].tobytes()) == everyone[r][k]) for r in range(world))
        worst = torch.tensor([good], device=device)
        dist.all_reduce(worst, op=dist.ReduceOp.MIN)
        gather_check = f"{int(worst.item())}/{len(names) * world} gathered slices byte-identical to the producing rank's buffers (min over ranks)"
    res = capi.to_host(out["results"], abi.KF_RESULT, (B,))
    peak, peak_src = load_peaks()
    value = world * share * args.steps / (ms * 1e-3)
    step_ms = ms / args.steps

    # ---- per-kernel table: event-pair time inside the timed steps, algorithmic bytes, fraction
    counts = {"BN": B * N, "T": n_tree_pts, "G": n_ground_pts, "V": n_vertex_pts}
    table, ksum = [], 0.0
    for name, tot_ms, runs in kernels:
        k_ms = tot_ms / max(runs, 1)
        bound, nbytes, formula = kernel_model(name, counts)
        ksum += k_ms
        row = {"kernel": name, "ms": k_ms, "bound": bound, "algorithmic_bytes": nbytes, "bytes_model": formula,
               "GBps": (nbytes / (k_ms * 1e-3) / 1e9) if nbytes and k_ms > 0 else None}
        row["frac"] = row["GBps"] / peak if row["GBps"] is not None else None
        table.append(row)
    for row in table:
        row["share_of_serial_step"] = row["ms"] / (ms_serial / args.steps / n_batches) if ms_serial else None
    # whole path, unfused = 61 N + 32 G per keyframe; ours = what the kernels above move by design
    survey_bytes = n_batches * (61.0 * B * N + 32.0 * n_ground_pts)
    own_bytes = n_batches * float(sum(r["algorithmic_bytes"] for r in table if r["algorithmic_bytes"]))
    dominant = max(table, key=lambda r: r["ms"]) if table else None
    if dominant is not None and dominant["algorithmic_bytes"]:
        roof = {"bound": "hbm", "kernel": dominant["kernel"], "achieved": dominant["GBps"], "peak": peak, "unit": "GB/s",
                "frac": dominant["frac"], "traffic": committed_traffic(dominant["kernel"], B, args.workload),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dominant["algorithmic_bytes"],
                "ms_per_launch": dominant["ms"], "share_of_step": dominant["ms"] / (ms_serial / args.steps / n_batches),
                "timed": "cudaEvent pairs around each kernel in a second timed region of the same steps run on one "
                         "lane and one stream (serial_ms_per_step); the kernel with the largest time is reported"}
    else:
        roof = {"bound": "hbm", "kernel": dominant["kernel"] if dominant else None, "achieved": None, "peak": peak,
                "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src}
    roof.update({"whole_step_frac": survey_bytes / (step_ms * 1e-3) / 1e9 / peak,
                 "whole_step_bytes_model": "SURVEY 8(d): 61 N + 32 G per keyframe (unfused path)",
                 "whole_step_frac_own_bytes": own_bytes / (step_ms * 1e-3) / 1e9 / peak,
                 "tree_points": n_tree_pts, "ground_points": n_ground_pts, "vertex_points": n_vertex_pts})

    line = {
        "metric": WORKLOADS[args.workload][2], "value": value, "unit": "keyframes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "strong" if args.total_keyframes > 0 else "weak", "vs_baseline": None, "dtype": "f32/f64",
        "data": "synthetic",
        "config": workload_config(args, p, n_scene),
        "run": {"lanes": args.lanes, "keyframes_per_step_per_gpu": share, "batches_per_step": n_batches,
                "total_keyframes": world * share,
                "gather": "sloam_b200_gather_results_dev (ncclAllGather, side stream)" if world > 1 else None,
                "gather_check": gather_check,
                "keyframes_ok": int((res["success"] == 1).sum()),
                "mean_landmarks": float(res["n_landmarks"].mean()),
                "lm_converged": int((res["lm_termination"][:, 0] == 0).sum()),
                "serial_ms_per_step": (ms_serial / args.steps) if ms_serial else None,
                "kernel_sum_ms": ksum},
        "clocks": sampler.summary(),
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "keyframes/s", "h2d_bytes_per_step": int(h2d_xyz),
                "d2h_bytes_per_step": int(d2h), "entry": "sloam_b200_run_keyframes_host_xyz (x, y, z cloud, 12 B/point)",
                "xyzi_entry": {"value": world * B * e2e_steps / e2e_xyzi_s, "h2d_bytes_per_step": int(h2d),
                               "entry": "sloam_b200_run_keyframes_host (x, y, z, intensity, 16 B/point)"}},
        "gpu_launches": int(launches),
        "roofline": roof,
        "kernels": table,
    }
    if rank == 0 and world == 1:
        # CPU baseline: the oracle, single thread, on a bounded sample of the same keyframes
        import ctypes as C
        import orc
        pts = capi.to_host(inp["points"], abi.POINT, (B, N))
        cres = np.zeros(B, abi.KF_RESULT)

        def oracle(lo, hi):
            return orc.lib().orc_time_keyframes(
                C.byref(p), 0, hi - lo, abi.ptr(pts[lo:hi]), abi.ptr(mk_h[lo:hi]), abi.ptr(host["pose_est"][lo:hi]),
                abi.ptr(host["first_scan"][lo:hi]), abi.ptr(host["map_models"][lo:hi]),
                abi.ptr(host["n_map_models"][lo:hi]), M, abi.ptr(host["prev_planes"][lo:hi]),
                abi.ptr(host["n_prev_planes"][lo:hi]), PP, abi.ptr(cres[lo:hi]))
        probe = min(8, B)
        secs = oracle(0, probe)
        S = int(max(probe, min(B, args.cpu_seconds / max(secs / probe, 1e-9))))
        secs = oracle(0, S)
        ints = ("status", "success", "n_ground", "n_planes", "n_trees", "n_landmarks", "n_tree_matches",
                "n_plane_matches", "lm_iterations", "lm_termination")
        agree = 0
        for k in range(S):
            ok = all(np.array_equal(cres[k][f], res[k][f]) for f in ints)
            ok = ok and np.max(np.abs(cres[k]["T_Map_Curr"]["t"] - res[k]["T_Map_Curr"]["t"])) <= 1e-5
            ok = ok and np.max(np.abs(cres[k]["T_Map_Curr"]["q"] - res[k]["T_Map_Curr"]["q"])) <= 1e-5
            agree += int(ok)
        line["cpu_baseline"] = {"value": S / secs, "unit": "keyframes/s", "cores": 1, "kind": "port",
                                "sample": f"first {S} keyframes of the same batch ({secs:.1f} s), oracle single thread",
                                "agree_with_gpu": f"{agree}/{S} keyframes (all integer result fields + pose within 1e-5)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

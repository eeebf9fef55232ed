#!/usr/bin/env python
"""bench.py -- keyframes/s of the SLOAM per-keyframe hot path on B200.

One step = one pass of the whole hot path (projection/split -> ground cells + plane fits
-> tree clustering + vertices -> cylinder models -> association -> LM pose -> projection
-> association) over one batch of synthetic keyframes.  Workload at N=1: BASELINE.json
configs[1] (synthetic VLP-16 sequence, 1000 keyframes, 50-tree submap).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload vlp-16|os1-64|os1-128|os1-64-dense] [--keyframes B]

Under torchrun (N > 1) every rank processes its own B keyframes (weak scaling, no data-path
collective); the per-keyframe result records are all-gathered over NCCL inside the timed
region; time is the max over ranks of CUDA-event time.
`--impl reference` times the CPU oracle (the restated reference path, oracle/) on all host
cores -- the reference itself cannot be compiled in this image.
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

METRIC = "keyframes/sec (synthetic forest, whole per-keyframe hot path)"


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


def k1_traffic(B, workload):
    """DRAM bytes of one split-kernel launch from the committed `ncu --set full` capture
    (profiles/k1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per keyframe of
    this workload), scaled to the batch; None when no capture is committed."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "k1_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f)
        return float(rec["dram_bytes_per_keyframe"]) * B if rec.get("workload") == workload else None
    except (OSError, KeyError, ValueError):
        return None


def make_inputs(capi, abi, ctx, p, cfg, B, k0, device):
    """Device-resident inputs of keyframes [k0, k0+B): generated on the GPU; prevGPlanes_ come
    from an untimed first-scan pass over the same keyframes (planes of keyframe k-1)."""
    import torch
    N, M, PP = p.img_h * p.img_w, p.max_map_models, p.max_prev_planes
    pts, mask = ctx.synth_generate_dev(cfg, k0, B)
    scene = capi.synth_scene(cfg)
    assert len(scene) <= M
    pose = np.array([capi.synth_pose(cfg, k0 + k)[1] for k in range(B)])
    maps = np.zeros((B, M), abi.CYLINDER)
    maps[:, :len(scene)] = scene
    nmap = np.full(B, len(scene), np.int32)
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
    return inp, out, host


def cpu_reference_arm(args, p, cfg, capi, abi):
    """--impl reference: the CPU restatement of the reference path on all host cores."""
    import orc
    cores = os.cpu_count() or 1
    sample = min(args.keyframes, 32 * cores)
    N, M, PP = p.img_h * p.img_w, p.max_map_models, p.max_prev_planes
    pts, mask = capi.synth_generate_host(cfg, 0, sample)
    scene = capi.synth_scene(cfg)
    pose = np.array([capi.synth_pose(cfg, k)[1] for k in range(sample)])
    maps = np.zeros((sample, M), abi.CYLINDER); maps[:, :len(scene)] = scene
    nmap = np.full(sample, len(scene), np.int32)
    prev = np.zeros((sample, PP), abi.PLANE); nprev = np.zeros(sample, np.int32)
    first = np.ones(sample, np.uint8)
    res = np.zeros(sample, abi.KF_RESULT)
    import ctypes as C

    errors = []

    def run_slice(lo, hi, first_arr):
        lo, hi = int(lo), int(hi)
        try:
            _run_slice(lo, hi, first_arr)
        except Exception as e:  # a failed thread must not look like a fast one
            errors.append(repr(e))

    def _run_slice(lo, hi, first_arr):
        orc.lib().orc_time_keyframes(C.byref(p), 0, hi - lo, abi.ptr(pts[lo:hi]), abi.ptr(mask[lo:hi]),
                                     abi.ptr(pose[lo:hi]), abi.ptr(first_arr[lo:hi]), abi.ptr(maps[lo:hi]),
                                     abi.ptr(nmap[lo:hi]), M, abi.ptr(prev[lo:hi]), abi.ptr(nprev[lo:hi]), PP,
                                     abi.ptr(res[lo:hi]))
    # prevGPlanes_: planes of the previous keyframe from a first-scan pass (untimed)
    for k in range(sample):
        o = orc.run_keyframe(p, pts[k], mask[k], pose[k:k + 1], True, maps[k, :0], prev[k, :0])
        if k + 1 < sample:
            prev[k + 1] = o.planes; nprev[k + 1] = o.n_planes
    first = np.zeros(sample, np.uint8); first[0] = 1
    bounds = np.linspace(0, sample, cores + 1).astype(int)

    def step():
        th = [threading.Thread(target=run_slice, args=(bounds[i], bounds[i + 1], first)) for i in range(cores)
              if bounds[i + 1] > bounds[i]]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0
    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    if errors or int((res["n_trees"] > 0).sum()) == 0:
        raise SystemExit(f"reference arm failed: {errors[:1]} (keyframes with trees: {int((res['n_trees'] > 0).sum())})")
    total = sum(times)
    value = sample * args.steps / total
    line = {"metric": METRIC, "value": value, "unit": "keyframes/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": args.workload, "keyframes_per_step": int(sample),
                       "note": "CPU oracle (restated reference path); the reference cannot be compiled here"},
            "cpu_baseline": {"value": value, "unit": "keyframes/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} keyframes of {args.workload} per step, {cores} threads"},
            "e2e": {"value": value, "unit": "keyframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vlp-16")
    ap.add_argument("--keyframes", type=int, default=1000, help="keyframes per step per GPU")
    ap.add_argument("--cpu-sample", type=int, default=512)
    ap.add_argument("--lanes", type=int, default=2, help="concurrent sub-batches of a fused run (1..4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # NCCL prints its version banner to stdout at VERSION level: keep stdout to the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from sloam_b200 import abi, capi, configs
    p, cfg = configs.make(capi, args.workload, max_trees=128, max_map_models=64)
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
    with torch.cuda.stream(ctx.stream):
        inp, out, host = make_inputs(capi, abi, ctx, p, cfg, B, rank * B, device)
    ctx.sync()
    gather_buf = None
    if world > 1:
        gather_buf = torch.empty(world * out["results"].numel(), dtype=torch.uint8, device=device)

    def step():
        ctx.run_keyframes_dev(B, inp, out)
        if world > 1:  # gather the per-keyframe result records (north star: NCCL only for this)
            dist.all_gather_into_tensor(gather_buf, out["results"])

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record(ctx.stream)
        for _ in range(steps):
            fn()
        ev1.record(ctx.stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.cuda.stream(ctx.stream):
        # label statistics of the batch (for the algorithmic bytes of the split kernel): one
        # untimed un-split run, pixel indices from the intermediates
        ctx.set_lanes(1)
        step()
        ctx.sync()
        it = ctx.intermediates()
        pix = capi.read_dev(it.pix, B * N * 4, device).view(np.int32).reshape(B, N)
        mk_h = capi.to_host(inp["mask"], np.uint8, (B, N))
        lab = np.take_along_axis(mk_h, pix, axis=1)
        n_tree_pts, n_ground_pts = int((lab == 255).sum()), int((lab == 1).sum())
        del pix, lab
        ctx.set_lanes(args.lanes)
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = ctx.launches()
        ctx.profile_enable(True)  # CUDA events around the split kernel of every fused run
        ms = timed(step, args.steps)
        k1_total_ms, k1_n = ctx.profile_read()
        ctx.profile_enable(False)
        launches = ctx.launches() - l0
        sampler.stop_flag = True
        sampler.join(timeout=2)

        # ---- roofline kernel: the fused project/split pass (K1), timed INSIDE the steps above
        # by the library's own event pairs on the launching stream (sloam_b200_profile_*).
        # Algorithmic bytes per launch: every point is read once (16 B +
        # 1 B mask), its pixel index (4 B) and range-image entry (4 B) are written, plus 16 B per
        # tree-labelled point, 1 bit per pixel of tree mask, 17 B per ground point (point + cell).
        k1_ms = k1_total_ms / max(k1_n, 1)
        k1_bytes = float(B * N * (16 + 1 + 4 + 4) + 16 * n_tree_pts + B * N // 8 + 17 * n_ground_pts)

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

        def e2e_step():
            ctx.run_keyframes_host(B, h_in, h_out)
        e2e_steps = max(2, min(args.steps, 5))
        e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=device)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_s = float(e2e_s.item())

    res = capi.to_host(out["results"], abi.KF_RESULT, (B,))
    peak, peak_src = load_peaks()
    value = world * B * args.steps / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "keyframes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": args.workload, "img_h": p.img_h, "img_w": p.img_w, "keyframes_per_step_per_gpu": B,
                   "submap_cylinders": int(capi.to_host(inp["n_map_models"], np.int32, (B,))[0]),
                   "two_step": bool(p.twoStepOptim), "lanes": args.lanes,
                   "l2": f"inputs {B * N * 17 / 1e6:.0f} MB per step > 126 MB L2, no flush needed"
                         if B * N * 17 > 2 * 126e6 else "inputs smaller than L2 (short run)",
                   "keyframes_ok": int((res["success"] == 1).sum()),
                   "mean_landmarks": float(res["n_landmarks"].mean()),
                   "lm_converged": int((res["lm_termination"][:, 0] == 0).sum())},
        "clocks": sampler.summary(),
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "keyframes/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "project_split_kernel<true,true>",
                     "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / peak, "traffic": k1_traffic(B, args.workload), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms,
                     "launches_timed": k1_n, "timed": "cudaEvent pairs around the kernel inside the timed steps",
                     "tree_points": n_tree_pts, "ground_points": n_ground_pts,
                     "share_of_step": k1_ms / (ms / args.steps)},
    }
    if rank == 0 and world == 1:
        # CPU baseline: the oracle, single thread, on a bounded sample of the same keyframes
        import ctypes as C
        import orc
        S = min(args.cpu_sample, B)
        pts = np.ascontiguousarray(capi.to_host(inp["points"], abi.POINT, (B, N))[:S])
        mk = np.ascontiguousarray(capi.to_host(inp["mask"], np.uint8, (B, N))[:S])
        cres = np.zeros(S, abi.KF_RESULT)
        secs = orc.lib().orc_time_keyframes(
            C.byref(p), 0, S, abi.ptr(pts), abi.ptr(mk), abi.ptr(host["pose_est"]), abi.ptr(host["first_scan"]),
            abi.ptr(host["map_models"]), abi.ptr(host["n_map_models"]), M, abi.ptr(host["prev_planes"]),
            abi.ptr(host["n_prev_planes"]), PP, abi.ptr(cres))
        agree = int(sum(int(cres[k]["n_landmarks"] == res[k]["n_landmarks"] and
                            cres[k]["n_tree_matches"] == res[k]["n_tree_matches"]) for k in range(S)))
        line["cpu_baseline"] = {"value": S / secs, "unit": "keyframes/s", "cores": 1, "kind": "port",
                                "sample": f"first {S} keyframes of the same batch, oracle single thread",
                                "agree_with_gpu": f"{agree}/{S} keyframes (landmark + match counts)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

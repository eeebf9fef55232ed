"""Copies what scripts/profile_round.sh left in gpurun_out/ into profiles/ (bench lines, launch
list, ncu tables and hot lines; the .ncu-rep files stay in gpurun_out/, they are too large for
the history) and writes profiles/<tag>_summary.md from them.
usage: python scripts/round_summary.py <tag> [scaling_tag]"""
import glob, json, os, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
stag = sys.argv[2] if len(sys.argv) > 2 else None
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def line(path):
    txt = open(path).read().strip().splitlines()
    return json.loads(txt[-1]) if txt else None


for f in sorted(glob.glob(os.path.join(G, tag + "_*"))):
    if f.endswith((".ncu-rep", ".err", ".log")):
        continue
    shutil.copy(f, P)
out = [f"# {tag}: measured state of the build (B200, one box, `scripts/profile_round.sh {tag}`)", ""]
out += ["## Bench lines (`bench.py --workload W`, defaults)", "",
        "| workload | keyframes/step | ms/step | keyframes/s (device-resident) | end to end, host buffers | oracle, 1 thread | "
        "oracle, 16 threads (`--impl reference`) | dominant kernel: frac of HBM peak | whole step: frac |",
        "|---|---|---|---|---|---|---|---|---|"]
for w in ["os1-64", "vlp-16", "os1-64-dense", "os1-128", "assoc-100k"]:
    p = os.path.join(G, f"{tag}_bench_{w}.json")
    if not os.path.exists(p):
        continue
    d = line(p)
    ref = os.path.join(G, f"{tag}_bench_{w}_reference.json")
    refv = f'{line(ref)["value"]:.0f}' if os.path.exists(ref) else "-"
    r = d.get("roofline", {})
    frac = f'{r.get("kernel", "")} {r["frac"]:.3f}' if r.get("frac") and r["frac"] > 1e-3 else \
        (f'fp64 {d["fp64"]["frac"]:.3f} of {d["fp64"]["peak_tflops"]:.1f} TFLOP/s' if "fp64" in d else "-")
    ws = r.get("whole_step_frac")
    out.append(f'| {w} | {d["config"]["keyframes_per_step_per_gpu"]} | {d["ms_per_step"]:.3f} | {d["value"]:.0f} | '
               f'{d["e2e"]["value"]:.0f} | {d["cpu_baseline"]["value"]:.1f} | {refv} | {frac} | {"%.3f" % ws if ws else "-"} |')
d = line(os.path.join(G, f"{tag}_bench_os1-64.json"))
out += ["", f'Clocks during the default run: {d["clocks"]}; kernels launched in the timed region: {d["gpu_launches"]}; '
        f'end to end: {d["e2e"]}.', "",
        "## Per-kernel table of the default workload (CUDA-event pairs, serial schedule: one lane, one stream)", "",
        f'Production step (2 lanes x 2 streams): {d["ms_per_step"]:.3f} ms; serial step: {d["run"]["serial_ms_per_step"]:.3f} ms; '
        f'sum of the kernel groups: {d["run"]["kernel_sum_ms"]:.3f} ms.', "",
        "| kernel group | us per launch | bound | algorithmic bytes | GB/s | frac of 6551.7 GB/s | share of the serial step |",
        "|---|---|---|---|---|---|---|"]
for k in d["kernels"]:
    gb = f'{k["GBps"]:.0f}' if k.get("GBps") else "-"
    fr = f'{k["frac"]:.3f}' if k.get("frac") else "-"
    ab = f'{k["algorithmic_bytes"] / 1e6:.1f} MB' if k.get("algorithmic_bytes") else "-"
    out.append(f'| `{k["kernel"]}` | {k["ms"] * 1e3:.1f} | {k["bound"]} | {ab} | {gb} | {fr} | {k["share_of_serial_step"]:.3f} |')
out += ["", "Byte models: `bench.py::kernel_model` (DESIGN.md section 4 states them per kernel).", ""]
ls = os.path.join(G, f"{tag}_launch_summary.txt")
if os.path.exists(ls):
    out += ["## ncu launch list of the same command (`--metrics gpu__time_duration.sum --clock-control none`)", "",
            f"`profiles/{tag}_launches_os1-64_k1024.csv`; largest launch per kernel (cold cache, serialised -- shares, not absolutes; "
            "`synth_*` generate the input, `project_split<0,1,0>` / `<1,0,0>`, `tree_fill`, `ground_compact` belong to the "
            "un-fused intermediates call that `bench.py` uses once for its label statistics):", "", "```"]
    out += open(ls).read().rstrip().splitlines() + ["```", ""]
for r, title in (("top", "OS1-64, 1024 keyframes, one lane"), ("dense", "OS1-64 dense forest (configs[2])"),
                 ("assoc", "association only (configs[4])")):
    p = os.path.join(G, f"{tag}_ncu_{r}.md")
    if os.path.exists(p):
        out += [f"## `ncu --set full --clock-control none`: {title}", ""] + open(p).read().rstrip().splitlines() + [""]
# DRAM traffic per keyframe of the kernel groups bench.py names, for its roofline.traffic
rep = os.path.join(G, f"{tag}_top.ncu-rep")
if os.path.exists(rep):
    import csv, io, subprocess
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    ik, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ur, uw = unit[rows[1][ir]], unit[rows[1][iw]]
    kf = d["config"]["keyframes_per_step_per_gpu"]
    per = {}
    for r in rows[2:]:
        name = r[ik].split("(")[0].strip()
        per[name[5:] if name.startswith("void ") else name] = (float(r[ir]) * ur + float(r[iw]) * uw) / kf
    groups = {"project_split_kernel": ["project_split_kernel<1, 1, 1>"], "range_finalize_kernel": ["range_finalize_kernel"],
              "cc_rows_kernel": ["cc_rows_kernel"], "cc_label_kernel": ["cc_label_kernel"], "vertex_kernel": ["vertex_kernel"],
              "ground_offsets+ground_scatter_kernel": ["ground_scatter_kernel"],
              "ground_cells_kernel<0>": ["ground_cells_kernel<0>"],
              "ground_cells_kernel<1> (tie replay)": ["ground_cells_kernel<1>"],
              "ground_fit+plane_finish+planes_compact": ["ground_fit_kernel"], "cylinder_kernel+compact": ["cylinder_kernel"],
              "lm_kernel": ["lm_kernel"], "build_matches_kernel": ["build_matches_kernel"]}
    tr = {g: sum(per[k] for k in ks) for g, ks in groups.items() if all(k in per for k in ks)}
    json.dump({"os1-64": tr, "source": f"gpurun_out/{tag}_top.ncu-rep (table: profiles/{tag}_ncu_top.md): ncu --set full "
               f"--clock-control none, bench.py --lanes 1 ({kf} OS1-64 keyframes per launch); "
               "dram__bytes_read.sum + dram__bytes_write.sum per keyframe"},
              open(os.path.join(P, "kernel_traffic.json"), "w"), indent=1)
    out += ["## DRAM traffic per keyframe against the algorithmic bytes (`profiles/kernel_traffic.json`)", "",
            "| kernel group | ncu DRAM bytes / keyframe | algorithmic bytes / keyframe | ratio |", "|---|---|---|---|"]
    alg = {k["kernel"]: k.get("algorithmic_bytes") for k in d["kernels"]}
    for g, v in tr.items():
        a = alg.get(g)
        out.append(f"| `{g}` | {v / 1e3:.1f} KB | {a / kf / 1e3:.1f} KB | {v / (a / kf):.2f} |" if a else f"| `{g}` | {v / 1e3:.1f} KB | - | - |")
    out.append("")
if stag:
    out += ["## Scaling (one box with 8 B200, one process per GPU, NCCL all-gather of the results inside the step)", "",
            "| run | GPUs | keyframes/s | ms/step | end to end keyframes/s |", "|---|---|---|---|---|"]
    for f in sorted(glob.glob(os.path.join(G, stag + "_*_n*.json"))):
        try:
            s = line(f)
            out.append(f'| {os.path.basename(f)[len(stag) + 1:-5]} ({s["scaling"]}) | {s["n_gpus"]} | {s["value"]:.0f} | '
                       f'{s["ms_per_step"]:.3f} | {s["e2e"]["value"]:.0f} |')
            shutil.copy(f, P)
        except Exception as e:  # a failed run has no JSON line
            out.append(f"| {os.path.basename(f)} | - | failed: {e} | | |")
    out.append("")
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

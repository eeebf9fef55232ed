"""Top CUDA source lines of one kernel in an .ncu-rep by warp-stall samples.
usage: python scripts/ncu_hot.py report.ncu-rep kernel_regex [n_lines]"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
by_inst = len(sys.argv) > 4 and sys.argv[4] == "inst"  # rank by instructions executed instead of stall samples
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.defaultdict(lambda: [0.0, collections.Counter(), ""])
fname, h = "", None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; si = h.index("Instructions Executed" if by_inst else "# Samples"); stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]; continue
    if h is None or len(r) != len(h) or not r[0].isdigit(): continue
    key = (fname, int(r[0]))
    a = agg[key]
    s = float(r[si] or 0)
    a[0] += s
    a[2] = r[1]
    for i in stalls:
        v = float(r[i] or 0)
        if v: a[1][h[i][6:]] += v
tot = sum(a[0] for a in agg.values())
print("total", "warp instructions" if by_inst else "samples", tot)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{100*a[0]/tot:5.1f}%  {key[0][:18]:18s}:{key[1]:<4d} {a[2].strip()[:95]:95s} " + " ".join(f"{k}:{int(v)}" for k, v in a[1].most_common(3)))

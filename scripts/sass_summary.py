"""Instruction-class counts per kernel from `cuobjdump -sass` of the built library
-> profiles/<tag>_sass_summary.md (+ the gzipped dump).
usage: python scripts/sass_summary.py <tag>            (e.g. r1)"""
import gzip, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
so = os.path.join(ROOT, "sloam_b200", "lib", "libsloam_b200.so")
dump = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
with gzip.open(os.path.join(ROOT, "profiles", f"{tag}_sass_libsloam_b200.txt.gz"), "wt") as f:
    f.write(dump)
CLASSES = [("LDG.E.128", r"^LDG\.E\.(ENL2\.)?128|^LDG\.E\.128"), ("STG.E.128", r"^STG\.E\.(ENL2\.)?128|^STG\.E\.128"),
           ("ATOMG/REDG (global)", r"^(ATOMG|REDG)"), ("ATOMS (smem)", r"^ATOMS"),
           ("UBLKCP (TMA bulk)", r"^UBLKCP"), ("SYNCS (mbarrier)", r"^SYNCS"),
           ("DADD/DMUL/DFMA", r"^(DADD|DMUL|DFMA)"), ("MUFU", r"^MUFU"), ("BAR", r"^BAR"),
           ("SHFL/VOTE/MATCH", r"^(SHFL|VOTE|MATCH)"), ("HMMA/UTCMMA", r"^(HMMA|UTC.*MMA|IMMA)")]
rows, name, counts, total = [], None, None, 0
ins = re.compile(r"^\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
for line in dump.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name: rows.append((name, total, counts))
        name, counts, total = m.group(1), [0] * len(CLASSES), 0
        continue
    m = ins.match(line)
    if m and name:
        total += 1
        for i, (_, pat) in enumerate(CLASSES):
            if re.search(pat, m.group(1)): counts[i] += 1
if name: rows.append((name, total, counts))
out = [f"# SASS evidence (cuobjdump -sass sloam_b200/lib/libsloam_b200.so, sm_100a)\n",
       f"Full dump: `{tag}_sass_libsloam_b200.txt.gz`; regenerate with `python scripts/sass_summary.py {tag}`.",
       "Instruction-class counts per kernel:\n",
       "| kernel | instrs | " + " | ".join(c for c, _ in CLASSES) + " |", "|---|---|" + "---|" * len(CLASSES)]
for name, total, counts in rows:
    short = re.sub(r"^_ZN2sb\d+", "", name)[:44]
    out.append(f"| `{short}` | {total} | " + " | ".join(str(c) for c in counts) + " |")
out += ["", "`UBLKCP.S.G` + `SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK.TRANS64.TRYWAIT` in `cylinder_kernel` are the "
        "`cp.async.bulk` + mbarrier staging of a tree's vertex records (csrc/k4_cylinder.cu).",
        "No `HMMA`/`UTC*MMA`: nothing on this path is a dense contraction (DESIGN.md section 4)."]
open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))

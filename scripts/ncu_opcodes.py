"""Dynamic opcode census of one kernel launch in an .ncu-rep (executed warp instructions and stall samples per SASS opcode).
usage: python scripts/ncu_opcodes.py REPORT.ncu-rep KERNEL_REGEX [launch_skip]"""
import collections, csv, re, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
si, ei, sm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ex, sa = collections.Counter(), collections.Counter()
for r in rows:
    if len(r) <= ei or not r[ei].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[si])
    if not m:
        continue
    full = m.group(2); op = full.split(".")[0]
    key = ".".join(full.split(".")[:2]) if op in ("MUFU", "SHFL", "LD", "ST") else op
    ex[key] += int(r[ei]); sa[key] += int(r[sm])
tot, tots = sum(ex.values()), sum(sa.values())
print(f"{kre}: executed warp instructions {tot:.4g}, stall samples {tots}")
for k, v in ex.most_common(30):
    print(f"  {k:14s} executed {100 * v / tot:5.1f} %   samples {100 * sa[k] / max(1, tots):5.1f} %")

"""Attribute the executed warp instructions of one kernel in an .ncu-rep to device functions and source lines.

usage: python scripts/ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [--lines N] [--id K | --name DEMANGLED_SUBSTRING]

Joins the SASS page of the report (instruction address, executed count, stall samples) with `nvdisasm --print-line-info`
of the cubin inside LIB.so (function labels of the __noinline__ device functions, `//## File ..., line N` markers).
"""
import collections, csv, os, re, subprocess, sys, tempfile


def disasm(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    txt = []
    for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):   # one cubin per translation unit
        t = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        if kernel_sub in t:
            txt = t.splitlines(); break
    # find the kernel's .text section
    start = None
    for i, l in enumerate(txt):
        if l.lstrip().startswith(".section") and ".text." in l:
            if start is not None:
                end = i; break
            if kernel_sub in l: start = i
    else:
        end = len(txt)
    fn, line, info = "kernel", ("?", 0), {}
    for l in txt[start:end]:
        m = re.match(r"^\$?([^\s:]+):\s*$", l)
        if m and not m.group(1).startswith(".L_"):
            name = m.group(1)
            fn = name.split("$")[-1] if "$" in name else "kernel"
            mm = re.search(r"_ZN3bbg?\d+([A-Za-z0-9]+?)I[df]", fn)
            if mm: fn = mm.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            info[int(m.group(1), 16)] = (fn, line, m.group(2).strip())
    return info


def sass_page(rep, kid):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
    blocks, cur = [], None
    for l in raw:
        if l.startswith('"Kernel Name"'):
            cur = {"name": l, "rows": []}; blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(l)
    b = [x for x in blocks if kid in x["name"]][0] if isinstance(kid, str) else blocks[kid]
    rows = list(csv.reader(b["rows"]))
    hdr = rows[0]
    ia, ie, isamp, ithr = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    ini = hdr.index("stall_no_inst")
    out = [(int(r[ia], 16), int(r[ie]), int(r[isamp]), int(r[ithr]), int(r[ini] or 0)) for r in rows[1:] if len(r) > ie and r[ia].startswith("0x")]
    base = out[0][0]
    return b["name"], [(a - base, e, s, t, n) for a, e, s, t, n in out]


if __name__ == "__main__":
    rep, lib, ksub = sys.argv[1:4]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 40
    kid = sys.argv[sys.argv.index("--name") + 1] if "--name" in sys.argv else (int(sys.argv[sys.argv.index("--id") + 1]) if "--id" in sys.argv else 0)
    info = disasm(lib, ksub)
    name, rows = sass_page(rep, kid)
    tot = sum(r[1] for r in rows); tots = sum(r[2] for r in rows); totn = sum(r[4] for r in rows)
    noi = collections.Counter()
    byfn = collections.Counter(); sfn = collections.Counter(); byline = collections.Counter(); sline = collections.Counter(); thr = collections.Counter()
    nins = collections.Counter()
    for off, e, s, t, n in rows:
        fn, line, _ = info.get(off, ("?", ("?", 0), ""))
        byfn[fn] += e; sfn[fn] += s; byline[(fn,) + line] += e; sline[(fn,) + line] += s; thr[fn] += t; nins[fn] += 1; noi[fn] += n
    print(f"{name[:120]}\nexecuted warp instructions {tot:.4g}, stall samples {tots}, static SASS instructions {len(rows)}")
    print(f"{'function':28s} {'static':>7s} {'executed':>12s} {'share':>7s} {'samples':>8s} {'lanes/inst':>10s} {'no_inst':>8s}")
    for fn, e in byfn.most_common():
        print(f"{fn[:28]:28s} {nins[fn]:7d} {e:12.4g} {100 * e / tot:6.1f}% {100 * sfn[fn] / max(tots, 1):7.1f}% {thr[fn] / max(e, 1):10.1f} {100 * noi[fn] / max(totn, 1):7.1f}%")
    print(f"\ntop {nlines} source lines (function, file:line, executed share, sample share)")
    for k, e in byline.most_common(nlines):
        print(f"  {k[0][:22]:22s} {k[1]}:{k[2]:<5d} {100 * e / tot:6.2f}% {100 * sline[k] / max(tots, 1):6.2f}%")

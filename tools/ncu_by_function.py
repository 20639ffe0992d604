#!/usr/bin/env python
"""Attribute ncu warp-stall samples / executed instructions of one kernel to the C++ functions of csrc/*.h.

  ncu -i rep.ncu-rep --page source --csv > src.csv           (SASS view of the profiled kernel)
  python tools/ncu_by_function.py src.csv <lib.so> <kernel-substring>   e.g. ILb1ELi7
Uses nvdisasm -g line info of the same cubin (the kernel lives in headers, which ncu's CUDA view does not show).
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def line_table(so, kern):
    tmp = Path(tempfile.mkdtemp())
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(so).resolve())], cwd=tmp, check=True, capture_output=True)
    cubin = [p for p in tmp.glob("*.cubin") if "segments" not in p.name][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
    out, cur, inside = [], None, False
    for ln in dis:
        if ln.startswith("//--------------------- .text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
            out.append((cur, ln.strip()))
    return out


def func_of(file, line, cache={}):
    if file not in cache:
        p = next(Path(__file__).resolve().parents[1].rglob(file), None)
        starts = []
        if p:
            for i, l in enumerate(p.read_text().splitlines(), 1):
                m = re.match(r"(?:SVAD_HD|__device__ __forceinline__|static|template.*>\s*)?\s*[\w:<>\*&\s]+?\b(\w+)\s*\([^;]*$", l)
                if m and not l.startswith((" ", "\t", "#", "//")) and "(" in l:
                    starts.append((i, m.group(1)))
                elif re.match(r"\s+(?:SVAD_HD|__device__ __forceinline__)\s+(?:static\s+)?[\w:<>\*&\s]+?\b(\w+)\s*\(", l):
                    starts.append((i, re.match(r"\s+(?:SVAD_HD|__device__ __forceinline__)\s+(?:static\s+)?[\w:<>\*&\s]+?\b(\w+)\s*\(", l).group(1)))
        cache[file] = starts
    name = "?"
    for s, n in cache[file]:
        if s <= line:
            name = n
    return f"{file}:{name}"


def main():
    src, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    tab = line_table(so, kern)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    assert len(body) == len(tab), (len(body), len(tab))
    agg = collections.defaultdict(lambda: collections.Counter())
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for (loc, sass), r in zip(tab, body):
        fn = func_of(*loc) if loc else "?"
        a = agg[fn]
        a["samples"] += int(r[col["# Samples"]] or 0)
        a["inst"] += int(r[col["Instructions Executed"]] or 0)
        op = sass.split("*/")[1].split()[0] if "*/" in sass else "?"
        if op.startswith("@"):
            op = sass.split("*/")[1].split()[1]
        op = op.split(".")[0]
        a["op_" + op] += int(r[col["Instructions Executed"]] or 0)
        for s in stalls:
            a[s] += int(r[col[s]] or 0)
        a["smem_wavefronts"] += int(float(r[col["L1 Wavefronts Shared"]] or 0))
        a["smem_excess"] += int(float(r[col["L1 Wavefronts Shared Excessive"]] or 0))
    tot = sum(a["samples"] for a in agg.values())
    toti = sum(a["inst"] for a in agg.values())
    print(f"total samples {tot}, instructions {toti}")
    for fn, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        top = sorted(((s, a[s]) for s in stalls), key=lambda x: -x[1])[:4]
        ops = sorted(((k[3:], v) for k, v in a.items() if k.startswith("op_")), key=lambda x: -x[1])[:5]
        print(f"{fn:42s} samples {100*a['samples']/tot:5.1f}%  inst {100*a['inst']/toti:5.1f}%  smem_wf {a['smem_wavefronts']:>10d} (+{a['smem_excess']} excess)  "
              + " ".join(f"{s[6:]}={100*v/max(a['samples'],1):.0f}%" for s, v in top) + "  | " + " ".join(f"{o}:{100*v/max(a['inst'],1):.0f}%" for o, v in ops))


if __name__ == "__main__":
    main()

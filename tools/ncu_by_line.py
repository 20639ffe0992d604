#!/usr/bin/env python
"""Warp-stall samples and executed instructions of one kernel by SOURCE LINE, from an ncu report and the library it profiled.

  ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv      (per SASS instruction, address order)
  tools/ncu_by_line.py sass.csv silero_vad_b200/lib/libsilero_vad_b200.so svad_fused_h16ILb1EfE [top]

The library is disassembled with nvdisasm -g (line info from -lineinfo); instructions are matched by order."""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def disasm_lines(so, func):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(so).resolve())], cwd=d, check=True, capture_output=True)
        cubin = max(Path(d).glob("*.cubin"), key=lambda p: p.stat().st_size)
        out = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
    lines, cur, on = [], None, False
    for l in out.splitlines():
        if l.startswith(".text.") and l.endswith(":"):
            on = func in l
            continue
        if not on:
            continue
        if l.startswith("//-----"):
            if lines:
                break
            continue
        if l.strip().startswith("//## File"):
            m = re.search(r'File "([^"]+)", line (\d+)', l)
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return lines


def main():
    sass_csv, so, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = disasm_lines(so, func)
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    assert len(data) == len(dis), (len(data), len(dis), "report and library are different builds")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    by = collections.defaultdict(lambda: collections.Counter())
    tot = 0
    for r, (_, loc, op) in zip(data, dis):
        s = int(r[ix["# Samples"]] or 0)
        tot += s
        c = by[loc]
        c["samples"] += s
        c["inst"] += int(r[ix["Instructions Executed"]] or 0)
        for h in stalls:
            c[h] += int(r[ix[h]] or 0)
    srcs = {}
    print(f"total samples {tot}")
    for loc, c in sorted(by.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        f, ln = loc if loc else ("?", 0)
        if f not in srcs:
            p = next(Path(__file__).resolve().parents[1].rglob(f), None)
            srcs[f] = p.read_text().splitlines() if p else []
        text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
        why = ", ".join(f"{h[6:]} {100 * c[h] / max(c['samples'], 1):.0f}%" for h in sorted(stalls, key=lambda h: -c[h])[:3] if c[h])
        print(f"{100 * c['samples'] / tot:5.1f}%  inst {c['inst']:10d}  {f}:{ln:<4d} {text}\n        [{why}]")


if __name__ == "__main__":
    main()

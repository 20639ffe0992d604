#!/usr/bin/env python
"""Summarise one ncu --set full capture of the fused kernel into profiles/ (tracked):
   python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_fused_fp32_b4096_t64 <kernel-substring> [traffic-key]
Writes <out>.md (key metrics + per-function stall attribution) and updates profiles/traffic.json."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.sum.per_cycle_elapsed", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, out, kern = sys.argv[1], Path(sys.argv[2]), sys.argv[3]
    tkey = sys.argv[4] if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = [f"# ncu summary: {Path(rep).name}", "", f"kernel: `{m.get('Kernel Name', ('?',))[0]}`", "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    def num(k):
        v, u = m[k]
        f = float(v.replace(",", ""))
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
        return f * mult
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    lines += ["", f"DRAM traffic per launch (read+write): {traffic/1e6:.2f} MB"]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = Path("/tmp/_src.csv"); tmp.write_text(src)
    byfn = subprocess.run([sys.executable, str(REPO / "tools/ncu_by_function.py"), str(tmp), str(REPO / "silero_vad_b200/lib/libsilero_vad_b200.so"), kern],
                          capture_output=True, text=True)
    lines += ["", "## warp-stall samples and executed instructions by source function", "", "```", byfn.stdout.strip() or byfn.stderr.strip(), "```"]
    out.with_suffix(".md").write_text("\n".join(lines) + "\n")
    if tkey:
        tp = REPO / "profiles" / "traffic.json"
        d = json.loads(tp.read_text()) if tp.exists() else {}
        d[tkey] = traffic
        tp.write_text(json.dumps(d, indent=1) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise one `ncu --set full` capture of svad_fused_h16 into profiles/<out>.md (tracked) and update profiles/traffic.json.
   python tools/summarize_ncu_r02.py gpurun_out/x.ncu-rep <library the capture ran> profiles/r02_h16_b4096_t64 [traffic-key]"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_op_read_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, so, out = sys.argv[1], sys.argv[2], Path(sys.argv[3])
    tkey = sys.argv[4] if len(sys.argv) > 4 else None
    func = sys.argv[5] if len(sys.argv) > 5 else "svad_fused_h16ILb1EfLb1E"   # 16 kHz, fp32 audio, CTA-pair mode
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    m = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    lines = [f"# ncu summary: {Path(rep).name}", "", f"kernel: `{m.get('Kernel Name', ('?',))[0]}`", "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")

    def num(k):
        v, u = m[k]
        return float(v.replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1}.get(u, 1)
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    lines += ["", f"DRAM traffic per launch (read + write): {traffic / 1e6:.2f} MB  (algorithmic: 4096 x 64 x 2052 B = 537.9 MB)",
              f"L2 -> SM bytes per launch: {num('l1tex__m_xbar2l1tex_read_bytes.sum') / 1e9:.2f} GB (the weight tapes, streamed by every CTA every chunk step)"]
    if "lts__t_sectors_srcunit_tex_op_read.sum" in m:
        lines.append(f"bytes read out of the L2 slices for the SMs: {num('lts__t_sectors_srcunit_tex_op_read.sum') * 32 / 1e9:.2f} GB "
                     "(CTA pairs fetch half a slab each and multicast it; 148 single CTAs: 8.58 GB, 148 CTAs in pairs: 6.14 GB, round 1: 16 GB)")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tmp = Path("/tmp/_sass.csv")
    tmp.write_text(src)
    by = subprocess.run([sys.executable, str(REPO / "tools/ncu_by_line.py"), str(tmp), so, func, "48"], capture_output=True, text=True)
    lines += ["", "## warp-stall samples and executed instructions by source line (top 48)", "",
              "(12 warps per CTA: most samples are warps WAITING at an mbarrier -- `mbar_wait*` lines -- which is what idle roles do)", "",
              "```", by.stdout.strip() or by.stderr.strip(), "```"]
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    ops = {}
    for key in ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTCATOMSWS", "SYNCS", "F2FP", "MUFU", "FFMA", "HADD2", "LDG", "STS", "LDS"):
        ops[key] = sum(1 for l in sass.splitlines() if f" {key}" in l)
    lines += ["", "## SASS opcode census of the whole library (all kernels, static counts)", "", "```", json.dumps(ops), "```"]
    out.with_suffix(".md").write_text("\n".join(lines) + "\n")
    if tkey:
        tp = REPO / "profiles" / "traffic.json"
        d = json.loads(tp.read_text()) if tp.exists() else {}
        d[tkey] = traffic
        tp.write_text(json.dumps(d, indent=1) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()

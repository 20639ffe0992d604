"""Print the headline fields of a bench.py JSON line: python tools/show_bench.py gpurun_out/bench.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.3e chunks/s  kernel_ms %.3f  fp32_frac %.3f  hbm_frac %.4f" % (d["value"], d.get("kernel_ms", 0), d["roofline"]["fp32_frac"], d["roofline"]["frac"]))
if "e2e" in d:
    print("e2e %.3e  pcm16 %.3e" % (d["e2e"]["value"], d["e2e"].get("pcm16", {}).get("value", 0)))
if "cpu_baseline" in d:
    print("cpu_baseline %.3e on %d threads" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
print("latency_b1", d.get("latency_b1"))
print("clocks", d.get("clocks"))

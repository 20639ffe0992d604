"""One fused-kernel launch on the bench workload shape (for ncu): python tools/profile_run.py [B] [T] [sr] [rows]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from silero_vad_b200 import load_silero_vad

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sr = int(sys.argv[3]) if len(sys.argv) > 3 else 16000
rows = int(sys.argv[4]) if len(sys.argv) > 4 else 0
kernel = sys.argv[5] if len(sys.argv) > 5 else "tc"
n = 512 if sr == 16000 else 256
m = load_silero_vad(device=0)
m.engine.set_tile_rows(rows)
if kernel == "small":
    m.engine.set_small_batch_max(1 << 30)
else:
    m.engine.set_kernel(kernel)
    m.engine.set_small_batch_max(0)
x = torch.randn(B, n * T, device="cuda") * 0.03
p = torch.empty(B, T, device="cuda")
for _ in range(3):
    m.engine.forward_device(sr, B, n * T, n * T, x.data_ptr(), 0, 0, 0, 0, p.data_ptr(), T, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m.engine.forward_device(sr, B, n * T, n * T, x.data_ptr(), 0, 0, 0, 0, p.data_ptr(), T, torch.cuda.current_stream().cuda_stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"kernel={kernel} B={B} T={T} sr={sr} rows={rows}: {ms:.3f} ms -> {B*T/ms*1e3:.3e} chunks/s")

"""clock64 stamps of one step of the small-batch cluster kernel (CTA 0): python tools/small_phase_times.py [B] [T]"""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from silero_vad_b200 import load_silero_vad, _cabi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m = load_silero_vad(device=0)
m.engine.set_small_batch_max(1 << 30)
L = _cabi.lib()
L.svad_engine_set_debug_buffer.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
L.svad_engine_set_debug_buffer(m.engine._h, dbg.data_ptr())
x = torch.randn(B, 512 * T, device="cuda") * 0.03
p = torch.empty(B, T, device="cuda")
for _ in range(3):
    m.engine.forward_device(16000, B, 512 * T, 512 * T, x.data_ptr(), 0, 0, 0, 0, p.data_ptr(), T, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
d = dbg.cpu().tolist()
names = ["window load", "STFT dots", "cluster.sync", "enc0 dots", "cluster.sync", "enc1..enc3 (+3 syncs)", "LSTM dots + gates", "cluster.sync", "head"]
for i, n in enumerate(names):
    print(f"{n:26s} {d[i+1]-d[i]:8d} cycles")
print(f"{'step total':26s} {d[9]-d[0]:8d} cycles")

"""Per-phase clock64 stamps of the middle step (t = T/2) of the tensor-core kernel (CTA 0, thread 0): python tools/tc_phase_times.py [B] [T] [rows]"""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from silero_vad_b200 import load_silero_vad, _cabi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 8
m = load_silero_vad(device=0)
m.engine.set_kernel("tc"); m.engine.set_tile_rows(rows)
L = _cabi.lib()
L.svad_engine_set_debug_buffer.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
L.svad_engine_set_debug_buffer(m.engine._h, dbg.data_ptr())
x = torch.randn(B, 512 * T, device="cuda") * 0.03
p = torch.empty(B, T, device="cuda")
for _ in range(3):
    m.engine.forward_device(16000, B, 512 * T, 512 * T, x.data_ptr(), 0, 0, 0, 0, p.data_ptr(), T, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
dbg.zero_()
m.engine.forward_device(16000, B, 512 * T, 512 * T, x.data_ptr(), 0, 0, 0, 0, p.data_ptr(), T, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
d = dbg.cpu().tolist()
names = ["STFT", "lo-stage enc0 + sync", "enc0 MMA issue loop", "enc0 MMA tail (acc_wait)", "enc0 epilogue", "enc1-3 (tensor core)", "lo-stage LSTM + sync",
         "LSTM MMA issue loop", "LSTM MMA tail", "LSTM epilogue + head"]
tot = d[10] - d[0]
for i, n in enumerate(names):
    print(f"{n:28s} {d[i+1]-d[i]:8d} cycles  {100*(d[i+1]-d[i])/tot:5.1f}%")
print(f"{'step total':28s} {tot:8d} cycles")
print(f"waiting for weights (warp 0): enc0 loop {d[11]} cycles, LSTM loop {d[13]} cycles")
print(f"enc1 MMA phase: {d[20]-d[5]} cycles, epilogue {d[21]-d[20]};  enc2: {d[22]-d[21]};  enc3: {d[6]-d[22]}")
print(f"enc2: slab_wait {d[23]-d[21]}, issue+commit {d[24]-d[23]}, acc_wait {d[25]-d[24]}, epilogue+sync {d[22]-d[25]};  enc3: slab_wait {d[26]-d[22]}, issue+commit {d[27]-d[26]}, stage_lo(h) {d[28]-d[27]}, acc_wait {d[29]-d[28]}, epilogue {d[6]-d[29]}")

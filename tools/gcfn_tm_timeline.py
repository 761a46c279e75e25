"""Pipeline timeline of the frames-as-M GCFN kernel (k_gcfn_tm; block 0, tile iterations 2-5) from in-kernel clock64 stamps.
    python tools/gcfn_tm_timeline.py [rows] [T] [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs, _lib
from sepreformer_b200.params import seeded_state, state_shapes
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
shape = MODEL_SHAPES["SepReformer_Base_WSJ0"]; F = shape.feat
m = Separator(**separator_kwargs(shape)); m.load_state_dict(seeded_state(state_shapes(m), seed=1)); m = m.cuda().eval()
m.gemm_path = 2; m.gcfn_tm = mode
pre = b"enc_stages.1.l_block_1.block.gcfn."
L = _lib.lib(); h = m.handle()
x = torch.randn(rows, T, F, device="cuda"); y = torch.empty_like(x)
clk = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    rc = L.sepref_debug_gcfn_timeline(h, pre, x.data_ptr(), rows, T, y.data_ptr(), clk.data_ptr(), st)
    assert rc == 0, L.sepref_last_error()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L.sepref_debug_gcfn_timeline(h, pre, x.data_ptr(), rows, T, y.data_ptr(), clk.data_ptr(), st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"k_gcfn_tm mode {mode}: {min(ts):.1f} us for rows={rows} T={T} ({rows*T/min(ts)*1e-3:.2f} G frames/s)")
c = clk.cpu().view(8, 64)
t0 = int(c[0][c[0] > 0].min())
names = {0: "g1:a_full ok", 16: "pro:a_empty ok", 17: "pro:LN done", 19: "drain:y_full ok", 18: "drain:done"}
for j in range(6):
    names[1 + j] = f"g1:chunk {j} issued"; names[8 + j] = f"g2:chunk {j} issued"
    names[24 + 4 * j] = f"epi{j}:acc_full"; names[25 + 4 * j] = f"epi{j}:h_empty ok"; names[26 + 4 * j] = f"epi{j}:tmem read done"; names[27 + 4 * j] = f"epi{j}:chunk done"
for it in range(2, 5):
    ev = sorted((int(c[it][k]) - t0, v) for k, v in names.items() if c[it][k] > 0)
    print(f"--- tile iteration {it}")
    for t, v in ev:
        print(f"{t:9d}  {v}")

"""Pipeline timeline of the CTA-pair GCFN kernel (block 0, first 8 tiles) from in-kernel clock64 stamps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs, _lib
from sepreformer_b200.params import seeded_state, state_shapes
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
shape = MODEL_SHAPES["SepReformer_Base_WSJ0"]; F = shape.feat
m = Separator(**separator_kwargs(shape)); m.load_state_dict(seeded_state(state_shapes(m), seed=1)); m = m.cuda().eval()
m.gemm_path = 2; setattr(m, "gcfn_" + (sys.argv[3] if len(sys.argv) > 3 else "pair"), 1)
pre = b"enc_stages.1.l_block_1.block.gcfn."
L = _lib.lib(); h = m.handle()
x = torch.randn(rows, T, F, device="cuda"); y = torch.empty_like(x)
clk = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    rc = L.sepref_debug_gcfn_timeline(h, pre, x.data_ptr(), rows, T, y.data_ptr(), clk.data_ptr(), st)
    assert rc == 0, L.sepref_last_error()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); L.sepref_debug_gcfn_timeline(h, pre, x.data_ptr(), rows, T, y.data_ptr(), clk.data_ptr(), st); e1.record(); torch.cuda.synchronize()
print(f"kernel {e0.elapsed_time(e1)*1e3:.1f} us for rows={rows} T={T}")
c = clk.cpu().view(8, 64)
t0 = int(c[0][c[0] > 0].min())
names = {0: "mma:b1_full", 1: "mma:G1(0) issued", 2: "mma:G1(1) issued", 3: "mma:G1(2) issued", 8: "mma:G2(0) issued", 9: "mma:G2(1) issued", 10: "mma:G2(2) issued",
         16: "pro:b1_empty ok", 17: "pro:LN done",
         48: "drain:start", 49: "drain:recv_free ok", 50: "drain:y_full ok", 51: "drain:sent", 52: "drain:recv_full ok", 53: "drain:done"}
names.update({24: "epi:tm_full", 26: "epi:tmem read done", 27: "epi:chunk done", 1: "mma:G1 issued", 8: "mma:G2 issued",
              48: "drain:start", 49: "drain:y_full ok", 50: "drain:sent", 51: "drain:recv_full ok", 52: "drain:done"} if len(sys.argv) > 3 and sys.argv[3] == "trio" else {})
for k in range(3 if not (len(sys.argv) > 3 and sys.argv[3] == "trio") else 0):
    for gate, nm in ((0, "val"), (4, "gate")):
        names[20 + 8 * k + gate] = f"{nm}{k}:c_full"; names[21 + 8 * k + gate] = f"{nm}{k}:b2_empty ok"
        names[22 + 8 * k + gate] = f"{nm}{k}:tmem read done"; names[23 + 8 * k + gate] = f"{nm}{k}:chunk done"
for it in range(2, 6):
    ev = sorted((int(c[it][k]) - t0, v) for k, v in names.items() if c[it][k] > 0)
    print(f"--- tile iteration {it}")
    for t, v in ev:
        print(f"{t:9d}  {v}")

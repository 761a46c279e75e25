"""Pipeline timeline of the fused GCFN kernel (block 0, first 8 tiles) from in-kernel clock64 stamps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs, _lib
from sepreformer_b200.params import seeded_state, state_shapes
name = sys.argv[1] if len(sys.argv) > 1 else "SepReformer_Base_WSJ0"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cl = int(sys.argv[4]) if len(sys.argv) > 4 else 2
T = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
shape = MODEL_SHAPES[name]; F = shape.feat
m = Separator(**separator_kwargs(shape), per_stage_split=shape.per_stage_split)
m.load_state_dict(seeded_state(state_shapes(m), seed=1)); m = m.cuda().eval(); m.cluster = cl; m.gemm_path = int(sys.argv[5]) if len(sys.argv) > 5 else 2
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
names = {0: "mma:b1_full", 1: "mma:S1(0) issued", 2: "mma:S1(1) issued", 3: "mma:S1(2) issued", 8: "mma:S2(0) issued", 9: "mma:S2(1) issued", 10: "mma:S2(2) issued",
         16: "pro:b1_empty ok", 17: "pro:LN done", 18: "pro:drain(prev) done", 19: "pro:y_full(prev) ok",
         24: "epi0:tm_full", 25: "epi0:b2_empty ok", 26: "epi0:tmem read done", 27: "epi0:done",
         28: "epi1:tm_full", 29: "epi1:b2_empty ok", 30: "epi1:tmem read done", 31: "epi1:done",
         32: "epi2:tm_full", 33: "epi2:b2_empty ok", 34: "epi2:tmem read done", 35: "epi2:done"}
for it in range(2, 6):
    ev = sorted((int(c[it][k]) - t0, v) for k, v in names.items() if c[it][k] > 0)
    print(f"--- tile iteration {it}")
    for t, v in ev:
        print(f"{t:9d}  {v}")

"""Pipeline timeline of one k_tok configuration (block 0, first tiles) while running a block through the C ABI."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs, _lib
from sepreformer_b200.params import seeded_state, state_shapes
tag = sys.argv[1] if len(sys.argv) > 1 else "gate"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 8000
shape = MODEL_SHAPES["SepReformer_Base_WSJ0"]; F = shape.feat
m = Separator(**separator_kwargs(shape)); m.load_state_dict(seeded_state(state_shapes(m), seed=1)); m = m.cuda().eval()
L = _lib.lib(); h = m.handle()
x = torch.randn(rows, T, F, device="cuda")
clk = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
kind, prefix, kw = {"gate": ("ega", "enc_stages.0.g_block_1.block.ega.", dict(td=int(os.environ.get("TD", T // 16)))),
                    "qkv_pool": ("ega", "enc_stages.0.g_block_1.block.ega.", dict(td=int(os.environ.get("TD", T // 16)))),
                    "cla_a": ("cla", "enc_stages.0.l_block_1.block.cla.", {}),
                    "cla_b": ("cla", "enc_stages.0.l_block_1.block.cla.", {}),
                    "qkv>": ("spk_attention", "dec_stages.3.spk_attn_1.", {}),
                    "proj_res": ("spk_attention", "dec_stages.3.spk_attn_1.", {}),
                    "fuse": ("fusion", "simple_fusion.3.", dict(x_low=torch.randn(rows, T // 2, F, device="cuda")))}[tag]
_lib.check(L.sepref_debug_tok_timeline(h, tag.encode(), clk.data_ptr()))
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m.raw_f16 = int(sys.argv[5]) if len(sys.argv) > 5 else 0
_lib.check(L.sepref_set_option(h, 99, flags))
for _ in range(3):
    m.run_block(kind, prefix, x, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.run_block(kind, prefix, x, **kw); e1.record(); torch.cuda.synchronize()
print(f"{tag}: block call {e0.elapsed_time(e1)*1e3:.1f} us (rows={rows}, T={T}, raw_f16={m.raw_f16})")
c = clk.cpu().view(8, 64)
names = {0: "mma:b1_full", 1: "mma:issued", 16: "pro:b1_empty ok", 17: "pro:done", 18: "pro:drain(prev) done",
         24: "epi0:tm_full", 27: "epi0:done", 28: "epi1:tm_full", 31: "epi1:done"}
for bq in range(8):
    names[32 + 3 * bq] = f"b{bq}:ld_issued"; names[33 + 3 * bq] = f"b{bq}:ld_done"; names[34 + 3 * bq] = f"b{bq}:stored"
for j in range(2):
    names[2 + 3 * j] = f"mma:s2({j}) enter"; names[3 + 3 * j] = f"mma:s2({j}) b2_full ok"; names[4 + 3 * j] = f"mma:s2({j}) issued"
for w in range(8):
    names[48 + w] = f"w{w}:done"
t0 = int(c[c > 0].min())
for it in range(int(os.environ.get("IT0", 2)), 7):
    ev = sorted((int(c[it][k]) - t0, v) for k, v in names.items() if c[it][k] > 0)
    print(f"--- tile iteration {it}: " + "  ".join(f"{v}@{t}" for t, v in ev))

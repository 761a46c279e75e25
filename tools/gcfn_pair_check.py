"""Parity and speed of the weights-resident GCFN kernels (CTA pair / cluster of three) against the streaming kernel.
    python tools/gcfn_pair_check.py [pair|trio]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import separator_oracle as O
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs
from sepreformer_b200.params import seeded_state, state_shapes
shape = MODEL_SHAPES["SepReformer_Base_WSJ0"]
m = Separator(**separator_kwargs(shape)); sd = seeded_state(state_shapes(m), seed=1); m.load_state_dict(sd); m = m.cuda().eval()
p = {k: v for k, v in sd.items() if v.is_floating_point()}
m.write_stage_outputs = False
prefix = "dec_stages.1.g_block_2.block.gcfn."
attr = "gcfn_" + (sys.argv[1] if len(sys.argv) > 1 else "pair")
worst = 0.0
for rows, T in ((1, 1), (1, 2), (1, 93), (1, 94), (1, 95), (2, 188), (1, 189), (2, 300), (3, 158), (2, 1000), (4, 8000), (7, 1234)):
    x = torch.randn(rows, T, shape.feat, device="cuda")
    setattr(m, attr, 0); y0 = m.run_block("gcfn", prefix, x)
    setattr(m, attr, 1); y1 = m.run_block("gcfn", prefix, x)
    torch.cuda.synchronize()
    d = float((y1 - y0).norm() / y0.norm())
    line = f"gcfn rows={rows} T={T}: pair vs streaming rel {d:.3e} max abs {float((y1-y0).abs().max()):.3e}"
    if rows * T <= 4000:
        with torch.no_grad():
            ref = O.gcfn(x.cpu(), p, prefix)
        e1 = float((y1.cpu() - ref).norm() / ref.norm()); e0 = float((y0.cpu() - ref).norm() / ref.norm())
        line += f"   vs oracle: pair {e1:.3e} streaming {e0:.3e}"
        worst = max(worst, e1)
    print(line, flush=True)
print("worst pair-vs-oracle", worst)
x = torch.randn(32, shape.feat, 7997, device="cuda")
for pair in (0, 1):
    setattr(m, attr, pair)
    y, _ = m(x); y, _ = m(x)
    torch.cuda.synchronize()
    if pair == 0: yref = y.clone()
    else: print("forward rel diff", float((y - yref).norm() / yref.norm()))
    prof = m.profile_kernels(x, steps=3)
    print("pair", pair, "gcfn_ms", round(prof["gcfn_ms"], 3), "sum", round(sum(v for k, v in prof.items() if k.endswith("_ms")), 3), flush=True)

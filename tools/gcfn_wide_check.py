"""Parity and speed of the 160-frame GCFN tile variant against the default kernel (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs
from sepreformer_b200.params import seeded_state, state_shapes
shape = MODEL_SHAPES["SepReformer_Base_WSJ0"]
m = Separator(**separator_kwargs(shape)); m.load_state_dict(seeded_state(state_shapes(m), seed=1)); m = m.cuda().eval()
m.write_stage_outputs = False
prefix = "dec_stages.1.g_block_2.block.gcfn."
for rows, T in ((2, 300), (3, 158), (1, 159), (2, 1000), (4, 8000)):
    x = torch.randn(rows, T, shape.feat, device="cuda")
    m.gcfn_wide = 0; y0 = m.run_block("gcfn", prefix, x)
    m.gcfn_wide = 1; y1 = m.run_block("gcfn", prefix, x)
    torch.cuda.synchronize()
    d = (y1 - y0).norm() / y0.norm()
    print(f"gcfn block rows={rows} T={T}: rel diff wide vs default {float(d):.3e}  max abs {float((y1-y0).abs().max()):.3e}")
x = torch.randn(32, shape.feat, 7997, device="cuda")
for wide in (0, 1):
    m.gcfn_wide = wide
    y, _ = m(x); y, _ = m(x)
    torch.cuda.synchronize()
    if wide == 0: yref = y.clone()
    else: print("forward rel diff", float((y - yref).norm() / yref.norm()))
    prof = m.profile_kernels(x, steps=2)
    print("wide", wide, "gcfn_ms", round(prof["gcfn_ms"], 3), "sum", round(sum(v for k, v in prof.items() if k.endswith("_ms")), 3))

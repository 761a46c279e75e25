"""Ad-hoc timing of the separator forward (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs
from sepreformer_b200.params import seeded_state, state_shapes

name = sys.argv[1] if len(sys.argv) > 1 else "SepReformer_Base_WSJ0"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
path = int(sys.argv[3]) if len(sys.argv) > 3 else 0
T = int(sys.argv[4]) if len(sys.argv) > 4 else 7997
shape = MODEL_SHAPES[name]
m = Separator(**separator_kwargs(shape), per_stage_split=shape.per_stage_split)
m.load_state_dict(seeded_state(state_shapes(m), seed=1))
m = m.cuda().eval()
m.gemm_path = path
x = torch.randn(B, shape.feat, T, device="cuda")
for _ in range(2):
    m(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n):
    m(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{name} B={B} T={T} path={path}: {ms:.2f} ms/forward, {B*T/ms*1e3:.3e} frames/s, launches={m.last_launch_count}")

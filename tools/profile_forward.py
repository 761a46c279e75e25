"""Per-kernel device time of one separator forward (CUDA events recorded by the library)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs
from sepreformer_b200.params import seeded_state, state_shapes
name = sys.argv[1] if len(sys.argv) > 1 else "SepReformer_Base_WSJ0"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
path = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cl = int(sys.argv[4]) if len(sys.argv) > 4 else 2
shape = MODEL_SHAPES[name]
m = Separator(**separator_kwargs(shape), per_stage_split=shape.per_stage_split)
m.load_state_dict(seeded_state(state_shapes(m), seed=1)); m = m.cuda().eval()
m.gemm_path = path; m.cluster = cl; m.write_stage_outputs = False
x = torch.randn(B, shape.feat, 7997, device="cuda")
for _ in range(2): m(x)
prof = m.profile_kernels(x, steps=3)
tot = sum(v for k, v in prof.items() if k.endswith("_ms"))
print(f"{name} B={B} path={path} cluster={cl}: sum of kernels {tot:.2f} ms")
for k, v in sorted(((k, v) for k, v in prof.items() if k.endswith("_ms")), key=lambda kv: -kv[1]):
    print(f"  {v:8.3f} ms  x{prof[k[:-3] + '_launches']:3d}  {k[:-3]}")

"""One warm-up + one separator forward (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs
from sepreformer_b200.params import seeded_state, state_shapes
name = sys.argv[1] if len(sys.argv) > 1 else "SepReformer_Base_WSJ0"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
shape = MODEL_SHAPES[name]
m = Separator(**separator_kwargs(shape), per_stage_split=shape.per_stage_split)
m.load_state_dict(seeded_state(state_shapes(m), seed=1))
m = m.cuda().eval()
m.write_stage_outputs = False
x = torch.randn(B, shape.feat, 7997, device="cuda")
for _ in range(n):
    m(x)
torch.cuda.synchronize()
print("launches per forward", m.last_launch_count)

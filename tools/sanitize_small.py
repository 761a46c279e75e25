"""Tiny forward on every kernel path, for compute-sanitizer (memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs
from sepreformer_b200.params import seeded_state, state_shapes
for name in ("SepReformer_Base_WSJ0", "SepReformer_Large_DM_WHAM"):
    shape = MODEL_SHAPES[name]
    m = Separator(**separator_kwargs(shape), per_stage_split=shape.per_stage_split)
    m.load_state_dict(seeded_state(state_shapes(m), seed=1)); m = m.cuda().eval()
    x = torch.randn(2, shape.feat, 203, device="cuda")
    for path in (2, 1, 0):
        m.gemm_path = path
        y, st = m(x)
        torch.cuda.synchronize()
        print(name, path, float(y.abs().mean()))

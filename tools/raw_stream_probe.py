"""Raw-stream GEMMs (split / fuse): kernel time and re-computation count, default (FP16 + conditional TF32) vs FP16 only vs TF32 path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs, _lib
from sepreformer_b200.params import seeded_state, state_shapes
shape = MODEL_SHAPES["SepReformer_Base_WSJ0"]
m = Separator(**separator_kwargs(shape)); m.load_state_dict(seeded_state(state_shapes(m), seed=1)); m = m.cuda().eval()
m.write_stage_outputs = False
L = _lib.lib()
x = torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 32, shape.feat, 7997, device="cuda")
for name, path, raw in (("f16 + conditional tf32", 2, 0), ("f16 only", 2, 1), ("tf32 path", 1, 0)):
    m.gemm_path = path; m.raw_f16 = raw
    y, _ = m(x); y, _ = m(x); torch.cuda.synchronize()
    n0 = L.sepref_range_rerun_count(m.handle())
    prof = m.profile_kernels(x, steps=3)
    n1 = L.sepref_range_rerun_count(m.handle())
    print(f"{name:24s} fuse {prof.get('tok<fuse>_ms', 0):.3f} ms  split {prof.get('tok<split>_ms', 0):.3f} ms  sum {sum(v for k, v in prof.items() if k.endswith('_ms')):.2f} ms  re-computations during 3 profiled forwards: {n1 - n0}", flush=True)

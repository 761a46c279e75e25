"""Relative L2 error of the separator output against every committed golden vector, per GEMM path (development aid)."""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from _util import load_golden, model_state, seeded_input, rel_l2, MODEL_SHAPES
from sepreformer_b200 import Separator, separator_kwargs

paths = [int(a) for a in sys.argv[1:]] or [2, 1, 0]
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for f in sorted(glob.glob(os.path.join(root, "sep_*.npz"))):
    tag = os.path.basename(f)[:-4]
    gold = load_golden(tag)
    name = str(gold["model"])
    shape = MODEL_SHAPES[name]
    m = Separator(**separator_kwargs(shape), per_stage_split=shape.per_stage_split)
    m.load_state_dict(model_state(name, int(gold["wseed"])), strict=True)
    m = m.cuda().eval()
    x = seeded_input(int(gold["xseed"]), int(gold["batch"]), shape.feat, int(gold["t_enc"]))
    out = []
    for p in paths:
        m.gemm_path = p
        with torch.no_grad():
            last, _ = m(x.cuda())
        out.append(f"path{p} {rel_l2(last.cpu()[..., ::int(gold['stride'])], gold['last']):.3e}")
    print(tag, "  ".join(out))

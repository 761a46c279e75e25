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
    # block-level entries on exact-size tensors whose token count ends early in a tile (tail clamping of the
    # prefetched residual rows must stay inside the tensor): 96-token tiles (cla_b), 128-token tiles (gate, proj)
    for path in (2, 1):
        m.gemm_path = path
        for rows, t in ((1, 96 * 3 + 20), (2, 128 + 16), (1, 1000)):
            xb = torch.randn(rows, t, shape.feat, device="cuda")
            m.run_block("cla", "dec_stages.1.l_block_1.block.cla.", xb)
            m.run_block("local_block", "enc_stages.2.l_block_2.", xb)
            m.run_block("spk_attention", "dec_stages.1.spk_attn_1.", torch.cat([xb, xb]))
            if t % 16 == 0:
                m.run_block("ega", "dec_stages.1.g_block_3.block.ega.", xb, td=t // 16)
        torch.cuda.synchronize()
        print(name, "blocks", path, "ok")
    if shape.feat == 128:
        # frames-as-M GCFN kernel (both modes) and the raw-stream range re-computation
        m.gemm_path = 2
        for mode in (1, 2):
            m.gcfn_tm = mode
            for rows, t in ((1, 1), (1, 31), (2, 121), (3, 250)):
                m.run_block("gcfn", "dec_stages.1.g_block_2.block.gcfn.", torch.randn(rows, t, shape.feat, device="cuda"))
            y, _ = m(x)
            torch.cuda.synchronize()
            print(name, "gcfn_tm", mode, float(y.abs().mean()))
        m.gcfn_tm = 0
        y, _ = m(x * 1.0e5)
        torch.cuda.synchronize()
        print(name, "raw-stream re-computation", float(y.abs().mean()))

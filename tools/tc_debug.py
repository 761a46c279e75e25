"""Bring-up aid for the tcgen05 GCFN kernel: compares h (GEMM1) and y against the CPU oracle, per cluster size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import separator_oracle as O
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs, _lib
from sepreformer_b200.params import seeded_state, state_shapes

name = sys.argv[1] if len(sys.argv) > 1 else "SepReformer_Base_WSJ0"
shape = MODEL_SHAPES[name]
F = shape.feat
m = Separator(**separator_kwargs(shape), per_stage_split=shape.per_stage_split)
sd = seeded_state(state_shapes(m), seed=1)
m.load_state_dict(sd)
m = m.cuda().eval()
p = {k: v for k, v in sd.items() if v.is_floating_point()}
pre = "enc_stages.1.l_block_1.block.gcfn."
L = _lib.lib()
for rows, T in ((1, 12), (2, 94), (3, 300), (5, 1000)):
    x = torch.randn(rows, T, F, generator=torch.Generator().manual_seed(T))
    h_ref = O.affine(O.layer_norm(x, p[pre + "net1.0.weight"], p[pre + "net1.0.bias"]), p[pre + "net1.1.weight"], p[pre + "net1.1.bias"])
    y_ref = O.gcfn(x, p, pre)
    xg = x.cuda()
    for cl, path in ((1, 1), (2, 1), (4, 1), (1, 2), (2, 2), (4, 2)):
        m.cluster = cl
        m.gemm_path = path
        h = m.handle()
        y = torch.zeros_like(xg)
        hbuf = torch.zeros(rows * T, 6 * F, device="cuda")
        rc = L.sepref_debug_gcfn_h(h, pre.encode(), xg.data_ptr(), rows, T, y.data_ptr(), hbuf.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, L.sepref_last_error()
        torch.cuda.synchronize()
        eh = ((hbuf.cpu().view(rows, T, -1).double() - h_ref.double()).norm() / h_ref.double().norm()).item()
        eb = ((y.cpu().double() - y_ref.double()).norm() / (y_ref.double() - x.double()).norm()).item()
        y2 = m.run_block("gcfn", pre, xg)          # fast (interior) epilogue path
        torch.cuda.synchronize()
        eb2 = ((y2.cpu().double() - y_ref.double()).norm() / (y_ref.double() - x.double()).norm()).item()
        print(f"rows={rows} T={T} cluster={cl} path={path}: rel err h={eh:.3e} branch(edge path)={eb:.3e} branch(fast path)={eb2:.3e}")

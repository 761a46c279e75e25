"""Bring-up aid for the tcgen05 GCFN kernel: compares h (GEMM1) and y against the CPU oracle."""
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
for rows, T in ((1, 12), (2, 94), (3, 300), (4, 1000)):
    x = torch.randn(rows, T, F, generator=torch.Generator().manual_seed(T))
    h_ref = O.affine(O.layer_norm(x, p[pre + "net1.0.weight"], p[pre + "net1.0.bias"]), p[pre + "net1.1.weight"], p[pre + "net1.1.bias"])
    y_ref = O.gcfn(x, p, pre)
    xg = x.cuda()
    y = torch.zeros_like(xg)
    hbuf = torch.zeros(rows * T, 6 * F, device="cuda")
    h = m.handle()
    rc = L.sepref_debug_gcfn_h(h, pre.encode(), xg.data_ptr(), rows, T, y.data_ptr(), hbuf.data_ptr(), torch.cuda.current_stream().cuda_stream)
    print("rc", rc, L.sepref_last_error() if rc else "")
    torch.cuda.synchronize()
    eh = ((hbuf.cpu().view(rows, T, -1).double() - h_ref.double()).norm() / h_ref.double().norm()).item()
    ey = ((y.cpu().double() - y_ref.double()).norm() / y_ref.double().norm()).item()
    ebranch = ((y.cpu().double() - x.double() - (y_ref.double() - x.double())).norm() / (y_ref.double() - x.double()).norm()).item()
    print(f"rows={rows} T={T}: rel err h={eh:.3e}  y={ey:.3e}  branch={ebranch:.3e}")
    if eh > 1e-2:
        d = (hbuf.cpu().view(rows, T, -1) - h_ref)
        print(" h sample got", hbuf.cpu().view(rows, T, -1)[0, 0, :8].tolist())
        print(" h sample ref", h_ref[0, 0, :8].tolist())
        print(" per-token err", d.norm(dim=-1)[0, :12].tolist())
        print(" per-chan-block err", [float(d[..., i*128:(i+1)*128].norm()) for i in range(6*F//128)])

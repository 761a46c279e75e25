"""Time the host-buffer entry point (copy/compute overlap) against the device-resident forward."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs, _lib
from sepreformer_b200.params import seeded_state, state_shapes
shape = MODEL_SHAPES["SepReformer_Base_WSJ0"]
m = Separator(**separator_kwargs(shape)); m.load_state_dict(seeded_state(state_shapes(m), seed=1)); m = m.cuda().eval()
B = 32
x = torch.randn(B, 128, 7997)
xp = x.pin_memory(); xd = x.cuda()
for chunk in (32, 16, 8, 4):
    _lib.check(_lib.lib().sepref_set_option(m.handle(), _lib.OPT_HOST_CHUNK, chunk))
    for _ in range(2): out, _ = m.forward_host(xp)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): out, _ = m.forward_host(xp)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"host entry, sub-batch {chunk:2d}: {dt*1e3:.2f} ms/step")
ref, _ = m(xd); torch.cuda.synchronize()
print("host == device result:", torch.equal(ref.cpu(), out))
for b in (32, 8):
    xb = xd[:b].contiguous()
    for _ in range(2): m(xb)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): m(xb)
    torch.cuda.synchronize(); print(f"device forward B={b}: {(time.perf_counter()-t)/5*1e3:.2f} ms")

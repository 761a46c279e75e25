"""Parity and speed of the frames-as-M GCFN kernel (k_gcfn_tm) against the streaming kernel and the CPU oracle.
    python tools/gcfn_tm_check.py [mode ...]      modes: 2 = single CTA, 1 = CTA pair, 3-5 = pair with bring-up flags
Every mode runs in its own process (a protocol bug traps the kernel and kills the CUDA context)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(mode, full):
    import torch
    from oracle import separator_oracle as O
    from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs
    from sepreformer_b200.params import seeded_state, state_shapes
    shape = MODEL_SHAPES["SepReformer_Base_WSJ0"]
    m = Separator(**separator_kwargs(shape)); sd = seeded_state(state_shapes(m), seed=1); m.load_state_dict(sd); m = m.cuda().eval()
    p = {k: v for k, v in sd.items() if v.is_floating_point()}
    m.write_stage_outputs = False
    prefix = "dec_stages.1.g_block_2.block.gcfn."
    worst = 0.0
    for rows, T in ((1, 1), (1, 2), (1, 29), (1, 30), (1, 31), (1, 60), (2, 61), (1, 120), (3, 121), (2, 300), (3, 158), (2, 1000), (4, 8000), (7, 1234)):
        x = torch.randn(rows, T, shape.feat, device="cuda")
        m.gcfn_tm = 0; y0 = m.run_block("gcfn", prefix, x)
        m.gcfn_tm = mode; y1 = m.run_block("gcfn", prefix, x)
        torch.cuda.synchronize()
        d = float((y1 - y0).norm() / y0.norm())
        line = f"mode {mode} rows={rows} T={T}: tm vs streaming rel {d:.3e} max abs {float((y1-y0).abs().max()):.3e}"
        if rows * T <= 4000:
            with torch.no_grad():
                ref = O.gcfn(x.cpu(), p, prefix)
            e1 = float((y1.cpu() - ref).norm() / ref.norm()); e0 = float((y0.cpu() - ref).norm() / ref.norm())
            line += f"   vs oracle: tm {e1:.3e} streaming {e0:.3e}"
            worst = max(worst, e1)
        if d > 1e-2:
            r = (y1 - y0)[0]
            line += f"   err by frame (first 8) {[round(float(v), 3) for v in r.norm(dim=1)[:8]]} by channel half {float(r[:, :64].norm()):.3f} {float(r[:, 64:].norm()):.3f}"
        print(line, flush=True)
    print(f"mode {mode} worst tm-vs-oracle", worst, flush=True)
    if not full:
        return
    x = torch.randn(32, shape.feat, 7997, device="cuda")
    for mm in (0, mode):
        m.gcfn_tm = mm
        y, _ = m(x); y, _ = m(x)
        torch.cuda.synchronize()
        if mm == 0: yref = y.clone()
        else: print("forward rel diff", float((y - yref).norm() / yref.norm()))
        prof = m.profile_kernels(x, steps=3)
        print("gcfn_tm", mm, "gcfn_ms", round(prof["gcfn_ms"], 3), "sum", round(sum(v for k, v in prof.items() if k.endswith("_ms")), 3), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(int(sys.argv[2]), len(sys.argv) > 3 and sys.argv[3] == "full")
    else:
        modes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2, 1]
        full = "full" in sys.argv
        for mode in modes:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(mode)] + (["full"] if full else []),
                               capture_output=True, text=True, timeout=600)
            print(r.stdout[-6000:], flush=True)
            if r.returncode != 0:
                print(f"mode {mode}: exit {r.returncode}\n{r.stderr[-1500:]}", flush=True)

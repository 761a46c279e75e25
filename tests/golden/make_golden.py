"""Generate the golden vectors under tests/golden/ by running the *reference itself*.

Run in the build container only (needs /root/reference; the GPU box has neither it nor this need):

    python tests/golden/make_golden.py

For every case the reference ``Separator`` (or one of its block classes) is instantiated from the
reference's own configs.yaml, loaded (strict) with ``sepreformer_b200.params.seeded_state(seed)`` -
weights that any machine can regenerate from the seed - and run in fp64 on a seeded input; the fp64
output is stored as float32.  A few weight/input checksums are stored too so that a silent change of
the random generator would be detected rather than misread as a parity failure.
"""
import importlib
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from loguru import logger  # noqa: E402

logger.remove()

from sepreformer_b200.configs import MODEL_SHAPES  # noqa: E402
from sepreformer_b200.params import ParamTree, separator_spec, seeded_state, state_shapes  # noqa: E402


def seeded_input(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def checksums(sd):
    keys = sorted(k for k in sd if not k.endswith("num_batches_tracked"))
    picks = [keys[0], keys[len(keys) // 3], keys[2 * len(keys) // 3], keys[-1]]
    return np.array([float(sd[k].double().abs().sum()) for k in picks]), np.array(picks)


def ref_separator(model_name):
    mod = importlib.import_module(f"models.{model_name}.modules.module")
    cfg = yaml.full_load(open(f"/root/reference/models/{model_name}/configs.yaml"))["config"]["model"]["module_separator"]
    return mod.Separator(**cfg).eval(), mod


def separator_case(tag, model_name, batch, t_enc, wseed, xseed, stride=1):
    shape = MODEL_SHAPES[model_name]
    ref, _ = ref_separator(model_name)
    sd = seeded_state(state_shapes(ParamTree(separator_spec(shape))), seed=wseed)
    ref.load_state_dict(sd, strict=True)
    ref = ref.double()
    x = seeded_input(xseed, batch, shape.feat, t_enc)
    with torch.no_grad():
        last, stages = ref(x.double())
    wsum, wkeys = checksums(sd)
    out = dict(model=np.array(model_name), batch=batch, t_enc=t_enc, wseed=wseed, xseed=xseed, stride=stride,
               wsum=wsum, wkeys=wkeys, xsum=float(x.double().abs().sum()),
               last=last[..., ::stride].float().numpy(),
               last_norm=float(last.norm()))
    for i, s in enumerate(stages):
        out[f"stage{i}"] = s[..., ::stride].float().numpy()
        out[f"stage{i}_norm"] = float(s.norm())
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print(tag, last.shape, float(last.norm()))


def block_cases(tag, model_name, wseed, xseed, batch=2, td=3):
    """One call of each reference block class on channels-last input; prefixes name the weights used."""
    shape = MODEL_SHAPES[model_name]
    ref, mod = ref_separator(model_name)
    sd = seeded_state(state_shapes(ParamTree(separator_spec(shape))), seed=wseed)
    ref.load_state_dict(sd, strict=True)
    ref = ref.double()
    f, r = shape.feat, 4
    t = td * r
    x = seeded_input(xseed, batch * 2, t, f).double()          # [B*S, T, F] channels-last
    pos = torch.arange(td)
    pos_k, _ = ref.pos_emb((pos[:, None] - pos[None, :]).long())
    out = dict(model=np.array(model_name), wseed=wseed, xseed=xseed, batch=batch, td=td, t=t)
    with torch.no_grad():
        dec = ref.dec_stages[1]
        enc = ref.enc_stages[2]
        out["gcfn"] = dec.g_block_2.block["gcfn"](x).float().numpy()                      # dec_stages.1.g_block_2.block.gcfn.
        out["cla"] = dec.l_block_1.block["cla"](x).float().numpy()                        # dec_stages.1.l_block_1.block.cla.
        out["ega"] = dec.g_block_3.block["ega"](x.transpose(1, 2), pos_k).float().numpy()  # dec_stages.1.g_block_3.block.ega.
        out["global"] = enc.g_block_1(x.transpose(1, 2), pos_k).transpose(1, 2).float().numpy()   # enc_stages.2.g_block_1.
        out["local"] = enc.l_block_2(x).float().numpy()                                    # enc_stages.2.l_block_2.
        out["spkattn"] = dec.spk_attn_1(x.transpose(1, 2), 2).transpose(1, 2).float().numpy()     # dec_stages.1.spk_attn_1.
        out["downconv"] = enc.downconv(x).float().numpy()                                  # enc_stages.2.downconv.
        split = ref.spk_split_blocks[1] if shape.per_stage_split else ref.spk_split_block
        out["spksplit"] = split(x[:batch].transpose(1, 2)).transpose(1, 2).float().numpy()
        low = seeded_input(xseed + 1, batch * 2, t // 2, f).double()
        up = torch.nn.functional.interpolate(low.transpose(1, 2), size=t)
        out["fusion"] = ref.simple_fusion[2](torch.cat([up, x.transpose(1, 2)], 1)).transpose(1, 2).float().numpy()
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print(tag, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 1})


if __name__ == "__main__":
    only = set(sys.argv[1:])        # optional: regenerate just the named cases

    def separator_case(tag, *a, _f=separator_case, **k):      # noqa: F811
        if not only or tag in only:
            _f(tag, *a, **k)

    def block_cases(tag, *a, _f=block_cases, **k):            # noqa: F811
        if not only or tag in only:
            _f(tag, *a, **k)

    separator_case("sep_base_small", "SepReformer_Base_WSJ0", batch=2, t_enc=157, wseed=1, xseed=11)
    separator_case("sep_base_exact16", "SepReformer_Base_WSJ0", batch=1, t_enc=96, wseed=2, xseed=12)
    separator_case("sep_base_medium", "SepReformer_Base_WSJ0", batch=1, t_enc=1997, wseed=1, xseed=13, stride=8)
    separator_case("sep_large_whamr_small", "SepReformer_Large_DM_WHAMR", batch=1, t_enc=150, wseed=3, xseed=14)
    separator_case("sep_large_wham_small", "SepReformer_Large_DM_WHAM", batch=2, t_enc=79, wseed=4, xseed=15)
    # F = 256 at a length where every persistent kernel walks several tiles per CTA at the full-rate stages
    separator_case("sep_large_medium", "SepReformer_Large_DM_WSJ0", batch=1, t_enc=2003, wseed=5, xseed=16, stride=8)
    block_cases("blocks_base", "SepReformer_Base_WSJ0", wseed=1, xseed=21)
    block_cases("blocks_large", "SepReformer_Large_DM_WSJ0", wseed=5, xseed=22)

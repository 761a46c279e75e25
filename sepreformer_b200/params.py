"""Parameter containers whose ``state_dict()`` keys and shapes equal the reference separator's.

The reference builds its separator out of ``torch.nn`` layers nested in ``ModuleDict`` /
``Sequential`` / ``ModuleList`` containers (``modules/network.py``, ``modules/module.py:38-188``); a
checkpoint addresses parameters through that nesting, e.g.
``enc_stages.0.g_block_1.block.gcfn.net1.1.weight``.  The B200 separator never *runs* these layers -
its forward is one C-ABI call - but it must own identically named parameters so that
``load_state_dict`` of a reference checkpoint (``utils/util_engine.py:43``, ``strict=False``) lands.

The tree is described by a small table (:func:`separator_spec`) and instantiated generically by
:class:`ParamTree`; leaves are stock ``torch.nn`` layers used purely as initialised storage.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch
from torch import nn

from .configs import SeparatorShape


class LayerScaleParam(nn.Module):
    """Holds ``layer_scale`` of shape ``[1,1,F]`` initialised to 1e-5 (reference network.py:7-18)."""

    def __init__(self, feat: int, init: float = 1.0e-5):
        super().__init__()
        self.layer_scale = nn.Parameter(torch.full((1, 1, feat), init))


class ParamTree(nn.Module):
    """A module whose children are created from a ``{name: spec}`` table.

    ``spec`` is either a ready ``nn.Module`` leaf or a nested dict.  Integer-like names ("0", "1", ...)
    reproduce the keys that ``Sequential``/``ModuleList`` would generate.
    """

    def __init__(self, table: dict):
        super().__init__()
        for name, spec in table.items():
            self.add_module(name, spec if isinstance(spec, nn.Module) else ParamTree(spec))


def _gcfn(f: int) -> dict:      # network.py:46-58
    return {"net1": {"0": nn.LayerNorm(f), "1": nn.Linear(f, 6 * f)},
            "depthwise": nn.Conv1d(6 * f, 6 * f, 3, padding=1, groups=6 * f),
            "net2": {"2": nn.Linear(3 * f, f)},
            "Layer_scale": LayerScaleParam(f)}


def _mha(f: int) -> dict:       # network.py:76-88
    return {"layer_norm": nn.LayerNorm(f), "linear_q": nn.Linear(f, f), "linear_k": nn.Linear(f, f),
            "linear_v": nn.Linear(f, f), "linear_out": nn.Linear(f, f), "Layer_scale": LayerScaleParam(f)}


def _global_block(f: int) -> dict:   # network.py:126-136,189-196
    return {"block": {"ega": {"block": {"self_attn": _mha(f),
                                        "linear": {"0": nn.LayerNorm(f), "1": nn.Linear(f, f)}}},
                      "gcfn": _gcfn(f)}}


def _local_block(f: int, k: int) -> dict:   # network.py:159-172,212-218
    cla = {"layer_norm": nn.LayerNorm(f), "linear1": nn.Linear(f, 2 * f),
           "dw_conv_1d": nn.Conv1d(f, f, k, padding="same", groups=f),
           "linear2": nn.Linear(f, 2 * f), "BN": nn.BatchNorm1d(2 * f),
           "linear3": {"1": nn.Linear(2 * f, f)}, "Layer_scale": LayerScaleParam(f)}
    return {"block": {"cla": cla, "gcfn": _gcfn(f)}}


def _enc_stage(s: SeparatorShape, down: bool) -> dict:   # module.py:59-86
    f = s.feat
    t = {"g_block_1": _global_block(f), "l_block_1": _local_block(f, s.cla_kernel),
         "g_block_2": _global_block(f), "l_block_2": _local_block(f, s.cla_kernel)}
    if down:
        t["downconv"] = {"down_conv": nn.Conv1d(f, f, s.down_kernel, stride=2,
                                                padding=(s.down_kernel - 1) // 2, groups=f),
                         "BN": nn.BatchNorm1d(f)}
    return t


def _spk_split(s: SeparatorShape) -> dict:   # module.py:110-118
    f, n = s.feat, s.num_spks
    return {"linear": {"0": nn.Conv1d(f, 4 * f * n, 1), "2": nn.Conv1d(2 * f * n, f * n, 1)},
            "norm": nn.GroupNorm(1, f, eps=1e-8)}


def _dec_stage(s: SeparatorShape) -> dict:   # module.py:127-143
    f = s.feat
    t = {}
    for n in (1, 2, 3):
        t[f"g_block_{n}"] = _global_block(f)
        t[f"l_block_{n}"] = _local_block(f, s.cla_kernel)
        t[f"spk_attn_{n}"] = {"self_attn": _mha(f), "feed_forward": _gcfn(f)}
    return t


def separator_spec(s: SeparatorShape) -> dict:
    """The whole tree, in the reference's registration order (module.py:172-188)."""
    r = s.num_stages
    t = {"pos_emb": {"pe_k": nn.Embedding(2 * s.maxlen, s.dk)},
         "enc_stages": {str(i): _enc_stage(s, True) for i in range(r)},
         "bottleneck_G": _enc_stage(s, False)}
    if s.per_stage_split:       # SepReformer_Large_DM_WHAM/modules/module.py:182-184
        t["spk_split_blocks"] = {str(i): _spk_split(s) for i in range(r + 1)}
    else:
        t["spk_split_block"] = _spk_split(s)
    t["simple_fusion"] = {str(i): nn.Conv1d(2 * s.feat, s.feat, 1) for i in range(r)}
    t["dec_stages"] = {str(i): _dec_stage(s) for i in range(r)}
    return t


# --------------------------------------------------------------------------- seeded weights
def seeded_state(shapes: Iterable[Tuple[str, torch.Size, torch.dtype]], seed: int = 1) -> Dict[str, torch.Tensor]:
    """Deterministic, *non-trivial* values for every entry of a separator ``state_dict``.

    At default init every transformer branch is scaled by LayerScale=1e-5 and BatchNorm statistics are
    (0,1), so a wrong kernel would still pass a 1e-3 test.  Parity runs therefore use
    these values instead: LayerScale ~ 0.3*(0.5+U), BN mean ~ N(0,0.1), BN var ~ U(0.5,1.5), norm gains
    ~ 1+0.1N, norm shifts ~ 0.1N, matrices/filters ~ U(+-1/sqrt(fan_in)) as torch would, embedding ~ N(0,1).
    Reproducible on any machine from ``seed`` alone, so golden vectors need not ship 56 MB of weights.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape, dtype in shapes:
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros(shape, dtype=dtype)
            continue
        u = lambda: torch.rand(shape, generator=g)
        n = lambda: torch.randn(shape, generator=g)
        parent = name.rsplit(".", 1)[0]
        is_norm = len(shape) == 1 and parent.endswith(
            ("layer_norm", "net1.0", ".BN", ".norm", ".ega.block.linear.0"))
        if leaf == "layer_scale":
            v = (u() + 0.5) * 0.3
        elif leaf == "running_mean":
            v = 0.1 * n()
        elif leaf == "running_var":
            v = 0.5 + u()
        elif name.startswith("pos_emb."):
            v = n()
        elif is_norm and leaf == "weight":
            v = 1.0 + 0.1 * n()
        elif is_norm and leaf == "bias":
            v = 0.1 * n()
        else:
            if leaf == "weight":
                fan_in = int(torch.tensor(shape[1:]).prod()) if len(shape) > 1 else int(shape[0])
            else:   # bias of a matrix / filter: torch uses the weight's fan_in; approximate by own length
                fan_in = max(int(shape[0]) // 4, 8)
            bound = 1.0 / fan_in ** 0.5
            v = (2.0 * u() - 1.0) * bound
        out[name] = v.to(dtype)
    return out


def state_shapes(module: nn.Module):
    return [(k, v.shape, v.dtype) for k, v in module.state_dict().items()]

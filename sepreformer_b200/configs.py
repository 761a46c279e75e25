"""Constructor kwargs of the four SepReformer model directories, restated.

The reference keeps one ``configs.yaml`` per model directory and splats the
``model.module_separator`` subtree into ``Separator(**kwargs)``
(reference ``models/SepReformer_Base_WSJ0/configs.yaml:46-83``, ``model.py:27``).
The GPU box has no ``/root/reference``, so the same kwargs are rebuilt here from
the few numbers that actually differ between the model directories.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class SeparatorShape:
    """The numbers a separator is specialised on (SURVEY.md section 8 notation)."""

    feat: int = 128            # F
    heads: int = 8             # H
    num_stages: int = 4        # R
    num_spks: int = 2
    cla_kernel: int = 65
    down_kernel: int = 5
    maxlen: int = 2000
    per_stage_split: bool = False   # Large_DM_WHAM keeps one SpkSplitStage per stage

    @property
    def dk(self) -> int:
        return self.feat // self.heads

    @property
    def chunk(self) -> int:
        return 2 ** self.num_stages


MODEL_SHAPES = {
    "SepReformer_Base_WSJ0": SeparatorShape(feat=128),
    "SepReformer_Large_DM_WSJ0": SeparatorShape(feat=256),
    "SepReformer_Large_DM_WHAMR": SeparatorShape(feat=256),
    "SepReformer_Large_DM_WHAM": SeparatorShape(feat=256, per_stage_split=True),
}


def separator_kwargs(shape: SeparatorShape, dropout_rate: float = 0.05) -> dict:
    """Build the ``module_separator`` kwargs dict (configs.yaml:46-83) for ``shape``."""
    f, h = shape.feat, shape.heads
    gb = dict(in_channels=f, num_mha_heads=h, dropout_rate=dropout_rate)
    lb = dict(in_channels=f, kernel_size=shape.cla_kernel, dropout_rate=dropout_rate)
    return dict(
        num_stages=shape.num_stages,
        relative_positional_encoding=dict(in_channels=f, num_heads=h, maxlen=shape.maxlen, embed_v=False),
        enc_stage=dict(global_blocks=dict(gb), local_blocks=dict(lb),
                       down_conv_layer=dict(in_channels=f, samp_kernel_size=shape.down_kernel)),
        spk_split_stage=dict(in_channels=f, num_spks=shape.num_spks),
        simple_fusion=dict(out_channels=f),
        dec_stage=dict(num_spks=shape.num_spks, global_blocks=dict(gb), local_blocks=dict(lb),
                       spk_attention=dict(gb)),
    )


def shape_from_kwargs(num_stages, relative_positional_encoding, enc_stage, spk_split_stage,
                      simple_fusion, dec_stage, per_stage_split=False) -> SeparatorShape:
    """Recover a :class:`SeparatorShape` from reference-style kwargs, validating what the kernels assume."""
    f = int(enc_stage["global_blocks"]["in_channels"])
    h = int(enc_stage["global_blocks"]["num_mha_heads"])
    shape = SeparatorShape(
        feat=f, heads=h, num_stages=int(num_stages), num_spks=int(spk_split_stage["num_spks"]),
        cla_kernel=int(enc_stage["local_blocks"]["kernel_size"]),
        down_kernel=int(enc_stage["down_conv_layer"]["samp_kernel_size"]),
        maxlen=int(relative_positional_encoding["maxlen"]), per_stage_split=bool(per_stage_split))
    for name, sub in (("dec_stage.global_blocks", dec_stage["global_blocks"]),
                      ("dec_stage.spk_attention", dec_stage["spk_attention"])):
        if int(sub["in_channels"]) != f or int(sub["num_mha_heads"]) != h:
            raise ValueError(f"{name}: channel/head count differs from enc_stage ({sub})")
    if int(simple_fusion["out_channels"]) != f or int(spk_split_stage["in_channels"]) != f:
        raise ValueError("simple_fusion/spk_split_stage channels differ from enc_stage")
    if int(dec_stage["local_blocks"]["kernel_size"]) != shape.cla_kernel:
        raise ValueError("enc/dec CLA kernel sizes differ")
    if relative_positional_encoding.get("embed_v", False):
        raise ValueError("embed_v=True is not used by any reference config and is not implemented")
    if int(relative_positional_encoding["in_channels"]) // int(relative_positional_encoding["num_heads"]) != f // h:
        raise ValueError("relative positional embedding width differs from head width")
    return shape

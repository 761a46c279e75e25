"""CPU-side checks of the drop-in boundary: the library loads, exports what include/sepref.h declares, the
module surface mirrors the reference's, and the product refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

import sepreformer_b200
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs
from sepreformer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sepref.h")).read()
    declared = sorted(set(re.findall(r"\b(sepref_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.sepref_version()


def test_header_cites_reference_for_each_forward_entry():
    header = open(os.path.join(ROOT, "include", "sepref.h")).read()
    for name in _lib.EXPORTS:
        if name.endswith("_forward"):
            idx = header.index("int " + name)
            assert re.search(r"modules/(network|module)\.py:\d+", header[max(0, idx - 600):idx]), name


def test_create_rejects_bad_config_and_missing_gpu():
    L = _lib.lib()
    p = C.c_void_p()
    bad = _lib.SeprefConfig(96, 8, 4, 2, 65, 5, 2000, 0)
    assert L.sepref_create(C.byref(bad), 0, C.byref(p)) == -1
    assert b"feat=96" in L.sepref_last_error()
    if not torch.cuda.is_available():
        ok = _lib.SeprefConfig(128, 8, 4, 2, 65, 5, 2000, 0)
        assert L.sepref_create(C.byref(ok), 0, C.byref(p)) == -3
        assert b"no CPU path" in L.sepref_last_error()


@pytest.mark.parametrize("name", sorted(MODEL_SHAPES))
def test_module_surface(name):
    shape = MODEL_SHAPES[name]
    m = Separator(**separator_kwargs(shape), per_stage_split=shape.per_stage_split)
    sd = m.state_dict()
    f = shape.feat
    assert sd["pos_emb.pe_k.weight"].shape == (4000, f // 8)
    assert sd["enc_stages.3.l_block_2.block.cla.dw_conv_1d.weight"].shape == (f, 1, 65)
    assert sd["dec_stages.2.spk_attn_3.feed_forward.net2.2.weight"].shape == (f, 3 * f)
    assert ("spk_split_blocks.4.norm.weight" in sd) == shape.per_stage_split
    assert ("spk_split_block.norm.weight" in sd) != shape.per_stage_split
    n_params = sum(v.numel() for k, v in sd.items() if v.is_floating_point() and "running" not in k)
    # separator parameters 13 974 528 (Base) / 54 753 280 (Large, shared split)
    if name == "SepReformer_Base_WSJ0":
        assert n_params == 13_974_528
    if name in ("SepReformer_Large_DM_WSJ0", "SepReformer_Large_DM_WHAMR"):
        assert n_params == 54_753_280
    assert m.padded_frames(7997) == 8000 and m.padded_frames(8000) == 8000


def test_no_cpu_fallback():
    m = Separator(**separator_kwargs(MODEL_SHAPES["SepReformer_Base_WSJ0"]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 128, 32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sepreformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.lower().replace("no oracle", ""), f"{fn} mentions the oracle"


def test_option_constants_match_the_header():
    """The ctypes binding's option ids are the header's SEPREF_OPT_* values (no silent drift)."""
    import re
    from sepreformer_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sepref.h")).read()
    defs = dict(re.findall(r"#define\s+SEPREF_(OPT_[A-Z_]+)\s+(\d+)", hdr))
    assert defs, "no SEPREF_OPT_* definitions found"
    for name, value in defs.items():
        assert getattr(_lib, name) == int(value), name


def test_forward_refuses_training_mode_and_grad_inputs():
    """the module is inference-only; silent no-grad training behind a drop-in surface is worse than an error.
    (The checks sit in front of the device check, so they are testable without a GPU.)"""
    m = Separator(**separator_kwargs(MODEL_SHAPES["SepReformer_Base_WSJ0"]))
    assert m.training
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 128, 32))                       # device check comes first for CPU tensors
    src = open(os.path.join(ROOT, "sepreformer_b200", "separator.py")).read()
    assert "inference-only" in src and "requires grad" in src


def test_data_parallel_replicas_share_packed_handles_and_owner_weights():
    """torch.nn.parallel.replicate() hands forward() shallow copies whose _parameters are empty: the
    replicas must share one handle table / lock and pack from the owner's state_dict, not from their own."""
    import copy
    m = Separator(**separator_kwargs(MODEL_SHAPES["SepReformer_Base_WSJ0"])).eval()
    rep = m._replicate_for_data_parallel()
    assert rep._is_replica and len(rep._parameters) == 0        # (replicate() does the same to every child module)
    assert rep._sh() is m._sh() and rep._sh().master() is m
    assert rep._sh().handles is m._sh().handles
    e0 = m._sh().epoch
    m.load_state_dict(m.state_dict())
    assert m._sh().epoch > e0 and rep._sh().epoch == m._sh().epoch   # a weight change is visible to every replica
    e1 = m._sh().epoch
    holder = torch.nn.ModuleDict({"separator": m})                   # a parent's load_state_dict never calls ours:
    holder.load_state_dict(holder.state_dict())                      # the post-hook still sees it
    assert m._sh().epoch > e1
    e2 = m._sh().epoch
    m.float()
    assert m._sh().epoch > e2
    clone = copy.deepcopy(m)
    assert clone._sh() is not m._sh() and clone._sh().master() is clone
    import pickle
    again = pickle.loads(pickle.dumps(m))
    assert again._sh().master() is again and again._sh().handles == {}

"""sepreformer_b200 - B200-native (sm_100a) separator hot path of SepReformer behind the reference's module surface.

``Separator`` mirrors ``models/SepReformer_*/modules/module.py:Separator`` (constructor kwargs, state_dict keys,
forward contract); its forward is one call into ``libsepref_b200.so`` (C ABI in ``include/sepref.h``).
"""
from .configs import MODEL_SHAPES, SeparatorShape, separator_kwargs, shape_from_kwargs  # noqa: F401
from .separator import Separator  # noqa: F401
from .model import Model  # noqa: F401


def install(model_module, per_stage_split: bool = False, level: str = "separator"):
    """Make a reference model package build the B200 code: ``install(models.SepReformer_Base_WSJ0.model)``.

    ``level="separator"`` rebinds the ``Separator`` symbol that ``Model.__init__`` looks up (reference ``model.py:27``):
    the reference's encoder / output layer / decoder stay PyTorch modules around the CUDA separator.
    ``level="model"`` rebinds ``Model`` itself (looked up by ``main.py:5,29``, so install before importing ``main``): the
    whole mixture -> waveforms path is one C-ABI call.  ``run.py``, ``main.py``, ``engine.py`` and the configs stay untouched.
    """
    import functools
    if level == "model":
        model_module.Model = functools.partial(Model, per_stage_split=True) if per_stage_split else Model
    elif level == "separator":
        model_module.Separator = functools.partial(Separator, per_stage_split=True) if per_stage_split else Separator
    else:
        raise ValueError("level must be 'separator' or 'model'")
    return model_module

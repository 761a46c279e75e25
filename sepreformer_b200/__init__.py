"""sepreformer_b200 - B200-native (sm_100a) separator hot path of SepReformer behind the reference's module surface.

``Separator`` mirrors ``models/SepReformer_*/modules/module.py:Separator`` (constructor kwargs, state_dict keys,
forward contract); its forward is one call into ``libsepref_b200.so`` (C ABI in ``include/sepref.h``).
"""
from .configs import MODEL_SHAPES, SeparatorShape, separator_kwargs, shape_from_kwargs  # noqa: F401
from .separator import Separator  # noqa: F401


def install(model_module, per_stage_split: bool = False):
    """Make a reference model package build the B200 separator: ``install(models.SepReformer_Base_WSJ0.model)``.

    Rebinds the ``Separator`` symbol that ``Model.__init__`` looks up (reference ``model.py:27``); ``run.py``,
    ``main.py``, ``engine.py`` and the configs stay untouched.
    """
    if per_stage_split:
        import functools
        model_module.Separator = functools.partial(Separator, per_stage_split=True)
    else:
        model_module.Separator = Separator
    return model_module

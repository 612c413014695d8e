"""Model registry, the reference's plugin point (code/register.py:40-47): MODELS['lgn'] is looked up
as MODELS[world.model_name](world.config, dataset).  Unlike the reference, importing this module
does not build a dataset; `install(reference_register_module)` swaps the B200 model into the
reference's own registry."""
from .model import LightGCN

MODELS = {'lgn': LightGCN}


def install(reference_register=None):
    """register.MODELS['lgn'] = lgcn_b200.LightGCN for the reference's own training loop."""
    if reference_register is None:
        import register as reference_register      # the reference's module, must be importable
    reference_register.MODELS['lgn'] = LightGCN
    return reference_register

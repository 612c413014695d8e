"""lgcn_b200 — B200-native LightGCN propagation / BPR training / full-ranking evaluation hot path
behind the Python API of saamiya225/Graph-and-sequential-recommendation-systems (LightGCN_work/code).

    import lgcn_b200 as lg
    lg.world.configure(latent_dim_rec=64, lightGCN_n_layers=3)
    ds = lg.Loader(lg.world.config, path='data/gowalla')
    model = lg.LightGCN(lg.world.config, ds)
    bpr = lg.utils.BPRLoss(model, lg.world.config)
    lg.Procedure.BPR_train_original(ds, model, bpr, epoch=0)
    lg.Procedure.Test(ds, model, epoch=0)

All device work is done by liblgcn_b200.so (csrc/, C ABI in include/lgcn_b200.h); importing this
package without the built library raises at first use — there is no fallback implementation.
"""
from . import _lib, world            # noqa: F401
from . import ops, utils, sampling   # noqa: F401
from . import dataloader, synth      # noqa: F401
from .dataloader import BasicDataset, InteractionDataset, Loader     # noqa: F401
from . import engine, model          # noqa: F401
from .model import LightGCN          # noqa: F401
from . import Procedure, register    # noqa: F401
from .register import MODELS         # noqa: F401

__all__ = ['world', 'ops', 'utils', 'dataloader', 'synth', 'engine', 'model', 'Procedure', 'register',
           'BasicDataset', 'InteractionDataset', 'Loader', 'LightGCN', 'MODELS']

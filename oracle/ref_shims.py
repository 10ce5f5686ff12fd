"""ORACLE SUPPORT (test infrastructure): import the UNMODIFIED reference modules with in-memory
shims for the five packages the image lacks.  None of the shims changes
arithmetic.  Two locations are tried: /root/reference (this container: the sources where they lie)
and `oracle/_ref/` (the GPU box: sourceless byte-code compiled from those sources by
`oracle/build_ref.py`; git-ignored, travels like the built `.so`; loaded module by module with SourcelessFileLoader).  Used by
`tests/golden/make_golden.py`, by the tests that run the reference's own classes, and by
`bench.py --impl reference` / its `cpu_baseline` leg.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch
from torch import nn

SRC_ROOT = "/root/reference/Unet_research"
_HERE = os.path.dirname(os.path.abspath(__file__))
PYC_ROOT = os.path.join(_HERE, "_ref")


def _pyc_complete() -> bool:
    from . import build_ref
    return build_ref.built()


def reference_location():
    """("source" | "bytecode" | None, root directory that contains unet_code/)."""
    if os.path.isdir(os.path.join(SRC_ROOT, "unet_code")):
        return "source", SRC_ROOT
    if _pyc_complete():
        return "bytecode", PYC_ROOT
    return None, None


def reference_available() -> bool:
    return reference_location()[0] is not None


REF_ROOT = reference_location()[1] or SRC_ROOT
REF_CODE = os.path.join(REF_ROOT, "unet_code")


def _install_shims():
    if not hasattr(np, "product"):
        np.product = np.prod                      # utils_modules.py:5 (removed in numpy 2)

    if "dropblock" not in sys.modules:
        m = types.ModuleType("dropblock")

        class LinearScheduler(nn.Module):
            """dropblock==0.3.0 scheduler restated (PARITY UNPINNED, see unet_oracle)."""

            def __init__(self, dropblock, start_value, stop_value, nr_steps):
                super().__init__()
                self.dropblock = dropblock
                self.i = 0
                self.drop_values = np.linspace(start=start_value, stop=stop_value, num=int(nr_steps))

            def forward(self, x):
                return self.dropblock(x)

            def step(self):
                if self.i < len(self.drop_values):
                    self.dropblock.drop_prob = self.drop_values[self.i]
                self.i += 1

        m.LinearScheduler = LinearScheduler
        sys.modules["dropblock"] = m

    if "fairscale" not in sys.modules:
        fs = types.ModuleType("fairscale")
        fsnn = types.ModuleType("fairscale.nn")
        fsnn.checkpoint_wrapper = lambda mod, *a, **k: mod      # identity: values unchanged
        fs.nn = fsnn
        sys.modules["fairscale"] = fs
        sys.modules["fairscale.nn"] = fsnn

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def log(self, *a, **k):
                pass

            @classmethod
            def load_from_checkpoint(cls, path, **kw):
                obj = cls(**kw)
                obj.load_state_dict(torch.load(path, map_location="cpu")["state_dict"])
                return obj

        class Trainer:
            def __init__(self, *a, **k):
                pass

            @staticmethod
            def add_argparse_args(p):
                return p

            @classmethod
            def from_argparse_args(cls, *a, **k):
                return cls()

        def seed_everything(seed, workers=False):
            import random
            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)
            return seed

        pl.LightningModule = LightningModule
        pl.Trainer = Trainer
        pl.seed_everything = seed_everything
        cb = types.ModuleType("pytorch_lightning.callbacks")
        cb.ModelCheckpoint = cb.EarlyStopping = cb.LearningRateMonitor = lambda *a, **k: None
        pl.callbacks = cb
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.callbacks"] = cb

    class _Fake(types.ModuleType):
        """Attribute sink: plotting / sklearn entry points the oracle never calls."""

        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return lambda *a, **k: None

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
                 "pandas", "sklearn", "sklearn.metrics"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Fake(name)


_LOADED = {}


def load_reference(prefer: str = None):
    """Returns a namespace with the reference's own classes: UNet, DropBlock2D,
    Dropblock2d_ichan, LinearScheduler, BaseUNetTraining, DropBlockEval, set_dropblock_on,
    RotationEval, UNetTraining (+ `.kind`: "source" or "bytecode", and the script modules `.mod_db`,
    `.mod_rot`, `.mod_train` for import-swap tests).  prefer="bytecode" forces oracle/_ref."""
    kind, root = reference_location()
    if prefer == "bytecode" and _pyc_complete():
        kind, root = "bytecode", PYC_ROOT
    if kind is None:
        raise RuntimeError("reference not available: neither " + SRC_ROOT + " nor a complete " + PYC_ROOT)
    if kind in _LOADED:
        return _LOADED[kind]
    if _LOADED:
        raise RuntimeError("the reference is already loaded from the other location in this process")
    code = os.path.join(root, "unet_code")
    _install_shims()
    cwd = os.getcwd()
    if kind == "source":
        if code not in sys.path:
            sys.path.insert(0, code)
        os.chdir(root)          # scripts do sys.path.append(os.getcwd() + '/unet_code')
        try:
            from utils import utils_unet, utils_modules, utils_training  # type: ignore

            def load_script(name, rel):
                spec = importlib.util.spec_from_file_location(name, os.path.join(code, rel))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                return mod

            db = load_script("ref_dropblock_uncertainty", "uncertainty_tests/Dropblock_Uncertainty.py")
            rot = load_script("ref_rotational_uncertainty", "uncertainty_tests/Rotational_Uncertainty.py")
            tr = load_script("ref_training", "base_model_tests/training.py")
        finally:
            os.chdir(cwd)
    else:
        # byte-code: no path finder involved -- every module is loaded explicitly, in dependency order, under the name
        # the reference's own `from utils.utils_x import ...` statements expect
        from importlib.machinery import SourcelessFileLoader
        from . import build_ref

        def load_bc(name, rel):
            path = os.path.join(code, rel[:-3] + build_ref.EXT)
            loader = SourcelessFileLoader(name, path)
            spec = importlib.util.spec_from_loader(name, loader, origin=path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            loader.exec_module(mod)
            return mod

        if "utils" not in sys.modules:
            pkg = types.ModuleType("utils")
            pkg.__path__ = []                       # a package without a search path: only what is registered below
            sys.modules["utils"] = pkg
        mods = {}
        for leaf in ("utils_modules", "utils_unet", "utils_training", "utils_general", "utils_dataset", "utils_metrics"):
            mods[leaf] = load_bc("utils." + leaf, f"utils/{leaf}.py")
            setattr(sys.modules["utils"], leaf, mods[leaf])
        utils_unet, utils_modules, utils_training = mods["utils_unet"], mods["utils_modules"], mods["utils_training"]
        db = load_bc("ref_dropblock_uncertainty", "uncertainty_tests/Dropblock_Uncertainty.py")
        rot = load_bc("ref_rotational_uncertainty", "uncertainty_tests/Rotational_Uncertainty.py")
        tr = load_bc("ref_training", "base_model_tests/training.py")
    ns = types.SimpleNamespace(
        UNet=utils_unet.UNet, DropBlock2D=utils_modules.DropBlock2D,
        Dropblock2d_ichan=utils_modules.Dropblock2d_ichan, LinearScheduler=utils_modules.LinearScheduler,
        BaseUNetTraining=utils_training.BaseUNetTraining, DropBlockEval=db.DropBlockEval,
        set_dropblock_on=db.set_dropblock_on, RotationEval=rot.RotationEval, UNetTraining=tr.UNetTraining, kind=kind,
        mod_db=db, mod_rot=rot, mod_train=tr, mod_unet=utils_unet, mod_modules=utils_modules, mod_training=utils_training)
    _LOADED[kind] = ns
    return ns


def build_reference_unet(ref, init_channels=1, filters=64, dropblock=None, drop_prob=0.15, block_size=7,
                         use_scheduler=False, num_groups=32, **sched):
    """The canonical configuration of R/base_model_tests/training.py:171-192."""
    unet = ref.UNet(init_channels=init_channels, filters=filters, output_channels=1, model_depth=4,
                    pool_mode="max", up_mode="upconv", connection="cat", same_padding=True,
                    conv_layers_per_block=2, checkpointing=True)
    unet.set_activation_function(nn.ReLU())
    if dropblock is not None:
        unet.set_dropblock(dropblock, block_size=block_size, drop_prob=drop_prob, use_scheduler=use_scheduler, **sched)
    unet.set_normalization(nn.GroupNorm, params={"num_groups": num_groups, "num_channels": "fill"})
    unet.create_model()
    return unet

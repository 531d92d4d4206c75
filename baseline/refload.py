"""Load the UNMODIFIED reference model classes from a copy of the reference's `models/` directory.

This repository ships a drop-in `models/` package (a regular package with an __init__.py) under the same module paths as
the reference, and a regular package always wins over the reference's namespace directories whatever the order of
sys.path -- so `import models.X.model` from the repository root resolves to the engine.  The reference copy is therefore
mounted under a private top-level name; its relative imports (FastTransformer/model.py: `from .utils import ...`) keep working.
"""
import importlib
import os
import sys
import types

_ALIAS = "tu_reference_models"


def reference_model_class(models_dir: str, name: str):
    """models_dir = .../models of the reference (baseline/_ref/models or /root/reference/models); name e.g. 'WindowTransformer'."""
    models_dir = os.path.abspath(models_dir)
    pkg = sys.modules.get(_ALIAS)
    if pkg is None or list(getattr(pkg, "__path__", [])) != [models_dir]:
        for k in [k for k in sys.modules if k == _ALIAS or k.startswith(_ALIAS + ".")]:
            del sys.modules[k]
        pkg = types.ModuleType(_ALIAS)
        pkg.__path__ = [models_dir]
        sys.modules[_ALIAS] = pkg
    mod = importlib.import_module(f"{_ALIAS}.{name}.model")
    assert os.path.abspath(mod.__file__).startswith(models_dir), mod.__file__
    return mod.TransformerModel

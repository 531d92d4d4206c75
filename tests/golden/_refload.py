"""Import the UNMODIFIED reference classes from /root/reference (build container only).

This repository ships a drop-in `models/` package with the same module paths as the reference, so a plain
`import models.X.model` from the repository root resolves to the engine; the golden generators need the reference."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ref_model(model):
    """an instance of the reference's TransformerModel (eval mode), not of this repository's drop-in alias"""
    saved = list(sys.path)
    sys.path[:] = ["/root/reference"] + [q for q in saved if os.path.abspath(q or ".") != ROOT and q not in ("", ".")]
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    try:
        cls = importlib.import_module(f"models.{model}.model").TransformerModel
        assert cls.__module__.startswith("models.") and "/root/reference" in sys.modules[cls.__module__].__file__
    finally:
        sys.path[:] = saved
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
    return cls().eval()



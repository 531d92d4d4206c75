"""Drop-in ``TransformerModel`` classes: ``transformerupscaler_b200.models.<Name>.model.TransformerModel``
(mirrored at the repo root as ``models.<Name>.model`` for the reference's importlib convention)."""

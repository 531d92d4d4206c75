"""Drop-in alias: reference path models/WindowTransformer/model.py -> transformerupscaler_b200.models.WindowTransformer.model."""
from transformerupscaler_b200.models.WindowTransformer.model import TransformerModel  # noqa: F401

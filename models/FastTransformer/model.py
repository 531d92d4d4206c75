"""Drop-in alias: reference path models/FastTransformer/model.py -> transformerupscaler_b200.models.FastTransformer.model."""
from transformerupscaler_b200.models.FastTransformer.model import TransformerModel  # noqa: F401

"""Drop-in alias: reference path models/ResidualTransformer/model.py -> transformerupscaler_b200.models.ResidualTransformer.model."""
from transformerupscaler_b200.models.ResidualTransformer.model import TransformerModel  # noqa: F401

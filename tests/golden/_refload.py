"""Import the UNMODIFIED reference classes from /root/reference (build container only); see baseline/refload.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from baseline.refload import reference_model_class      # noqa: E402


def ref_model(model):
    """an instance of the reference's TransformerModel (eval mode), not of this repository's drop-in alias"""
    return reference_model_class("/root/reference/models", model)().eval()

"""`efficient_models` drop-in (the reference's package __init__ is empty)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_here)) if os.path.dirname(_here) not in sys.path else None
from _extend import extend_path  # noqa: E402

extend_path("efficient_models", __path__)

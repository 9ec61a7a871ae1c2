"""`models` drop-in (reference: /root/reference/models/__init__.py:1-3 exports XVLMBase, build_mlp, load_pretrained)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_here) not in sys.path:
    sys.path.insert(0, os.path.dirname(_here))
from _extend import extend_path  # noqa: E402

extend_path("models", __path__)
from models.xvlm import XVLMBase  # noqa: E402,F401
from models.xvlm import build_mlp  # noqa: E402,F401
from models.xvlm import load_pretrained  # noqa: E402,F401

"""Importable alias of the product package.

The product lives in ``gradient-boosted-normalizing-flows_b200/`` (the name the build contract fixes); a hyphen is
not a legal Python identifier, so this package only extends its search path to that directory:
``import gbnf_b200.boosted_flow`` loads ``gradient-boosted-normalizing-flows_b200/boosted_flow.py``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "gradient-boosted-normalizing-flows_b200")
__path__.append(_PKG_DIR)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401

"""Drop-in shim: ``import UNet`` resolves to the sm_100a implementation (put this directory first on
sys.path; the reference scripts' ``from UNet import ...`` lines then work unchanged)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from anoddpm_b200.UNet import *  # noqa: F401,F403,E402
from anoddpm_b200.UNet import __dict__ as _d  # noqa: E402

globals().update({k: v for k, v in _d.items() if k.startswith("_") and not k.startswith("__")})

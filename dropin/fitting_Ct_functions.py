"""Drop-in for SpinRelax's `fitting_Ct_functions` module: the same names, implemented by spinrelax_b200.fitct (device stages on the GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import fitct as _impl  # noqa: E402
from spinrelax_b200.fitct import *  # noqa: E402,F401,F403

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})

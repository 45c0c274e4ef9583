"""Drop-in for SpinRelax's `transforms3d_supplement` module: the same names, implemented by spinrelax_b200.qs (device stages on the GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import qs as _impl  # noqa: E402
from spinrelax_b200.qs import *  # noqa: E402,F401,F403

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})

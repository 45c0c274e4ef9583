#!/usr/bin/env python
"""Drop-in for SpinRelax's `calculate-Ct-from-traj.py` (same flags and output files): the work is done by spinrelax_b200.cli_ct on the GPU.
Put this directory where run-all.bash's $script_loc points (or copy these files over the reference's)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200.cli_ct import main  # noqa: E402

if __name__ == '__main__':
    main()

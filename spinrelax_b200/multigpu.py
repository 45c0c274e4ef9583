"""Single-process multi-GPU execution of the hot path (BASELINE configs 4 and 5).

Every stage of the path is a set of independent units -- bond vectors for C(t), the histogram and S2
(calculate-Ct-from-traj.py:222-228: one column per vector), residues for the fits and the relaxation grid, lags for
the dq moments -- so the public functions shard their units over the selected devices and stitch the results
together: one host thread per GPU (ctypes and the large NumPy copies release the GIL), each with its own current
device and stream, no data-path collective.  Stages that *do* end in a reduction (frames or replica trajectories
spread over ranks) use torch.distributed instead, see shard.py and pipeline.py.

Device selection: SPINRELAX_GPUS = "all" | N (the first N visible devices) | "0,2,3" (explicit list).  Unset = only
the current device, i.e. exactly the single-GPU behaviour.
"""
import os
from concurrent.futures import ThreadPoolExecutor

from . import _lib
from .shard import split_range

_OVERRIDE = None


def set_devices(devs):
    """Programmatic override of SPINRELAX_GPUS (None restores the environment's choice)."""
    global _OVERRIDE
    _OVERRIDE = None if devs is None else [int(d) for d in devs]


def devices():
    """Device indices the sharded stages may use (always at least the current device)."""
    torch = _lib.require_cuda()
    n = torch.cuda.device_count()
    if _OVERRIDE is not None:
        devs = list(_OVERRIDE)
    else:
        spec = os.environ.get("SPINRELAX_GPUS", "").strip().lower()
        if not spec:
            return [torch.cuda.current_device()]
        if spec == "all":
            devs = list(range(n))
        elif "," in spec:
            devs = [int(x) for x in spec.split(",") if x.strip()]
        else:
            devs = list(range(min(int(spec), n)))
    bad = [d for d in devs if d < 0 or d >= n]
    if bad or not devs:
        raise _lib.SpinRelaxError("SPINRELAX_GPUS selects devices %s but %d are visible" % (devs, n))
    return devs


def plan(n_units, min_per_device=1):
    """[(device, first, last)] contiguous balanced blocks of range(n_units) over the selected devices; a device gets a
    block only if every block can hold at least `min_per_device` units."""
    devs = devices()
    k = max(1, min(len(devs), n_units // max(1, min_per_device)))
    out = []
    for r in range(k):
        a, b = split_range(n_units, k, r)
        if b > a:
            out.append((devs[r], a, b))
    return out


def run(blocks, fn):
    """fn(device, first, last) for every block, each on its own host thread with that device current; results in
    block order.  A single block runs inline on the calling thread."""
    torch = _lib.require_cuda()
    if len(blocks) == 1:
        d, a, b = blocks[0]
        with torch.cuda.device(d):
            return [fn(d, a, b)]

    def work(blk):
        d, a, b = blk
        with torch.cuda.device(d):
            return fn(d, a, b)

    with ThreadPoolExecutor(max_workers=len(blocks)) as pool:
        return list(pool.map(work, blocks))

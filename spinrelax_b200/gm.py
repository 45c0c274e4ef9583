"""Mirror of the part of the reference's ``general_maths`` module that the trajectory path calls.

``xyz_to_rtp`` (general_maths.py:118-158) is applied by calculate-Ct-from-traj.py:588 to the whole
(frames, bonds, 3) array before the ``_vecPhiTheta`` outputs; it runs in ``sr_xyz_to_rtp_f32/_f64`` in the
precision of its input, as NumPy does.  ``rtp_to_xyz`` (:160-208) is only ever applied to the 72 x 36 bin
centres of a histogram (calculate-relaxations-from-Ct.py:412) and is evaluated on the host.
"""
import sys

import numpy as np

from . import _lib


def _perturb_tuple(t, mod, axis):
    l = list(t)
    l[axis] += mod
    return tuple(l)


def xyz_to_rtp_device(v, bUnit=False):
    """torch CUDA tensor (..., 3), float32 or float64, contiguous -> (..., 3) [or (..., 2) if bUnit]."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    if v.dtype not in (torch.float32, torch.float64) or v.shape[-1] != 3 or not v.is_contiguous():
        raise ValueError("xyz_to_rtp_device: need a contiguous float32/float64 (..., 3) CUDA tensor")
    out = torch.empty(v.shape[:-1] + ((2,) if bUnit else (3,)), dtype=v.dtype, device=v.device)
    if v.numel():
        fn = lib.sr_xyz_to_rtp_f32 if v.dtype == torch.float32 else lib.sr_xyz_to_rtp_f64
        _lib.check(fn(v.data_ptr(), v.numel() // 3, out.data_ptr(), 1 if bUnit else 0, _lib.current_stream_ptr()),
                   "sr_xyz_to_rtp")
    return out


def xyz_to_rtp(uv, vaxis=-1, bUnit=False):
    """gm.xyz_to_rtp: X/Y/Z -> R/Phi/Theta with 0 <= Theta <= pi from +Z; vaxis is -1 (last) or 0 (first).
    The bUnit form returns (Phi, arccos(z / Phi)) exactly as shipped (:131-133)."""
    uv = np.asarray(uv)
    if vaxis not in (-1, 0) and uv.ndim > 1:
        print("= = ERROR encountered in vec-to-rtp in general_maths.py, vaxis only accepts arguments of -1 or 0 for now.",
              file=sys.stderr)
        raise UnboundLocalError("rtp")          # the reference falls through to `return rtp` unassigned
    torch = _lib.require_cuda()
    dt = uv.dtype if uv.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    src = uv if (vaxis == -1 or uv.ndim == 1) else np.moveaxis(uv, 0, -1)
    src = np.ascontiguousarray(src, dtype=dt)
    out = xyz_to_rtp_device(torch.from_numpy(src).cuda(), bUnit).cpu().numpy()
    if vaxis == 0 and uv.ndim > 1:
        out = np.ascontiguousarray(np.moveaxis(out, -1, 0))
    return out.astype(uv.dtype, copy=False) if uv.dtype.kind == 'f' else out


def rtp_to_xyz(rtp, vaxis=-1, bUnit=False):
    """gm.rtp_to_xyz (:160-208); bUnit expects (Phi, Theta) only.  Host-side: its callers pass the bin centres
    of one histogram (calculate-relaxations-from-Ct.py:412).  In the non-unit array forms the shipped code scales
    by ``rtp[0]`` (the first entry along axis 0, not the R column, :197-203); that is kept."""
    rtp = np.asarray(rtp)
    if rtp.ndim > 1 and vaxis not in (-1, 0):
        print("= = ERROR encountered in rtp-to-vec in general_maths.py, vaxis only accepts arguments of -1 or 0 for now.",
              file=sys.stderr)
        raise UnboundLocalError("uv")
    last = rtp.ndim == 1 or vaxis == -1
    comp = (lambda a, k: a[..., k]) if last else (lambda a, k: a[k, ...])
    k0 = 0 if bUnit else 1
    phi, theta = comp(rtp, k0), comp(rtp, k0 + 1)
    scale = 1 if bUnit else rtp[0]
    xyz = (scale * np.cos(phi) * np.sin(theta), scale * np.sin(phi) * np.sin(theta), scale * np.cos(theta)) if not bUnit \
        else (np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta))
    return np.stack(xyz, axis=-1 if last else 0).astype(rtp.dtype, copy=False)

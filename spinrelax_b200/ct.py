"""Host mirror of the C(t) functions of calculate-Ct-from-traj.py, backed by the CUDA path.

Same names, argument meaning and return layout as the reference:
  calculate_Ct_Palmer   calculate-Ct-from-traj.py:200-238
  reformat_vecs_by_tau  calculate-Ct-from-traj.py:245-275
  calculate_dt          calculate-Ct-from-traj.py:240-243
All numerics run in libspinrelax_b200.so (sr_ct_palmer_*); nothing here computes C(t) on the CPU.
"""
import ctypes
import sys

import numpy as np

from . import _lib


def calculate_dt(dt, tau):
    """Time axis of the C(t) points (calculate-Ct-from-traj.py:240-243): (1..int(0.5*tau/dt)) * dt."""
    nPts = int(0.5 * tau / dt)
    return (np.arange(nPts) + 1.0) * dt


def reformat_vecs_by_tau(vecs, dt, tau):
    """Cut every trajectory to a whole number of tau-long chunks and stack them.

    vecs: list (or array) of per-file arrays (frames, bonds, 3); frame counts may differ between files
    (the reference's own np.array(list) at :498 cannot hold ragged input on NumPy >= 1.24, the function
    at :245-275 can).  Returns (nChunk, int(tau/dt), bonds, 3) with the dtype of the first file.
    """
    nFiles = len(vecs)
    nFramesPerChunk = int(tau / dt)
    if nFramesPerChunk < 1:
        raise ValueError("reformat_vecs_by_tau: tau/dt < 1 frame per chunk")
    used = [int(v.shape[0] / nFramesPerChunk) * nFramesPerChunk for v in vecs]
    nFramesTot = int(sum(used))
    first = vecs[0]
    out = np.zeros((nFramesTot, first.shape[1], first.shape[2]), dtype=first.dtype)
    start = 0
    for i in range(nFiles):
        end = start + used[i]
        out[start:end] = vecs[i][: used[i]]
        start = end
    return out.reshape((nFramesTot // nFramesPerChunk, nFramesPerChunk, first.shape[1], first.shape[2]))


def _check_4d(vecs):
    sh = vecs.shape
    if len(sh) != 4 or sh[-1] != 3:
        # reference: message to stderr + sys.exit(1) (calculate-Ct-from-traj.py:213-216)
        print("= = = ERROR: The input vectors to calculate_Ct_Palmer is not of the expected "
              "4-dimensional form! %s" % (sh,), file=sys.stderr)
        sys.exit(1)
    if sh[1] < 50:
        print("= = = WARNING: there are less than 50 frames per block of memory-time!", file=sys.stderr)
    return sh


def ct_palmer_device(vecs, workspace=None, return_workspace=False):
    """C(t), dC(t) from a CUDA float32 tensor (nC, nF, nR, 3). Returns two (nF//2, nR) float32 tensors."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    if not (vecs.is_cuda and vecs.dtype == torch.float32 and vecs.is_contiguous()):
        raise _lib.SpinRelaxError("ct_palmer_device: need a contiguous float32 CUDA tensor")
    nC, nF, nR, _ = _check_4d(vecs)
    L = nF // 2
    need = lib.sr_ct_workspace_bytes(nC, nF, nR)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=vecs.device)
    Ct = torch.empty((L, nR), dtype=torch.float32, device=vecs.device)
    dCt = torch.empty((L, nR), dtype=torch.float32, device=vecs.device)
    rc = lib.sr_ct_palmer_device(vecs.data_ptr(), nC, nF, nR, Ct.data_ptr(), dCt.data_ptr(),
                                 workspace.data_ptr(), workspace.numel(), _lib.current_stream_ptr())
    _lib.check(rc, "sr_ct_palmer_device")
    if return_workspace:
        return Ct, dCt, workspace
    return Ct, dCt


def ct_lag_sums_device(vecs):
    """Raw FP64 lag sums S[nR, nC, L] = sum_t (u(t).u(t+delta))^2 (test / diagnostics entry)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    nC, nF, nR, _ = _check_4d(vecs)
    L = nF // 2
    pitch = lib.sr_ct_row_pitch(nF)
    packed = torch.empty((nR, nC, 3, pitch), dtype=torch.float32, device=vecs.device)
    S = torch.empty((nR, nC, L), dtype=torch.float64, device=vecs.device)
    st = _lib.current_stream_ptr()
    _lib.check(lib.sr_pack_vectors_f32(vecs.data_ptr(), nC, nF, nR, None, packed.data_ptr(), pitch, st),
               "sr_pack_vectors_f32")
    _lib.check(lib.sr_ct_lag_sums(packed.data_ptr(), pitch, nC, nF, nR, L, S.data_ptr(), st), "sr_ct_lag_sums")
    return S


def calculate_Ct_Palmer_quiet(vecs):
    """calculate_Ct_Palmer without the reference's debug print."""
    return calculate_Ct_Palmer(vecs, _verbose=False)


def calculate_Ct_Palmer(vecs, _verbose=True):
    """Drop-in for calculate_Ct_Palmer (calculate-Ct-from-traj.py:200-238).

    vecs: (nReplicates, nFrames, nResidues, 3) array.  Returns (Ct, dCt), each (nFrames//2, nResidues)
    with vecs' dtype; first row is delta = 1.  Input is taken through the host-buffer C ABI
    (sr_ct_palmer_host): H2D, CUDA kernels, D2H.  float64 input is computed from its float32 rounding
    (the reference's own pipeline only ever produces float32 vectors, obtain_XHvecs :64-86).
    """
    vecs = np.asarray(vecs)
    sh = _check_4d(vecs)
    _lib.require_cuda()
    lib = _lib.load()
    if _verbose:
        print("= = = Debug of calculate_Ct_Palmer confirming the dimensions of vecs:", sh)
    out_dtype = vecs.dtype if vecs.dtype in (np.float32, np.float64) else np.float32
    v32 = np.ascontiguousarray(vecs, dtype=np.float32)
    nC, nF, nR, _ = sh
    L = int(nF / 2)
    Ct = np.empty((L, nR), dtype=np.float32)
    dCt = np.empty((L, nR), dtype=np.float32)
    if L == 0:
        return Ct.astype(out_dtype), dCt.astype(out_dtype)
    from . import multigpu, resident
    blocks = multigpu.plan(nR)
    if resident._REG.get(resident._key(v32)) is not None:
        # the array is being kept on the device for the stages that follow (resident.keep_on_device): compute from the
        # device copies, which the histogram / S2 / average-vector stages then reuse
        def work_dev(dev, a, b):
            blk = resident.device_block(v32.reshape(nC * nF, nR, 3), dev, a, b).view(nC, nF, b - a, 3)
            c, d = ct_palmer_device(blk)
            return c.cpu().numpy(), d.cpu().numpy()

        for (dev, a, b), (c, d) in zip(blocks, multigpu.run(blocks, work_dev)):
            Ct[:, a:b], dCt[:, a:b] = c, d
        return Ct.astype(out_dtype, copy=False), dCt.astype(out_dtype, copy=False)
    if len(blocks) == 1:
        with _lib.require_cuda().cuda.device(blocks[0][0]):
            rc = lib.sr_ct_palmer_host(v32.ctypes.data_as(ctypes.c_void_p), nC, nF, nR,
                                       Ct.ctypes.data_as(ctypes.c_void_p), dCt.ctypes.data_as(ctypes.c_void_p))
            _lib.check(rc, "sr_ct_palmer_host")
        return Ct.astype(out_dtype, copy=False), dCt.astype(out_dtype, copy=False)

    # bond vectors are independent (:222-228): every selected GPU takes a contiguous block of them through the same
    # host-buffer entry point, on its own host thread
    def work(dev, a, b):
        sub = np.ascontiguousarray(v32[:, :, a:b, :])
        c, d = np.empty((L, b - a), dtype=np.float32), np.empty((L, b - a), dtype=np.float32)
        rc = lib.sr_ct_palmer_host(sub.ctypes.data_as(ctypes.c_void_p), nC, nF, b - a,
                                   c.ctypes.data_as(ctypes.c_void_p), d.ctypes.data_as(ctypes.c_void_p))
        _lib.check(rc, "sr_ct_palmer_host (device %d)" % dev)
        return c, d

    for (dev, a, b), (c, d) in zip(blocks, multigpu.run(blocks, work)):
        Ct[:, a:b], dCt[:, a:b] = c, d
    return Ct.astype(out_dtype, copy=False), dCt.astype(out_dtype, copy=False)


def _block_moments(vecs3, frames_per_block):
    """GPU sums of x,y,z and the six second moments per (block, vector): (nBlocks, nR, 9) float64."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    from . import multigpu, resident
    v = np.asarray(vecs3, dtype=np.float32)
    nFr, nR, _ = v.shape
    nB = -(-nFr // frames_per_block)

    def work(dev, a, b):
        vd = resident.device_block(v, dev, a, b)
        out = torch.empty((nB, b - a, 9), dtype=torch.float64, device=vd.device)
        _lib.check(lib.sr_vec_block_moments(vd.data_ptr(), nFr, b - a, int(frames_per_block), out.data_ptr(),
                                            _lib.current_stream_ptr()), "sr_vec_block_moments")
        return out.cpu().numpy()

    blocks = multigpu.plan(nR)
    return np.concatenate(multigpu.run(blocks, work), axis=1)


def _s2_from_moments(m, n):
    """1.5 sum_ij <v_i v_j>^2 - 0.5 from packed (xx,xy,xz,yy,yz,zz) sums over n frames."""
    a = m / n
    return 1.5 * (a[..., 0] ** 2 + a[..., 3] ** 2 + a[..., 5] ** 2
                  + 2.0 * (a[..., 1] ** 2 + a[..., 2] ** 2 + a[..., 4] ** 2)) - 0.5


def calculate_S2_by_outerProduct(vecs, delta_t=-1, tau_memory=-1):
    """Order parameter S2 = 3/2 sum_ij <e_i e_j>^2 - 1/2 (calculate-Ct-from-traj.py:96-145) for (frames, nR, 3)
    or (frames, 3) vectors; with delta_t and tau_memory the mean and std/(sqrt(nBlocks)-1) over tau-long blocks."""
    vecs = np.asarray(vecs)
    single = (vecs.ndim == 2)
    v3 = vecs[:, None, :] if single else vecs
    if v3.ndim != 3:
        print("= = = ERROR in calculate_S2_by_outerProduct: unsupported number of dimensions! vecs.shape: ",
              vecs.shape, file=sys.stderr)
        sys.exit(1)
    nFr = v3.shape[0]
    if delta_t < 0 or tau_memory < 0:
        s2 = _s2_from_moments(_block_moments(v3, nFr)[0, :, 3:], nFr)
        return s2[0] if single else s2
    per = int(tau_memory / delta_t)
    nB = int(nFr / per)
    m = _block_moments(v3[: nB * per], per)[:, :, 3:]
    s2 = _s2_from_moments(m, per)
    out = np.stack((s2.mean(axis=0), s2.std(axis=0) / (np.sqrt(nB) - 1.0)), axis=-1)
    return out[0] if single else out


def average_vectors(vecs3, q_rot=None):
    """--vecAvg (:579-583): normalised mean vector per bond, optionally in the PAF frame.  The mean of the
    rotated vectors is the rotated mean, so only the (nR, 3) sums leave the GPU."""
    from . import qs
    v3 = np.asarray(vecs3)
    mean = _block_moments(v3, v3.shape[0])[0, :, :3] / v3.shape[0]
    if q_rot is not None:
        mean = qs.rotate_vector_simd(mean, np.asarray(q_rot, dtype=np.float64))
    return mean / np.linalg.norm(mean, axis=-1, keepdims=True)

"""One pass of the bond-vector hot path over a batch, as calculate-Ct-from-traj.py runs it with
`--Ct --vecRot q --vecHist` (run-all.bash:475-481): C(t) with Palmer statistics (:527-531), then PAF
rotation and the Lambert-cylindrical histogram of the same vectors (:535-630).

CtHistStep keeps the device workspaces alive between steps, launches the kernels on torch's current
stream through the C ABI, and (optionally) brackets every launch with CUDA events so bench.py can
report per-kernel durations.  With world > 1 every rank owns a shard of the bond vectors; the only
communication is the gather of the per-vector result rows to rank 0.
"""
import numpy as np

from . import _lib, ct


class CtHistStep:
    def __init__(self, nC, nF, nR, q_rot=None, device=None, world=1, rank=0, hist_bins=72):
        self.torch = _lib.require_cuda()
        torch = self.torch
        self.lib = _lib.load()
        self.nC, self.nF, self.nR, self.L = nC, nF, nR, nF // 2
        self.q_rot = None if q_rot is None else np.asarray(q_rot, dtype=np.float64)
        self.dev = device if device is not None else torch.device("cuda")
        self.world, self.rank = world, rank
        self.nbx, self.nby = hist_bins, hist_bins // 2
        self.has_hist = hasattr(self.lib, "sr_sphere_hist") and q_rot is not None
        self.pitch = self.lib.sr_ct_row_pitch(nF)
        self.packed = torch.empty((nR, nC, 3, self.pitch), dtype=torch.float32, device=self.dev)
        self.S = torch.empty((nR, nC, self.L), dtype=torch.float64, device=self.dev)
        self.Ct = torch.empty((self.L, nR), dtype=torch.float32, device=self.dev)
        self.dCt = torch.empty((self.L, nR), dtype=torch.float32, device=self.dev)
        if self.has_hist:
            from . import hist as _hist
            self._hist = _hist.SphereHistogram(nR, self.nbx, device=self.dev)
        self._events = {}
        self._launches = 3 + (1 if self.has_hist else 0)

    # -------------------------------------------------------------------------------------------
    def _timed(self, name, fn, on):
        if not on:
            fn()
            return
        torch = self.torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        self._events.setdefault(name, []).append((a, b))

    def run_device(self, v_dev, time_kernels=False):
        lib, st = self.lib, _lib.current_stream_ptr()
        nC, nF, nR, L = self.nC, self.nF, self.nR, self.L
        self._timed("pack_kernel", lambda: _lib.check(
            lib.sr_pack_vectors_f32(v_dev.data_ptr(), nC, nF, nR, None, self.packed.data_ptr(), self.pitch, st),
            "sr_pack_vectors_f32"), time_kernels)
        self._timed("ct_lag_kernel", lambda: _lib.check(
            lib.sr_ct_lag_sums(self.packed.data_ptr(), self.pitch, nC, nF, nR, L, self.S.data_ptr(), st),
            "sr_ct_lag_sums"), time_kernels)
        self._timed("ct_finalize_kernel", lambda: _lib.check(
            lib.sr_ct_palmer_finalize(self.S.data_ptr(), nC, nF, nR, L, self.Ct.data_ptr(), self.dCt.data_ptr(), st),
            "sr_ct_palmer_finalize"), time_kernels)
        hist = None
        if self.has_hist:
            self._timed("sphere_hist_kernel",
                        lambda: self._hist.accumulate_device(v_dev.view(nC * nF, nR, 3), self.q_rot, reset=True),
                        time_kernels)
            hist = self._hist.finish(v_dev, self.q_rot)     # ambiguous-sample tie-break + D2H of the counts
        if self.world > 1:
            # every rank owns nR vectors of the global set: gather the (2L, nR) result columns on rank 0
            from . import shard
            both = self.torch.cat((self.Ct, self.dCt), dim=0)
            self.gathered = shard.gather_columns(both, self.nR * self.world, dst=0)
        return self.Ct, self.dCt, hist

    def run_host(self, v_np):
        """Public host-buffer path: NumPy (nC, nF, nR, 3) float32 in, NumPy out.  One H2D of the vectors
        (pinned memory is copied asynchronously), the same kernels as run_device, D2H of Ct, dCt, counts."""
        torch = self.torch
        v = np.ascontiguousarray(v_np, dtype=np.float32)
        v_dev = torch.from_numpy(v).to(self.dev, non_blocking=True)
        Ct, dCt, hist = self.run_device(v_dev)
        out = torch.stack((Ct, dCt)).cpu().numpy()
        if hist is not None:
            hist = hist.astype(np.float64)
        return out[0], out[1], hist

    # -------------------------------------------------------------------------------------------
    def reset_kernel_timers(self):
        self._events = {}

    def kernel_times_ms(self):
        self.torch.cuda.synchronize()
        return {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in self._events.items()}

    def launches_per_step(self):
        return self._launches

    def d2h_bytes(self):
        n = 2 * self.L * self.nR * 4
        if self.has_hist:
            n += self.nR * self.nbx * self.nby * 4
        return n

    @staticmethod
    def ncu_traffic_bytes():
        """DRAM bytes per ct_lag_kernel launch from the committed ncu capture (profiles/), or None."""
        import json
        import os
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
        if os.path.exists(p):
            with open(p) as fp:
                return json.load(fp).get("ct_lag_kernel_dram_bytes_per_launch")
        return None

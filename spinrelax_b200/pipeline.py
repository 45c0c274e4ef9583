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
    def __init__(self, nC, nF, nR, q_rot=None, device=None, world=1, rank=0, hist_bins=72, n_total=None):
        """nR = bond vectors held by THIS rank; n_total = vectors of the whole job (default nR * world), partitioned
        over the ranks by shard.split_range (so nR must be this rank's share of it)."""
        self.torch = _lib.require_cuda()
        torch = self.torch
        self.lib = _lib.load()
        self.nC, self.nF, self.nR, self.L = nC, nF, nR, nF // 2
        self.q_rot = None if q_rot is None else np.asarray(q_rot, dtype=np.float64)
        self.dev = device if device is not None else torch.device("cuda")
        self.world, self.rank = world, rank
        self.n_total = nR * world if n_total is None else int(n_total)
        if world > 1:
            from . import shard
            if shard.split_sizes(self.n_total, world)[rank] != nR:
                raise _lib.SpinRelaxError("CtHistStep: rank %d holds %d vectors but the partition of %d over %d ranks "
                                          "gives it %d" % (rank, nR, self.n_total, world,
                                                           shard.split_sizes(self.n_total, world)[rank]))
        self.gathered = self.gathered_hist = None
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
        self._launches = 3 + (3 if self.has_hist else 0)   # pack, lag sums, finalize (+ mark, histogram, resolve)

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
            self._gather(hist)
        return self.Ct, self.dCt, hist

    def run_host(self, v_np):
        """Public host-buffer path: NumPy (nC, nF, nR, 3) float32 in, NumPy out.

        Chunks are contiguous in the reference layout, so the H2D copy is pipelined with the kernels: chunk
        0 is copied first and its K2 / K1 run while the remaining chunks are copied on a side stream
        (asynchronously when the caller's array is pinned; sr_pack_vectors_f32_chunks, sr_ct_lag_sums_chunks).
        Finalize, the histogram and the D2H of Ct, dCt and the counts follow once the last chunk is done.  The
        returned Ct / dCt are views of a pinned staging buffer that the next call overwrites."""
        torch, lib = self.torch, self.lib
        nC, nF, nR, L = self.nC, self.nF, self.nR, self.L
        v = np.ascontiguousarray(v_np, dtype=np.float32)
        if v.shape != (nC, nF, nR, 3):
            raise _lib.SpinRelaxError("run_host: expected shape %s, got %s" % ((nC, nF, nR, 3), v.shape))
        v_host = torch.from_numpy(v)
        if getattr(self, "_v_dev", None) is None:
            self._v_dev = torch.empty((nC, nF, nR, 3), dtype=torch.float32, device=self.dev)
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        v_dev = self._v_dev
        main = torch.cuda.current_stream(self.dev)
        self._copy_stream.wait_stream(main)            # the previous step may still be reading v_dev
        # two stages: chunk 0 alone, then the rest in one launch (every extra launch costs a tail of ~1 wave)
        groups = [(0, 1)] + ([(1, nC - 1)] if nC > 1 else [])
        ready = []
        with torch.cuda.stream(self._copy_stream):
            for c0, n in groups:
                v_dev[c0:c0 + n].copy_(v_host[c0:c0 + n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                ready.append(ev)
        st = _lib.current_stream_ptr()
        for (c0, n), ev in zip(groups, ready):
            main.wait_event(ev)
            _lib.check(lib.sr_pack_vectors_f32_chunks(v_dev[c0].data_ptr(), nC, c0, n, nF, nR, None, self.packed.data_ptr(),
                                                      self.pitch, st), "sr_pack_vectors_f32_chunks")
            _lib.check(lib.sr_ct_lag_sums_chunks(self.packed.data_ptr(), self.pitch, nC, c0, n, nF, nR, L,
                                                 self.S.data_ptr(), st), "sr_ct_lag_sums_chunks")
        _lib.check(lib.sr_ct_palmer_finalize(self.S.data_ptr(), nC, nF, nR, L, self.Ct.data_ptr(), self.dCt.data_ptr(), st),
                   "sr_ct_palmer_finalize")
        if getattr(self, "_out_host", None) is None:
            self._out_host = torch.empty((2, L, nR), dtype=torch.float32).pin_memory()
        self._out_host[0].copy_(self.Ct, non_blocking=True)      # D2H overlaps the histogram pass
        self._out_host[1].copy_(self.dCt, non_blocking=True)
        hist = None
        if self.has_hist:
            self._hist.accumulate_device(v_dev.view(nC * nF, nR, 3), self.q_rot, reset=True)
            hist = self._hist.finish(v_dev, self.q_rot)
        if self.world > 1:
            self._gather(hist)
        main.synchronize()
        out = self._out_host.numpy()          # pinned staging buffer, overwritten by the next call
        if hist is not None:
            hist = hist.astype(np.float64)
        return out[0], out[1], hist

    def _gather(self, hist):
        """Every rank owns its block of the job's bond vectors: the (2L, nR_local) result columns and the
        (nR_local, nbx, nby) histogram counts are gathered on rank 0 (no other communication on this path)."""
        from . import shard
        both = self.torch.cat((self.Ct, self.dCt), dim=0)
        self.gathered = shard.gather_columns(both, self.n_total, dst=0)
        if hist is not None:
            h = self.torch.from_numpy(np.ascontiguousarray(hist)).to(self.dev)
            self.gathered_hist = shard.gather_rows(h, self.n_total, dst=0)

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

    # tile configuration of the ct_lag_kernel instantiation the library launches for long chunks (csrc/ct.cu: CtLong)
    CT_LAG_CONFIG = "CtCfg<R=23,MB=9,FB=9,NW=12,MINB=1,NS=2,FLUSH=2,ORDER=1,SYNC=2>"

    @classmethod
    def ncu_traffic_bytes(cls):
        """DRAM bytes per ct_lag_kernel launch from the committed ncu capture (profiles/ncu_traffic.json), or None.
        The capture records the tile configuration it was taken with; a capture of another configuration is not
        reported as this build's traffic."""
        import json
        import os
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
        if not os.path.exists(p):
            return None
        with open(p) as fp:
            rec = json.load(fp)
        if rec.get("kernel_config") != cls.CT_LAG_CONFIG:
            return None
        return rec.get("ct_lag_kernel_dram_bytes_per_launch")

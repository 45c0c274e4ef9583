"""CLI mirror of calculate-Ct-from-traj.py for the part of it that is on the hot path.

    python -m spinrelax_b200.cli_ct -s ref.pdb -f vecs.npy [vecs2.npy ...] --dt 10 --tau 5000 -o rotdif \
           --vecRot "qw qx qy qz" --vecHist --binary --vecAvg --S2 --Ct

Flags and defaults are the reference's (:303-345).  What is NOT reproduced is mdtraj itself (trajectory file
formats and the atom-selection language, :396-498 -- out of scope).  `-f` takes, per
trajectory, either
  * the unit X-H vector trajectory: a `.npy` (frames, bonds, 3) array or an `.npz` with `vecs` [, `vecs_unfitted`,
    `names`, `dt`], i.e. exactly what obtain_XHvecs (:64-86) returns after the fit; or
  * Cartesian coordinates: an `.npz` with `xyz` (frames, atoms, 3), `indexH`, `indexX` [, `fit`, `names`, `dt`]
    (the index arrays mdtraj's `topology.select(--Hsel / --Xsel / --fitsel)` would return).  The vectors are then
    extracted on the GPU (spinrelax_b200.traj, :83-84) before and after the least-squares superposition onto the
    reference given with `-s` (`.npy` (atoms, 3) or `.npz` with `xyz`; default: the trajectory's first frame), which
    is what `trj.center_coordinates(); trj.superpose(ref, frame=0, atom_indices=fit_indices)` (:466-467) does.
Everything downstream is the reference's flow: reformat by tau (:513-516), C(t) (:527-531),
reshape (:535-536), PAF rotation (:567), average vector (:579-583), spherical histogram (:585-630), S2 (:638-646),
written to the same files in the same formats.
"""
import argparse
import sys
import time

import numpy as np

from . import ct, hist, io_formats, resident


def build_parser():
    p = argparse.ArgumentParser(description='Obtain the unit X-H vectors from one of more trajectories,'
                                'and conduct calculations on it, such as S^2, C(t), and others analyses.',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('-s', type=str, dest='topfn', required=False, nargs='+', default=[],
                   help='Reference structure(s) for the superposition of coordinate inputs (.npy / .npz); ignored '
                        'for X-H vector inputs.')
    p.add_argument('-f', '--infn', type=str, dest='infn', required=True, nargs='+',
                   help='One or more X-H vector trajectories (.npy / .npz). Multiple trajectories are analysed '
                        'separately in C(t)-calculations, but otherwise aggregated.')
    p.add_argument('-o', '--outpref', type=str, dest='out_pref', default='out', help='Output file prefix.')
    p.add_argument('--split', type=int, dest='nSplitFrames', default=-1, help='Accepted, unused.')
    p.add_argument('--dt', type=float, dest='dt', default=None,
                   help='Time between frames (the reference reads it from the trajectory); default: npz `dt` or 1.0.')
    p.add_argument('-t', '--tau', type=float, dest='tau', default=None,
                   help='An estimate of the global tumbling time that is used to ignore internal motions '
                        'on timescales larger than can be measured by NMR relaxation.')
    p.add_argument('--prefact', type=float, dest='zeta', default=(1.02 / 1.04) ** 6,
                   help='MD-specific prefactor that accounts for librations of the XH-vector not seen in classical MD.')
    p.add_argument('--S2', dest='bDoS2', action='store_true', default=False, help='Calculate order parameters S2.')
    p.add_argument('--Ct', dest='bDoCt', action='store_true', default=False, help='Calculate autocorrelation Ct.')
    p.add_argument('--vecDist', dest='bDoVecDistrib', action='store_true', default=False,
                   help='Print the vectors distribution in spherical coordinates.')
    p.add_argument('--binary', action='store_true', default=False, help='Store distributions as numpy binaries.')
    p.add_argument('--vecHist', dest='bDoVecHist', action='store_true', default=False,
                   help='Print the 2D-histogram rather than just the collection of vecs.')
    p.add_argument('--histBin', type=int, default=72,
                   help='Number of bins along phi (-pi,pi); the number of bins along cos(theta) is half of it.')
    p.add_argument('--vecAvg', dest='bDoVecAverage', action='store_true', default=False,
                   help='Print the average unit XH-vector.')
    p.add_argument('--vecRot', dest='vecRotQ', type=str, default='',
                   help='Rotation quaternion to be applied to the vector to transform it into PAF frame.')
    p.add_argument('--Hsel', '--selection', type=str, dest='Hseltxt', default='name H', help='Accepted, unused.')
    p.add_argument('--Xsel', type=str, dest='Xseltxt', default='name N and not resname PRO', help='Accepted, unused.')
    p.add_argument('--fitsel', type=str, dest='fittxt', default='custom occupancy', help='Accepted, unused.')
    p.add_argument('--help_sel', action='store_true', help='Display help for selection texts and exit.')
    return p


def print_selection_help():
    """calculate-Ct-from-traj.py:17-19, plus what replaces the selection texts on this path."""
    print("Notes: This python program uses MDTraj as its underlying engine to analyse trajectories and select atoms.")
    print("It uses selection syntax such as 'chain A and resname GLY and name HA1 HA2', in a manner similar to GROMACS and VMD.")
    print("On this path coordinates enter as arrays: give the atom index lists as `indexH`, `indexX` and `fit` in the input .npz;"
          " --Hsel / --Xsel / --fitsel are accepted and ignored.")


def _load_reference(fn):
    ref = np.load(fn, allow_pickle=False)       # our own input format: plain arrays only, never unpickle user files
    if not isinstance(ref, np.ndarray):
        ref = ref['xyz']
    ref = np.asarray(ref, dtype=np.float32)
    return ref[0] if ref.ndim == 3 else ref


def _from_coordinates(z, ref_fn):
    """obtain_XHvecs before and after centre + superpose (:461-469) for one coordinate trajectory."""
    import torch
    from . import traj
    xyz = np.ascontiguousarray(z['xyz'], dtype=np.float32)
    index_h, index_x = np.asarray(z['indexH']), np.asarray(z['indexX'])
    if len(index_h) == 0 or len(index_h) != len(index_x):
        print("= = = ERROR: selection text failed to find atoms!", file=sys.stderr)
        sys.exit(1)
    fit = np.asarray(z['fit']) if 'fit' in z else np.arange(xyz.shape[1])
    ref = _load_reference(ref_fn) if ref_fn else xyz[0]
    print("= = = File loaded - it has %i atoms and %i frames." % (xyz.shape[1], xyz.shape[0]))
    xd = torch.from_numpy(xyz).cuda()
    unfit = traj.xh_vectors_device(xd, index_h, index_x).cpu().numpy()
    fitv = traj.xh_vectors_superposed_device(xd, ref, fit, index_h, index_x).cpu().numpy()
    print("= = = Molecule centered and fitted.")
    return fitv, unfit


def _load(fn, ref_fn=None):
    if fn.endswith('.npy'):
        return np.load(fn), None, None, None
    if fn.endswith('.npz'):
        z = np.load(fn, allow_pickle=False)     # `names` must be a unicode or integer array, not an object array
        names = list(z['names']) if 'names' in z else None
        dt = float(z['dt']) if 'dt' in z else None
        if 'xyz' in z:
            fitv, unfit = _from_coordinates(z, ref_fn)
            return fitv, unfit, names, dt
        return z['vecs'], z['vecs_unfitted'] if 'vecs_unfitted' in z else None, names, dt
    print("= = = ERROR: %s: only .npy/.npz X-H vector trajectories are accepted on this path "
          "(trajectory reading via mdtraj is out of scope)." % fn, file=sys.stderr)
    sys.exit(2)


def _spherical_by_residue(frames, q_rot=None):
    """(frames, nR, 3) float32 -> (nR, frames, 3) r/phi/theta as calculate-Ct-from-traj.py:567,588,600 produce it:
    float32 without --vecRot, float64 after the PAF rotation (rotate_vector_simd promotes)."""
    import ctypes
    from . import _lib, gm, qs
    torch = _lib.require_cuda()
    lib = _lib.load()
    vd = torch.from_numpy(np.ascontiguousarray(frames, dtype=np.float32)).cuda()
    if q_rot is not None:
        q = qs.vecnorm_NDarray(np.asarray(q_rot, dtype=np.float64))
        rot = torch.empty(vd.shape, dtype=torch.float64, device=vd.device)
        _lib.check(lib.sr_rotate_vectors_f32_f64(vd.data_ptr(), vd.numel() // 3, (ctypes.c_double * 4)(*q.tolist()),
                                                 rot.data_ptr(), _lib.current_stream_ptr()), "sr_rotate_vectors_f32_f64")
        vd = rot
    rtp = gm.xyz_to_rtp_device(vd)
    return rtp.permute(1, 0, 2).contiguous().cpu().numpy()


def main(argv=None):
    argv = sys.argv[1:] if argv is None else list(argv)
    if '--help_sel' in argv:                           # upstream -s and -f are required, but --help_sel exits first (:350-352)
        print_selection_help()
        sys.exit(0)
    args = build_parser().parse_args(argv)
    time_start = time.time()
    tau_memory = args.tau
    if args.bDoCt and tau_memory is None:
        print("= = = Refusing to do C(t)-analysis without using a block averaging over memory_time tau!", file=sys.stderr)
        sys.exit(1)
    bDoVecDistrib = args.bDoVecDistrib or args.bDoVecHist
    histBinX = args.histBin
    bRotVec = args.vecRotQ != ''
    q_rot = None
    if bRotVec:
        q_rot = np.array([float(v) for v in args.vecRotQ.split()])
        if len(q_rot) != 4 or not np.allclose(np.dot(q_rot, q_rot), 1):
            print("= = = ERROR: input rotation quaternion is malformed!", q_rot)
            sys.exit(23)

    vecXH, vecXHfit, resXH, deltaT = [], [], None, args.dt
    n_refs = len(args.topfn)
    if n_refs > 1 and n_refs != len(args.infn):
        print("= = ERROR: When giving multiple reference files, you must have one for each trajecfile file given!",
              file=sys.stderr)                                            # :411-414
        sys.exit(1)
    for i_fn, fn in enumerate(args.infn):
        ref_fn = None
        if n_refs and args.topfn[0].endswith(('.npy', '.npz')):
            ref_fn = args.topfn[i_fn] if n_refs > 1 else args.topfn[0]
        fit, unfit, names, dt = _load(fn, ref_fn)
        fit = np.asarray(fit, dtype=np.float32)
        if fit.ndim != 3 or fit.shape[-1] != 3:
            print("= = = ERROR: %s does not hold a (frames, bonds, 3) array." % fn, file=sys.stderr)
            sys.exit(1)
        print("= = = File loaded - it has %i bonds and %i frames." % (fit.shape[1], fit.shape[0]))
        if resXH is None:
            resXH = names if names is not None else list(range(1, fit.shape[1] + 1))
        elif fit.shape[1] != len(resXH):
            print("= = = ERROR: Differences in trajectories have been detected! Aborting.", file=sys.stderr)
            sys.exit(1)
        if deltaT is None:
            deltaT = dt
        vecXHfit.append(fit)
        vecXH.append(np.asarray(unfit, dtype=np.float32) if unfit is not None else None)
    if deltaT is None:
        deltaT = 1.0
    if tau_memory is not None and deltaT > 0.5 * tau_memory:
        print("= = = ERROR: delta-t form the trajectory is too small relative to tau! %g vs. %g" % (deltaT, tau_memory),
              file=sys.stderr)
        sys.exit(1)
    print("= = Loading finished.")
    nBonds = len(resXH)
    have_ext = all(v is not None for v in vecXH)

    if tau_memory is not None:
        print("= = Reformatting all vecXH information into chunks of tau ( %g ) " % tau_memory)
        fit4 = ct.reformat_vecs_by_tau(vecXHfit, deltaT, tau_memory)
        ext4 = ct.reformat_vecs_by_tau(vecXH, deltaT, tau_memory) if have_ext else None
    else:
        cat = np.concatenate(vecXHfit, axis=0)
        fit4, ext4 = cat[None], None

    # C(t), the average vector, the histogram and S2 all read the fitted vectors: one upload for all of them
    with resident.keep_on_device(fit4):
        if args.bDoCt:
            dt = ct.calculate_dt(deltaT, tau_memory)
            if ext4 is not None:
                print("= = = Conducting Ct_external using Palmer's approach.")
                Ct, dCt = ct.calculate_Ct_Palmer(ext4)
                io_formats.print_sxylist(args.out_pref + '_Ctext.dat', resXH, dt, np.stack((Ct.T, dCt.T), axis=-1))
            print("= = = Conducting Ct_internal using Palmer's approach.")
            Ct, dCt = ct.calculate_Ct_Palmer(fit4)
            io_formats.print_sxylist(args.out_pref + '_Ctint.dat', resXH, dt, np.stack((Ct.T, dCt.T), axis=-1))

        sh = fit4.shape
        frames = fit4.reshape((sh[0] * sh[1], sh[-2], sh[-1]))          # :535-536

        if args.bDoVecAverage:
            avg = ct.average_vectors(frames, q_rot)
            io_formats.print_xylist(args.out_pref + '_avgvec.dat', resXH, np.array(avg).T, True)

        if bDoVecDistrib:
            if not args.bDoVecHist:
                # calculate-Ct-from-traj.py:586-607: spherical coordinates of every sample, residue first
                print("= = = Converting vectors into spherical coordinates.")
                rtp = _spherical_by_residue(frames, q_rot)
                print("= = = Debug: shape of the spherical vector distribution:", rtp.shape)
                if args.binary:
                    np.savez_compressed(args.out_pref + '_vecPhiTheta.npz', names=resXH, dataType='PhiTheta',
                                        axisLabels=['phi', 'theta'], bHistogram=False, data=rtp[..., 1:3])
                else:
                    io_formats.print_s3d(args.out_pref + '_vecPhiTheta.dat', resXH, rtp, (1, 2))
        if args.bDoVecHist:
            print("= = = Histgrams will use Lambert Cylindrical projection by converting Theta spanning (0,pi) to "
                  "cos(Theta) spanning (-1,1)")
            hist_list, edges = hist.sphere_histogram(frames, q_rot, histBinX)
            if args.binary:
                hist.save_vec_histogram(args.out_pref + '_vecHistogram.npz', resXH, hist_list, edges)
            else:
                for i in range(nBonds):
                    ofile = args.out_pref + '_vecXH_' + str(resXH[i]) + '.hist'
                    io_formats.print_gplot_hist(ofile, hist_list[i], edges,
                                                header='# Lamber Cylindrical Histogram over phi,cos(theta).', bSphere=True)
                    print("= = = Written to output: ", ofile)

        if args.bDoS2:
            if tau_memory is not None:
                print("= = = Conducting S2 analysis using memory time to chop input-trajectories", tau_memory, "ps")
                S2 = ct.calculate_S2_by_outerProduct(frames, deltaT, tau_memory)
            else:
                print("= = = Conducting S2 analysis directly from trajectories.")
                S2 = ct.calculate_S2_by_outerProduct(frames)
            io_formats.print_xylist(args.out_pref + '_S2.dat', resXH, (S2.T) * args.zeta, True)
            print("      ...complete.")

    print("= = Finished. Total seconds elapsed: %g" % (time.time() - time_start))


if __name__ == '__main__':
    main()

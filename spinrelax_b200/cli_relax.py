"""CLI mirror of calculate-relaxations-multi-field.py: prediction and optimisation against experiment.

    python -m spinrelax_b200.cli_relax -f rotdif_fittedCt.dat --distfn rotdif_vecHistogram.npz \
           -D "Diso" --aniso a [--zeta z] [--csa v|file] [--opt Diso,rsCSA --cycles 10 --tol 1e-6] \
           -o rotdif expt1.dat expt2.dat ...

Flags follow the reference (:41-105).  Output: `<o>_<A><B>_<MHz>MHz_<Type>.xvg` per experiment file
(spectral_densities.py:775-780, :1178-1194) and, after a residue-specific CSA optimisation, `<o>_CSA_opt.dat`
(:219-225).  All J(omega) / R1 / R2 / NOE arithmetic runs on the GPU; `--localopt powell` replays the reference's
residue-by-residue Powell searches instead of the batched solver.
"""
import argparse
import sys
import time
from re import split as regexp_split

import numpy as np

from . import fitct, specdens as sd


def build_parser():
    p = argparse.ArgumentParser(description='Prediction of spin relaxation parameters at multiple fields from the '
                                'simulated local and global tumbling characteristics. All internal units are picoseconds.',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('expFiles', type=str, nargs='+',
                   help='Experiment files: header lines `# Type R1|R2|NOE`, `# NucleiA 15N`, `# NucleiB 1H`, '
                        '`# Frequency 600.133`, then `name value [error]` rows.')
    p.add_argument('-o', '--outpref', type=str, dest='out_pref', default='out', help='Output file prefix.')
    p.add_argument('-f', '--infn', type=str, dest='in_Ct_fn', required=True, help='Fitted C_internal(t) parameter file.')
    p.add_argument('--refpdb', type=str, dest='refPDBFile', default=None, help='Not supported on this path (needs mdtraj).')
    p.add_argument('--distfn', type=str, dest='distfn', default=None,
                   help='Vector orientation histogram (.npz) in the principal axes frame.')
    p.add_argument('--tau', type=float, dest='tau', default=None, help='Isotropic relaxation time constant.')
    p.add_argument('--aniso', type=float, dest='aniso', default=None, help='Diffusion anisotropy (prolate/oblate).')
    p.add_argument('-D', '--DTensor', type=str, dest='D', default=None,
                   help='Diffusion tensor: one value = Diso, two values = Dpar,Dperp.')
    p.add_argument('--zeta', type=float, default=0.890023, help='Zero-point vibration scaling of the C(t) magnitudes.')
    p.add_argument('--csa', type=str, default=None, help='Average CSA value, or a file with `residue CSA` lines.')
    p.add_argument('--opt', '--fit', type=str, dest='listOptParams', default=None,
                   help='Perform optimisation against all given experimental data, over the following possible '
                        'parameters %s (comma-separated, in the order used in the optimisation loop).'
                        % sd.spinRelaxationExperiments.listAllowedOptimisationVariables)
    p.add_argument('--cycles', type=int, default=10,
                   help='Maximum number of global/local refinement cycles when both are optimised.')
    p.add_argument('--tol', type=float, default=1e-6,
                   help='Tolerance for terminating the global/local cycles early, as a fractional change.')
    p.add_argument('--localopt', type=str, default='batched', choices=['batched', 'powell'],
                   help='Residue-specific CSA solver: all residues at once on the GPU, or the reference\'s '
                        'residue-by-residue Powell.')
    return p


def parse_rotdif_params(D=None, tau=None, aniso=None):
    """(:13-37) isotropic or axisymmetric model from -D / --tau / --aniso."""
    if D is None:
        if tau is None:
            print("= = ERROR: No global tumbling parameters given!", file=sys.stderr)
            sys.exit(1)
        Diso = 1.0 / (6 * tau)
        if aniso is None or aniso == 1.0:
            return sd.globalRotationalDiffusion_Isotropic(D=Diso)
        return sd.globalRotationalDiffusion_Axisymmetric(D=[Diso, aniso])
    tmp = [float(x) for x in regexp_split('[, ]', D) if len(x) > 0]
    if len(tmp) == 1:
        if aniso is None:
            return sd.globalRotationalDiffusion_Isotropic(D=tmp[0])
        return sd.globalRotationalDiffusion_Axisymmetric(D=[tmp[0], aniso])
    if len(tmp) == 2:
        return sd.globalRotationalDiffusion_Axisymmetric(D=tmp, bConvert=True)
    print("WARNING: fully anisotropic global rotdif not implemented.", file=sys.stderr)
    return None


def main(argv=None):
    time_start = time.time()
    args = build_parser().parse_args(argv)
    models = fitct.read_fittedCt_parameters(args.in_Ct_fn)
    if models.nModels == 0:
        print("= = = ERROR: The fitted-Ct file %s was read, but did not yield any usable parameters!" % args.in_Ct_fn)
        sys.exit(1)
    rot = parse_rotdif_params(args.D, args.tau, args.aniso)
    if rot is None:
        sys.exit(1)
    if args.distfn is not None:
        rot.import_frame_vectors(args.distfn)
    elif args.refPDBFile is not None:
        print("= = = ERROR: --refpdb needs mdtraj, which is outside this path; use --distfn.", file=sys.stderr)
        sys.exit(2)
    ex = sd.spinRelaxationExperiments(rot, models, local_mode=args.localopt)
    for f in args.expFiles:
        ex.add_experiment(f)
    if args.zeta != 1.0:
        print(" = = Applying scaling of all C(t) magnitudes to account for zero-point QM vibrations (zeta) of %g" % args.zeta)
        ex.set_global_zeta(args.zeta)
    ex.map_experiment_peaknames_to_models()
    ex.report_maps()
    if args.csa is None:
        print("= = = Using default CSA value respective to each experiment.")
    else:
        try:
            tab = np.loadtxt(args.csa, ndmin=2)
            resid, vals = [str(int(x)) for x in tab[:, 0]], tab[:, 1].copy()
            if np.fabs(vals[0]) > 1.0:
                vals *= 1e-6
            ex.initialise_CSA_array(resid, vals)
        except OSError:
            try:
                v = float(args.csa)
            except ValueError:
                print("= = = ERROR at parsing the --csa argument!", file=sys.stderr)
                sys.exit(1)
            if np.fabs(v) > 1.0:
                v *= 1e-6
            ex.initialise_CSA_array(models.get_names(), np.repeat(v, models.nModels))
    if args.listOptParams is None:
        ex.eval_all(bVerbose=True)
        ex.export_xvg(args.out_pref, bIncludeExpt=False)
        print("= = Finished. Total seconds elapsed: %g" % (time.time() - time_start))
        return
    ex.parse_optimisation_params(args.listOptParams.split(','))
    print("= = = Parsed optimiser input %s." % args.listOptParams)
    print("    ... conducting global optimisations over parameters %s ..." % (ex.listUpdateVariables))
    if ex.bDoLocalOpt:
        print("    ... conducting local optimisations over residue-specific CSA...")
    chisq = ex.perform_optimisation(maxCycles=args.cycles, tol=args.tol)
    print("= = = Optimisation complete. Final chi-value: %g" % np.sqrt(chisq))
    ex.export_xvg(args.out_pref, bIncludeExpt=True)
    if ex.bDoLocalOpt and ex.bOptCompleted:
        with open(args.out_pref + '_CSA_opt.dat', 'w') as fp:
            for x, y in zip(ex.localCtModels.get_names(), ex.get_first_csa()):
                print("%s %g" % (x, y), file=fp)
    print("= = Finished. Total seconds elapsed: %g" % (time.time() - time_start))


if __name__ == '__main__':
    main()

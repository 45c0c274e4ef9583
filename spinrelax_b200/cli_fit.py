"""CLI mirror of calculate-fitted-Ct.py (flags :64-82): reads `<prefix>_Ctint.dat`, fits every residue's C(t)
with the multi-exponential ladder on the GPU (one batched launch per rung) and writes `<prefix>_fittedCt.dat`.

    python -m spinrelax_b200.cli_fit -f rotdif_Ctint.dat -o rotdif [--nc N] [--nofast]

Only the single-input-file form used by run-all.bash (:488-491) is supported; the multi-file averaging branch of
the reference (:110-146) references undefined names upstream and is not reproduced.
"""
import argparse
import sys
import time

import numpy as np

from . import fitct, io_formats


def build_parser():
    p = argparse.ArgumentParser(description='This script reads in the raw autocorrelation functions C(t) '
                                'and fits one or a set of simple exponential decay compnonets to them.',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('-f', '--infn', type=str, dest='in_Ct_fn', nargs='+',
                   help='File containing the autocorrelation functions, one xmgrace set per residue (legend = residue).')
    p.add_argument('-o', '--outpref', type=str, dest='out_pref', default='out', help='Output file prefix.')
    p.add_argument('--nc', type=int, default=-1,
                   help='number of transient components to fit; -1 searches for the best number.')
    p.add_argument('--nofast', dest='bNoFast', action='store_true', default=False,
                   help='Do not permit an S_fast component, so that C(0) must be one.')
    return p


def main(argv=None):
    time_start = time.time()
    args = build_parser().parse_args(argv)
    if not args.in_Ct_fn or len(args.in_Ct_fn) != 1:
        print("= = = ERROR: exactly one C(t) input file is supported on this path.", file=sys.stderr)
        sys.exit(1)
    print("= = = Found %d input C(t) files." % 1)
    legs, dt, Ct, Cterr = io_formats.load_sxydylist(args.in_Ct_fn[0], 'legend')
    legs = [int(x) for x in legs]
    if len(Cterr) == 0:
        Cterr = None
    ac = fitct.autoCorrelations()
    ac.import_target_array(keys=legs, DeltaT=dt, Decay=Ct, dDecay=Cterr)
    use_fast = not args.bNoFast
    for k in ac.DeltaT.keys():
        ac.add_model(k)
    if args.nc == -1:
        ac.fit_all_residues(listDoG=(2, 3, 5, 7, 9) if use_fast else (2, 4, 6, 8), chiSqThreshold=0.5)
    else:
        n = 2 * args.nc + (1 if use_fast else 0)
        ac.fit_all_residues(listDoG=(n,), chiSqThreshold=0.5, single=True)
    ac.export(fileName=args.out_pref + '_fittedCt.dat', style='xmgrace')
    print(" = = Completed C(t)-fits.")
    print("= = Finished. Total seconds elapsed: %g" % (time.time() - time_start))


if __name__ == '__main__':
    main()

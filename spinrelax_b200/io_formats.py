"""Boundary text formats either side of the hot path (readers/writers only, no numerics):

  read_from_plumedprint   plumedcolvario.py:24-81   (PLUMED PRINT file -> float32 (nfields, ndata))
  print_xylist            general_scripts.py:246-273 (-aniso_q.dat, -tensor.dat, _S2.dat, _avgvec.dat)
  print_sxylist           general_scripts.py:275-290 (_Ctint.dat / _Ctext.dat, xmgrace sets with legends)
  load_sxydylist          general_scripts.py:182-213 (reads them back for the fitting stage)
"""
import numpy as np


def read_from_plumedprint(fname):
    """Returns [field_names, data] with data float32 of shape (nfields, ndata); every value is rounded
    to float32 exactly like the reference's per-token np.float32() (quirk G2: including the time column)."""
    names = None
    rows = []
    with open(fname) as fp:
        for line in fp:
            if line == '\n':
                continue
            if line.startswith("#"):
                tok = line.split()
                if len(tok) > 1 and tok[1] == "FIELDS":
                    found = tok[2:]
                    if names is None:
                        names = found
                    elif any(a != b for a, b in zip(names, found)):
                        print('= = ERROR: Multiple FIELD headers are present to indicate parallel trajectoreies, '
                              'but their entries do not agree!')
                        print(names)
                        print(found)
                        return -1
                continue
            if names is None:
                print('= = ERROR: Data-like line encountered before a FIELDS definition! Line as follows:')
                print(line)
                return -1
            tok = line.split()
            if len(tok) != len(names):
                print('= = ERROR: Data-like line does not have the same number of fields as defined in FIELDS! ( %i )'
                      % len(names))
                print(tok)
                return -1
            rows.append(tok)
    data = np.array(rows, dtype=np.float64).astype(np.float32).T.copy(order='F') if rows else \
        np.zeros((len(names or []), 0), dtype=np.float32)
    print('= = Input file %s has been read: Found %i data-like lines in input plumed FES file.' % (fname, data.shape[1]))
    print('= = = %i field entries discovered. Field entries are as follows:' % len(names))
    print(str(names).strip('[]'))
    return names, data


def load_xys(fn):
    """general_scripts.load_xys (:58-67): xmgrace/xvg text -> x (n,), y (n, ncol-1); '#', '@', '&' lines skipped."""
    x, y = [], []
    with open(fn) as fp:
        for l in fp:
            if l == "" or l[0] in "#@&":
                continue
            v = [float(i) for i in l.split()]
            x.append(v[0])
            y.append(v[1:])
    return np.array(x), np.array(y)


def print_xylist(fn, x, ylist, bCols=False, header=""):
    """x (nvals), ylist (nplots, nvals); bCols puts all plots on one line (`%g` columns)."""
    ylist = np.array(ylist)
    with open(fn, 'w') as fp:
        if header != "":
            print(header, file=fp)
        if ylist.ndim == 1:
            for j in range(len(x)):
                print(x[j], ylist[j], file=fp)
            print("&", file=fp)
        elif ylist.ndim == 2:
            if bCols:
                for j in range(ylist.shape[1]):
                    print("%g " % x[j] + " ".join("%g" % ylist[i][j] for i in range(ylist.shape[0])), file=fp)
                print("&", file=fp)
            else:
                for i in range(ylist.shape[0]):
                    for j in range(len(x)):
                        print(x[j], ylist[i][j], file=fp)
                    print("&", file=fp)


def print_sxylist(fn, legend, x, ylist, header=[]):
    """One xmgrace set per legend entry; each row is `x` followed by str() of the y-row without brackets."""
    ylist = np.array(ylist)
    with open(fn, 'w') as fp:
        for line in header:
            print("%s" % line, file=fp)
        for s in range(len(ylist)):
            print("@s%d legend \"%s\"" % (s, legend[s]), file=fp)
            for j in range(len(x)):
                print(x[j], str(ylist[s][j]).strip('[]'), file=fp)
            print("&", file=fp)


def print_s3d(fn, legend, arr, cols, header=[]):
    """gs.print_s3d (general_scripts.py:292-307): one xmgrace set per arr[i], rows of "%g" of the chosen columns.
    The shipped loop reuses its set counter for the row string and stops with a TypeError after the first set;
    this writes every set, numbered as the first one is there.  Rows are formatted a set at a time."""
    arr = np.asarray(arr)
    cols = list(cols)
    fmt = " ".join(["%g"] * len(cols))
    with open(fn, 'w') as fp:
        for line in header:
            print("%s" % line, file=fp)
        for s in range(arr.shape[0]):
            print("@s%d legend \"%s\"" % (s, legend[s]), file=fp)
            block = arr[s][:, cols]
            if len(block):
                fp.write("\n".join(fmt % tuple(row) for row in block.tolist()))
                fp.write("\n")
            print("&", file=fp)


def load_sxydylist(fn, key="legend"):
    """Reads xmgrace sets `x y [dy]` separated by `&`; returns legends, x, y, dy arrays (dy=[] when absent)."""
    legs, xs, ys, dys = [], [], [], []
    x, y, dy = [], [], []
    with open(fn) as fp:
        for l in fp:
            if l == "" or l == "\n":
                continue
            tok = l.split()
            if l[0] in "#@":
                if key in l:
                    legs.append(tok[-1].strip('"'))
                continue
            if l[0] == "&":
                xs.append(x); ys.append(y)
                if len(dy) > 0:
                    dys.append(dy)
                x, y, dy = [], [], []
                continue
            x.append(float(tok[0])); y.append(float(tok[1]))
            if len(tok) > 2:
                dy.append(float(tok[2]))
    if x != []:
        xs.append(x); ys.append(y); dys.append(dy)
    if dys != []:
        return legs, np.array(xs), np.array(ys), np.array(dys)
    return legs, np.array(xs), np.array(ys), []


def print_gplot_hist(fn, hist, edges, header='', bSphere=False):
    """Gnuplot text form of a histogram, bin centres (general_scripts.py:327-381).  With bSphere the 2-D
    (phi, cos theta) map is closed: pole caps at the first/last cos(theta) edge and a repeat of the first phi
    row at +2 pi."""
    nb = hist.shape
    dim = len(nb)
    ctr = [0.5 * (np.asarray(edges[i])[:-1] + np.asarray(edges[i])[1:]) for i in range(dim)]
    with open(fn, 'w') as fp:
        if header != '':
            print('%s' % header, file=fp)
        print('# DIMENSIONS: %i' % dim, file=fp)
        print("# BINWIDTH: " + " ".join("%g" % ((edges[i][-1] - edges[i][0]) / nb[i]) for i in range(dim)), file=fp)
        print("# NBINS: " + " ".join("%g" % (nb[i]) for i in range(dim)), file=fp)
        if not bSphere:
            for index, val in np.ndenumerate(hist):
                print(" ".join("%g" % ctr[i][index[i]] for i in range(dim)) + " %g" % val, file=fp)
                if index[-1] == nb[-1] - 1:
                    print('', file=fp)
            return
        if dim != 2:
            import sys
            print("= = = ERROR: histogram data is not in 2D, but spherical histogram plotting is requested!", file=sys.stderr)
            sys.exit(1)
        ymin, ymax = edges[1][0], edges[1][-1]

        def row(x, h):
            print('%g %g %g' % (x, ymin, h[0]), file=fp)
            for j in range(nb[1]):
                print('%g %g %g' % (x, ctr[1][j], h[j]), file=fp)
            print('%g %g %g' % (x, ymax, h[-1]), file=fp)
            print('', file=fp)

        for i in range(nb[0]):
            row(ctr[0][i], hist[i])
        row(ctr[0][0] + 2 * np.pi, hist[0])


def write_to_dx(fname, data, dims, orig, abc, units='A', bScaleDat=True):
    """OpenDX scalar grid (dxio.py:79-120): header, then the values in C order, three per line."""
    if units == 'A':
        scale = 10.0
    elif units == 'nm':
        scale = 1.0
    else:
        print("= = ERROR in dxio.py - write_to_dx: units argument of write_to_dx accepts only 'nm' and 'A'!")
        return
    if dims[0] != data.shape[0] or dims[1] != data.shape[1] or dims[2] != data.shape[2]:
        print("= = ERROR in dxio.py - write_to_dx: Data dimensions do not match with the matrix dimenstions!")
        print(dims, data.shape)
        return
    outabc = np.multiply(scale, abc)
    outorig = np.multiply(scale, orig)
    with open(fname, 'w') as fp:
        print('#DX-file written by dxio.py', file=fp)
        print('object 1 class gridpositions counts %i %i %i' % (dims[0], dims[1], dims[2]), file=fp)
        print('origin %g %g %g' % (outorig[0], outorig[1], outorig[2]), file=fp)
        for i in range(3):
            print('delta %g %g %g' % (outabc[i, 0], outabc[i, 1], outabc[i, 2]), file=fp)
        print('object 2 class gridpositions counts %i %i %i' % (dims[0], dims[1], dims[2]), file=fp)
        print('object 3 class array type double rank 0 items %i data follows' % (dims[0] * dims[1] * dims[2]), file=fp)
        flat = np.multiply(1.0 / scale ** 3, data.flatten(order='C')) if bScaleDat else data.flatten(order='C')
        for pos in range(0, len(flat), 3):
            print(" ".join("%g" % v for v in flat[pos:pos + 3]), file=fp)
        print('', file=fp)
        print('object "density [%s^-3]" class field' % units, file=fp)

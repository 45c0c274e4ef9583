"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

Restatement of the J(omega) -> R1/R2/NOE evaluation of spectral_densities.py (the classes
angularFrequencies, globalRotationalDiffusion_{Isotropic,Axisymmetric}, spinRelaxation{R1,R2,NOE}) and of
the `npufunc.Jomega` ufunc (Jomega/Jomega.c).  Pinned by tests/golden, generated from the reference
classes with the gcc-built reference ufunc (oracle/_ref).
"""
import numpy as np

GAMMA = {"1H": 267.513e6, "13C": 67.262e6, "15N": -27.116e6, "17O": -36.264e6, "19F": 251.662e6,
         "31P": 108.291e6}                                  # gyromag.set_gamma, spectral_densities.py:50-67
DEFAULT_CSA = {"15N": -170e-6, "13C": -130e-6}                # gyromag.reset_csa, :39-48
TIME_FACT_PS = 1.0e-12                                       # _return_time_fact('ps')
R_NH_NM = 1.02e-1                                            # angularFrequencies.rAB, :164
DD_CONST = 1.1121216813552401e-82                            # (mu0 hbar / 4 pi)^2, :239


def jomega(x, y):
    """npufunc.Jomega, Jomega/Jomega.c:49-66 (double loop): x / (x*x + y*y), elementwise with broadcasting."""
    x = np.asarray(x)
    y = np.asarray(y)
    return x / (x * x + y * y)


def omegas(field_mhz, nucA="15N", nucB="1H"):
    """angularFrequencies.__init__ + set_magnetic_field('MHz'), spectral_densities.py:153-175,187-195.
    Returns (omega[5] in rad/ps, B0 in T)."""
    B0 = 2.0 * np.pi * field_mhz / 267.513
    om = np.zeros(5)
    om[1] = -1.0 * GAMMA[nucA] * B0 * TIME_FACT_PS
    om[3] = -1.0 * GAMMA[nucB] * B0 * TIME_FACT_PS
    om[2] = om[3] - om[1]
    om[4] = om[3] + om[1]
    return om, B0


def factor_dd(nucA="15N", nucB="1H"):
    """get_factor_DD, :225-239 (dist_fact for nm = 1e-9)."""
    return 0.10 * DD_CONST * GAMMA[nucA] ** 2.0 * GAMMA[nucB] ** 2.0 * (R_NH_NM * 1e-9) ** -6.0


def factor_csa(B0, csa, nucA="15N"):
    """get_factor_CSA, :241-243."""
    return 2.0 / 15.0 * np.asarray(csa) ** 2.0 * (GAMMA[nucA] * B0) ** 2


def hist_to_vectors(hist, edges):
    """convert_LambertCylindricalHist_to_vecs, spectral_densities.py:2334-2350 + gm.rtp_to_xyz(bUnit=True),
    general_maths.py:176-180.  Returns bin vectors (B,3) [identical for every residue] and weights (nR,B)."""
    phis = 0.5 * (edges[0][:-1] + edges[0][1:])
    thetas = np.arccos(0.5 * (edges[1][:-1] + edges[1][1:]))
    P, T = np.meshgrid(phis, thetas, indexing="ij")
    vec = np.stack((np.cos(P) * np.sin(T), np.sin(P) * np.sin(T), np.cos(T)), axis=-1).reshape(-1, 3)
    return vec, np.reshape(hist, (hist.shape[0], -1))


def a_coefficients(vec, prolate=True):
    """update_A_coefficients, spectral_densities.py:503-523."""
    z2 = np.square(vec[..., -1] if prolate else vec[..., 0])
    om = 1 - z2
    return np.stack((3.0 * z2 * om, 0.75 * np.square(om), 0.25 * np.square(3.0 * z2 - 1.0)), axis=-1)


def d_coefficients(Diso, Dani):
    """transform_D (:535-540) + D_coefficients_symmtop (:1874-1884)."""
    Dperp = 3.0 * Diso / (2.0 + Dani)
    Dpar = Dani * Dperp
    return np.array([5 * Dperp + Dpar, 2 * Dperp + 4 * Dpar, 6 * Dperp])


def j_axisymmetric(om, A_J, D_J, S2, C, tau, zeta=1.0):
    """calc_Jomega_one, :552-557 with _do_Jsum, :1961-1972.  A_J (...,3) -> J (...,5)."""
    J = np.einsum("...j,jk", zeta * S2 * A_J, jomega(D_J[:, None], om[None, :]))
    for c, t in zip(C, tau):
        Dk = D_J + 1.0 / t
        J = J + np.einsum("...j,jk", zeta * c * A_J, jomega(Dk[:, None], om[None, :]))
    return J


def j_isotropic(om, Diso, S2, C, tau, zeta=1.0):
    """globalRotationalDiffusion_Isotropic.calc_Jomega_one, :430-443."""
    tg = 1.0 / (6.0 * Diso)
    J = zeta * S2 * tg / (1.0 + (om * tg) ** 2.0)
    for c, t in zip(C, tau):
        k = 1.0 / tg + 1.0 / t
        J = J + zeta * c * k / (k ** 2.0 + om ** 2.0)
    return J


def r1_of_j(f_dd, f_csa, J):
    """spinRelaxationR1.func, :824-829."""
    return TIME_FACT_PS * (f_dd * (J[..., 2] + 3 * J[..., 1] + 6 * J[..., 4]) + f_csa * J[..., 1])


def r2_of_j(f_dd, f_csa, J):
    """spinRelaxationR2.func, :859-864."""
    return TIME_FACT_PS * (0.5 * f_dd * (4 * J[..., 0] + J[..., 2] + 3 * J[..., 1] + 6 * J[..., 4] + 6 * J[..., 3])
                           + 1.0 / 6.0 * f_csa * (4 * J[..., 0] + 3 * J[..., 1]))


def noe_of_j(f_dd, R1, J, nucA="15N", nucB="1H"):
    """spinRelaxationNOE.func, :888-892."""
    return 1.0 + TIME_FACT_PS * GAMMA[nucB] / (GAMMA[nucA] * R1) * f_dd * (6 * J[..., 4] - J[..., 2])


def wavg_std(x, w):
    """check_and_calculate_average for one residue, :751-763 (weights = histogram counts)."""
    v = np.average(x, weights=w)
    return v, np.sqrt(np.average((x - v) ** 2.0, weights=w))


def relax_axisymmetric(field_mhz, Diso, Dani, vec, weights, models, csa=None, zeta=1.0):
    """eval() of R1, R2, NOE for every residue (spinRelaxation*.eval, :831-853, :866-875, :894-907) with an
    axisymmetric global tumbling model and a bin-vector distribution.  models = list of (S2, C[], tau[]);
    weights (nR, B); csa scalar or (nR,).  NOE uses the bin-averaged R1 (quirk G8).
    Returns dict name -> (values (nR,), errors (nR,))."""
    om, B0 = omegas(field_mhz)
    f_dd = factor_dd()
    prolate = Dani > 1
    A = a_coefficients(vec, prolate)
    D_J = d_coefficients(Diso, Dani)
    nR = len(models)
    csa = np.broadcast_to(np.asarray(DEFAULT_CSA["15N"] if csa is None else csa, dtype=float), (nR,))
    out = {k: (np.zeros(nR), np.zeros(nR)) for k in ("R1", "R2", "NOE")}
    for i, (S2, C, tau) in enumerate(models):
        J = j_axisymmetric(om, A, D_J, S2, C, tau, zeta)
        f_csa = factor_csa(B0, csa[i])
        r1 = r1_of_j(f_dd, f_csa, J)
        v1, e1 = wavg_std(r1, weights[i])
        v2, e2 = wavg_std(r2_of_j(f_dd, f_csa, J), weights[i])
        vn, en = wavg_std(noe_of_j(f_dd, v1, J), weights[i])
        for k, (v, e) in (("R1", (v1, e1)), ("R2", (v2, e2)), ("NOE", (vn, en))):
            out[k][0][i] = v
            out[k][1][i] = e
    return out


def relax_isotropic(field_mhz, Diso, models, csa=None, zeta=1.0):
    """Same for isotropic tumbling (no vector averaging; errors are None in the reference)."""
    om, B0 = omegas(field_mhz)
    f_dd = factor_dd()
    nR = len(models)
    csa = np.broadcast_to(np.asarray(DEFAULT_CSA["15N"] if csa is None else csa, dtype=float), (nR,))
    J = np.array([j_isotropic(om, Diso, S2, C, tau, zeta) for (S2, C, tau) in models])
    f_csa = factor_csa(B0, csa)
    r1 = r1_of_j(f_dd, f_csa, J)
    return {"R1": r1, "R2": r2_of_j(f_dd, f_csa, J), "NOE": noe_of_j(f_dd, r1, J)}

"""Cubed-sphere topology for the device library: global node ids and the
covector re-basing matrices at panel seams.

Panel orientation follows CubedSphereTrans::XYZFromXYP (reference
src/atm/CubedSphereTrans.cpp:25-83): with gnomonic (X, Y) = (tan alpha,
tan beta) the point of panel p lies in the direction

    p0 (1, X, Y)   p1 (-X, 1, Y)   p2 (-1, -X, Y)   p3 (X, -1, Y)
    p4 (-Y, X, 1)  p5 (Y, X, -1)

The reference locates coincident nodes through its exchange-buffer topology
(Grid.cpp:1066-1573); here every GLL node gets the integer point of the cube
surface it sits on as a global id, so duplicates are simply equal ids.
"""
import numpy as np


def gll_points(n, x0=0.0, x1=1.0):
    """Gauss-Lobatto-Legendre nodes and weights on [x0, x1]
    (GaussLobattoQuadrature::GetPoints, reference
    src/base/GaussLobattoQuadrature.cpp)."""
    if n < 2:
        raise ValueError("at least two GLL points")
    # interior nodes are the roots of P'_{n-1}
    c = np.zeros(n)
    c[-1] = 1.0
    dp = np.polynomial.legendre.legder(c)
    xi = np.concatenate(([-1.0], np.sort(np.polynomial.legendre.legroots(dp)), [1.0]))
    # polish with Newton on P'_{n-1} and evaluate weights 2/(n(n-1)P_{n-1}^2)
    for _ in range(3):
        d1 = np.polynomial.legendre.legval(xi[1:-1], dp)
        d2 = np.polynomial.legendre.legval(xi[1:-1], np.polynomial.legendre.legder(dp))
        xi[1:-1] -= d1 / d2
    pn = np.polynomial.legendre.legval(xi, c)
    w = 2.0 / (n * (n - 1) * pn * pn)
    x = x0 + 0.5 * (x1 - x0) * (xi + 1.0)
    return x, 0.5 * (x1 - x0) * w


def unique_index(g, np_):
    """Repeated (element-local) global index -> unique node index."""
    g = np.asarray(g)
    return (g // np_) * (np_ - 1) + (g % np_)


def lattice_point(panel, s, t, n):
    """Integer cube-surface point of (panel, s, t) with s,t,n as in
    node_ids; returns (x, y, z) arrays."""
    nn = np.full_like(s, n)
    if panel == 0:
        return nn, s, t
    if panel == 1:
        return -s, nn, t
    if panel == 2:
        return -nn, -s, t
    if panel == 3:
        return s, -nn, t
    if panel == 4:
        return -t, s, nn
    if panel == 5:
        return t, s, -nn
    raise ValueError("invalid panel")


def panel_coords(panel, x, y, z):
    """Inverse of lattice_point for a point that lies on `panel`."""
    if panel == 0:
        return y, z
    if panel == 1:
        return -x, z
    if panel == 2:
        return -y, z
    if panel == 3:
        return x, z
    if panel == 4:
        return y, -x
    if panel == 5:
        return y, x
    raise ValueError("invalid panel")


def panels_containing(x, y, z, n):
    out = []
    if x == n:
        out.append(0)
    if y == n:
        out.append(1)
    if x == -n:
        out.append(2)
    if y == -n:
        out.append(3)
    if z == n:
        out.append(4)
    if z == -n:
        out.append(5)
    return out


def node_ids(panel, nelem_a, nelem_b, elem_a0, elem_b0, ne, np_):
    """Global ids [nelem_a*np][nelem_b*np] of a patch whose first element is
    (elem_a0, elem_b0) on a panel of ne x ne elements."""
    n = (np_ - 1) * ne
    ga = elem_a0 * np_ + np.arange(nelem_a * np_)
    gb = elem_b0 * np_ + np.arange(nelem_b * np_)
    s = 2 * unique_index(ga, np_) - n
    t = 2 * unique_index(gb, np_) - n
    S, T = np.meshgrid(s, t, indexing="ij")
    x, y, z = lattice_point(panel, S, T, n)
    m = 2 * n + 1
    return ((x + n).astype(np.int64) * m + (y + n)) * m + (z + n)


def _direction(panel, X, Y):
    one = 1.0
    return {
        0: np.array([one, X, Y]),
        1: np.array([-X, one, Y]),
        2: np.array([-one, -X, Y]),
        3: np.array([X, -one, Y]),
        4: np.array([-Y, X, one]),
        5: np.array([Y, X, -one]),
    }[panel]


def _ddirection(panel):
    """d(direction)/dX and d(direction)/dY."""
    return {
        0: (np.array([0., 1., 0.]), np.array([0., 0., 1.])),
        1: (np.array([-1., 0., 0.]), np.array([0., 0., 1.])),
        2: (np.array([0., -1., 0.]), np.array([0., 0., 1.])),
        3: (np.array([1., 0., 0.]), np.array([0., 0., 1.])),
        4: (np.array([0., 1., 0.]), np.array([-1., 0., 0.])),
        5: (np.array([0., 1., 0.]), np.array([1., 0., 0.])),
    }[panel]


def covariant_basis(panel, X, Y):
    """Tangent vectors d r / d alpha, d r / d beta on the unit sphere."""
    d = _direction(panel, X, Y)
    dX, dY = _ddirection(panel)
    r = np.sqrt(d @ d)
    ex = dX / r - d * (d @ dX) / r ** 3
    ey = dY / r - d * (d @ dY) / r ** 3
    return ex * (1.0 + X * X), ey * (1.0 + Y * Y)


def seam_matrix(p_dst, Xd, Yd, p_src, Xs, Ys):
    """2x2 M with (u_alpha, u_beta)_dst = M (u_alpha, u_beta)_src for a
    covector at one physical point seen from two panels.  Same linear map as
    CubedSphereTrans::CoVecPanelTrans (reference CubedSphereTrans.h:1751-2275),
    derived from the geometry instead of the 18 per-panel-pair formulas."""
    ea_d, eb_d = covariant_basis(p_dst, Xd, Yd)
    ea_s, eb_s = covariant_basis(p_src, Xs, Ys)
    g = np.array([[ea_s @ ea_s, ea_s @ eb_s], [eb_s @ ea_s, eb_s @ eb_s]])
    gi = np.linalg.inv(g)
    ca_s = gi[0, 0] * ea_s + gi[0, 1] * eb_s     # contravariant basis of src
    cb_s = gi[1, 0] * ea_s + gi[1, 1] * eb_s
    return np.array([[ea_d @ ca_s, ea_d @ cb_s], [eb_d @ ca_s, eb_d @ cb_s]])


def alpha_nodes(elem0, nelem, ne, np_):
    """alpha (or beta) of the element-local nodes of nelem elements starting
    at element elem0 (GridSpacingGaussLobattoRepeated::GetNode, reference
    src/atm/GridSpacing.cpp:180-197)."""
    delta = 0.5 * np.pi / ne
    g, _ = gll_points(np_, 0.0, delta)
    e = elem0 + np.arange(nelem)
    return ((-0.25 * np.pi) + delta * e[:, None].astype(np.float64) + g[None, :]).reshape(-1)


def seam_transforms(panel, nelem_a, nelem_b, elem_a0, elem_b0, ne, np_,
                    anode=None, bnode=None):
    """Seam entries of one patch: (ia, ib, src_panel, M[4]) for every interior
    node lying on a panel edge, for every other panel containing the point.
    anode / bnode: alpha, beta of the patch's interior nodes (defaults to
    alpha_nodes)."""
    n = (np_ - 1) * ne
    if anode is None:
        anode = alpha_nodes(elem_a0, nelem_a, ne, np_)
    if bnode is None:
        bnode = alpha_nodes(elem_b0, nelem_b, ne, np_)
    ga = elem_a0 * np_ + np.arange(nelem_a * np_)
    gb = elem_b0 * np_ + np.arange(nelem_b * np_)
    sa = 2 * unique_index(ga, np_) - n
    tb = 2 * unique_index(gb, np_) - n
    # tan of the unique-node angle for any unique index (symmetric table)
    full = alpha_nodes(0, ne, ne, np_)
    tan_of = {}
    for g, a in zip(np.arange(ne * np_), full):
        tan_of[int(2 * unique_index(g, np_) - n)] = np.tan(a)
    ia_l, ib_l, sp_l, m_l = [], [], [], []
    for ia in range(len(sa)):
        for ib in range(len(tb)):
            if abs(sa[ia]) != n and abs(tb[ib]) != n:
                continue
            x, y, z = lattice_point(panel, np.array(sa[ia]), np.array(tb[ib]), n)
            x, y, z = int(x), int(y), int(z)
            Xd, Yd = np.tan(anode[ia]), np.tan(bnode[ib])
            for q in panels_containing(x, y, z, n):
                if q == panel:
                    continue
                s, t = panel_coords(q, x, y, z)
                M = seam_matrix(panel, Xd, Yd, q, tan_of[int(s)], tan_of[int(t)])
                ia_l.append(ia)
                ib_l.append(ib)
                sp_l.append(q)
                m_l.append(M.reshape(4))
    return (np.array(ia_l, dtype=np.int32), np.array(ib_l, dtype=np.int32),
            np.array(sp_l, dtype=np.int32),
            np.array(m_l, dtype=np.float64).reshape(-1, 4))

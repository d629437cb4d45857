"""CPU oracle for the aperture-field assembly (hot path B).  TEST INFRASTRUCTURE ONLY.

Float64 numpy restatement of the reference's ``nearfield.py`` (build_nearfield,
build_nearfield_big, good_fft_number) and of the two third-party pieces it calls:
scipy ``RegularGridInterpolator`` in linear mode (scipy/interpolate/_rgi.py
``_find_indices`` / ``_evaluate_linear``; version in the dev image 1.18.1, unpinned by
the reference) and ``scipy.spatial.cKDTree.query`` (used here as is).

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import
this module.  Parity pinning: the reference has no tests for this path; the oracle is
pinned by running the unmodified reference on the synthetic lens of
``tests/synth_lens.py`` (``tests/golden/make_nearfield_golden.py``) and is checked against
those fixtures in ``tests/test_oracle_nearfield.py``.

The oracle works on flat arrays of lens points instead of the reference's 2-D boolean
masks, but performs the same arithmetic per point; lines cited are nearfield.py.
"""
import math

import numpy as np
from scipy.spatial import cKDTree

MU0 = 4e-7 * math.pi
C0 = 299792458.0
Z0 = MU0 * C0
NM = 1e-9
inf = float("inf")


def good_fft_number(goal):
    """Smallest 2^a 3^b 5^c >= goal (nearfield.py:30-36)."""
    assert goal < 1e5
    best = None
    p2 = 1
    while p2 < 2e5:
        p3 = p2
        while p3 < 2e5:
            p5 = p3
            while p5 < 2e5:
                if p5 >= goal and (best is None or p5 < best):
                    best = p5
                p5 *= 5
            p3 *= 3
        p2 *= 2
    return best


def trilinear(grid, values, pts):
    """scipy RegularGridInterpolator(method='linear') restated.

    Per axis: interval i with g[i] <= x < g[i+1], clipped to [0, n-2] (so x == g[-1]
    falls in the last interval); t = (x - g[i]) / (g[i+1] - g[i]); value = sum over the 8
    corners of v * prod(weights)."""
    pts = np.asarray(pts, dtype=float)
    idx, t = [], []
    for d in range(3):
        g = np.asarray(grid[d], dtype=float)
        i = np.clip(np.searchsorted(g, pts[:, d], side="right") - 1, 0, g.size - 2)
        den = g[i + 1] - g[i]
        idx.append(i)
        t.append((pts[:, d] - g[i]) / den)
    out = np.zeros(pts.shape[0], dtype=complex)
    for a in (0, 1):
        wa = t[0] if a else 1 - t[0]
        for b in (0, 1):
            wb = t[1] if b else 1 - t[1]
            for c in (0, 1):
                wc = t[2] if c else 1 - t[2]
                out += values[idx[0] + a, idx[1] + b, idx[2] + c] * (wa * wb * wc)
    return out


def _orders(grating_list):
    """Set of (ox,oy) present in any table row (nearfield.py:264, :390) -- same set
    expression as the reference, so the iteration order is the same in one process."""
    return {(e['ox'], e['oy']) for g in grating_list for e in g.data}


def _check_bounds(vals, lo, hi, what, scale=1.0):
    """The six ValueErrors of nearfield.py:294-305 / four of :412-419."""
    if vals.min() < lo:
        raise ValueError('need to calculate at smaller %s!' % what, vals.min() / scale, lo / scale)
    if vals.max() > hi:
        raise ValueError('need to calculate at bigger %s!' % what, vals.max() / scale, hi / scale)


def _accumulate(acc, interpolators, key_head, pts, Hw_x, Hw_y, kx, ky, kz, k_glass, n_glass, phase):
    """Inner polarisation/amplitude loop, nearfield.py:306-327 (periphery) and :420-441 (centre).
    acc = [E_a, E_b, H_a, H_b] in whatever frame the caller works in."""
    for pol, Hw in (('x', Hw_x), ('y', Hw_y)):
        Ew = Hw * Z0                                                       # :308
        for which in ('ampfy', 'ampfx'):
            f = interpolators[key_head + (pol, which)]
            amps = trilinear(f.grid, f.values, pts)                        # :310-311
            if which == 'ampfy':
                acc[0] += Ew * amps * kx * ky / (k_glass * kz) / n_glass * phase            # :313-315
                acc[1] += Ew * amps * (-kx ** 2 - kz ** 2) / (k_glass * kz) / n_glass * phase   # :316-318
                acc[2] += Hw * amps * phase                                               # :319
            else:
                acc[0] += Ew * amps * (ky ** 2 + kz ** 2) / (k_glass * kz) / n_glass * phase   # :321-323
                acc[1] += Ew * amps * -kx * ky / (k_glass * kz) / n_glass * phase           # :324-326
                acc[3] += Hw * amps * phase                                               # :327


def build_nearfield(source_x, source_y, source_z, source_pol, wavelength,
                    lens_periphery_summary, lens_center_summary, hexgridset,
                    x_pts=None, y_pts=None, dipole_moment=1e-30, n_glass_table=None):
    """Same contract as the reference's build_nearfield (nearfield.py:66-480)."""
    assert source_z < 0                                                    # :84
    assert source_pol in ('x', 'y', 'z')                                   # :85
    wl_nm = int(round(wavelength / NM))                                    # :86
    P = lens_periphery_summary
    r_min = np.asarray(P['r_min_list'], float)
    r_max = np.asarray(P['r_max_list'], float)
    r_cen = np.asarray(P['r_center_list'], float)
    gc_of_ring = np.asarray(P['gratingcollection_index_here_list'])
    n_around = np.asarray(P['num_around_circle_list'])
    gp_of_ring = np.asarray(P['grating_period_list'], float)
    gcs = P['gratingcollection_list']
    R = r_max[-1]
    if x_pts is None:                                                      # :95-99
        x_pts = np.linspace(-R, R, num=good_fft_number(2 * R / (wavelength / 2.2)))
    if y_pts is None:                                                      # :100-104
        y_pts = np.linspace(-R, R, num=good_fft_number(2 * R / (wavelength / 2.2)))
    x_pts, y_pts = np.asarray(x_pts, float), np.asarray(y_pts, float)
    for l in (x_pts, y_pts):                                               # :106-109
        d = np.diff(l)
        assert 0 < d[0] < wavelength / 2
        assert d.max() - d.min() <= 1e-9 * np.abs(d).max()
    n_glass = gcs[0].grating_list[0].n_glass                               # :111
    if n_glass == 0:
        if n_glass_table is None:
            raise ValueError('n_glass == 0 needs a dispersion table')
        n_glass = n_glass_table(wl_nm)                                     # :113
    k_glass = 2 * math.pi * n_glass / wavelength
    kvac = 2 * math.pi / wavelength

    X, Y = np.meshgrid(x_pts, y_pts, indexing='ij')                        # :117
    shape = X.shape
    x, y = X.ravel(), Y.ravel()
    r = (x ** 2 + y ** 2) ** 0.5
    phi = np.arctan2(y, x)
    ring = np.searchsorted(np.hstack((r_min, R)), r) - 1                   # :125-126
    in_center = ring == -1                                                 # :127
    ring[ring == len(r_min)] = -1                                          # :128
    in_periph = ring >= 0

    out = [np.zeros(x.size, dtype=complex) for _ in range(4)]              # Ex, Ey, Hx, Hy
    if not in_periph.any() and not in_center.any():                        # :130-134
        z = np.zeros(shape, dtype=complex)
        return z, z, z, z, x_pts, y_pts, 0, n_glass

    # ---- incident dipole / plane-wave field at every sample (:172-228)
    dx, dy, dz = x - source_x, y - source_y, 0 - source_z
    if source_z == -inf:
        assert source_pol != 'z'                                           # :224
        ux, uy, uz = np.zeros_like(x), np.zeros_like(x), np.ones_like(x)
        pv = {'x': (1, 0, 0), 'y': (0, 1, 0)}[source_pol]
        dEx = pv[0] * dipole_moment * np.ones_like(x)
        dEy = pv[1] * dipole_moment * np.ones_like(x)
        dHx = -pv[1] * dipole_moment / Z0 * np.ones_like(x)
        dHy = pv[0] * dipole_moment / Z0 * np.ones_like(x)
    else:
        dist = (dx ** 2 + dy ** 2 + dz ** 2) ** 0.5
        ux, uy, uz = dx / dist, dy / dist, dz / dist
        H_coef = C0 * (2 * math.pi / wavelength) ** 2 * dipole_moment / (4 * math.pi)   # :213
        pv = {'x': (1, 0, 0), 'y': (0, 1, 0), 'z': (0, 0, 1)}[source_pol]
        amp = H_coef * uz ** 0.5 / dist                                    # Lambert sqrt(cos), :216
        dHx = (uy * pv[2] - uz * pv[1]) * amp
        dHy = (uz * pv[0] - ux * pv[2]) * amp
        dHz = (ux * pv[1] - uy * pv[0]) * amp
        dEx = (dHy * uz - dHz * uy) * Z0                                   # :221
        dEy = (dHz * ux - dHx * uz) * Z0                                   # :222

    # ---- periphery (:148-354)
    if in_periph.any():
        ip = np.nonzero(in_periph)[0]
        rg = ring[ip]
        gp = gp_of_ring[rg]                                                # :152
        apg = 2 * math.pi / n_around[rg]                                   # :161
        rc = r_cen[rg]                                                     # :162
        lat = rc * apg                                                     # :165
        rot = (phi[ip] / apg).round() * apg                                # :167  (round half to even)
        c, s = np.cos(rot), np.sin(rot)
        uxp = ux[ip] * c + uy[ip] * s                                      # :195
        uyp = -ux[ip] * s + uy[ip] * c                                     # :196
        xp = x[ip] * c + y[ip] * s - rc                                    # :200
        yp = -x[ip] * s + y[ip] * c                                        # :201
        Hxp_w = dHx[ip] * c + dHy[ip] * s                                  # :231-232
        Hyp_w = -dHx[ip] * s + dHy[ip] * c                                 # :233-234
        H_xp_weight, H_yp_weight = Hyp_w, Hxp_w                            # :246-247
        acc = [np.zeros(ip.size, dtype=complex) for _ in range(4)]         # Exp, Eyp, Hxp, Hyp
        which_gc = gc_of_ring[rg]
        for gi, gc in enumerate(gcs):                                      # :263
            for ox, oy in _orders(gc.grating_list):                        # :264-265
                kxp = kvac * uxp + ox * 2 * math.pi / gp                   # :268
                kyp = kvac * uyp + oy * 2 * math.pi / lat                  # :269
                m = (kxp ** 2 + kyp ** 2 <= kvac ** 2) & (which_gc == gi)  # :279-280
                if not m.any():
                    continue
                kx_, ky_ = kxp[m], kyp[m]
                kz_ = (k_glass ** 2 - kx_ ** 2 - ky_ ** 2) ** 0.5         # :287
                phase = np.exp(1j * (kx_ * xp[m] + ky_ * yp[m]))           # :291
                b = gc.interpolator_bounds
                _check_bounds(uxp[m], b[0], b[1], 'ux')                    # :294-297
                _check_bounds(uyp[m], b[2], b[3], 'uy')                    # :298-301
                _check_bounds(gp[m], b[4], b[5], 'grating_period', NM)     # :302-305
                pts = np.stack((uxp[m], uyp[m], gp[m]), axis=1)            # :293
                sub = [np.zeros(int(m.sum()), dtype=complex) for _ in range(4)]
                _accumulate(sub, gc.interpolators, (wl_nm, (ox, oy)), pts, H_xp_weight[m], H_yp_weight[m],
                            kx_, ky_, kz_, k_glass, n_glass, phase)
                for a, sacc in zip(acc, sub):
                    a[m] += sacc
        if source_z > -inf:                                                # :337-346
            gx, gy = rc * np.cos(rot), rc * np.sin(rot)                    # :170-171
            path = ((gx - source_x) ** 2 + (gy - source_y) ** 2 + source_z ** 2) ** 0.5
            eikr = np.exp(1j * kvac * path)
            acc = [a * eikr for a in acc]
        out[0][ip] = acc[0] * c - acc[1] * s                               # :351
        out[1][ip] = acc[0] * s + acc[1] * c                               # :352
        out[2][ip] = acc[2] * c - acc[3] * s                               # :353
        out[3][ip] = acc[2] * s + acc[3] * c                               # :354

    # ---- centre (:359-466)
    if in_center.any():
        ic = np.nonzero(in_center)[0]
        cells = np.asarray(lens_center_summary, float)
        nearest = cKDTree(cells[:, 0:2]).query(np.stack((x[ic], y[ic]), axis=1))[1]   # :363-364
        cx, cy = cells[nearest, 0], cells[nearest, 1]
        which = cells[nearest, 2].astype(int)                              # :367
        Hw_x, Hw_y = dHy[ic], dHx[ic]                                      # :375-376
        ux_c, uy_c = ux[ic], uy[ic]
        x_period = hexgridset.grating_list[0].grating_period               # :391
        y_period = hexgridset.grating_list[0].lateral_period               # :392
        acc = [np.zeros(ic.size, dtype=complex) for _ in range(4)]
        for ox, oy in _orders(hexgridset.grating_list):                    # :390, :393
            kx = kvac * ux_c + ox * 2 * math.pi / x_period                 # :395
            ky = kvac * uy_c + oy * 2 * math.pi / y_period                 # :396
            m = kx ** 2 + ky ** 2 <= kvac ** 2                             # :398
            if not m.any():
                continue
            kx_, ky_ = kx[m], ky[m]
            kz_ = (k_glass ** 2 - kx_ ** 2 - ky_ ** 2) ** 0.5              # :404
            phase = np.exp(1j * (kx_ * (x[ic][m] - cx[m]) + ky_ * (y[ic][m] - cy[m])))   # :408-409
            b = hexgridset.interpolator_bounds
            _check_bounds(ux_c[m], b[0], b[1], 'ux')                       # :412-415
            _check_bounds(uy_c[m], b[2], b[3], 'uy')                       # :416-419
            pts = np.stack((ux_c[m], uy_c[m], which[m].astype(float)), axis=1)   # :411
            sub = [np.zeros(int(m.sum()), dtype=complex) for _ in range(4)]
            _accumulate(sub, hexgridset.interpolators, (wl_nm, (ox, oy)), pts, Hw_x[m], Hw_y[m],
                        kx_, ky_, kz_, k_glass, n_glass, phase)
            for a, sacc in zip(acc, sub):
                a[m] += sacc
        if source_z > -inf:                                                # :453-461
            path = ((cx - source_x) ** 2 + (cy - source_y) ** 2 + source_z ** 2) ** 0.5
            eikr = np.exp(1j * kvac * path)
            acc = [a * eikr for a in acc]
        for o, a in zip(out, acc):
            o[ic] += a                                                     # :463-466

    inside = in_periph | in_center                                         # :475
    power = ((dEx * dHy - dEy * dHx)[inside].sum()
             * (x_pts[1] - x_pts[0]) * (y_pts[1] - y_pts[0]))              # :474-477
    Ex, Ey, Hx, Hy = (o.reshape(shape) for o in out)
    return Ex, Ey, Hx, Hy, x_pts, y_pts, power, n_glass


def build_nearfield_big(source_x, source_y, source_z, source_pol, wavelength,
                        lens_periphery_summary, lens_center_summary, hexgridset,
                        x_pts=None, y_pts=None, dipole_moment=1e-30, pts_at_a_time=1e7, **kw):
    """y-slab loop of nearfield.py:482-516 (x_pts / y_pts are required there, :489)."""
    per = int(pts_at_a_time / x_pts.size)
    parts, power, n_glass = [], 0, None
    for start in range(0, y_pts.size, per):
        res = build_nearfield(source_x, source_y, source_z, source_pol, wavelength, lens_periphery_summary,
                              lens_center_summary, hexgridset, x_pts=x_pts, y_pts=y_pts[start:start + per],
                              dipole_moment=dipole_moment, **kw)
        parts.append(res[:4])
        power += res[6]
        n_glass = res[7]
    Ex, Ey, Hx, Hy = (np.concatenate([p[i] for p in parts], axis=1) for i in range(4))
    return Ex, Ey, Hx, Hy, x_pts, y_pts, power, n_glass

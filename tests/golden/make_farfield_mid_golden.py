"""Far-field golden at the reference's own typical grid size: 675 x 675 (= good_fft_number() of the SURVEY probe lens,
3^3 5^2: odd, not a power of two), 580 nm, focusing-lens aperture, from the UNMODIFIED reference.  The full 675^2 float64
map would be 3.6 MB; the fixture keeps a seeded random subset of 40000 bins (finite and NaN alike), total_P, the NaN
count and the axes.  Dev container only.

    python tests/golden/make_farfield_mid_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import reference_loader  # noqa: E402
import apertures  # noqa: E402

ref = reference_loader.load()["nearfield_farfield"]
WL = 580e-9
NG = apertures.N_GLASS[580]
M = 675
for name, rot in (("lens675_seed9", False), ("lens675_seed9_rot", True)):
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 9, WL, NG, rotate=rot)
    f = [np.fft.fft2(np.fft.fftshift(a.astype(complex))) for a in (Ex, Ey, Hx, Hy)]
    with contextlib.redirect_stdout(io.StringIO()):
        P, total, ux, uy, dux, duy = ref.farfield_from_nearfield(f[0], f[1], f[2], f[3], list(x), list(y), WL, NG)
    rng = np.random.default_rng(11)
    idx = np.sort(rng.choice(M * M, size=40000, replace=False)).astype(np.int32)
    out = dict(index=idx, P=P.ravel()[idx], P_max=np.nanmax(P), nan_count=int(np.isnan(P).sum()), total_P=total,
               ux=ux.ravel(), uy=uy.ravel(), dux=dux, duy=duy, wavelength=WL, n_glass=NG, shape=np.array(P.shape))
    path = os.path.join(HERE, "farfield_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, P.shape, "nan bins", out["nan_count"], "total_P %.6e" % total, "%.2f MB" % (os.path.getsize(path) / 1e6))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mle_b200, mle_b200.kl as kl
from oracle import oracle as O
from conftest import splitmix_uniforms, ulp_diff
eng = mle_b200.Engine(0)
N01 = (1, [0.0, 1.0])
s, R, n = 100, 16, 2048
z = O.genvec("K_3600_32", s)
lat = eng.lattice(z, s); d = eng.dists([N01] * s)
sh = splitmix_uniforms(2024, R * s).reshape(R, s)
got = eng.points(lat, d, sh, 3, n); want = O.points(z, [N01] * s, sh, 3, n)
u = 2 * O.points(z, None, sh, 3, n) - 1
ul = ulp_diff(got, want)
for name, mask in (("central", np.abs(u) <= .75), ("middle", (np.abs(u) > .75) & (np.abs(u) <= .9375)), ("tail", np.abs(u) > .9375)):
    print(name, mask.sum(), "max ulp", ul[mask].max(), "hist", np.bincount(np.minimum(ul[mask], 10))[:11])
i = np.unravel_index(np.argmax(ul), ul.shape)
print("worst", u[i], got[i], want[i])
# lognormal
m = 100
for level in (2, 4, 6):
    gm = eng.model("lognormal1d", 1, [16]); om = O.Model(O.MODEL_LOGNORMAL1D, 1, [16.0])
    phi = kl.kl_basis_1d(level, m=m); gm.set_basis(level, phi); om.set_basis(level, phi)
    sh2 = sh[:2]
    g, sg = eng.sample_batch(gm, lat, d, [level], sh2, 0, 130, want_samples=True)
    w, sw = O.sample_batch(om, z, [N01] * m, [level], sh2, 0, 130, want_samples=True)
    print("level", level, "Qf rel", np.max(np.abs(sg[:, 0, 1] - sw[:, 0, 1]) / np.abs(sw[:, 0, 1])), "dQ abs", np.max(np.abs(sg[:, 0, 0] - sw[:, 0, 0])),
          "dQ scale", np.abs(sw[:, 0, 0]).max(), "mean rel", np.max(np.abs(g[..., 1] - w[..., 1]) / np.abs(w[..., 1])), "M2 rel", np.max(np.abs(g[..., 2] - w[..., 2]) / w[..., 2]))

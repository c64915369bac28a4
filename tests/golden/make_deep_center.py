#!/usr/bin/env python
"""Generator of the deep-zoom centre used by BASELINE configs[3] ("synthetic 1e-120 deep
zoom"): the Misiurewicz point M(23,2) of the seahorse valley,
c ~ -0.77568377 + 0.13646737i, refined by Newton's method on
g(c) = f_c^(k+p)(0) - f_c^k(0), k = 23, p = 2, to 180 digits with mpmath.  A
pre-periodic point lies on the boundary of the set, so a 1e-120 wide view centred on it
is full of structure and every pixel escapes after a few hundred to a few thousand
iterations.  The digits are pasted into tests/views.py (DEEP120)."""
import mpmath as mp

mp.mp.dps = 220
K, P = 23, 2


def g_and_dg(c):
    z, dz = mp.mpc(0), mp.mpc(0)
    zk = dzk = None
    for i in range(K + P):
        dz = 2 * z * dz + 1
        z = z * z + c
        if i + 1 == K:
            zk, dzk = z, dz
    return z - zk, dz - dzk


c = mp.mpc("-0.77568377", "0.13646737")
for it in range(40):
    g, dg = g_and_dg(c)
    step = g / dg
    c -= step
    if abs(step) < mp.mpf(10) ** -200:
        break
g, _ = g_and_dg(c)
print("iterations", it, "residual", mp.nstr(abs(g), 5))
print("re", mp.nstr(c.real, 170))
print("im", mp.nstr(c.imag, 170))

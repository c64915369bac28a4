#!/usr/bin/env python
"""Generator of the deep-zoom view used by bench.py's default workload and by the full-size
parity tests for BASELINE configs[3] ("synthetic 1e-120 deep zoom", SURVEY 8d config 4:
"a minibrot nucleus refined by Newton iteration with mpmath so that part of the frame
reaches maxiter (imbalanced load)").

Method (mpmath, 300 digits).  Around a Misiurewicz point m the set is self-similar under
the multiplier mu of the cycle m lands on, and carries minibrots at every scale: one at
distance d from m has a size of the order of d^2 and a period of about ln(1/d) / ln|mu|
cycle lengths.  So a copy of the set a tenth of a 1e-120 wide frame across sits ~1e-61
away from m, and pixels around it escape after about eight times its period.
  1. m = M(7,2) = -1.02004618 + 0.36748404i (pre-period 7, period 2, |mu| = 1.472; picked
     so that a frame costs a few seconds on one B200), by Newton's method on
     f_c^9(0) - f_c^7(0);
  2. period of the lowest-period nucleus inside a square box of radius 1e-61 centred
     2e-61 to the right of m: iterate the four corners until the quadrilateral of the
     iterates surrounds the origin (the "box period" method) -> 707;
  3. Newton's method on f_c^707(0) = 0 from the box centre;
  4. complex size of the copy, 1 / (beta * lambda^2) with lambda = prod 2 z_k and
     beta = 1 + sum 1 / lambda_k over the periodic orbit (R. Munafo's estimate):
     |size| = 8.80e-122, rotated by 2.18 rad;
  5. view centre = nucleus + size * (-0.5): the middle of the copy (its cardioid and disc run from
     w = 0.25 to w = -1.25) in the middle of the frame.
Prints the digits pasted into tests/views.py (MINIBROT120, MINIBROT120_NUCLEUS)."""
import mpmath as mp

mp.mp.dps = 300
K, P = 7, 2


def misiurewicz(c):
    for _ in range(80):
        z, dz = mp.mpc(0), mp.mpc(0)
        zk = dzk = None
        for i in range(K + P):
            dz = 2 * z * dz + 1
            z = z * z + c
            if i + 1 == K:
                zk, dzk = z, dz
        step = (z - zk) / (dz - dzk)
        c -= step
        if abs(step) < mp.mpf(10) ** -290:
            break
    return c


def box_period(c0, r, maxp):
    corners = [c0 + mp.mpc(-r, -r), c0 + mp.mpc(r, -r), c0 + mp.mpc(r, r), c0 + mp.mpc(-r, r)]
    z = [mp.mpc(0)] * 4
    for p in range(1, maxp + 1):
        z = [z[k] * z[k] + corners[k] for k in range(4)]
        # does the polygon z0 z1 z2 z3 surround 0 ?  (crossing number of the ray y = 0, x > 0)
        cross = 0
        for k in range(4):
            a, b = z[k], z[(k + 1) % 4]
            if (a.imag > 0) != (b.imag > 0):
                x = a.real - a.imag * (b.real - a.real) / (b.imag - a.imag)
                if x > 0:
                    cross += 1
        if cross & 1:
            return p
    return 0


def newton(c, p, tol):
    for it in range(200):
        z, dz = mp.mpc(0), mp.mpc(0)
        for _ in range(p):
            dz = 2 * z * dz + 1
            z = z * z + c
        step = z / dz
        c = c - step
        if abs(step) < tol:
            return c, it
    return None, it


def size_of(c, p):
    z, lam, beta = mp.mpc(0), mp.mpc(1), mp.mpc(1)
    for _ in range(1, p):
        z = z * z + c
        lam = 2 * z * lam
        beta += 1 / lam
    return 1 / (beta * lam * lam)


def main():
    m = misiurewicz(mp.mpc("-1.020046182", "0.367484035"))
    d = mp.mpf("2e-61")
    centre, r = m + d, d / 2
    p = box_period(centre, r, 40000)
    c, its = newton(centre, p, mp.mpf(10) ** -290)
    if c is None:
        raise SystemExit("Newton did not converge")
    s = size_of(c, p)
    print("period", p, "|size|", mp.nstr(abs(s), 8), "arg", mp.nstr(mp.arg(s), 6), "newton steps", its)
    print("nucleus_re", mp.nstr(c.real, 175))
    print("nucleus_im", mp.nstr(c.imag, 175))
    view = c + s * mp.mpf("-0.5")
    print("cx", mp.nstr(view.real, 175))
    print("cy", mp.nstr(view.imag, 175))


if __name__ == "__main__":
    main()

"""Named views used by the parity tests and bench.py."""
from mdz_b200 import (ImageView, FAMILY_MANDEL, FAMILY_JULIA, MANDELBROT, BURNING_SHIP,
                      GENERALIZED_CELTIC, VARIANT)
from mdz_b200.coords import center_to_rect, rect_to_gmp
from mdz_b200.mp import Mpfr


def make_view(cx, cy, size, w, h, *, mode="mpfr", precision=80, depth=300, aa=1,
              family=FAMILY_MANDEL, fractal=MANDELBROT, julia=None, fixed_re=True):
    """Centre/size view the way MDZ's host would hand it to the render pool."""
    xmin, xmax, ymax, width, crect = center_to_rect(cx, cy, size, w, h, precision)
    v = ImageView(use_multi_prec=(mode != "ld"), use_rounding=(mode == "mpfr"),
                  precision=precision, family=family, fractal=fractal, depth=depth,
                  user_width=w, user_height=h, aa_factor=aa,
                  xmin=xmin, xmax=xmax, ymax=ymax, width=width)
    if mode == "gmp":
        v.gxmin, v.gymax, v.gwidth = rect_to_gmp(crect, precision, fixed_re)
    if julia is not None:
        ip = max(precision, 80)
        v.julia_re, v.julia_im = Mpfr(ip, julia[0]), Mpfr(ip, julia[1])
    return v


# BASELINE.json configs[1]: full M-set, "double" (= long double) precision, maxiter 10k
def config2(w=1920, h=1080, depth=10000):
    return make_view("-0.5", "0.0", "4.0", w, h, mode="ld", depth=depth)


SEAHORSE = ("-0.743643887037158704752191506114774", "0.131825904205311970493132056385139")


def rect_view(xmin, xmax, ymax, w, h, *, mode="mpfr", precision=80, depth=300, aa=1,
              family=FAMILY_MANDEL, fractal=MANDELBROT):
    """A view given directly as the rect the hot path reads (img->xmin, xmax,
    ymax; width = RN(xmax - xmin) as coords_rect_to_center computes it)."""
    from mdz_b200.mp import mpfr
    ip = max(precision, 80)
    a, b, c = Mpfr(ip, xmin), Mpfr(ip, xmax), Mpfr(ip, ymax)
    wd = Mpfr(ip)
    mpfr.mpfr_sub(wd.ref, b.ref, a.ref, 0)
    return ImageView(use_multi_prec=(mode != "ld"), use_rounding=(mode == "mpfr"),
                     precision=precision, family=family, fractal=fractal, depth=depth,
                     user_width=w, user_height=h, aa_factor=aa,
                     xmin=a, xmax=b, ymax=c, width=wd)


# gallery/deep_embedded_julia.mdz as its author meant it:
# MPFR-320, depth 10000, a 3.79e-39 wide window
DEEP_EMBEDDED_JULIA = dict(
    xmin="-1.9990958622795566830287308905472659490602008969484573074469166957563201773778385691107409249442388",
    xmax="-1.9990958622795566830287308905472659490564113825918759307447952539214217882449639333585621571026900",
    ymax="1.2251442643252650599959235781025334506324529252744858584464673344525174648352769071170032669073514e-5")


def deep_embedded_julia(w=1920, h=1080, depth=10000, precision=320):
    d = DEEP_EMBEDDED_JULIA
    return rect_view(d["xmin"], d["xmax"], d["ymax"], w, h, precision=precision, depth=depth)


# gallery/honeytrace.mdz: MPFR-176, depth 25000
HONEYTRACE = ("-7.66701995256949964922178305994168934712886433081717606e-1",
              "1.00264595952135512543352958560880514410135965558537508e-1",
              "2.94676302638774725686015831604561721063962123995835471e-19")


def honeytrace(w=96, h=72, depth=25000):
    return make_view(HONEYTRACE[0], HONEYTRACE[1], HONEYTRACE[2], w, h, precision=176, depth=depth)


# BASELINE configs[3] "synthetic 1e-120 deep zoom": the Misiurewicz point M(23,2) of the
# seahorse valley to 150 digits (tests/golden/make_deep_center.py); every pixel of a
# 1e-120 wide view escapes after ~12 800 iterations
DEEP120 = ("-0.77661059259970185656403950255299474932817032143986600096924365785204686090459870878696142803205041429132937443851811918717130866762136492225227431666272",
           "0.13460896167502816605673727023305780954118749622040361733997923275543996926182434686374673753720867571346542156176411558667716488800448659989231766107654")


def config4(w=3840, h=2160, depth=100000, mode="gmp", precision=512):
    return make_view(DEEP120[0], DEEP120[1], "1e-120", w, h, mode=mode, precision=precision, depth=depth)


# The same configuration on a view that holds a minibrot, as SURVEY 8(d) config 4 asks ("a minibrot
# nucleus refined by Newton iteration ... so that part of the frame reaches maxiter (imbalanced load)"):
# the period-707 nucleus 2e-61 away from the Misiurewicz point M(7,2) = -1.02004618 + 0.36748404i,
# size 8.80e-122 (tests/golden/make_minibrot_center.py).  In a 1e-120 wide 16:9 frame the copy of the
# set covers ~2 % of the pixels (they run to depth); the rest escapes after a few thousand iterations.
# The view is centred on the middle of the copy (nucleus + size * -0.5).
MINIBROT120_NUCLEUS = ("-1.020046182259385217091865714835682856485530230077341008574363862552092767130751929513454984842547245021348825939975602029114599645677072089656815758406526854037775728211996565",
                       "0.3674840351832474379034844057678939574299590406994822242761426141816237076291546180151431726933838064685031011319634748792295760419129867435812162956210919108393798335111629735")
MINIBROT120_PERIOD = 707
MINIBROT120 = ("-1.020046182259385217091865714835682856485530230077341008574363862552092767130751929513454984842547245021348825939975602029089382380660505102952920023722220432497713672616646017",
               "0.3674840351832474379034844057678939574299590406994822242761426141816237076291546180151431726933838064685031011319634748791935173372535570498321917665285018130137968923407166453")


def config4m(w=3840, h=2160, depth=100000, mode="mpfr", precision=512):
    return make_view(MINIBROT120[0], MINIBROT120[1], "1e-120", w, h, mode=mode, precision=precision, depth=depth)


# BASELINE configs[4]: Burning Ship / generalized Celtic, 7680x4320 with 3x3 anti-aliasing
def config5(fractal, w=7680, h=4320, aa=3, depth=1000, mode="ld", precision=64):
    cy = "-0.5" if fractal == BURNING_SHIP else "0.0"
    return make_view("-0.5", cy, "4.0", w, h, mode=mode, precision=precision, depth=depth, aa=aa, fractal=fractal)


def gmp_close_path_view(precision, w=8, h=6, depth=3000, julia=("-1", "0.3")):
    """A GMP-mode Julia view 1e-130 wide whose top left pixel is x = 2^-32, y = 2^-32 - 2^-98 exactly (ix = 0 and
    line 0 reproduce gxmin and gymax bit for bit, src/fractal.c:310-328) and whose other pixels differ from it only
    some 370 bits further down.  For every pixel x^2 = 1:0:0:... (a one, then zero limbs) and y^2 = ff..ff:7f..f:...
    one limb below it, so the first wre2 - wim2 of EVERY pixel is GMP's one-limb-gap "close" subtraction (SURVEY
    Appendix E) -- an operand pattern ordinary views never produce -- with different low limbs each time."""
    from mdz_b200 import FAMILY_JULIA
    v = make_view("0", "0", "1e-130", w, h, mode="gmp", precision=precision, depth=depth, family=FAMILY_JULIA, julia=julia)
    def put(m, limbs, exp):
        for i, l in enumerate(limbs):
            m.s.d[i] = l
        m.s.size = len(limbs)
        m.s.exp = exp
    put(v.gxmin, [1 << 32], 0)
    put(v.gymax, [(1 << 64) - (1 << 30), (1 << 32) - 1], 0)
    return v

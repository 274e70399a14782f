"""The oracle's restatement of Go's math functions (oracle/gomath.c) against high-precision references.
A wrong polynomial constant shows up as an error of many ulps; the algorithms themselves are <1 ulp."""
import math

import numpy as np
from scipy import special


def ulps(a, b):
    if a == b:
        return 0.0
    return abs(a - b) / math.ulp(b)


def test_log_exp_pow(orc):
    L = orc.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([10 ** rng.uniform(-300, 300, 2000), rng.uniform(0.5, 2.0, 2000), [1.0, 2.0, 10.0, 0.1]])
    for x in xs:
        x = float(x)
        assert ulps(L.gm_log(x), math.log(x)) <= 1.0
        # Go's Log10 = Log2(x)*(Ln2/Ln10) with Log2 = Log(frac)/Ln2 + exp: cancels near x~1, so bound the absolute error
        assert abs(L.gm_log10(x) - math.log10(x)) <= 2.3e-16 * max(1.0, abs(math.log2(x)))
    for x in rng.uniform(-700, 700, 3000):
        assert ulps(L.gm_exp(float(x)), math.exp(float(x))) <= 1.0
    for i in range(0, 94):
        y = i / -10.0
        assert ulps(L.gm_pow(10.0, y), 10.0 ** y) <= 6.0
    for y in rng.uniform(-9.3, 0, 2000):
        assert ulps(L.gm_pow(10.0, float(y)), 10.0 ** float(y)) <= 6.0  # Go pow = exp(yf*log x) * x^yi by squaring: a few ulps
    assert L.gm_pow(10.0, -1.0) == 0.1 and L.gm_pow(10.0, 0.0) == 1.0 and L.gm_pow(10.0, -0.5) == 1 / math.sqrt(10.0)
    assert L.gm_log2(8.0) == 3.0 and L.gm_log10(1.0) == 0.0


def test_lgamma(orc):
    L = orc.lib()
    ns = list(range(1, 400)) + [10 ** k for k in range(3, 10)] + [2 ** 31 - 1, 2 ** 31, 123456789]
    for n in ns:
        ref = float(special.gammaln(float(n)))
        got = L.gm_lgamma(float(n))
        assert got == 0.0 if n in (1, 2) else ulps(got, ref) <= 4.0, n
    rng = np.random.default_rng(1)
    for x in np.concatenate([rng.uniform(0.01, 8, 3000), rng.uniform(8, 1e6, 1000)]):
        ref = float(special.gammaln(float(x)))
        got = L.gm_lgamma(float(x))
        assert abs(got - ref) <= 4 * math.ulp(ref) + 4 * math.ulp(1.0), x  # absolute bound near the zeros at 1 and 2


def test_round(orc):
    L = orc.lib()
    for x, e in [(0.5, 1.0), (-0.5, -1.0), (2.5, 3.0), (2.4999999999999996, 2.0), (-2.5, -3.0), (39.5, 40.0)]:
        assert L.gm_round(x) == e

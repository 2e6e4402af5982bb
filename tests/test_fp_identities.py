"""The floating-point identities the instruction trims of the fast composite kernels rest on (CPU, numpy float32 = IEEE binary32 with one
rounding per operation, like the kernels' FMUL / FADD; products of two float32 are exact in float64).

The trims (csrc/ts2d_fast.cuh: eval_fast; csrc/ts2d_render_fwd_fast.cu; csrc/ts2d_render_bwd_fast.cu) replace

    power = -0.5 * pw;  G = ex2(power * log2e)            by   G = ex2(pw * (-0.5 * log2e))
    fma(c1, |power|, c0)                                   by   fma(0.5 * c1, pw, c0)
    -3 * dL_dpower * 2 * power * rcp        (gamma == 1)   by   ((3 * dL_dpower) * pw) * rcp

and are claimed to keep every bit because a factor of two commutes with rounding.  That holds away from the subnormal range; the
kernels evaluate these expressions only for pairs that contribute (pw in [1e-8, 100], alpha >= 1/255), far from it."""
import numpy as np

F = np.float32
LOG2E = F(1.4426950408889634)


def _samples(n=1_000_000, seed=0):
    rng = np.random.default_rng(seed)
    pw = np.exp(rng.uniform(np.log(1e-8), np.log(100.0), n)).astype(F)  # ecc^(2 gamma) of a contributing pair
    edge = np.array([1e-8, 1.0, 2.0, 11.0, 11.0900001, 99.99, np.nextafter(F(1), F(2)), np.nextafter(F(4), F(0))], dtype=F)
    return np.concatenate([pw, edge])


def _bits(x):
    return np.asarray(x, dtype=F).view(np.uint32)


def test_exp_argument_in_one_multiply():
    pw = _samples()
    old = (F(-0.5) * pw) * LOG2E
    new = pw * (F(-0.5) * LOG2E)
    assert F(-0.5) * LOG2E == F(-0.5 * float(LOG2E))  # the folded constant is exact
    assert np.array_equal(_bits(old), _bits(new))


def test_error_model_term_from_pw():
    pw = _samples(seed=1)
    for c1 in (F(3.84e-6), F(2.7e-5), F(1.8e-4)):  # terr_c1 for gamma = 1, 7, 50
        # fma rounds the exact product + addend once: equal exact products => equal results
        old = np.float64(c1) * np.float64(np.abs(F(-0.5) * pw))
        new = np.float64(F(0.5) * c1) * np.float64(pw)
        assert np.array_equal(old, new)


def test_gamma_one_chain_rule_factor():
    rng = np.random.default_rng(2)
    pw = _samples(seed=3)
    n = pw.size
    d = (rng.standard_normal(n) * np.exp(rng.uniform(-30, 5, n))).astype(F)  # dL/dpower: any sign, many magnitudes
    d[:8] = [0.0, -0.0, 1.0, -1.0, 3e-38, -3e-38, 1e30, -1e30]
    rcp = (1.0 / (np.sqrt(pw.astype(np.float64)) + 1e-8)).astype(F)  # ~ 1 / (ecc + eps)
    power = F(-0.5) * pw
    with np.errstate(over="ignore", under="ignore"):
        old = (((F(-3.0) * d) * F(2.0)) * power) * rcp
        new = ((F(3.0) * d) * pw) * rcp
    normal = np.abs(new.astype(np.float64)) > 1e-36  # away from the subnormal range every bit agrees
    assert np.array_equal(_bits(old[normal]), _bits(new[normal]))
    assert np.allclose(old[~normal], new[~normal], rtol=0, atol=1e-37)

"""The shared numeric contract (pht_math.h, pht_philox.h), exercised on the host build."""
import ctypes as C

import numpy as np

from oracle import pyoracle as po


def _philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); out = (C.c_uint32 * 4)()
    po.oracle().lib.pho_philox(c, k, out)
    return [int(x) for x in out]


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert _philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_exp_log_within_one_ulp_of_libm():
    L = po.oracle()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-745, 709, 40000), rng.uniform(-2, 2, 40000), rng.normal(0, 1e-3, 10000)])
    got = np.array([L.pho_exp(float(x)) for x in xs]); ref = np.exp(xs)
    assert (np.abs(got - ref) <= np.spacing(ref)).all()
    xs = np.concatenate([rng.uniform(0, 1, 40000) ** 8 * 10, rng.uniform(0.5, 2, 40000), np.exp(rng.uniform(-700, 700, 20000))])
    xs = xs[xs > 0]
    got = np.array([L.pho_log(float(x)) for x in xs]); ref = np.log(xs)
    assert (np.abs(got - ref) <= np.spacing(np.abs(ref) + 1e-300)).all()
    assert L.pho_exp(710.0) == np.inf and L.pho_exp(-746.0) == 0.0 and L.pho_exp(0.0) == 1.0
    assert L.pho_log(0.0) == -np.inf and np.isnan(L.pho_log(-1.0)) and L.pho_log(1.0) == 0.0
    assert abs(L.pho_log(5e-324) - np.log(5e-324)) < 1e-12


def test_exp_fast_path_is_bit_identical_to_the_general_path():
    """pht_exp scales by the exponent field for |x| <= 708 and falls back to the two-step scaling elsewhere; wherever
    both apply they must agree to the last bit (the golden vectors were produced with the general path)."""
    L = po.oracle()
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(-708, 708, 60000), rng.uniform(-1, 1, 20000), np.nextafter(708.0, 0) * np.array([1.0, -1.0]),
                         np.array([708.0, -708.0, 0.0, -0.0, 1e-300, -1e-300, 0.5 * np.log(2), -0.5 * np.log(2)]),
                         np.log(2.0) * rng.integers(-1021, 1021, 2000) + rng.normal(0, 1e-12, 2000)])
    for x in xs:
        a = L.pho_exp(float(x)); b = L.pho_exp_general(float(x))
        assert a == b and np.signbit(a) == np.signbit(b), x
    for x in (708.0000001, -708.0000001, 709.5, -740.0, -745.2, 710.0, np.inf, -np.inf):
        assert L.pho_exp(x) == L.pho_exp_general(x)
    assert np.isnan(L.pho_exp(np.nan))


def test_uniform_single_subtraction_is_exact():
    """pht_u01 builds (x + 0.5) 2^-52 as 1.x - (1 - 2^-53); check against exact integer arithmetic."""
    L = po.oracle()
    from fractions import Fraction
    for k in range(300):
        c = (C.c_uint32 * 4)(k, 0, 5, 1); key = (C.c_uint32 * 2)(7, 0); out = (C.c_uint32 * 4)()
        L.lib.pho_philox(c, key, out)
        x = ((int(out[1]) << 32) | int(out[0])) >> 12
        want = Fraction(2 * x + 1, 2 ** 53)
        assert Fraction(L.pho_unif_at(7, 1, 5, 0, 2 * k)) == want


def test_uniforms_are_open_interval_and_keyed():
    L = po.oracle()
    u = np.array([L.pho_unif_at(7, 1, k, 0, d) for k in range(200) for d in range(6)])
    assert (u > 0).all() and (u < 1).all() and abs(u.mean() - 0.5) < 0.03
    assert L.pho_unif_at(7, 1, 5, 3, 2) == L.pho_unif_at(7, 1, 5, 3, 2)
    assert L.pho_unif_at(7, 1, 5, 3, 2) != L.pho_unif_at(7, 1, 5, 4, 2)
    assert L.pho_unif_at(7, 1, 5, 3, 2) != L.pho_unif_at(8, 1, 5, 3, 2)


def test_gamma_sampler_moments():
    L = po.oracle()
    for shape, scale in ((0.5, 2.0), (2.0, 0.25), (180.0, 1 / 16.0), (1e5, 1e-5)):
        x = np.array([L.pho_rgamma_at(3, 1, k, shape, scale) for k in range(20000)])
        assert (x > 0).all()
        assert abs(x.mean() - shape * scale) < 5 * np.sqrt(shape) * scale / np.sqrt(20000) + 1e-12
        assert abs(x.var() / (shape * scale * scale) - 1) < 0.1

"""r2p2_b200/csrc/r2p2_math.h against this box's libm (the arithmetic the reference reaches through CPython `math`)."""
import ctypes as C

import pytest

from oracle import oracle as orc


@pytest.fixture(scope="module")
def mc():
    L = C.CDLL(orc.build()[2])
    L.mathcheck_sincos.restype = C.c_long
    L.mathcheck_sincos.argtypes = [C.c_long, C.c_uint64, C.c_double, C.c_double, C.POINTER(C.c_double)]
    L.mathcheck_sincos_degrees.restype = C.c_long
    L.mathcheck_sincos_degrees.argtypes = [C.c_long, C.c_uint64, C.c_double]
    L.mathcheck_maxulp.restype = C.c_double
    L.mathcheck_maxulp.argtypes = [C.c_int, C.c_long, C.c_uint64]
    return L


@pytest.mark.parametrize("lo,hi", [(0, 1e-7), (0, 0.126), (0.126, 0.855469), (0.85, 0.86), (0.855469, 2.426265), (2.42, 2.43),
                                   (2.426265, 10), (10, 1000), (1000, 1e6), (1e6, 1.05e8)])
def test_sincos_bit_identical_to_libm(mc, lo, hi):
    """Every branch of glibc 2.39 __sin_fma/__cos_fma: 2e6 arguments per range, zero bit mismatches
    (2.2e8 arguments were checked when the restatement was written)."""
    bad_x = C.c_double()
    assert mc.mathcheck_sincos(2_000_000, 12345, lo, hi, C.byref(bad_x)) == 0, bad_x.value


def test_sincos_on_simulator_arguments(mc):
    assert mc.mathcheck_sincos_degrees(3_000_000, 99, 4000.0) == 0      # (a + theta) * pi/180, |theta| < 2000 degrees
    assert mc.mathcheck_sincos_degrees(1_000_000, 7, 720.0) == 0


def test_own_functions_stay_close_to_libm(mc):
    """atan2 / tanh / logistic are library-independent restatements: not bit-identical, <= 4 ulp."""
    assert mc.mathcheck_maxulp(0, 2_000_000, 1) <= 4.0
    assert mc.mathcheck_maxulp(1, 2_000_000, 2) <= 4.0
    assert mc.mathcheck_maxulp(2, 2_000_000, 3) <= 6.0

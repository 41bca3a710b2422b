"""CPU: exhaustive checks of the two division sequences the CUDA kernels use instead of IEEE divides (the checker is
oracle/div_check.c, test infrastructure).

(1) x / c for the integer counts c = 2..16 of the TIES disjoint mean (csrc/ties.cu: div_by_count_fast) and of PCB's final
    / n (csrc/pcb.cu: pcb_div_count): reciprocal multiply + one FMA correction -- bit-identical to IEEE division for
    every fp32 mantissa, both signs, at the exponents around 2^0 and at both ends of the range the kernels allow
    ([2^-100, 2^100]; outside it they take the IEEE divide).
(2) x / y with a prepared reciprocal and two FMA corrections (csrc/pcb.cu: pcb_div_by): correctly rounded for every fp32
    mantissa of x, for random divisors and for the classical hard ones (all-ones mantissa, 1 + ulp, near sqrt 2, ...).
    (The checker also counts the quotients that are wrong after ONE correction: none on these divisors, and none on
    10^11 quotients of a wider random search (6,000 divisors x all mantissas x 2 binades) -- the kernels keep the second correction because only the two-step form is
    covered by Markstein's theorem: q0 = RN(x r) may be 1.5 ulp off, i.e. not faithful.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


@pytest.fixture(scope="module")
def chk():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    lib = ctypes.CDLL(os.path.join(ORACLE_DIR, "_build", "libdivcheck.so"))
    lib.orc_check_div_count.restype = ctypes.c_uint64
    lib.orc_check_div_count.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.orc_check_div_chain.restype = ctypes.c_uint64
    lib.orc_check_div_chain.argtypes = [ctypes.c_float, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]
    return lib


@pytest.mark.parametrize("c", list(range(2, 17)))
def test_divide_by_count_is_ieee_for_every_mantissa(chk, c):
    for exp2 in (0, 1, -100, 99):        # 2 x 2^23 quotients each: x = +-(1.m) 2^exp2
        assert chk.orc_check_div_count(c, exp2) == 0, (c, exp2)


def _divisors():
    rng = np.random.default_rng(20261019)
    mant = rng.integers(0, 1 << 23, size=40, dtype=np.uint32)
    special = np.array([0, 1, 2, (1 << 23) - 1, (1 << 23) - 2, 1 << 22, (1 << 22) - 1, (1 << 22) + 1,
                        0x3504F3, 0x3504F4, 0x2AAAAA, 0x2AAAAB, 0x555555, 0x555556], dtype=np.uint32)   # sqrt 2, 4/3, 5/3 ...
    bits = (np.uint32(127) << np.uint32(23)) | np.concatenate([mant, special])
    ys = bits.view(np.float32).tolist()
    return ys + [y * 2.0 ** -40 for y in ys[:4]] + [y * 2.0 ** 50 for y in ys[:4]]


def test_prepared_reciprocal_two_corrections_is_correctly_rounded(chk):
    for y in _divisors():
        for exp2 in (0, -1):             # quotients on both sides of 1
            one = ctypes.c_uint64(0)
            e = exp2 + int(np.floor(np.log2(y)))          # keep x within a binade of y, as in the kernels (0 <= x <= y)
            assert chk.orc_check_div_chain(ctypes.c_float(y), e, ctypes.byref(one)) == 0, (y, exp2)

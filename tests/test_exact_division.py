"""The particle kernels form t = (p - lo) / delta with `div_exact` (spacecharge.jl_b200/csrc/particle_common.cuh): a
multiplication by the host-rounded reciprocal followed by two Markstein correction steps, five FP64 instructions
instead of the division sequence.  The cell index floor(t) must be bit-identical to the reference's true division
(src/deposition.jl:39-41, src/interpolation.jl:31-33), so the sequence is restated here in exact rational arithmetic
(every fused multiply-add rounded once, like the hardware) and compared with the correctly rounded quotient: random
operands, quotients next to integers (where a wrong last bit would change the cell), spacings whose significand is all
ones or a power of two, and the ranges a mesh can produce."""
from fractions import Fraction

import numpy as np


def fma(x, y, z):
    return float(Fraction(x) * Fraction(y) + Fraction(z))   # one rounding (Fraction -> float is correctly rounded)


def div_exact(a, d):
    rinv = 1.0 / d
    q0 = a * rinv
    e0 = fma(-q0, d, a)
    q1 = fma(e0, rinv, q0)
    e1 = fma(-q1, d, a)
    return fma(e1, rinv, q1)


def check(a, d):
    want = float(Fraction(a) / Fraction(d))
    assert want == a / d                    # IEEE division is the correctly rounded quotient
    got = div_exact(a, d)
    assert got == want, (a.hex(), d.hex(), got.hex(), want.hex())


def test_random_operands():
    rng = np.random.default_rng(1)
    a = rng.random(4000) * 10.0 ** rng.integers(-12, 3, 4000)
    d = (0.5 + rng.random(4000)) * 10.0 ** rng.integers(-9, 0, 4000)
    for x, y in zip(a, d):
        check(float(x), float(y))
        check(-float(x), float(y))


def test_quotients_next_to_cell_boundaries():
    rng = np.random.default_rng(2)
    for _ in range(3000):
        d = float((0.5 + rng.random()) * 10.0 ** rng.integers(-8, -2))
        k = int(rng.integers(0, 1024))
        a = k * d                           # rounded product: quotient within an ulp of the integer k
        for _ in range(3):
            check(a, d)
            a = float(np.nextafter(a, np.inf))
        a = float(np.nextafter(k * d, -np.inf))
        check(a, d)


def test_special_spacings():
    ones = float.fromhex("0x1.fffffffffffffp-14")    # significand of all ones
    for d in (ones, 2.0 ** -20, 1e-6, float(np.float32(3.3e-5)), float(np.nextafter(2.0 ** -10, 1.0))):
        rng = np.random.default_rng(3)
        for a in rng.random(500) * 300 * d:
            check(float(a), d)
        check(0.0, d)
        check(d, d)
        check(255 * d, d)


# ---- the same sequence in Float32 (all-Float32 kernels: Float32 particles on a Float32 mesh) --------------------------
def f32(v):
    return np.float32(v)


def fma32(x, y, z):
    return f32(float(Fraction(float(x)) * Fraction(float(y)) + Fraction(float(z))))   # exact, then ONE rounding to Float32


def div_exact32(a, d):
    rinv = f32(1.0) / d
    q0 = a * rinv
    e0 = fma32(-q0, d, a)
    q1 = fma32(e0, rinv, q0)
    e1 = fma32(-q1, d, a)
    return fma32(e1, rinv, q1)


def check32(a, d):
    a, d = f32(a), f32(d)
    want = f32(float(Fraction(float(a)) / Fraction(float(d))))
    assert want == a / d
    got = div_exact32(a, d)
    assert got == want, (float(a).hex(), float(d).hex(), float(got).hex(), float(want).hex())


def test_float32_sequence():
    rng = np.random.default_rng(4)
    for _ in range(4000):
        d = f32((0.5 + rng.random()) * 10.0 ** rng.integers(-7, -1))
        check32(f32(rng.random() * 10.0 ** rng.integers(-8, 0)), d)
        k = int(rng.integers(0, 1024))
        a = f32(k) * d
        check32(a, d)
        check32(np.nextafter(a, f32(np.inf)), d)
        check32(np.nextafter(a, f32(-np.inf)), d)
    for d in (f32(float.fromhex("0x1.fffffep-14")), f32(2.0 ** -20), f32(1e-6)):
        for a in rng.random(300).astype(np.float32) * f32(300) * d:
            check32(a, d)

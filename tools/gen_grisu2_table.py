"""Prints the cached powers of ten of csrc/host/Grisu2.h: for k = -300, -292, ..., 324 the 64-bit significand f and
binary exponent e with 10^k ~= f * 2^e, 2^63 <= f < 2^64, f rounded to nearest (exact rational arithmetic)."""
from fractions import Fraction


def entry(k):
    x = Fraction(10) ** k
    e = (x.numerator.bit_length() - x.denominator.bit_length()) - 63
    while x / Fraction(2) ** e >= 2 ** 64:
        e += 1
    while x / Fraction(2) ** e < 2 ** 63:
        e -= 1
    s = x / Fraction(2) ** e
    f = s.numerator // s.denominator
    if (s - f) * 2 >= 1:
        f += 1
    if f == 2 ** 64:
        f >>= 1; e += 1
    return f, e


if __name__ == "__main__":
    for k in range(-300, 325, 8):
        f, e = entry(k)
        print(f"    {{0x{f:016X}ULL, {e}, {k}}},")

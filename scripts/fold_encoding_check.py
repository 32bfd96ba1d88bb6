#!/usr/bin/env python
"""Arithmetic check of the planned exact in-contraction screen - design evidence, not product code.

A 1000-bit hash leaves K positions 1000..1023 of the 1024-wide contraction free.  With the row operand's real bits at 2.0,
the row's free positions holding 6.0 (x23) and 2.0 (x1), and the column's free positions holding C - pc(j) spelled in e2m1
digits, the tensor core returns 2 dot - pc(j) + C.  This script checks, for every popcount 0..1000 and C = 800, that the digits
exist (at most 23 + 1), that every digit is an e2m1 value, and that the sum is exact.
"""
E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]          # magnitudes of the 8 e2m1 codes
REM = {0: (), 1: (0.5,), 2: (1.0,), 3: (1.5,), 4: (2.0,), 5: (2.0, 0.5), 6: (3.0,), 7: (3.0, 0.5), 8: (4.0,), 9: (4.0, 0.5),
       10: (4.0, 1.0), 11: (4.0, 1.5)}                    # m mod 12 as a sum of (2 x digit) from {1,2,3,4,6,8}
C = 800


def digits(x: int):
    """-> (23 digits paired with the row's 6.0, 1 digit paired with the row's 2.0), signs included"""
    sign = -1.0 if x < 0 else 1.0
    a = abs(x)
    m, r = divmod(a, 3)                                    # |x| = 3 m + r ;  6 * d = 3 * (2 d)  ->  sum of (2 d) must be m
    big = [6.0] * (m // 12) + list(REM[m % 12])
    assert len(big) <= 23, (x, len(big))
    big += [0.0] * (23 - len(big))
    last = r / 2.0                                         # 2 * last = r
    return [sign * d for d in big], sign * last


def main():
    worst = 0
    for pc in range(0, 1001):
        x = C - pc
        big, last = digits(x)
        assert all(abs(d) in E2M1 for d in big) and abs(last) in E2M1
        total = sum(6.0 * d for d in big) + 2.0 * last
        assert total == x, (pc, total)
        worst = max(worst, sum(1 for d in big if d))
    # the accumulator: 2 * dot + (C - pc) with dot <= 1000 stays far inside fp32's exact-integer range, every term is an integer
    print(f"ok: C - pc(j) representable for pc = 0..1000 with C = {C}; at most {worst} of 23 wide digits used; |value| <= {max(C, 1000 - C)}")


if __name__ == "__main__":
    main()

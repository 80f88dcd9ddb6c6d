"""div_by_const (csrc/gpd_math.cuh): the FP64 kernels divide the force by the drone mass (BaseAviary.py:855) with a
host-computed correctly rounded reciprocal and two FMAs.  Markstein's theorem says the result is the correctly rounded
quotient; this test checks the identity bit for bit against the plain division on random operands, for every drone mass
of the shipped models, with the same C arithmetic (fma from libm, no contraction)."""
import os
import subprocess
import tempfile

import gpd_b200  # noqa: F401
from gpd_b200.params import load_drone_params
from gpd_b200.utils.enums import DroneModel

SRC = r"""
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t nxt(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static double div_by_const(double a, double b, double rb)
{
    const double m = fabs(a);
    if (!(m > 1e-280 && m < 1e280)) return a / b;
    const double q0 = a * rb;
    const double r = fma(-q0, b, a);
    return fma(r, rb, q0);
}
int main(int argc, char** argv)
{
    long bad = 0, n = atol(argv[1]);
    for (int k = 2; k < argc; ++k) {
        const double b = atof(argv[k]);
        const volatile double rb = 1.0 / b;
        for (long i = 0; i < n; ++i) {
            uint64_t u = nxt();
            /* random sign, random mantissa, exponent spread over 2^-80 .. 2^80 (forces of 1e-24 .. 1e24 N), plus raw bit patterns */
            double a;
            if (i % 8 == 7) { memcpy(&a, &u, 8); }
            else {
                uint64_t e = 1023 - 80 + (nxt() % 161);
                uint64_t bits = (u & 0x800FFFFFFFFFFFFFull) | (e << 52);
                memcpy(&a, &bits, 8);
            }
            double q = div_by_const(a, b, rb), t = a / b;
            if (memcmp(&q, &t, 8) != 0 && !(q != q && t != t)) { if (bad < 5) printf("b=%.17g a=%.17g q=%.17g t=%.17g\n", b, a, q, t); ++bad; }
        }
    }
    printf("bad=%ld\n", bad);
    return bad != 0;
}
"""


def test_div_by_const_equals_ieee_division_bit_for_bit():
    masses = sorted({repr(float(load_drone_params(m).M)) for m in DroneModel})
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "dc.c"), os.path.join(d, "dc")
        open(c, "w").write(SRC)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, c, "-lm"], check=True)
        out = subprocess.run([exe, "10000000"] + masses + ["60.0", "0.027", "0.83"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("bad=0"), out.stdout[-2000:]

// b2_math.cuh -- the transcendental function of the model (sin/cos of the heading, src/models/robot_model.py:39-40) as an
// explicit sequence of IEEE-754 operations, shared verbatim by the device build and by the CPU mirror build of the kernels
// (oracle/simt/): libdevice's sincos and glibc's differ in the last bits, this one is the same instructions on both sides.
//
// Cody-Waite reduction by pi/2 with a three-part constant and fused multiply-adds (error of the reduced argument
// < 1e-33 |n|, fine for every heading an MPC iterate can take: |psi| < 1e9), then the minimax polynomials of
// fdlibm's __kernel_sin / __kernel_cos on [-pi/4, pi/4] (public domain coefficients, the pair stays within 1.6 ulp).
// Checked against numpy on the CPU in tests/test_mirror.py (max error 1.6 ulp of the result over |x| < 1e5).
#pragma once
#include <math.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline __attribute__((always_inline))
#endif
#endif

__host__ __device__ __forceinline__ void b2_sincos(double x, double* sn, double* cs)
{
    const double n = rint(x * 0.63661977236758138);                 // 2/pi
    double r = fma(-n, 1.5707963267948966, x);                      // pi/2 = hi + mid + lo
    r = fma(-n, 6.123233995736766e-17, r);
    r = fma(-n, -1.4973849048591698e-33, r);
    const double z = r * r;
    // sin(r) = r + r^3 (S1 + z (S2 + ... ))
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s = fma(z * r, ps, r);
    // cos(r) = 1 - z/2 + z^2 (C1 + z (C2 + ... ))
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;
    const double c = w + (((1.0 - w) - hz) + (z * z) * pc);
    // quadrant: n mod 4 (n is integral; |n| < 2^31 for every sane heading, otherwise the result is garbage like any sincos)
    const int q = (int)n & 3;
    const double s1 = (q & 1) ? c : s, c1 = (q & 1) ? s : c;
    *sn = (q & 2) ? -s1 : s1;
    *cs = ((q + 1) & 2) ? -c1 : c1;
}

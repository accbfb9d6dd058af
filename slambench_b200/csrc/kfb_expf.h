// Bit-faithful expf for the bilateral filter.
//
// The reference's bilateralFilterKernel (kfusion/src/cpp/kernels.cpp:186-188) calls the host
// libm `expf`.  glibc (>= 2.27; 2.39 in this image) implements expf with the ARM
// optimized-routines algorithm: the argument is widened to double, split as
// x*32/ln2 = k + r, and exp(x) = 2^(k/32) * P(r) is evaluated in double with a 32-entry
// table and a cubic, then rounded once to float — correctly rounded in all but
// astronomically rare cases.  CUDA's own expf is a 2-ulp fp32 routine, which would make
// the filtered depth (and everything downstream: pyramid, vertex/normal maps, ICP) differ
// from the reference in the last bit.  Restating the published algorithm in fp64 keeps
// the whole preprocessing stage bit-exact; the cost (25 exp/pixel, ~10 DP ops each) is
// negligible next to the HBM traffic.
//
// Verified exhaustively on the CPU against glibc expf (FMA ifunc variant) for every float in [-104.5, 0]
// (tests/test_host_math.py::test_expf_exhaustive) — the only domain the filter uses:
// the argument is -(d0-d1)^2 / (2 e_d^2) <= 0.
#ifndef KFB_EXPF_H
#define KFB_EXPF_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define KFB_HD __host__ __device__ __forceinline__
#else
#define KFB_HD static inline
#endif

// T[i] = bits(2^(i/32)) - (i << 47)
#define KFB_EXPF_TAB_INIT \
{ \
	0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
	0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
	0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
	0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
	0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
	0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
	0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
	0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL \
}

static const uint64_t kfb_exp2f_tab_host[32] = KFB_EXPF_TAB_INIT;
#if defined(__CUDACC__)
// global (not __constant__) memory: lanes index it divergently, which the constant cache would serialise
__device__ const uint64_t kfb_exp2f_tab_dev[32] = KFB_EXPF_TAB_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define KFB_EXPF_TAB kfb_exp2f_tab_dev
#else
#define KFB_EXPF_TAB kfb_exp2f_tab_host
#endif

KFB_HD uint64_t kfb_d2u(double d) {
#if defined(__CUDA_ARCH__)
	return (uint64_t) __double_as_longlong(d);
#else
	uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
KFB_HD double kfb_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
	return __longlong_as_double((long long) u);
#else
	double d; memcpy(&d, &u, 8); return d;
#endif
}

// expf for x <= 0 (and small positive x); not a general replacement: no overflow branch.
KFB_HD float kfb_expf_nonpos(float x) {
	if (x < -0x1.9fe368p6f) return 0.0f;  // underflow: glibc returns +0 here
	const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
	const double Shift = 0x1.8p+52;
	const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
	const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
	const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
	const double xd = (double) x;
	double z = InvLn2N * xd;
	double kd = z + Shift;
	const uint64_t ki = kfb_d2u(kd);
	kd -= Shift;
	const double r = fma(InvLn2N, xd, -kd);  // fused in glibc's FMA build (exhaustive check confirms)
	uint64_t t = KFB_EXPF_TAB[ki % 32];
	t += ki << (52 - 5);
	const double s = kfb_u2d(t);
	// glibc selects its FMA build of this routine at run time on every x86-64 CPU that
	// has FMA (sysdeps/x86_64/fpu/multiarch/e_expf.c), so these steps (and r above) are fused.
	z = fma(C0, r, C1);
	const double r2 = r * r;
	double y = fma(C2, r, 1.0);
	y = fma(z, r2, y);
	y = y * s;
	return (float) y;
}

#endif

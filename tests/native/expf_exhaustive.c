/* CPU check of slambench_b200/csrc/kfb_expf.h against the host libm expf (the function
 * the reference's bilateral filter calls): every float in [-104.5, -0.0] plus +0.
 * Prints "mismatches N of M"; exit code 0 iff N == 0.  Built and run by tests/test_host_math.py. */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "kfb_expf.h"

int main(int argc, char** argv) {
	float lo = -104.5f;
	uint32_t ulo, stride = argc > 1 ? (uint32_t) atoi(argv[1]) : 1;
	memcpy(&ulo, &lo, 4);
	long long bad = 0, total = 0;
#pragma omp parallel for reduction(+:bad,total) schedule(static)
	for (long long u = 0x80000000LL; u <= (long long) ulo; u += stride) {
		uint32_t b = (uint32_t) u;
		float x; memcpy(&x, &b, 4);
		float a = expf(x), c = kfb_expf_nonpos(x);
		uint32_t ua, uc; memcpy(&ua, &a, 4); memcpy(&uc, &c, 4);
		if (ua != uc) { if (bad < 5) fprintf(stderr, "x=%a libm=%a ours=%a\n", x, a, c); bad++; }
		total++;
	}
	if (kfb_expf_nonpos(0.0f) != 1.0f) bad++;
	printf("mismatches %lld of %lld\n", bad, total);
	return bad != 0;
}

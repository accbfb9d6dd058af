/* Evidence for profiles/r1_v17_summary.md ("running average with a table reciprocal"): for every divisor d = 1..128 and
 * EVERY fp32 significand of the numerator n, q = n*y, q + fma(-d, q, n)*y with y = RN(1/d) equals the IEEE quotient n/d
 * (the quotient scales exactly with the numerator's binade, so one binade suffices for normal operands).
 *     gcc -O2 -ffp-contract=off tools/div_small_exhaustive.c -lm -o /tmp/div_small && /tmp/div_small     (~15 s)
 * The experiment that used it (k_integrate_run's update, -14 instructions per voxel) was measured slower and is not in
 * the tree; the program stays as the record of the claim. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
int main(void) {
	long bad = 0, n = 0;
	for (int b = 1; b <= 128; b++) {
		const float den = (float) b, y = 1.0f / den;
		for (uint32_t m = 0; m < (1u << 23); m++) {
			const float num = asf(0x3f800000u | m);
			const float q0 = num * y;
			const float q1 = fmaf(fmaf(-den, q0, num), y, q0);
			n++;
			if (q1 != num / den) bad++;
		}
	}
	printf("checked %ld quotients, mismatches %ld\n", n, bad);
	return bad != 0;
}

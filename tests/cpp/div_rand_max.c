/* Exhaustive check of the division-free x / RAND_MAX used by msc-project_b200/csrc/rng.cu
 * (div_rand_max): for every integer 0 <= x < 2^31 the FMA-corrected product equals the correctly
 * rounded quotient x / 2147483647.0 that the reference computes (edmd.cpp:206-207).
 * Prints the number of mismatches (must be 0) and, for contrast, of the uncorrected product. */
#include <math.h>
#include <stdio.h>

int main(void) {
    const double M = 2147483647.0, c = 1.0 / 2147483647.0;
    long bad = 0, bad_plain = 0;
#pragma omp parallel for reduction(+ : bad, bad_plain)
    for (long x = 0; x < 2147483648L; x++) {
        const double xd = (double)x;
        const double q0 = xd * c;
        const double r = fma(-q0, M, xd);
        const double q = fma(r, c, q0);
        const double ref = xd / M;
        if (q != ref) bad++;
        if (q0 != ref) bad_plain++;
    }
    printf("%ld %ld\n", bad, bad_plain);
    return bad != 0;
}

/* mathcheck.c -- TEST INFRASTRUCTURE.  Compares the restated glibc sin/cos of
 * r2p2_b200/csrc/r2p2_math.h (the routines the CUDA kernels run) against this box's libm, and reports
 * the ulp distance of the library-independent atan2/tanh/logistic to libm. */
#include <math.h>
#include <stdint.h>
#include "../r2p2_b200/csrc/r2p2_math.h"

static inline uint64_t xs(uint64_t* s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
static inline double u01(uint64_t* s) { return (double)(xs(s) >> 11) * (1.0 / 9007199254740992.0); }

/* n random arguments, |x| uniform in [lo, hi), random sign; returns the number of bit mismatches of
 * sin, cos or the fused sincos against libm. first_bad receives one offending argument. */
long mathcheck_sincos(long n, uint64_t seed, double lo, double hi, double* first_bad) {
    uint64_t s = seed ? seed : 88172645463325252ULL;
    long bad = 0;
    for (long i = 0; i < n; ++i) {
        double x = lo + (hi - lo) * u01(&s);
        if (xs(&s) & 1) x = -x;
        int ok = 1;
        double ps = r2p2_sin(x, r2p2_h_sincostab, &ok), pc = r2p2_cos(x, r2p2_h_sincostab, &ok), s2, c2;
        r2p2_sincos(x, r2p2_h_sincostab, &s2, &c2, &ok);
        const double ls = sin(x), lc = cos(x);
        if (r2p2_bits(ps) != r2p2_bits(ls) || r2p2_bits(pc) != r2p2_bits(lc) || r2p2_bits(s2) != r2p2_bits(ls) || r2p2_bits(c2) != r2p2_bits(lc)) {
            if (!bad && first_bad) *first_bad = x;
            bad++;
        }
    }
    return bad;
}

/* the arguments the simulator really produces: (a + theta) * pi/180 */
long mathcheck_sincos_degrees(long n, uint64_t seed, double span_deg) {
    uint64_t s = seed ? seed : 1;
    const double RAD = 3.14159265358979323846 / 180.0;
    long bad = 0;
    for (long i = 0; i < n; ++i) {
        const double x = ((u01(&s) - 0.5) * span_deg) * RAD;
        int ok = 1;
        if (r2p2_bits(r2p2_sin(x, r2p2_h_sincostab, &ok)) != r2p2_bits(sin(x)) || r2p2_bits(r2p2_cos(x, r2p2_h_sincostab, &ok)) != r2p2_bits(cos(x))) bad++;
    }
    return bad;
}

static double ulps(double a, double b) {
    if (a == b) return 0;
    if (a != a || b != b) return (a != a && b != b) ? 0 : 1e300;
    const double u = fabs(nextafter(b, INFINITY) - b);
    return fabs(a - b) / u;
}

/* which: 0 atan2 (y, x uniform in [-10,10], some scaled tiny), 1 tanh, 2 logistic.  returns max ulp distance */
double mathcheck_maxulp(int which, long n, uint64_t seed) {
    uint64_t s = seed ? seed : 7;
    double mx = 0;
    for (long i = 0; i < n; ++i) {
        double d;
        if (which == 0) {
            double y = (u01(&s) - 0.5) * 20, x = (u01(&s) - 0.5) * 20;
            if (i % 7 == 0) y *= 1e-6;
            if (i % 11 == 0) x *= 1e-6;
            d = ulps(r2p2_atan2(y, x), atan2(y, x));
        } else if (which == 1) {
            const double x = (u01(&s) - 0.5) * (i % 3 == 0 ? 60 : (i % 3 == 1 ? 4 : 0.2));
            d = ulps(r2p2_tanh(x), tanh(x));
        } else {
            const double x = (u01(&s) - 0.5) * (i % 2 ? 80 : 8);
            d = ulps(r2p2_logistic(x), 1.0 / (1.0 + exp(-x)));
        }
        if (d > mx) mx = d;
    }
    return mx;
}

double mathcheck_atan2(double y, double x) { return r2p2_atan2(y, x); }
double mathcheck_tanh(double x) { return r2p2_tanh(x); }

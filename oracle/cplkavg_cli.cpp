// TEST INFRASTRUCTURE ONLY.  Caller of the reference's `double cplkavg(double, double, double)` through the
// reference's own header (cplkavg.h:7): reads "wvllo wvlhi T" triples from stdin, prints one %.17g per line.
// oracle/Makefile links it against the reference's cplkavg.cpp (cplkavg_cli_ref) and against librcm_b200.so
// (cplkavg_cli_b200) - the exact-signature drop-in check of SURVEY.md section 8(b).
#include <cstdio>

#include "cplkavg.h"

int main() {
    double lo, hi, t;
    while (std::scanf("%lf %lf %lf", &lo, &hi, &t) == 3) std::printf("%.17g\n", cplkavg(lo, hi, t));
    return 0;
}

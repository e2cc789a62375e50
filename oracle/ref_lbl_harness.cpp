// TEST INFRASTRUCTURE ONLY (oracle).  Exposes the two vendored libRadtran helpers of the
// reference as C entry points: cplkavg() (cplkavg.cpp:124-243, band-integrated Planck) and
// ASCII_file2xy2D() (lbl.arts/ascii.cpp:1631-1691, the line-by-line table reader).
// Separate translation unit from ref_harness.cpp because main.cpp defines its own cplkavg
// overloads.  Sources are compiled from /root/reference unmodified.
#include <cstdlib>
#include <cstring>

#include "cplkavg.h"
#include "lbl.arts/ascii.h"

extern "C" {

double ref_cplkavg(double wvllo, double wvlhi, double t) { return cplkavg(wvllo, wvlhi, t); }

void ref_cplkavg_many(int n, const double* lo, const double* hi, const double* t, double* out) {
    for (int i = 0; i < n; ++i) out[i] = cplkavg(lo[i], hi[i], t[i]);
}

// Two-call protocol: pass x == NULL to query sizes, then again with buffers x[nx], y[nx*ny].
int ref_ascii_file2xy2D(const char* filename, int* nx, int* ny, double* x, double* y) {
    double* xx = NULL;
    double** yy = NULL;
    int n1 = 0, n2 = 0;
    int status = ASCII_file2xy2D(const_cast<char*>(filename), &n1, &n2, &xx, &yy);
    if (status != 0) return status;
    *nx = n1;
    *ny = n2;
    if (x && y) {
        for (int i = 0; i < n1; ++i) {
            x[i] = xx[i];
            std::memcpy(y + (size_t)i * n2, yy[i], n2 * sizeof(double));
        }
    }
    if (xx) std::free(xx);
    if (yy) ASCII_free_double(yy, n1);
    return 0;
}

}  // extern "C"

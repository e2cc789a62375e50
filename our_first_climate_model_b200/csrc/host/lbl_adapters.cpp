// Reference-signature adapters of the line-by-line side (include/rcm_b200_adapters.hpp) over the C ABI.
//   ASCII_file2xy2D     <- lbl.arts/ascii.cpp:1631-1691   (ascii.h:63; only call site lbl.arts/testlblarts.cpp:24-26)
//   ASCII_free_double   <- lbl.arts/ascii.cpp:955-965     (how the caller releases y, ascii.cpp:1612-1613)
//   cplkavg             <- cplkavg.cpp:124-243            (cplkavg.h:7)
// Same names, linkage, ownership and error behaviour as the reference, so lbl.arts/testlblarts.cpp links against
// librcm_b200.so unchanged (oracle/Makefile: _ref/testlblarts_b200).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../../include/rcm_b200_adapters.hpp"
#include "rcm_internal.h"

extern "C" int ASCII_file2xy2D(char* filename, int* nx, int* ny, double** x, double*** y) {
    int rows = 0, cols = 0;
    double *xf = nullptr, *yf = nullptr;
    const int st = rcm_ascii_file2xy2D(filename, &rows, &cols, &xf, &yf);
    if (st != 0) return st < 0 ? st : -1;  // the reference's negative codes (ascii.h:34-38)
    // the reference's layout (ASCII_calloc_double, ascii.cpp:463-481): one calloc'ed row per wavelength, so that the
    // caller's free(x) / ASCII_free_double(y, nx) release it
    double** yy = rows > 0 ? (double**)std::calloc((size_t)rows, sizeof(double*)) : nullptr;
    bool ok = rows == 0 || yy != nullptr;
    for (int r = 0; ok && r < rows && cols > 0; ++r) {
        yy[r] = (double*)std::calloc((size_t)cols, sizeof(double));
        if (!yy[r]) ok = false; else std::memcpy(yy[r], yf + (size_t)r * cols, (size_t)cols * sizeof(double));
    }
    rcm_free(yf);
    if (!ok) {
        if (yy) {
            for (int r = 0; r < rows; ++r) std::free(yy[r]);
            std::free(yy);
        }
        rcm_free(xf);
        return -2;  // ASCII_NO_MEMORY
    }
    *nx = rows;
    *ny = cols;
    *x = xf;  // calloc'ed by rcm_ascii_file2xy2D: free(x) is valid
    *y = yy;
    return 0;
}

extern "C" int ASCII_free_double(double** value, int rows) {
    for (int i = 0; i < rows; ++i) std::free(value[i]);
    std::free(value);
    return 0;
}

// Error behaviour of the reference's c_errmsg (cplkavg.cpp:29-48): bad arguments print and exit(1); the two
// warnings go to stderr, at most 100 of them.
double cplkavg(double wvllo, double wvlhi, double t) {
    static int n_warn = 0;
    int st = 0;
    const double v = rcm_cplkavg_host(wvllo, wvlhi, t, &st);
    if (st == 1) {
        std::fprintf(stderr, "\n ******* ERROR >>>>>>  planck_func1--temperature or wavenums. wrong\n");
        std::exit(1);
    }
    if (st != 0 && ++n_warn <= 100)
        std::fprintf(stderr, "\n ******* WARNING >>>>>>  %s\n", st == 2 ? "planck_func1--Simpson rule didn't converge"
                                                                        : "planck_func1--returns zero; possible underflow");
    return v;
}

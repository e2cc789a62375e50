// C++ adapters with EXACTLY the reference's call signatures, implemented on top of the C ABI
// (rcm_b200.h).  Linking librcm_b200.so instead of repwvl_thermal.cpp makes the reference driver
// (main.cpp) run its optical-depth build on the GPU unchanged; its radiative_transfer() can be
// swapped the same way (INTEGRATION.md shows the link recipe).
#ifndef RCM_B200_ADAPTERS_HPP
#define RCM_B200_ADAPTERS_HPP
#include <vector>

// Replaces read_tau of repwvl_V2.01_cpp/repwvl_thermal.cpp:49-262 (declared in repwvl_thermal.h:3-7).
// Same ownership: *wvl, *weight, *tau and every (*tau)[i] are calloc'ed, the caller frees them.
// Errors (unreadable table, nLev != 21) are reported on stderr and leave *nWvl = 0.
void read_tau(const char* reducedLkpPath, int nLev, std::vector<double>& plevel, std::vector<double>& Tvector,
              double* H20_VMR, double* CO2_VMR, double* O3_VMR, double* N2O_VMR, double* CO_VMR, double* CH4_VMR,
              double* O2_VMR, double* HNO3_VMR, double* N2_VMR, double*** tau, double** wvl, double** weight,
              int* nWvl, int prop_at_Lev);

// Replaces radiative_transfer of main.cpp:320-344 (B and alpha are scratch in the reference; untouched here).
void radiative_transfer(std::vector<double>& B, std::vector<double>& alpha, std::vector<double>& E_down,
                        std::vector<double>& E_up, std::vector<double>& dE, const double solar_irr,
                        std::vector<double>& mu, const double& dmu, std::vector<double>& Tlayer,
                        const double& T_surface, double** tau, double* weight, int& nwvl, double* wvl);

// Replaces ASCII_file2xy2D of lbl.arts/ascii.cpp:1631-1691 (declared extern "C" in lbl.arts/ascii.h:63).  Same
// ownership: *x is calloc'ed (free it), *y is an array of nx calloc'ed rows of ny doubles (release it with
// ASCII_free_double(y, nx), ascii.cpp:955-965 - exported here too, so a caller that links only this library can).
// Returns 0 or the reference's codes: -1 file not found, -2 no memory, -5 not a rectangular matrix (ascii.h:34-38).
// Unlike the reference's reader (strtok, 1 MiB stack buffers) it is thread-safe and single-pass.
extern "C" int ASCII_file2xy2D(char* filename, int* nx, int* ny, double** x, double*** y);
extern "C" int ASCII_free_double(double** value, int rows);

// Replaces cplkavg of cplkavg.cpp:124-243 (cplkavg.h:7): Planck radiance integrated between two wavelengths [nm],
// W/m2/sr.  Error behaviour of the reference: bad arguments print "planck_func1--temperature or wavenums. wrong"
// and exit(1) (cplkavg.cpp:144-146, :32-35); non-convergence / underflow print a warning.  Do not include this header
// in a translation unit that also defines main.cpp's own cplkavg overloads.
double cplkavg(double wvllo, double wvlhi, double t);
#endif

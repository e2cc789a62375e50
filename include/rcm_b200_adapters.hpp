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
#endif

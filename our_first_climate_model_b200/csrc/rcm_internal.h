// Internal declarations shared by the host sources and the CUDA translation unit.
#ifndef RCM_INTERNAL_H
#define RCM_INTERNAL_H

#include <vector>

#include "../../include/rcm_b200.h"

struct rcm_table {
    int n_tpert = 0, n_species = 0, n_wvl = 0, n_p = 0;
    std::vector<double> xsec, wvl, weight, p_grid, t_ref, t_pert, vmrs_ref;
};

// LowerPos of the reference (repwvl_thermal.cpp:19-45), shared by host setup code.
long rcm_lowerpos_impl(const double* a, int n, double x);

#endif

// TEST INFRASTRUCTURE ONLY.  Caller of the reference's read_tau through the reference's own header
// (repwvl_thermal.h:3-7) for inputs the reference driver never exercises - above all prop_at_Lev != 0
// (repwvl_thermal.cpp:219-224: temperatures and mixing ratios given at the 21 LEVELS).
//   read_tau_cli <table> <in.bin> <out.bin> <prop_at_Lev>
// in.bin: double plevel[21], T[21], vmr[9][21] (for prop_at_Lev == 0 only the first 20 of each are used);
// out.bin: int32 nwvl, then tau[nwvl][20], wvl[nwvl], weight[nwvl].
// oracle/Makefile links it against the reference's repwvl_thermal.cpp (+ netcdf shim) and against librcm_b200.so.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "repwvl_thermal.h"

int main(int argc, char** argv) {
    if (argc != 5) return 2;
    FILE* f = std::fopen(argv[2], "rb");
    if (!f) return 3;
    std::vector<double> plevel(21), T(21);
    double vmr[9][21];
    if (std::fread(plevel.data(), 8, 21, f) != 21 || std::fread(T.data(), 8, 21, f) != 21 ||
        std::fread(vmr, 8, 9 * 21, f) != 9 * 21)
        return 4;
    std::fclose(f);
    double **tau = nullptr, *wvl = nullptr, *weight = nullptr;
    int nwvl = 0;
    read_tau(argv[1], 21, plevel, T, vmr[0], vmr[1], vmr[2], vmr[3], vmr[4], vmr[5], vmr[6], vmr[7], vmr[8], &tau, &wvl,
             &weight, &nwvl, std::atoi(argv[4]));
    if (nwvl <= 0) return 5;
    f = std::fopen(argv[3], "wb");
    if (!f) return 6;
    std::fwrite(&nwvl, 4, 1, f);
    for (int i = 0; i < nwvl; ++i) std::fwrite(tau[i], 8, 20, f);
    std::fwrite(wvl, 8, nwvl, f);
    std::fwrite(weight, 8, nwvl, f);
    std::fclose(f);
    return 0;
}

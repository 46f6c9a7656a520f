// pfx_dct.hpp -- cosine-transform form of the forward spectra and of the admissibility transform (pfx_dct.cu).
#pragma once
#include <utility>
#include <vector>

#include "pfx_plan.hpp"

namespace pfx {

struct DctState {
    double *v = nullptr;    // [2][nx][ny]  re-ordered grids (forward) / complex-to-real outputs (prepare)
    double2 *V = nullptr;   // [2][nx][ny/2+1]  half spectra of the re-ordered grids / Q planes
    double *D = nullptr;    // [2][nx][ny]  D[a][b] = sum x 2cos 2cos: the spectra pass 1 reads
    double2 *phx = nullptr, *phy = nullptr;  // [NX], [NY]  e^{i pi a/N} at slot a mod N (signed a)
    double2 *hx = nullptr, *hy = nullptr;    // [nx+1], [ny+1]  e^{-i pi a/(2n)}
    cufftHandle fwd = 0, inv = 0;
    bool fwd_ok = false, inv_ok = false;
};

// true when the padded array is exactly the even reflection (NX = 2 nx, NY = 2 ny) and PLATEFLEX_B200_DCT != "off"
bool dct_applicable(const pfx_plan *pl);
int dct_init(pfx_plan *pl, DctState **out, std::vector<std::pair<void *, size_t>> &owned);
void dct_free(DctState *st);
// D of grid `slot` (device pointer, C order) -- the counterpart of forward_spectrum
int dct_forward(pfx_plan *pl, DctState *st, const double *grid_dev, int slot);
// zc[j][i] = (Z_0, Z_1)(i, j) -- the counterpart of zmul + inverse FFT + zcrop
int dct_prepare(pfx_plan *pl, DctState *st, int ngrids, double2 *zc);

}  // namespace pfx

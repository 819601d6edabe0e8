// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
// Array-based entry point onto the UNMODIFIED reference
// computeOffDiagonalTerms (/root/reference/bem/icqLaplaceMatrices.cpp:37):
// wraps (xyz, tri) into the stand-in vtkPolyData and forwards.
#include <icqLaplaceMatrices.h>

extern "C" void icqref_offdiag(const double* xyz, const long long* tri,
                               long long ntri, double* gMat) {
    vtkPolyData pd(xyz, tri, ntri);
    computeOffDiagonalTerms(&pd, gMat);
}

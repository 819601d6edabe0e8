// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
// Minimal stand-in for vtkPolyData: GetPolys()/GetPoints() only
// (/root/reference/bem/icqLaplaceMatrices.cpp:54-55).
#ifndef ICQ_SHIM_VTKPOLYDATA_H
#define ICQ_SHIM_VTKPOLYDATA_H
#include <vtkCellArray.h>
#include <vtkPoints.h>
class vtkPolyData {
public:
    vtkPolyData(const double* xyz, const long long* tri, long long ntri)
        : points_(xyz), polys_(tri, ntri) {}
    vtkCellArray* GetPolys() { return &polys_; }
    vtkPoints* GetPoints() { return &points_; }
private:
    vtkPoints points_;
    vtkCellArray polys_;
};
#endif

// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
// Minimal stand-in for vtkCellArray (GetNumberOfCells / InitTraversal /
// GetNextCell as used at /root/reference/bem/icqLaplaceMatrices.cpp:54-70).
#ifndef ICQ_SHIM_VTKCELLARRAY_H
#define ICQ_SHIM_VTKCELLARRAY_H
#include <vtkIdList.h>
class vtkCellArray {
public:
    vtkCellArray(const long long* tri, long long n) : tri_(tri), n_(n), cur_(0) {}
    long long GetNumberOfCells() const { return n_; }
    void InitTraversal() { cur_ = 0; }
    int GetNextCell(vtkIdList* ids) {
        if (cur_ >= n_) return 0;
        ids->Set3(tri_[3 * cur_], tri_[3 * cur_ + 1], tri_[3 * cur_ + 2]);
        ++cur_;
        return 1;
    }
private:
    const long long* tri_;
    long long n_, cur_;
};
#endif

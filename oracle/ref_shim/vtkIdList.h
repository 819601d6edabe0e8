// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
// Minimal stand-in for VTK's vtkIdList: only the members that
// /root/reference/bem/icqLaplaceMatrices.cpp:60-71 touches.
#ifndef ICQ_SHIM_VTKIDLIST_H
#define ICQ_SHIM_VTKIDLIST_H
#include <vector>
class vtkIdList {
public:
    static vtkIdList* New() { return new vtkIdList(); }
    void Delete() { delete this; }
    long long GetId(int i) const { return ids_[i]; }
    void Set3(long long a, long long b, long long c) { ids_.assign({a, b, c}); }
private:
    std::vector<long long> ids_;
};
#endif

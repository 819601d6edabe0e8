// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
// Minimal stand-in for vtkPoints (GetPoint(id, double*) as used at
// /root/reference/bem/icqLaplaceMatrices.cpp:81-83,89-91).
#ifndef ICQ_SHIM_VTKPOINTS_H
#define ICQ_SHIM_VTKPOINTS_H
class vtkPoints {
public:
    explicit vtkPoints(const double* xyz) : xyz_(xyz) {}
    void GetPoint(long long id, double* out) const {
        out[0] = xyz_[3 * id + 0];
        out[1] = xyz_[3 * id + 1];
        out[2] = xyz_[3 * id + 2];
    }
private:
    const double* xyz_;
};
#endif

// Stand-in for Thirdparty/g2o/g2o/types/sim3.h (g2o::Sim3 on Eigen) with the accessors the drop-in Optimizer uses:
// rotation() -> quaternion with x() y() z() w(), translation() -> indexable 3-vector, scale().  Test scaffolding only; a real
// build includes the reference's own header (Eigen present).
#pragma once
namespace g2o
{
struct MockQuat { double c[4] = {0, 0, 0, 1}; double &x() { return c[0]; } double &y() { return c[1]; } double &z() { return c[2]; } double &w() { return c[3]; }
                  double x() const { return c[0]; } double y() const { return c[1]; } double z() const { return c[2]; } double w() const { return c[3]; } };
struct MockVec3 { double v[3] = {0, 0, 0}; double &operator[](int i) { return v[i]; } const double &operator[](int i) const { return v[i]; } };
struct Sim3 {
    MockQuat r; MockVec3 t; double s = 1;
    MockQuat &rotation() { return r; }
    MockVec3 &translation() { return t; }
    double &scale() { return s; }
    const MockQuat &rotation() const { return r; }
    const MockVec3 &translation() const { return t; }
    const double &scale() const { return s; }
};
}  // namespace g2o

// 3-D point of the reconstruction (reference: viso/point3d.hh:5-9; single precision on purpose, the reference stores and
// updates its points in float).
#ifndef VISOB_POINT3D_H
#define VISOB_POINT3D_H
struct Point3d {
  float x, y, z;
  Point3d() {}
  Point3d(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
#endif

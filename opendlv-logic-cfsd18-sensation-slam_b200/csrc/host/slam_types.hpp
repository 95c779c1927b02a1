// slam_types.hpp -- the few value types the Slam/Cone interface needs.
//
// Inside the reference tree (compile with -DSLAM_B200_WITH_EIGEN and the reference's include path)
// these are the reference's own Eigen types, so the class below is a drop-in for src/slam.hpp.
// Stand-alone (this repository's tests; Eigen is not shipped here) they are minimal stand-ins with
// the same access syntax: v(i), m(r,c), m.cols(), m.data() (column-major, like Eigen::MatrixXd).
#pragma once
#include <cstddef>
#include <vector>

#ifdef SLAM_B200_WITH_EIGEN
#include <Eigen/Dense>
namespace slamtypes {
typedef Eigen::Vector3d Vector3d;
typedef Eigen::MatrixXd MatrixXd;
}
#else
namespace slamtypes {
struct Vector3d {
  double v[3];
  Vector3d() : v{0, 0, 0} {}
  Vector3d(double a, double b, double c) : v{a, b, c} {}
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  const double* data() const { return v; }
};
struct MatrixXd {  // column-major, dynamic
  int r = 0, c = 0;
  std::vector<double> d;
  MatrixXd() {}
  MatrixXd(int rows, int cols) : r(rows), c(cols), d((size_t)rows * cols, 0.0) {}
  double& operator()(int i, int j) { return d[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return d[(size_t)j * r + i]; }
  long cols() const { return c; }
  long rows() const { return r; }
  const double* data() const { return d.data(); }
  double* data() { return d.data(); }
};
}
#endif

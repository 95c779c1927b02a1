// wgs84.cpp -- see wgs84.hpp.  Same arithmetic, operation for operation, as the reference header
// (the test compares bit for bit on the build host), laid out as a small projector object whose
// series coefficients are computed once per reference point.
#include "wgs84.hpp"

#include <cmath>
#include <limits>

namespace slamwgs84 {
namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kDegToRad = kPi / 180.0;
constexpr double kHalfPi = kPi / 2.0;
constexpr double kEquatorRadius = 6378137.0;                   // WGS84 semi-major axis
constexpr double kFlattening = 1.0 / 298.257223563;
constexpr double kE2 = 2.0 * kFlattening - kFlattening * kFlattening;  // squared first eccentricity

// meridional arc length (unit ellipsoid) as a series in sin^2(phi): WGS84toCartesian.hpp:52-79
struct MeridianArc {
  double r0, r1, r2, r3, r4;
  constexpr MeridianArc()
      : r0(1.0 - kE2 * (0.25 + kE2 * (0.046875 + kE2 * (0.01953125 + kE2 * 0.01068115234375)))),
        r1(kE2 * (0.75 - kE2 * (0.046875 + kE2 * (0.01953125 + kE2 * 0.01068115234375)))),
        r2((kE2 * kE2) * (0.46875 - kE2 * (0.01302083333333333333 + kE2 * 0.00712076822916666666))),
        r3(((kE2 * kE2) * kE2) * (0.36458333333333333333 - kE2 * 0.00569661458333333333)),
        r4(((kE2 * kE2) * kE2) * kE2 * 0.3076171875) {}
  double operator()(double phi) const {
    const double sp = std::sin(phi);
    const double cs = std::cos(phi) * sp;
    const double s2 = sp * sp;
    return r0 * phi - cs * (r1 + s2 * (r2 + s2 * (r3 + s2 * r4)));
  }
};

struct Projector {
  MeridianArc arc;
  double ml0;      // arc length of the reference latitude
  double lon0;     // reference longitude, radians
  explicit Projector(const double ref[2]) : arc(), ml0(arc(ref[0] * kDegToRad)), lon0(ref[1] * kDegToRad) {}

  // unit-ellipsoid projection of (phi, dlon): WGS84toCartesian.hpp:85-94
  void unit(double phi, double dlon, double out[2]) const {
    out[0] = dlon;
    out[1] = -1.0 * ml0;
    if (std::abs(phi) < 1.0e-10) return;
    const double sp = std::sin(phi);
    double ms = 0.0;
    if (std::abs(sp) > 1.0e-10) ms = (std::cos(phi) / std::sqrt(1.0 - kE2 * sp * sp)) / sp;
    dlon *= sp;
    out[0] = ms * std::sin(dlon);
    out[1] = (arc(phi) - ml0) + ms * (1.0 - std::cos(dlon));
  }

  // WGS84toCartesian.hpp:96-109
  void forward(double phi, double lam, double out[2]) const {
    const double beyond = std::abs(phi) - kHalfPi;
    if (beyond > 1.0e-12 || std::abs(lam) > 10.0) { out[0] = 0.0; out[1] = 0.0; return; }
    if (std::abs(beyond) < 1.0e-12) phi = (phi < 0.0) ? -1.0 * kHalfPi : kHalfPi;
    double u[2];
    unit(phi, lam - lon0, u);
    out[0] = kEquatorRadius * u[0];
    out[1] = kEquatorRadius * u[1];
  }
};

}  // namespace

void toCartesian(const double ref[2], const double pos[2], double out[2]) {
  const Projector p(ref);
  p.forward(pos[0] * kDegToRad, pos[1] * kDegToRad, out);
}

void fromCartesian(const double ref[2], const double xy[2], double out[2]) {
  const Projector p(ref);
  const double step = 1e-5;  // degrees
  const double tol = 1.0e-2; // metres
  double guess[2] = {ref[0], ref[1]};
  double c[2];
  p.forward(guess[0] * kDegToRad, guess[1] * kDegToRad, c);
  // axis 1 (north, latitude) first, then axis 0 (east, longitude): WGS84toCartesian.hpp:135-152
  for (int pass = 0; pass < 2; pass++) {
    const int cart = pass == 0 ? 1 : 0;  // Cartesian component driven in this pass
    const int geo = pass == 0 ? 0 : 1;   // geodetic component stepped in this pass
    const int sign = xy[cart] < 0 ? -1 : 1;
    double prev = std::numeric_limits<double>::max();
    double d = std::abs(xy[cart] - c[cart]);
    while (d < prev && d > tol) {
      guess[geo] = guess[geo] + sign * step;
      p.forward(guess[0] * kDegToRad, guess[1] * kDegToRad, c);
      prev = d;
      d = std::abs(xy[cart] - c[cart]);
    }
  }
  out[0] = guess[0];
  out[1] = guess[1];
}

double headingFromNorth(float northHeading) {
  const double PI_REF = 3.14159265f;  // slam.hpp:134: a float literal widened to double
  double heading = northHeading;
  heading = heading - PI_REF;
  heading = (heading > PI_REF) ? (heading - 2 * PI_REF) : heading;
  heading = (heading < -PI_REF) ? (heading + 2 * PI_REF) : heading;
  return heading;
}

}  // namespace slamwgs84

// cone.cpp -- see cone.hpp.  Behaviour follows src/cone.cpp:22-85 of the reference.
#include "cone.hpp"

#include <cmath>

namespace {
const double kRad2Deg = 57.295779513082325;  // cone.hpp:55
}

Cone::Cone(double x, double y, int type, int id) : m_x(x), m_y(y), m_type(type), m_id(id) {}

// bearing of the cone seen from `pose`, degrees, minus the heading scaled by 1/RAD2DEG
// (src/cone.cpp:34-44 -- the reference subtracts heading*(1/RAD2DEG) from a value in degrees)
ConeDirection Cone::getDirection(slamtypes::Vector3d pose) {
  const double dx = m_x - pose(0), dy = m_y - pose(1);
  const double heading = pose(2) * (1 / kRad2Deg);
  double az = std::atan2(dy, dx) * kRad2Deg;
  az -= heading;
  ConeDirection d;
  d.zenithAngle = 0;
  d.azimuthAngle = static_cast<float>(az);
  return d;
}

// range of the cone seen from `pose` (src/cone.cpp:46-53)
ConeDistance Cone::getDistance(slamtypes::Vector3d pose) {
  const double dx = m_x - pose(0), dy = m_y - pose(1);
  ConeDistance d;
  d.distance = static_cast<float>(std::sqrt(dx * dx + dy * dy));
  return d;
}

double Cone::getX() { return m_x; }
double Cone::getY() { return m_y; }
int Cone::getType() { return m_type; }
int Cone::getId() { return m_id; }
void Cone::setX(double x) { m_x = x; }
void Cone::setY(double y) { m_y = y; }
void Cone::setType(int type) { m_type = type; }
void Cone::setId(int id) { m_id = id; }

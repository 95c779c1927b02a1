// cone.hpp -- map element of the cone map; mirrors the reference's class Cone (src/cone.hpp:29-57)
// member for member so callers (Slam, the viewer's Drawer) compile unchanged.  getDirection /
// getDistance build the outgoing ObjectDirection / ObjectDistance payloads (src/cone.cpp:34-53);
// stand-alone they return plain structs carrying the same float fields.
#pragma once
#include "slam_types.hpp"

struct ConeDirection { float azimuthAngle; float zenithAngle; };  // opendlv.logic.perception.ObjectDirection
struct ConeDistance { float distance; };                           // opendlv.logic.perception.ObjectDistance

class Cone {
 public:
  Cone(double x, double y, int type, int id);
  ~Cone() = default;

  ConeDirection getDirection(slamtypes::Vector3d pose);
  ConeDistance getDistance(slamtypes::Vector3d pose);

  double getX();
  double getY();
  int getType();
  int getId();

  void setX(double x);
  void setY(double y);
  void setType(int type);
  void setId(int id);

 private:
  double m_x;
  double m_y;
  int m_type;
  int m_id;
};

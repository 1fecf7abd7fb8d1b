// Stand-in for the `state` carrier (reference faster/include/faster_types.hpp:79-165: pos/vel/accel/jerk + yaw, with the
// setters the planner uses), so that solverGurobi.hpp builds and is tested without the reference tree.  Inside the
// reference tree solverGurobi.hpp finds the reference's own faster_types.hpp first and this file is not used.
#pragma once
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#else
#include "fq_compat.hpp"
#endif
#include <iostream>

struct state
{
  Eigen::Vector3d pos = Eigen::Vector3d::Zero();
  Eigen::Vector3d vel = Eigen::Vector3d::Zero();
  Eigen::Vector3d accel = Eigen::Vector3d::Zero();
  Eigen::Vector3d jerk = Eigen::Vector3d::Zero();
  double yaw = 0;
  double dyaw = 0;

  void setPos(const double x, const double y, const double z) { pos = Eigen::Vector3d(x, y, z); }
  void setVel(const double x, const double y, const double z) { vel = Eigen::Vector3d(x, y, z); }
  void setAccel(const double x, const double y, const double z) { accel = Eigen::Vector3d(x, y, z); }
  void setJerk(const double x, const double y, const double z) { jerk = Eigen::Vector3d(x, y, z); }
  void setPos(const Eigen::Vector3d& d) { pos = d; }
  void setVel(const Eigen::Vector3d& d) { vel = d; }
  void setAccel(const Eigen::Vector3d& d) { accel = d; }
  void setJerk(const Eigen::Vector3d& d) { jerk = d; }
  void setYaw(const double& d) { yaw = d; }
  void setZero()
  {
    pos = vel = accel = jerk = Eigen::Vector3d::Zero();
    yaw = 0;
    dyaw = 0;
  }
  void printPos() { std::cout << "Pos= " << pos.x() << " " << pos.y() << " " << pos.z() << std::endl; }
};

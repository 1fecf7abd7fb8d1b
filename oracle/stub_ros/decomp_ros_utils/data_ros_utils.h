// TEST INFRASTRUCTURE ONLY: faster/include/solverGurobi.hpp includes the ROS message helpers of DecompROS; what the solver uses
// from them are DecompUtil's own types (LinearConstraint3D, vec_Vecf), taken from the reference's DecompUtil headers.
#pragma once
#include <decomp_geometry/polyhedron.h>

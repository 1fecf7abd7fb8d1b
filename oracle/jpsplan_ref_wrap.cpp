// TEST INFRASTRUCTURE ONLY -- C wrapper around the REFERENCE's own JPS planner layer: JPSPlanner<3>::plan
// (thirdparty/jps3d/src/jps_planner/jps_planner.cpp:196-295: cells -> world, removeLinePts :83-105, removeCornerPts :36-80
// forwards and backwards) over JPS::MapUtil<3> (jps_collision/map_util.h: floatToInt/intToFloat :334-347, rayTrace/isBlocked
// :349-383) and the graph search, all compiled unmodified from /root/reference by oracle/Makefile (oracle/stub_eigen,
// stub_boost, stub_ros, stub_pcl stand in for the libraries this image lacks).  Output: oracle/_ref/libjpsplan_ref.so.
#include <jps_planner/jps_planner/jps_planner.h>

#include <cstring>

namespace
{
std::shared_ptr<JPS::MapUtil<3>> make_map(const char* map, int xd, int yd, int zd, const double* origin, double res)
{
  auto mu = std::make_shared<JPS::MapUtil<3>>();
  JPS::Tmap m(map, map + (size_t)xd * yd * zd);
  mu->setMap(Vec3f(origin[0], origin[1], origin[2]), Vec3i(xd, yd, zd), m, res);
  return mu;
}
int put(const vec_Vecf<3>& p, double* out, int cap)
{
  const int n = (int)p.size();
  for (int i = 0; i < n && i < cap; i++) { out[3 * i] = p[i](0); out[3 * i + 1] = p[i](1); out[3 * i + 2] = p[i](2); }
  return n;
}
}  // namespace

extern "C" {
// map: x fastest, then y, then z; 0 free, 100 occupied, -1 unknown.  Returns the number of points of the simplified path
// (0: no path; status_out = the planner's status), path_out: up to cap xyz triples start -> goal, raw_out likewise.
int jpsplanref_plan(const char* map, int xd, int yd, int zd, const double* origin, double res, const double* start, const double* goal,
                    int use_jps, double* path_out, int cap, double* raw_out, int* n_raw, int* status_out)
{
  auto mu = make_map(map, xd, yd, zd, origin, res);
  JPSPlanner3D planner(false);
  planner.setMapUtil(mu);
  planner.updateMap();
  const bool ok = planner.plan(Vec3f(start[0], start[1], start[2]), Vec3f(goal[0], goal[1], goal[2]), 1, use_jps != 0);
  if (status_out) *status_out = planner.status();
  if (!ok) { if (n_raw) *n_raw = 0; return 0; }
  if (n_raw) *n_raw = put(planner.getRawPath(), raw_out, cap);
  return put(planner.getPath(), path_out, cap);
}
// the post-processing alone on a caller-supplied raw path (world coordinates, start -> goal): jps_planner.cpp:289-293
int jpsplanref_simplify(const char* map, int xd, int yd, int zd, const double* origin, double res, const double* raw, int n_raw,
                        double* path_out, int cap)
{
  auto mu = make_map(map, xd, yd, zd, origin, res);
  JPSPlanner3D planner(false);
  planner.setMapUtil(mu);
  vec_Vecf<3> p;
  for (int i = 0; i < n_raw; i++) p.push_back(Vec3f(raw[3 * i], raw[3 * i + 1], raw[3 * i + 2]));
  p = planner.removeLinePts(p);
  p = planner.removeCornerPts(p);
  std::reverse(std::begin(p), std::end(p));
  p = planner.removeCornerPts(p);
  std::reverse(std::begin(p), std::end(p));
  return put(p, path_out, cap);
}
// MapUtil's ray-traced line of sight (map_util.h:371-383)
int jpsplanref_blocked(const char* map, int xd, int yd, int zd, const double* origin, double res, const double* p1, const double* p2)
{
  auto mu = make_map(map, xd, yd, zd, origin, res);
  return mu->isBlocked(Vec3f(p1[0], p1[1], p1[2]), Vec3f(p2[0], p2[1], p2[2])) ? 1 : 0;
}
}

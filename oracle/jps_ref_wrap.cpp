// TEST INFRASTRUCTURE ONLY -- C wrapper around the REFERENCE's own JPS3D graph search
// (thirdparty/jps3d/src/jps_planner/graph_search.cpp, compiled from where it lies under /root/reference by oracle/Makefile
// with oracle/stub_boost standing in for boost::heap).  Output goes to oracle/_ref/ (git-ignored, travels to the GPU box).
#include <jps_planner/jps_planner/graph_search.h>

#include <cmath>
#include <cstring>

extern "C" {
// map: x fastest, then y, then z (coordToId, graph_search.cpp:46-48); 0 free, >0 occupied, <0 unknown.
// path_out: up to cap (x,y,z) triples from START to GOAL (the reference returns goal->start; reversed here).
// returns the number of path points (0 = no path), *cost = sum of Euclidean step lengths in cells.
int jpsref_plan(const char* map, int xd, int yd, int zd, int xs, int ys, int zs, int xg, int yg, int zg, int use_jps,
                int max_expand, int* path_out, int cap, double* cost, int* n_closed)
{
  JPS::GraphSearch gs(map, xd, yd, zd, 1.0, false);
  const bool ok = gs.plan(xs, ys, zs, xg, yg, zg, use_jps != 0, max_expand);
  if (n_closed) *n_closed = (int)gs.getCloseSet().size();
  if (!ok) { if (cost) *cost = INFINITY; return 0; }
  const auto path = gs.getPath();
  const int n = (int)path.size();
  double c = 0;
  for (int i = 0; i < n; i++)
  {
    const auto& s = path[n - 1 - i];
    if (i < cap) { path_out[3 * i] = s->x; path_out[3 * i + 1] = s->y; path_out[3 * i + 2] = s->z; }
    if (i > 0)
    {
      const auto& p = path[n - i];
      c += std::sqrt((double)((s->x - p->x) * (s->x - p->x) + (s->y - p->y) * (s->y - p->y) + (s->z - p->z) * (s->z - p->z)));
    }
  }
  if (cost) *cost = c;
  return n;
}

// the reference's neighbour tables (JPS3DNeib, graph_search.h:104-136): ns[27][3][26], f1[27][3][12], f2[27][3][12]
void jpsref_tables(int* ns, int* f1, int* f2)
{
  JPS::JPS3DNeib t;
  std::memcpy(ns, t.ns, sizeof(t.ns));
  std::memcpy(f1, t.f1, sizeof(t.f1));
  std::memcpy(f2, t.f2, sizeof(t.f2));
}
}

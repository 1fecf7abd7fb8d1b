// TEST INFRASTRUCTURE ONLY -- C wrapper around the REFERENCE's own convex decomposition: DecompUtil's EllipsoidDecomp3D
// (thirdparty/DecompROS/DecompUtil/include/decomp_util/ellipsoid_decomp.h and the headers it pulls in), compiled unmodified
// from where it lies under /root/reference by oracle/Makefile, with oracle/stub_eigen standing in for Eigen (absent from this
// image).  The call sequence is JPS_Manager::cvxEllipsoidDecomp's (faster/src/jps_manager.cpp:80-127).  Output goes to
// oracle/_ref/ (git-ignored, travels to the GPU box).
#include <decomp_util/ellipsoid_decomp.h>

extern "C" {
// path: n_pts x 3 (n_pts - 1 segments); obs: n_obs x 3; bbox: the local bounding box (jps_manager.cpp:100 passes (2, 2, 1));
// inflate = drone radius (:102); z_ground: the extra face appended last (:118-122).
// face_ofs: n_pts entries; Ab: rows [Ax Ay Az b] of A x <= b.  Returns the number of rows, -1 if cap_rows is too small.
int decompref_cvx(const double* path, int n_pts, const double* obs, int n_obs, const double* bbox, double inflate, double z_ground,
                  int* face_ofs, double* Ab, int cap_rows)
{
  vec_Vecf<3> p, o;
  for (int i = 0; i < n_pts; i++) p.push_back(Vec3f(path[3 * i], path[3 * i + 1], path[3 * i + 2]));
  for (int i = 0; i < n_obs; i++) o.push_back(Vec3f(obs[3 * i], obs[3 * i + 1], obs[3 * i + 2]));
  EllipsoidDecomp3D util;
  util.set_obs(o);                                             // :92-98
  util.set_local_bbox(Vec3f(bbox[0], bbox[1], bbox[2]));       // :100
  util.set_inflate_distance(inflate);                          // :102
  util.dilate(p);                                              // :103
  const auto polys = util.get_polyhedrons();                   // :109
  int rows = 0;
  face_ofs[0] = 0;
  for (size_t i = 0; i + 1 < p.size(); i++)
  {
    const Vec3f pt_inside = (p[i] + p[i + 1]) / 2;             // :115
    LinearConstraint3D cs(pt_inside, polys[i].hyperplanes());  // :116
    const auto A = cs.A();
    const auto b = cs.b();
    if (rows + A.rows() + 1 > cap_rows) return -1;
    for (int f = 0; f < A.rows(); f++)
    {
      double* r = Ab + 4 * (size_t)rows++;
      r[0] = A(f, 0); r[1] = A(f, 1); r[2] = A(f, 2); r[3] = b(f);
    }
    double* g = Ab + 4 * (size_t)rows++;                       // "above the ground" (:118-122)
    g[0] = 0; g[1] = 0; g[2] = -1; g[3] = -z_ground;
    face_ofs[i + 1] = rows;
  }
  return rows;
}
}

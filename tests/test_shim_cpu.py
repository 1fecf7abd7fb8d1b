"""CPU-only checks of the drop-in class surface (include/solverGurobi.hpp): it compiles against an Eigen / DecompUtil
API mock with Eigen's column-major storage (polyhedron.h:114-185, data_type.h:50-80), packs polytope rows correctly from
it, keeps setDistances' signature (solverGurobi.hpp:100), and the abort flag behaves like the reference's
(solverGurobi.cpp:30-39,:445,:474): StopExecution() before genNewTraj() => false, trials_ == 0, flag reset."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, built_lib, extra_includes, defines=()):
    exe = str(tmp_path / "mock_compile")
    libdir = os.path.dirname(built_lib)
    cmd = ["g++", "-std=c++17", "-Wall", "-Werror"] + ["-D" + d for d in defines] + sum((["-I", i] for i in extra_includes), []) + \
          ["-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mock_compile.cpp"), "-o", exe, "-L", libdir,
           "-lfaster_b200", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return exe


def test_header_against_eigen_api_mock(tmp_path, built_lib):
    exe = _build(tmp_path, built_lib, [os.path.join(ROOT, "tests", "cpp", "mock_eigen")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


def test_header_with_eigen_but_without_decomputil(tmp_path, built_lib):
    """<Eigen/Dense> on the include path, DecompUtil not: the look-alike LinearConstraint3D must still be defined
    (ADVICE round 1).  The mock's Eigen directory alone plays that role."""
    inc = tmp_path / "only_eigen"
    (inc / "Eigen").mkdir(parents=True)
    for f in ("Dense", "StdVector"):
        (inc / "Eigen" / f).write_text(open(os.path.join(ROOT, "tests", "cpp", "mock_eigen", "Eigen", f)).read())
    src = tmp_path / "t.cpp"
    src.write_text('#include "solverGurobi.hpp"\nint main(){ LinearConstraint3D c; SolverGurobi s; std::vector<LinearConstraint3D> v(1, c);'
                   ' s.setPolytopes(v); vec_Vecf<3> q; s.setDistances(q, std::vector<double>()); return 0; }\n')
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I", str(inc), "-I", os.path.join(ROOT, "include"), str(src), "-o",
                           str(tmp_path / "t"), "-L", libdir, "-lfaster_b200", "-Wl,-rpath," + libdir])
    assert subprocess.run([str(tmp_path / "t")], timeout=60).returncode == 0


def test_header_against_the_reference_own_headers(tmp_path, built_lib):
    """The drop-in header picks up the reference's own `state` (faster/include/faster_types.hpp) and DecompUtil's own
    LinearConstraint3D / vec_Vecf (decomp_geometry/polyhedron.h, decomp_basis/data_type.h) when they are on the include path
    -- as they are inside the reference tree -- and the same program as above behaves the same.  Eigen itself is not in this
    image: oracle/stub_eigen supplies the arithmetic; the types and their member functions are the reference's."""
    import pytest
    ref_inc = "/root/reference/faster/include"
    dec_inc = "/root/reference/thirdparty/DecompROS/DecompUtil/include"
    if not (os.path.exists(os.path.join(ref_inc, "faster_types.hpp")) and os.path.exists(os.path.join(dec_inc, "decomp_geometry", "polyhedron.h"))):
        pytest.skip("needs /root/reference")
    only_types = tmp_path / "ref_types"                       # faster_types.hpp alone: the directory also holds the reference's solverGurobi.hpp
    only_types.mkdir()
    os.symlink(os.path.join(ref_inc, "faster_types.hpp"), only_types / "faster_types.hpp")
    exe = _build(tmp_path, built_lib, [os.path.join(ROOT, "oracle", "stub_eigen"), dec_inc, str(only_types)], ["FQ_EXPECT_REFERENCE_TYPES"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr

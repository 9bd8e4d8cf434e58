"""Type-checks the shipped gcransac::ScoringFunction adapter (include/pxb200_scoring_adapter.h)
against the REFERENCE'S OWN declaration of the interface: the `Score` struct and the abstract `ScoringFunction` class are
extracted verbatim from graph-cut-ransac/src/pygcransac/include/scoring_function.h into a scratch translation unit (never
into the repository), the stand-in headers of oracle/shim/ supply cv::Mat / Eigen::MatrixXd, and g++ compiles an
instantiation -- the `override` specifiers fail the build if a signature drifts from the reference's virtuals.
Needs /root/reference (absent on the GPU box: skipped there)."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/graph-cut-ransac/src/pygcransac/include/scoring_function.h")


def test_scoring_adapter_overrides_the_reference_interface(tmp_path):
    if not REF.exists():
        pytest.skip("needs /root/reference")
    text = REF.read_text()
    start = text.index("namespace gcransac")
    end = text.index("class MSACScoringFunction")
    end = text.rindex("template<class _ModelEstimator>", start, end)
    extract = text[start:end] + "\n}\n"  # Score + ScoringFunction, then close the namespace
    assert "virtual OLGA_INLINE Score getScore" in extract and "virtual void initialize" in extract
    tu = tmp_path / "adapter_check.cpp"
    tu.write_text(f'''
#include <cstddef>
#include <vector>
#include "mini_eigen.h"
#include "mini_cv.h"
#define OLGA_INLINE inline
namespace Eigen {{ template <class T, int R, int C> inline const T *data_of(const Matrix<T, R, C> &m) {{ return m.data_.data(); }} }}
namespace gcransac {{ struct Model {{ Eigen::MatrixXd descriptor; }}; }}            // gcr/model.h: descriptor only
namespace progx {{ template <class E> struct Model : gcransac::Model {{}}; }}      // px/include/progx_model.h
// ---- verbatim from the reference's scoring_function.h ----
{extract}
// ----------------------------------------------------------
struct VectorWithData : Eigen::VectorXd {{ const double *data() const {{ return data_.data(); }} }};
#define VectorXd VectorXdShim
namespace Eigen {{ typedef ::VectorWithData VectorXdShim; }}
#include "pxb200_scoring_adapter.h"
#undef VectorXd
struct DummyEstimator {{}};
int main() {{
    GpuScoringWithCompoundModel<DummyEstimator> scoring(PXB_MODEL_HOMOGRAPHY);
    gcransac::ScoringFunction<DummyEstimator> *base = &scoring;                     // usable through the reference's seam
    base->initialize(9.0, 100);
    double pts[400] = {{0}};
    cv::Mat points(100, 4, CV_64F, pts);
    gcransac::Model model;
    model.descriptor.resize(3, 3);
    std::vector<size_t> inliers;
    DummyEstimator est;
    gcransac::Score s = base->getScore(points, model, est, 2.0, inliers);
    return (int)s.inlier_number;
}}
''')
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror=overloaded-virtual", f"-I{ROOT / 'include'}",
           f"-I{ROOT / 'oracle' / 'shim'}", str(tu)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]

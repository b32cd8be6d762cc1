// ORACLE/_ref — TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE'S OWN base/Math.h (FastAtan2, Square), compiled from the file where it
// lies under /root/reference (never copied into this repository): the one source file of the hot path that needs nothing but the standard library.
// Everything else on the path pulls in Eigen / Ceres / PCL / OpenCV / Boost headers: those files are compiled against the stand-ins of oracle/shim (ref_path_wrap.cpp,
// ref_assoc_wrap.cpp, ref_camlidar_wrap.cpp).
// Built by `make -C oracle ref` into oracle/_ref/libpvo_ref.so with the reference's own flags (CMakeLists.txt:4-13: -std=gnu++14 -fopenmp, Release =
// -O3 -DNDEBUG, no -march: plain x86-64, so no FMA contraction).  Used to pin the oracle's restatement of FastAtan2 bit for bit
// (tests/test_oracle_pinning.py) and to generate tests/golden/ref_fast_atan2.npz (tests/make_golden.py).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include REF_MATH_H

extern "C" {
void ref_fast_atan2_f(long n, const float* y, const float* x, float* out) { for (long i = 0; i < n; ++i) out[i] = FastAtan2<float>(y[i], x[i]); }
void ref_fast_atan2_d(long n, const double* y, const double* x, double* out) { for (long i = 0; i < n; ++i) out[i] = FastAtan2<double>(y[i], x[i]); }
void ref_square_d(long n, const double* a, double* out) { for (long i = 0; i < n; ++i) out[i] = Square<double>(a[i]); }
}

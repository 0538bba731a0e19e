// TEST INFRASTRUCTURE — not product code.
//
// C-callable shim around the UNMODIFIED mLib vector/triangle headers of the reference tree
// (/root/reference/external/mLib/include/core-math/vec3.h, core-graphics/triangle.h), compiled where they lie by
// oracle/Makefile into oracle/_ref/libref_mlib.so.  The headers are normally pulled in through mLibCore.h, which needs the
// whole (Windows-oriented) library; the few names they expect from it are declared below (empty serialisation base,
// integer typedefs, three math helpers that the functions under test never call).
//
//   ref_tri_area         Trianglef::getArea                      triangle.h:23-35   (used by Segmentation.h:138-140)
//   ref_vertex_normals   the loop of MeshData::computeVertexNormals (meshData.h:758-782) written with the real
//                        vec3f operators (^, +=, normalize) — MeshData itself drags in the rest of mLib.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <string>
#include <vector>
#define MLIB_ASSERT(x)
typedef uint64_t UINT64; typedef int64_t INT64; typedef unsigned int UINT; typedef unsigned char UCHAR; typedef unsigned short USHORT;
namespace ml {
template <class T> class BinaryDataSerialize {};
namespace math {
template <class T> T randomUniform(T a, T) { return a; }
template <class T> T radiansToDegrees(T x) { return x; }
template <class T> T clamp(T x, T a, T b) { return x < a ? a : (x > b ? b : x); }
}  // namespace math
}  // namespace ml
#include "core-math/vec3.h"
#include "core-graphics/triangle.h"

extern "C" {

float ref_tri_area(const float* a, const float* b, const float* c) {
  ml::Trianglef t(ml::vec3f(a[0], a[1], a[2]), ml::vec3f(b[0], b[1], b[2]), ml::vec3f(c[0], c[1], c[2]));
  return t.getArea();
}

void ref_vertex_normals(const float* xyz, int64_t nV, const uint32_t* tri, int64_t nF, float* out) {
  std::vector<ml::vec3f> V((size_t)nV), N((size_t)nV, ml::vec3f(0, 0, 0));
  for (int64_t i = 0; i < nV; ++i) V[i] = ml::vec3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  for (int64_t f = 0; f < nF; ++f) {
    const uint32_t* face = tri + 3 * f;
    ml::vec3f n(0, 0, 0);
    n += (V[face[1]] - V[face[0]]) ^ (V[face[2]] - V[face[0]]);
    n.normalize();
    for (int k = 0; k < 3; ++k) N[face[k]] += n;
  }
  for (auto& n : N) n.normalize();
  for (int64_t i = 0; i < nV; ++i) { out[3 * i] = N[i].x; out[3 * i + 1] = N[i].y; out[3 * i + 2] = N[i].z; }
}

}  // extern "C"

/* TEST INFRASTRUCTURE — CPU restatement ("oracle") of the segs.json consumers.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library; the product never links it.
 *
 * What is restated, and from where:
 *   - segment -> vertex lists            /root/reference/AnnotationTools/common/Segmentation.h:68-75
 *   - surface area per segment           Segmentation.h:113-147, Trianglef::getArea external/mLib/include/core-graphics/triangle.h:23-35
 *   - object id per vertex               AnnotationTools/ProjectAnnotations/Visualizer.cpp:284-297
 *   - vertex normals                     external/mLib/include/core-mesh/meshData.h:758-782 (vec3 ops core-math/vec3.h:154-160,191-243)
 *   - annotation propagation             Visualizer.cpp:308-377 — the reference queries an APPROXIMATE FLANN kd-tree
 *                                        (not in the tree); this is the exact 3-nearest-neighbour statement of the same rule,
 *                                        brute force, ties by source index.
 * Arithmetic: IEEE binary32, no contraction, association order as in the source.
 * Parity status: area and normal arithmetic pinned against the real mLib operators (oracle/_ref/libref_mlib.so, compiled
 * from the headers where they lie); propagation is "parity unpinned" (no FLANN, no fixture in the reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

float oracle_tri_area_mlib(const float* a, const float* b, const float* c) {
  const float abx = b[0] - a[0], aby = b[1] - a[1], abz = b[2] - a[2];
  const float acx = c[0] - a[0], acy = c[1] - a[1], acz = c[2] - a[2];
  const float lab = sqrtf(abx * abx + aby * aby + abz * abz), lac = sqrtf(acx * acx + acy * acy + acz * acz);
  const float len = lab * lac;
  const float ct = (abx * acx + aby * acy + abz * acz) / len;
  if (fabs(ct + 1) < 0.00001f || fabs(ct - 1) < 0.00001f) return 0.f;      /* fabs(double) of a float sum, as the C++ resolves it */
  const float th = acosf(ct);
  return 0.5f * len * sinf(th);
}

static int cmp_pair(const void* x, const void* y) {
  const uint64_t a = *(const uint64_t*)x, b = *(const uint64_t*)y;
  return a < b ? -1 : a > b;
}

/* outputs sized by the caller: seg_ids[nV], offsets[nV+1], vert_ids[nV], area[nV] (first *n_segs entries used) */
int oracle_segs_aggregate(const uint32_t* seg, int64_t nV, const float* xyz, const uint32_t* tri, int64_t nF,
                          uint32_t* seg_ids, int64_t* n_segs, uint64_t* offsets, uint32_t* vert_ids, float* area) {
  uint64_t* pr = (uint64_t*)malloc((size_t)(nV > 0 ? nV : 1) * 8);
  for (int64_t i = 0; i < nV; ++i) pr[i] = ((uint64_t)seg[i] << 32) | (uint64_t)i;
  qsort(pr, (size_t)nV, 8, cmp_pair);
  int64_t nS = 0;
  int64_t* dense = (int64_t*)malloc((size_t)(nV > 0 ? nV : 1) * 8);
  for (int64_t i = 0; i < nV; ++i) {
    const uint32_t s = (uint32_t)(pr[i] >> 32), v = (uint32_t)pr[i];
    if (i == 0 || s != (uint32_t)(pr[i - 1] >> 32)) { seg_ids[nS] = s; offsets[nS] = (uint64_t)i; ++nS; }
    vert_ids[i] = v; dense[v] = nS - 1;
  }
  offsets[nS] = (uint64_t)nV;
  *n_segs = nS;
  if (area && xyz && tri) {
    double* acc = (double*)calloc((size_t)(nS > 0 ? nS : 1), 8);
    for (int64_t f = 0; f < nF; ++f) {                                    /* a face counts iff all 3 corners are in the segment */
      const uint32_t i0 = tri[3 * f], i1 = tri[3 * f + 1], i2 = tri[3 * f + 2];
      if (i0 >= nV || i1 >= nV || i2 >= nV) continue;
      if (dense[i0] == dense[i1] && dense[i0] == dense[i2]) acc[dense[i0]] += (double)oracle_tri_area_mlib(xyz + 3 * i0, xyz + 3 * i1, xyz + 3 * i2);
    }
    for (int64_t s = 0; s < nS; ++s) area[s] = (float)acc[s];
    free(acc);
  }
  free(dense); free(pr);
  return 0;
}

void oracle_objects_per_vertex(const uint32_t* seg, int64_t nV, const uint32_t* group_segs, const uint64_t* group_offsets,
                               int64_t n_groups, uint32_t* obj) {
  for (int64_t v = 0; v < nV; ++v) obj[v] = 0;
  for (int64_t g = 0; g < n_groups; ++g)                                   /* later groups overwrite earlier ones */
    for (uint64_t j = group_offsets[g]; j < group_offsets[g + 1]; ++j)
      for (int64_t v = 0; v < nV; ++v) if (seg[v] == group_segs[j]) obj[v] = (uint32_t)(g + 1);
}

void oracle_vertex_normals_mlib(const float* xyz, int64_t nV, const uint32_t* tri, int64_t nF, float* nrm) {
  for (int64_t i = 0; i < 3 * nV; ++i) nrm[i] = 0.f;
  for (int64_t f = 0; f < nF; ++f) {
    const float* p0 = xyz + 3 * (size_t)tri[3 * f]; const float* p1 = xyz + 3 * (size_t)tri[3 * f + 1]; const float* p2 = xyz + 3 * (size_t)tri[3 * f + 2];
    const float ux = p1[0] - p0[0], uy = p1[1] - p0[1], uz = p1[2] - p0[2];
    const float vx = p2[0] - p0[0], vy = p2[1] - p0[1], vz = p2[2] - p0[2];
    float n[3] = {0.f, 0.f, 0.f};
    n[0] += uy * vz - uz * vy; n[1] += uz * vx - ux * vz; n[2] += ux * vy - uy * vx;
    const float val = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] *= val; n[1] *= val; n[2] *= val;
    for (int k = 0; k < 3; ++k) { float* d = nrm + 3 * (size_t)tri[3 * f + k]; d[0] += n[0]; d[1] += n[1]; d[2] += n[2]; }
  }
  for (int64_t v = 0; v < nV; ++v) {
    float* d = nrm + 3 * v;
    const float val = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    d[0] *= val; d[1] *= val; d[2] *= val;
  }
}

/* borderline[i] (optional) = 1 when some normal test of vertex i was within 1e-5 rad of the threshold (acosf differs by ulps
 * between libm implementations, so those vertices are excluded from exact comparison). */
void oracle_propagate_labels(const float* sxyz, const float* snrm, const uint32_t* sobj, int64_t nS, const float* dxyz,
                             const float* dnrm, int64_t nD, float normal_thresh, uint32_t* out, uint8_t* borderline) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = 0; i < nS; ++i) for (int k = 0; k < 3; ++k) { const float v = sxyz[3 * i + k]; if (v < mn[k]) mn[k] = v; if (v > mx[k]) mx[k] = v; }
  float ext = 0.f; for (int k = 0; k < 3; ++k) { const float e = mx[k] - mn[k]; if (e > ext) ext = e; }
  const float e1 = ext * 0.01f; const float max_thresh = e1 > 0.05f ? e1 : 0.05f;
  #pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < nD; ++i) {
    float bd[3] = {INFINITY, INFINITY, INFINITY}; int64_t bi[3] = {-1, -1, -1};
    const float px = dxyz[3 * i], py = dxyz[3 * i + 1], pz = dxyz[3 * i + 2];
    for (int64_t s = 0; s < nS; ++s) {
      if (!(sobj[s] > 0)) continue;
      const float dx = sxyz[3 * s] - px, dy = sxyz[3 * s + 1] - py, dz = sxyz[3 * s + 2] - pz;
      const float d = sqrtf(dx * dx + dy * dy + dz * dz);
      if (!(d < max_thresh)) continue;
      for (int k = 0; k < 3; ++k)
        if (d < bd[k]) {                                                 /* strict: equal distance keeps the lower source index first */
          for (int j = 2; j > k; --j) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; }
          bd[k] = d; bi[k] = s; break;
        }
    }
    uint32_t res = 0; uint8_t edge = 0;
    if (bi[0] >= 0) {
      const uint32_t val = sobj[bi[0]];
      int all_same = 1, best = -1;
      for (int k = 0; k < 3; ++k) {
        if (bi[k] >= 0) {
          const float* sn = snrm + 3 * bi[k];
          float dot = sn[0] * dnrm[3 * i] + sn[1] * dnrm[3 * i + 1] + sn[2] * dnrm[3 * i + 2];
          dot = dot < -1.0f ? -1.0f : (dot > 1.0f ? 1.0f : dot);
          const float ang = acosf(dot);
          if (fabsf(ang - normal_thresh) < 1e-5f) edge = 1;
          if (ang < normal_thresh) { best = k; break; }
          if (sobj[bi[k]] != val) all_same = 0;
        } else all_same = 0;
      }
      if (best >= 0) res = sobj[bi[best]]; else if (all_same) res = val;
    }
    out[i] = res; if (borderline) borderline[i] = edge;
  }
}

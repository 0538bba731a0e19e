/* TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference Segmentator.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libscannet_b200.so) never links it.
 *
 * What is restated, and from where:
 *   - vertex normals + edge list       /root/reference/Segmentator/segmentator.cpp:185-208
 *   - edge weights                      segmentator.cpp:211-229
 *   - sort by weight                    segmentator.cpp:67-72  -> std::sort of libstdc++
 *       std::sort is NOT in the reference tree: it is libstdc++ (GCC 13.3.0,
 *       GLIBCXX_3.4.33, bits/stl_algo.h:1848-1952 + bits/stl_heap.h) and its unstable
 *       tie order is observable in segIndices.  The published
 *       introsort algorithm is restated below (median-of-3 Hoare partition to 16-element
 *       leaves, depth limit 2*floor(log2 n) with heap-sort fallback, final insertion
 *       sort).  Pinned by tests/test_oracle_pinning.py against the reference's own
 *       segment_graph() (oracle/_ref/libref_segmentator.so) on tie-heavy inputs.
 *   - Kruskal with adaptive threshold   segmentator.cpp:71-91, universe :24-60
 *   - small-segment merge               segmentator.cpp:236-243
 *   - output ids                        segmentator.cpp:245-250
 *
 * Arithmetic: IEEE binary32, no contraction (compile with -ffp-contract=off; the
 * reference binary contains no FMA instructions), association order as in the source.
 *
 * Parity status: pinned against the compiled reference (gates381.ply golden hash
 * b57dfeed… from BASELINE.md, synthetic meshes, adversarial meshes).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float w; int32_t a, b; } oracle_edge;   /* 12 B, as segmentator.cpp:62-65 */

/* ------------------------------------------------------------------ normals + edges */
/* segmentator.cpp:185-208.  points[] is only written for referenced vertices, but an
 * edge endpoint is always referenced, so xyz can be read directly in the weight loop. */
static void face_pass(const float* xyz, int64_t nV, const uint32_t* tri, int64_t nF,
                      float* nrm /*3*nV, zeroed*/, oracle_edge* e) {
  int32_t* counts = (int32_t*)calloc((size_t)(nV > 0 ? nV : 1), sizeof(int32_t));
  for (int64_t f = 0; f < nF; ++f) {
    const uint32_t i1 = tri[3 * f], i2 = tri[3 * f + 1], i3 = tri[3 * f + 2];
    const float* p1 = xyz + 3 * (size_t)i1;
    const float* p2 = xyz + 3 * (size_t)i2;
    const float* p3 = xyz + 3 * (size_t)i3;
    e[3 * f + 0].a = (int32_t)i1; e[3 * f + 0].b = (int32_t)i2;
    e[3 * f + 1].a = (int32_t)i1; e[3 * f + 1].b = (int32_t)i3;
    e[3 * f + 2].a = (int32_t)i3; e[3 * f + 2].b = (int32_t)i2;
    /* cross(p2-p1, p3-p1), normalised (segmentator.cpp:107-112) */
    const float ux = p2[0] - p1[0], uy = p2[1] - p1[1], uz = p2[2] - p1[2];
    const float vx = p3[0] - p1[0], vy = p3[1] - p1[1], vz = p3[2] - p1[2];
    float cx = uy * vz - uz * vy;
    float cy = uz * vx - ux * vz;
    float cz = ux * vy - uy * vx;
    const float len = sqrtf(cx * cx + cy * cy + cz * cz);
    cx /= len; cy /= len; cz /= len;
    /* three running-average updates, counts bumped only afterwards (segmentator.cpp:203-207) */
    const uint32_t idx[3] = { i1, i2, i3 };
    for (int k = 0; k < 3; ++k) {
      float* n = nrm + 3 * (size_t)idx[k];
      const float v = 1.0f / ((float)counts[idx[k]] + 1.0f);
      const float u = 1.0f - v;
      n[0] = v * cx + u * n[0];
      n[1] = v * cy + u * n[1];
      n[2] = v * cz + u * n[2];
    }
    counts[i1]++; counts[i2]++; counts[i3]++;
  }
  free(counts);
}

/* segmentator.cpp:211-229 */
static void weight_pass(const float* xyz, const float* nrm, oracle_edge* e, int64_t nE) {
  for (int64_t i = 0; i < nE; ++i) {
    const float* n1 = nrm + 3 * (size_t)e[i].a;
    const float* n2 = nrm + 3 * (size_t)e[i].b;
    const float* p1 = xyz + 3 * (size_t)e[i].a;
    const float* p2 = xyz + 3 * (size_t)e[i].b;
    float dx = p2[0] - p1[0], dy = p2[1] - p1[1], dz = p2[2] - p1[2];
    const float dd = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= dd; dy /= dd; dz /= dd;
    const float dot  = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2];
    const float dot2 = n2[0] * dx + n2[1] * dy + n2[2] * dz;
    float ww = 1.0f - dot;
    if (dot2 > 0) ww = ww * ww;
    e[i].w = ww;
  }
}

void oracle_seg_build_edges(const float* xyz, int64_t nV, const uint32_t* tri, int64_t nF,
                            oracle_edge* edges /*3*nF*/, float* normals_out /*3*nV or NULL*/) {
  float* nrm = (float*)calloc((size_t)(nV > 0 ? 3 * nV : 1), sizeof(float));
  face_pass(xyz, nV, tri, nF, nrm, edges);
  weight_pass(xyz, nrm, edges, 3 * nF);
  if (normals_out) memcpy(normals_out, nrm, (size_t)nV * 3 * sizeof(float));
  free(nrm);
}

/* ------------------------------------------------------------------ libstdc++ std::sort */
/* comparator: segmentator.cpp:67-69 */
#define LT(x, y) ((x).w < (y).w)
static inline void eswap(oracle_edge* x, oracle_edge* y) { oracle_edge t = *x; *x = *y; *y = t; }

/* bits/stl_heap.h __push_heap / __adjust_heap / __make_heap / __pop_heap / __sort_heap */
static void heap_sift(oracle_edge* a, int64_t hole, int64_t len, oracle_edge val) {
  const int64_t top = hole;
  int64_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (LT(a[child], a[child - 1])) child--;
    a[hole] = a[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a[hole] = a[child - 1];
    hole = child - 1;
  }
  int64_t parent = (hole - 1) / 2;
  while (hole > top && LT(a[parent], val)) {
    a[hole] = a[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a[hole] = val;
}
static void heap_sort_range(oracle_edge* a, int64_t len) {   /* __partial_sort(first,last,last) */
  if (len >= 2) {
    for (int64_t parent = (len - 2) / 2;; --parent) {
      heap_sift(a, parent, len, a[parent]);
      if (parent == 0) break;
    }
  }
  for (int64_t last = len; last > 1;) {
    --last;
    oracle_edge v = a[last];
    a[last] = a[0];
    heap_sift(a, 0, last, v);
  }
}

/* bits/stl_algo.h:85-108,1871-1900 */
static int64_t partition_pivot(oracle_edge* a, int64_t first, int64_t last) {
  const int64_t mid = first + (last - first) / 2;
  oracle_edge *r = a + first, *x = a + first + 1, *y = a + mid, *z = a + last - 1;
  if (LT(*x, *y)) {
    if (LT(*y, *z)) eswap(r, y); else if (LT(*x, *z)) eswap(r, z); else eswap(r, x);
  } else if (LT(*x, *z)) eswap(r, x);
  else if (LT(*y, *z)) eswap(r, z);
  else eswap(r, y);
  int64_t lo = first + 1, hi = last;
  const oracle_edge* piv = a + first;
  for (;;) {
    while (LT(a[lo], *piv)) ++lo;
    --hi;
    while (LT(*piv, a[hi])) --hi;
    if (!(lo < hi)) return lo;
    eswap(a + lo, a + hi);
    ++lo;
  }
}
static void introsort_loop(oracle_edge* a, int64_t first, int64_t last, int depth) {
  while (last - first > 16) {
    if (depth == 0) { heap_sort_range(a + first, last - first); return; }
    --depth;
    const int64_t cut = partition_pivot(a, first, last);
    introsort_loop(a, cut, last, depth);
    last = cut;
  }
}
static void linear_insert_unguarded(oracle_edge* a, int64_t pos) {
  oracle_edge v = a[pos];
  int64_t nx = pos - 1;
  while (LT(v, a[nx])) { a[pos] = a[nx]; pos = nx; --nx; }
  a[pos] = v;
}
static void insertion_sort_guarded(oracle_edge* a, int64_t first, int64_t last) {
  if (first == last) return;
  for (int64_t i = first + 1; i != last; ++i) {
    if (LT(a[i], a[first])) {
      oracle_edge v = a[i];
      memmove(a + first + 1, a + first, (size_t)(i - first) * sizeof(oracle_edge));
      a[first] = v;
    } else linear_insert_unguarded(a, i);
  }
}
/* forced_depth > 0 overrides the depth limit (test hook to reach the heap-sort fallback) */
void oracle_seg_sort_edges_depth(oracle_edge* a, int64_t n, int forced_depth) {
  if (n <= 0) return;
  int lg = 0; for (int64_t t = n; t > 1; t >>= 1) ++lg;            /* std::__lg */
  introsort_loop(a, 0, n, forced_depth > 0 ? forced_depth : 2 * lg);
  if (n > 16) {
    insertion_sort_guarded(a, 0, 16);
    for (int64_t i = 16; i != n; ++i) linear_insert_unguarded(a, i);
  } else insertion_sort_guarded(a, 0, n);
}
void oracle_seg_sort_edges(oracle_edge* a, int64_t n) { oracle_seg_sort_edges_depth(a, n, 0); }

/* ------------------------------------------------------------------ union-find + Kruskal */
typedef struct { int32_t rank, p, size; } uf_elt;                  /* segmentator.cpp:18-22 */
static inline int32_t uf_find(uf_elt* u, int32_t x) {              /* :36-42, one-step compression */
  int32_t y = x;
  while (y != u[y].p) y = u[y].p;
  u[x].p = y;
  return y;
}
static inline void uf_join(uf_elt* u, int32_t x, int32_t y) {      /* :43-54 */
  if (u[x].rank > u[y].rank) { u[y].p = x; u[x].size += u[y].size; }
  else { u[x].p = y; u[y].size += u[x].size; if (u[x].rank == u[y].rank) u[y].rank++; }
}

/* Full pipeline over raw arrays.  Optional outputs may be NULL.
 * edges_presort/edges_sorted: 3*nF records each. */
int oracle_segment_arrays(const float* xyz, int64_t nV, const uint32_t* tri, int64_t nF,
                          float kthr, int32_t seg_min_verts, int32_t* seg_out /*nV*/,
                          oracle_edge* edges_presort, oracle_edge* edges_sorted,
                          int32_t* roots_after_kruskal /*nV*/) {
  const int64_t nE = 3 * nF;
  oracle_edge* e = (oracle_edge*)malloc((size_t)(nE > 0 ? nE : 1) * sizeof(oracle_edge));
  if (!e) return -1;
  oracle_seg_build_edges(xyz, nV, tri, nF, e, NULL);
  if (edges_presort) memcpy(edges_presort, e, (size_t)nE * sizeof(oracle_edge));
  oracle_seg_sort_edges(e, nE);
  if (edges_sorted) memcpy(edges_sorted, e, (size_t)nE * sizeof(oracle_edge));

  uf_elt* u = (uf_elt*)malloc((size_t)(nV > 0 ? nV : 1) * sizeof(uf_elt));
  float* thr = (float*)malloc((size_t)(nV > 0 ? nV : 1) * sizeof(float));
  for (int64_t i = 0; i < nV; ++i) { u[i].rank = 0; u[i].size = 1; u[i].p = (int32_t)i; thr[i] = kthr; }
  for (int64_t i = 0; i < nE; ++i) {                               /* :77-89 */
    int32_t a = uf_find(u, e[i].a), b = uf_find(u, e[i].b);
    if (a != b && e[i].w <= thr[a] && e[i].w <= thr[b]) {
      uf_join(u, a, b);
      a = uf_find(u, a);
      thr[a] = e[i].w + (kthr / (float)u[a].size);
    }
  }
  if (roots_after_kruskal)
    for (int64_t q = 0; q < nV; ++q) { int32_t y = (int32_t)q; while (y != u[y].p) y = u[y].p; roots_after_kruskal[q] = y; }
  for (int64_t j = 0; j < nE; ++j) {                               /* :237-243 */
    const int32_t a = uf_find(u, e[j].a), b = uf_find(u, e[j].b);
    if (a != b && (u[a].size < seg_min_verts || u[b].size < seg_min_verts)) uf_join(u, a, b);
  }
  for (int64_t q = 0; q < nV; ++q) seg_out[q] = uf_find(u, (int32_t)q);   /* :247-249 */
  free(thr); free(u); free(e);
  return 0;
}

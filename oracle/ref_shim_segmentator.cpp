// TEST INFRASTRUCTURE — not product code.
//
// C-callable shim around the UNMODIFIED reference Segmentator.  This file is
// linked (by oracle/Makefile) against object code compiled directly from
// /root/reference/Segmentator/segmentator.cpp (with -Dmain=ref_segmentator_main)
// and tinyply.cpp, where those sources lie.  No reference source is copied into
// this repository; only the resulting oracle/_ref/libref_segmentator.so travels
// to the GPU box.
//
// The reference has two non-static seams we bind to (C++ linkage):
//   std::vector<int> segment(const std::string&, float, int)   segmentator.cpp:123
//   universe* segment_graph(int, int, edge*, float)            segmentator.cpp:71
// The struct/class layouts below restate segmentator.cpp:17-65 only as far as
// the ABI needs (field order and sizes), so the shim can walk the result.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

// --- ABI mirror of segmentator.cpp:17-65 (layout only) ---------------------
typedef struct { int rank; int p; int size; } uni_elt;   // segmentator.cpp:18-22
class universe;                                          // segmentator.cpp:24-60 (opaque here)
typedef struct { float w; int a, b; } edge;              // segmentator.cpp:62-65

// Defined in the reference translation unit.
std::vector<int> segment(const std::string& meshFile, const float kthr, const int segMinVerts);
universe* segment_graph(int num_vertices, int num_edges, edge* edges, float c);
int ref_segmentator_main(int argc, const char** argv);

// The reference's universe members are defined inline in the class body, so the
// reference TU need not emit them; walk the parent array ourselves using the
// data layout {uni_elt* elts; int num;} (segmentator.cpp:56-59).
struct universe_view { uni_elt* elts; int num; };

extern "C" {

// Runs reference segment() on a mesh file.  out must hold nV ints; returns the
// number of vertices the reference reported (or -1 if out_cap is too small).
int64_t ref_segment_file(const char* path, float kthr, int seg_min_verts,
                         int32_t* out, int64_t out_cap) {
  std::vector<int> r = segment(std::string(path), kthr, seg_min_verts);
  if ((int64_t)r.size() > out_cap) return -1;
  std::memcpy(out, r.data(), r.size() * sizeof(int));
  return (int64_t)r.size();
}

// Runs reference segment_graph() (libstdc++ std::sort + Kruskal-with-threshold)
// on caller-supplied 12-byte {w,a,b} records.  `edges` is sorted IN PLACE exactly
// as the reference does; roots_out[v] = find(v) after the Kruskal phase;
// sizes_out[v] = size(root(v)).
void ref_segment_graph(int32_t n_verts, int32_t n_edges, void* edges, float c,
                       int32_t* roots_out, int32_t* sizes_out) {
  universe* u = segment_graph(n_verts, n_edges, (edge*)edges, c);
  universe_view* uv = (universe_view*)u;
  for (int v = 0; v < n_verts; ++v) {
    int y = v;
    while (y != uv->elts[y].p) y = uv->elts[y].p;
    if (roots_out) roots_out[v] = y;
    if (sizes_out) sizes_out[v] = uv->elts[y].size;
  }
  delete[] uv->elts;          // what ~universe() does (segmentator.cpp:34)
  ::operator delete((void*)u);
}

// The reference CLI itself (argv handling, stdout lines, segs.json writer).
int ref_segmentator_cli(int argc, const char** argv) {
  return ref_segmentator_main(argc, argv);
}

}  // extern "C"

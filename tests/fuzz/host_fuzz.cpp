// TEST INFRASTRUCTURE.  Mutation fuzzer for the host-side parsers (.sens container, inflate, JPEG, PNG, PLY/OBJ, segs.json),
// built with AddressSanitizer + UBSan from the product sources (tests/test_fuzz_cpu.py compiles and runs it).
// Every mutated input must come back as a status code: no crash, no sanitizer report, no hang.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scannet_b200.h"

// the CUDA side of the library is not linked here
namespace scn {
std::string& last_error_ref() { static thread_local std::string e; return e; }
int fail(int code, const char* fmt, ...) { char b[512]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap); last_error_ref() = b; return code; }
}
extern "C" {
const char* scn_last_error(void) { return scn::last_error_ref().c_str(); }
void scn_free(void* p) { free(p); }
int scn_cuda_warmup(void) { return 0; }
int scn_segment_last_timings(float*) { return 0; }
int scn_segment_mesh(const float*, uint64_t, const uint32_t*, uint64_t, float, int32_t, int32_t*, int) { return SCN_ERR_CUDA; }
}

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

static std::vector<uint8_t> slurp(const char* p) {
  std::vector<uint8_t> d; FILE* f = fopen(p, "rb"); if (!f) return d;
  fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET); d.resize(n > 0 ? n : 0);
  if (n > 0 && fread(d.data(), 1, (size_t)n, f) != (size_t)n) d.clear();
  fclose(f); return d;
}
static void spit(const std::string& p, const std::vector<uint8_t>& d) { FILE* f = fopen(p.c_str(), "wb"); if (f) { fwrite(d.data(), 1, d.size(), f); fclose(f); } }

static std::vector<uint8_t> mutate(const std::vector<uint8_t>& seed, size_t skip) {
  std::vector<uint8_t> d = seed;
  const int ops = 1 + rnd() % 6;
  const bool body_only = rnd() % 4 != 0;                                       // mostly leave the container header alone: reach the decoders
  for (int o = 0; o < ops && !d.empty(); ++o) {
    const size_t lo = body_only && d.size() > 2 * skip ? skip : 0;
    const size_t pos = lo + rnd() % (d.size() - lo);
    switch (rnd() % 6) {
      case 0: d[pos] ^= (uint8_t)(1u << (rnd() % 8)); break;
      case 1: d[pos] = (uint8_t)rnd(); break;
      case 2: d[pos] = (rnd() & 1) ? 0xFF : 0x00; break;
      case 3: d.resize(pos); break;                                            // truncate
      case 4: { const size_t n = 1 + rnd() % 16; d.insert(d.begin() + pos, n, (uint8_t)rnd()); } break;
      default: { const size_t n = std::min<size_t>(1 + rnd() % 32, d.size() - pos); d.erase(d.begin() + pos, d.begin() + pos + n); } break;
    }
  }
  return d;
}

int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "cache") {
    // read-ahead cache stress (built with ThreadSanitizer by tests/test_fuzz_cpu.py): many cache sizes / thread counts, early destroys
    scn_sens* s = nullptr;
    if (scn_sens_open(argv[2], &s)) { fprintf(stderr, "%s\n", scn_last_error()); return 2; }
    scn_sens_info_t in; scn_sens_info(s, &in);
    std::vector<uint16_t> dep((size_t)in.depth_width * in.depth_height), ref(dep.size()); std::vector<uint8_t> col((size_t)in.color_width * in.color_height * 3);
    long frames = 0;
    for (int round = 0; round < 24; ++round) {
      scn_sens_cache* c = nullptr;
      if (scn_sens_cache_create(s, 1 + round % 7, 1 + round % 5, &c)) return 3;
      const uint64_t stop_at = round % 3 == 2 ? in.n_frames / 2 : in.n_frames + 1;      // every third round abandons the stream half way
      uint64_t i = 0; uint64_t td, tc;
      while (i < stop_at && scn_sens_cache_next(c, dep.data(), col.data(), &td, &tc) == 1) {
        scn_sens_frame_depth_u16(s, i, ref.data());
        if (memcmp(ref.data(), dep.data(), dep.size() * 2)) { fprintf(stderr, "frame %llu differs\n", (unsigned long long)i); return 4; }
        ++i; ++frames;
      }
      scn_sens_cache_destroy(c);
    }
    scn_sens_close(s);
    printf("cache: %ld frames delivered in order\n", frames);
    return 0;
  }
  if (argc < 5) { fprintf(stderr, "usage: host_fuzz <kind: sens|mesh|segs> <seed file> <work dir> <iterations>  |  host_fuzz cache <file.sens>\n"); return 2; }
  const std::string kind = argv[1], work = argv[3]; const int iters = atoi(argv[4]);
  const std::vector<uint8_t> seed = slurp(argv[2]);
  if (seed.empty()) { fprintf(stderr, "empty seed\n"); return 2; }
  const std::string ext = std::string(argv[2]).substr(std::string(argv[2]).find_last_of('.'));
  int ok = 0, rejected = 0;
  for (int it = 0; it < iters; ++it) {
    const std::vector<uint8_t> d = it == 0 ? seed : mutate(seed, kind == "sens" ? 400 : 0);
    const std::string p = work + "/case" + ext;
    spit(p, d);
    if (kind == "sens") {
      scn_sens* s = nullptr;
      if (scn_sens_open(p.c_str(), &s)) { ++rejected; continue; }
      scn_sens_info_t in; scn_sens_info(s, &in);
      const uint64_t px_d = (uint64_t)in.depth_width * in.depth_height, px_c = (uint64_t)in.color_width * in.color_height;
      bool all = true;
      if (px_d < (1u << 24) && px_c < (1u << 24)) {
        std::vector<uint16_t> dep(px_d ? px_d : 1); std::vector<uint8_t> col(px_c ? px_c * 3 : 1);
        for (uint64_t i = 0; i < in.n_frames && i < 8; ++i) {
          if (scn_sens_frame_depth_u16(s, i, dep.data())) all = false;
          if (scn_sens_frame_color_rgb8(s, i, col.data())) all = false;
        }
        char buf[4096]; scn_sens_describe(s, buf, sizeof buf);
      } else all = false;
      scn_sens_close(s);
      all ? ++ok : ++rejected;
    } else if (kind == "mesh") {
      float* xyz = nullptr; uint32_t* tri = nullptr; uint64_t nv = 0, nf = 0;
      if (scn_mesh_load(p.c_str(), &xyz, &nv, &tri, &nf)) ++rejected; else { ++ok; scn_free(xyz); scn_free(tri); }
    } else {
      uint32_t* seg = nullptr; uint64_t n = 0; float k; uint32_t m; char sid[64];
      if (scn_segs_load(p.c_str(), &seg, &n, &k, &m, sid, sizeof sid)) ++rejected; else { ++ok; scn_free(seg); }
    }
  }
  printf("%s: %d inputs, %d accepted, %d rejected\n", kind.c_str(), iters, ok, rejected);
  return 0;
}

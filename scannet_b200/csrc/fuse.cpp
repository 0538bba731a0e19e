// `fuse` — the reconstruction stage contract of the reference pipeline in library form:
//   <exe> <params.txt> [<params2.txt>] <file.sens> [out.ply]        (Server/scan_processor.py:27-35,123-138)
// Reads the .sens stream (SensReader), fuses every frame that has a valid pose into the hashed TSDF volume
// (poses come from the file: camera tracking / bundle adjustment is not part of this path), extracts the
// surface with marching cubes and writes `<base>_vh.ply` in the VCGLIB layout the `segment` stage reads
// (Server/config/scan_stages.json:33-37).  Host decode (inflate / JPEG) runs on a thread pool one chunk ahead
// of the GPU; frames are handed over in pinned memory.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "scn_common.h"

namespace {

// Persistent worker pool: run(fn) executes fn() on every worker and on the caller, and returns when all are done
// (spawning 64 threads per 128-frame chunk cost ~3 ms of a ~10 ms chunk).
class Pool {
 public:
  explicit Pool(unsigned workers) { for (unsigned i = 0; i < workers; ++i) th_.emplace_back([this]() { loop(); }); }
  ~Pool() { { std::lock_guard<std::mutex> l(m_); stop_ = true; ++gen_; } cv_.notify_all(); for (auto& t : th_) t.join(); }
  void run(const std::function<void()>& fn) {
    { std::lock_guard<std::mutex> l(m_); fn_ = &fn; pending_ = (unsigned)th_.size(); ++gen_; }
    cv_.notify_all();
    fn();
    std::unique_lock<std::mutex> l(m_); done_.wait(l, [this]() { return pending_ == 0; });
  }
 private:
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void()>* fn;
      { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&]() { return gen_ != seen; }); seen = gen_; if (stop_) return; fn = fn_; }
      (*fn)();
      { std::lock_guard<std::mutex> l(m_); if (--pending_ == 0) done_.notify_one(); }
    }
  }
  std::vector<std::thread> th_; std::mutex m_; std::condition_variable cv_, done_;
  const std::function<void()>* fn_ = nullptr; unsigned pending_ = 0; unsigned long gen_ = 0; bool stop_ = false;
};

struct Chunk { uint16_t* depth = nullptr; uint8_t* rgb = nullptr; std::vector<float> poses; uint32_t n = 0; int rc = 0; std::string err; };

// colour pixel for every depth pixel: K_c * K_d^-1 * (x, y, 1), nearest (sensorData.h:1577-1586 with identity extrinsics)
void build_color_lut(const scn_sens_info_t& in, std::vector<int32_t>& lut) {
  const uint32_t W = in.depth_width, H = in.depth_height;
  lut.resize((size_t)W * H);
  const float fxd = in.depth_intrinsic[0], cxd = in.depth_intrinsic[2], fyd = in.depth_intrinsic[5], cyd = in.depth_intrinsic[6];
  const float fxc = in.color_intrinsic[0], cxc = in.color_intrinsic[2], fyc = in.color_intrinsic[5], cyc = in.color_intrinsic[6];
  for (uint32_t y = 0; y < H; ++y) for (uint32_t x = 0; x < W; ++x) {
    const float u = ((float)x - cxd) / fxd * fxc + cxc, v = ((float)y - cyd) / fyd * fyc + cyc;
    const long cx = lrintf(u), cy = lrintf(v);
    lut[(size_t)y * W + x] = (cx >= 0 && cy >= 0 && cx < (long)in.color_width && cy < (long)in.color_height) ? (int32_t)(cy * in.color_width + cx) : -1;
  }
}

void decode_chunk(const scn_sens* s, const scn_sens_info_t& in, const std::vector<int32_t>& lut, bool use_color, bool want_depth, uint64_t f0, uint32_t n, Chunk& c, Pool& pool) {
  const size_t px = (size_t)in.depth_width * in.depth_height;
  c.n = n; c.rc = 0; c.poses.assign((size_t)n * 16, 0.f);
  std::atomic<uint32_t> next{0}; std::atomic<int> rc{0};
  std::mutex err_m;
  auto work = [&]() {
    static thread_local std::vector<uint8_t> col;                               // kept per pool thread: no 1-4 MB allocation per chunk
    col.resize(use_color ? (size_t)in.color_width * in.color_height * 3 : 0);
    for (;;) {
      const uint32_t i = next.fetch_add(1);
      if (i >= n || rc.load()) break;
      scn_sens_frame_meta(s, f0 + i, &c.poses[(size_t)i * 16], nullptr, nullptr, nullptr, nullptr);
      if (c.poses[(size_t)i * 16] == -INFINITY) continue;                        // skipped by the integrator anyway
      int r = want_depth ? scn_sens_frame_depth_u16(s, f0 + i, c.depth + (size_t)i * px) : 0;
      if (!r && use_color) {
        r = scn_sens_frame_color_rgb8(s, f0 + i, col.data());
        if (!r) { uint8_t* o = c.rgb + (size_t)i * px * 3;
          for (size_t p = 0; p < px; ++p) { const int32_t q = lut[p]; if (q >= 0) { o[3 * p] = col[3 * q]; o[3 * p + 1] = col[3 * q + 1]; o[3 * p + 2] = col[3 * q + 2]; } else o[3 * p] = o[3 * p + 1] = o[3 * p + 2] = 0; } }
      }
      // the error text is thread-local: capture it on the worker that failed
      if (r) { std::lock_guard<std::mutex> l(err_m); if (!rc.load()) { c.err = scn_last_error(); rc.store(r); } }
    }
  };
  pool.run(work);
  c.rc = rc.load();
}

}  // namespace

extern "C" int scn_fuse_main(int argc, const char** argv) {
  std::vector<std::string> params; std::string sens_path, out_path;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a.size() > 5 && a.substr(a.size() - 5) == ".sens") sens_path = a;
    else if (!sens_path.empty()) out_path = a;
    else params.push_back(a);
  }
  if (sens_path.empty()) { printf("Usage: fuse <params.txt> [<params2.txt>] <file.sens> [out.ply]\n"); return 255; }
  if (out_path.empty()) out_path = sens_path.substr(0, sens_path.size() - 5) + "_vh.ply";
  std::thread warm([]() { scn_cuda_warmup(); });          // context creation overlaps reading the .sens file
  scn_sens* s = nullptr;
  const int open_rc = scn_sens_open(sens_path.c_str(), &s);
  warm.join();
  if (open_rc) { fprintf(stderr, "%s\n", scn_last_error()); return 1; }
  scn_sens_info_t in; scn_sens_info(s, &in);
  scn_tsdf_params p; scn_tsdf_default_params(&p);
  p.max_blocks = 1ull << 22; p.hash_slots = 1ull << 24;      // 16 GiB of voxel blocks unless the parameter file says otherwise (s_hashNumSDFBlocks)
  for (const std::string& f : params) if (scn_tsdf_params_from_file(f.c_str(), &p)) { fprintf(stderr, "%s\n", scn_last_error()); scn_sens_close(s); return 1; }
  p.width = in.depth_width; p.height = in.depth_height; p.depth_shift = in.depth_shift;   // integrate at the stream's depth resolution
  p.batch_frames = 16;
  const bool use_color = in.color_compression <= 2;           // raw, PNG, JPEG (sensorData.h:600-616)
  printf("fusing %s: %llu frames %ux%u, voxel %.4f m, truncation %.3f+%.3f*d\n", sens_path.c_str(), (unsigned long long)in.n_frames, in.depth_width,
         in.depth_height, p.voxel_size, p.trunc_base, p.trunc_scale);
  scn_tsdf* vol = nullptr;
  if (scn_tsdf_create(&p, 0, &vol)) { fprintf(stderr, "%s\n", scn_last_error()); scn_sens_close(s); return 1; }
  const size_t px = (size_t)in.depth_width * in.depth_height;
  // host decode is the bottleneck of this tool (one frame = ~2 ms inflate + ~3 ms JPEG on one core, the GPU fuses a frame in
  // ~30 us): use up to 64 cores, and chunks of two frames per worker so that thread start-up is amortised
  const unsigned threads = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  const uint32_t CH = std::max(32u, 2 * threads);
  Pool pool(threads - 1);
  Chunk ch[2];
  for (Chunk& c : ch) { c.depth = (uint16_t*)scn_host_alloc(CH * px * 2); c.rgb = use_color ? (uint8_t*)scn_host_alloc(CH * px * 3) : nullptr;
    if (!c.depth || (use_color && !c.rgb)) { fprintf(stderr, "%s\n", scn_last_error()); return 1; } }
  std::vector<int32_t> lut; if (use_color) build_color_lut(in, lut);
  // Depth decode.  Host pool, or (default for long streams with compressed colour; SCN_FUSE_DECODE=gpu|host overrides) the GPU: the compressed depth payloads of up to 8192 frames at a time
  // are uploaded and inflated in HBM in ONE launch (one warp per frame — a deflate stream is sequential, so the GPU only pays
  // off with thousands of frames in flight; see csrc/inflate.cu) while the host pool decodes the colour of the same frames
  // into a second HBM-resident array; the super-chunk is then fused from device memory.
  const char* dec_env = getenv("SCN_FUSE_DECODE");
  // measured on 1000 frames with JPEG colour: 4.5 k frames/s with the GPU inflate against 3.4 k with everything on the host pool
  // (profiles/r01i_pipeline_demo.json); a depth-only or short stream is faster on the pool
  const bool gpu_default = in.n_frames >= 512 && (in.color_compression == 1 || in.color_compression == 2);
  const bool gpu_decode = in.depth_compression == 1 && (dec_env ? !strcmp(dec_env, "gpu") : gpu_default);
  printf("depth decode: %s\n", gpu_decode ? "GPU inflate (one warp per frame)" : "host thread pool");
  const auto t0 = std::chrono::steady_clock::now();
  int rc = 0; uint64_t f = 0; int cur = 0;
  if (gpu_decode) {
    const uint64_t SUPER = std::min<uint64_t>(8192, in.n_frames);
    uint16_t* d_depth = (uint16_t*)scn_device_alloc((size_t)SUPER * px * 2);
    uint8_t* d_rgb = use_color ? (uint8_t*)scn_device_alloc((size_t)SUPER * px * 3) : nullptr;
    void* dec_stream = nullptr; void* up_stream = nullptr;
    if (!d_depth || (use_color && !d_rgb) || scn_stream_create(&dec_stream) || scn_stream_create(&up_stream)) { fprintf(stderr, "%s\n", scn_last_error()); rc = 1; }
    std::vector<float> poses;
    while (f < in.n_frames && !rc) {
      const uint32_t n = (uint32_t)std::min<uint64_t>(SUPER, in.n_frames - f);
      int dec_rc = 0; std::string dec_err; std::atomic<bool> depth_done{false};
      const auto ts0 = std::chrono::steady_clock::now();
      double t_depth = 0, t_color = 0;
      auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - ts0).count(); };
      std::thread dec([&]() { dec_rc = scn_sens_decode_depth_device(s, f, n, d_depth, dec_stream); if (dec_rc) dec_err = scn_last_error(); t_depth = since(); depth_done.store(true); });
      poses.assign((size_t)n * 16, 0.f);
      uint32_t fused = 0;                                                      // frames of this super-chunk already handed to the integrator
      auto fuse_upto = [&](uint32_t upto) {
        if (upto <= fused) return 0;
        const int r = scn_tsdf_integrate_device(vol, upto - fused, d_depth + (size_t)fused * px, use_color ? d_rgb + (size_t)fused * px * 3 : nullptr,
                                                &poses[(size_t)fused * 16], in.depth_intrinsic);
        fused = upto; return r;
      };
      for (uint32_t c0 = 0; c0 < n && !rc; c0 += CH) {                       // colour (and poses) of this super-chunk, CH frames at a time
        const uint32_t cn = std::min<uint32_t>(CH, n - c0);
        decode_chunk(s, in, lut, use_color, false, f + c0, cn, ch[0], pool);
        if (ch[0].rc) { fprintf(stderr, "%s\n", ch[0].err.c_str()); rc = 1; break; }
        memcpy(&poses[(size_t)c0 * 16], ch[0].poses.data(), (size_t)cn * 64);
        if (use_color && scn_memcpy_h2d(d_rgb + (size_t)c0 * px * 3, ch[0].rgb, (size_t)cn * px * 3, up_stream)) { fprintf(stderr, "%s\n", scn_last_error()); rc = 1; break; }
        // once the depth of the whole super-chunk is in HBM, fusion follows the colour decode chunk by chunk
        if (depth_done.load() && !dec_rc && fuse_upto(c0 + cn)) { fprintf(stderr, "%s\n", scn_last_error()); rc = 1; }
      }
      t_color = since();
      const uint32_t fused_early = fused;
      dec.join();
      if (dec_rc) { fprintf(stderr, "%s\n", dec_err.c_str()); rc = 1; }
      if (!rc && (fuse_upto(n) || scn_tsdf_sync(vol))) { fprintf(stderr, "%s\n", scn_last_error()); rc = 1; }
      if (getenv("SCN_TIMING")) fprintf(stderr, "[timing] super-chunk of %u frames: depth on GPU done at %.3f s, colour pool done at %.3f s (%u frames already fused), all fused at %.3f s\n",
                                        n, t_depth, t_color, fused_early, since());
      f += n;
    }
    scn_device_free(d_depth); scn_device_free(d_rgb); scn_stream_destroy(dec_stream); scn_stream_destroy(up_stream);
  } else {
    if (in.n_frames) decode_chunk(s, in, lut, use_color, true, 0, (uint32_t)std::min<uint64_t>(CH, in.n_frames), ch[0], pool);
    while (f < in.n_frames && !rc) {
      Chunk& c = ch[cur];
      if (c.rc) { fprintf(stderr, "%s\n", c.err.c_str()); rc = 1; break; }
      // the GPU consumes chunk `cur` asynchronously while the host decodes the next one into the other buffer
      if (scn_tsdf_integrate_batch(vol, c.n, c.depth, c.rgb, c.poses.data(), in.depth_intrinsic)) { fprintf(stderr, "%s\n", scn_last_error()); rc = 1; break; }
      f += c.n;
      if (f < in.n_frames) decode_chunk(s, in, lut, use_color, true, f, (uint32_t)std::min<uint64_t>(CH, in.n_frames - f), ch[cur ^ 1], pool);
      if (scn_tsdf_sync(vol)) { fprintf(stderr, "%s\n", scn_last_error()); rc = 1; break; }      // chunk `cur` may be overwritten next round
      cur ^= 1;
    }
  }
  const double fuse_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (!rc) {
    scn_tsdf_stats_t st; scn_tsdf_stats(vol, &st);
    printf("integrated %llu frames (%llu skipped: invalid pose) in %.3f s = %.1f frames/s incl. decode; %llu blocks, %llu voxel updates\n",
           (unsigned long long)st.frames_integrated, (unsigned long long)st.frames_skipped, fuse_s, st.frames_integrated / std::max(fuse_s, 1e-9),
           (unsigned long long)st.blocks_allocated, (unsigned long long)st.voxels_updated);
    float* xyz = nullptr; uint8_t* rgb = nullptr; uint32_t* tri = nullptr; uint64_t nV = 0, nF = 0;
    if (scn_tsdf_extract_mesh(vol, &xyz, &rgb, &tri, &nV, &nF) || scn_mesh_save_ply(out_path.c_str(), xyz, use_color ? rgb : nullptr, nV, tri, nF)) { fprintf(stderr, "%s\n", scn_last_error()); rc = 1; }
    else printf("mesh written to %s with %llu vertices, %llu faces\n", out_path.c_str(), (unsigned long long)nV, (unsigned long long)nF);
    scn_free(xyz); scn_free(rgb); scn_free(tri);
  }
  for (Chunk& c : ch) { scn_host_free(c.depth); scn_host_free(c.rgb); }

  scn_tsdf_destroy(vol); scn_sens_close(s);
  return rc;
}

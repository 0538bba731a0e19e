// `fuse` — the reconstruction stage contract of the reference pipeline in library form:
//   <exe> <params.txt> [<params2.txt>] <file.sens> [out.ply]        (Server/scan_processor.py:27-35,123-138)
//   <exe> --gpus N <params.txt> ... <a.sens> <b.sens> ...           one scene per GPU at a time (Server/process.py:75 serialises
//                                                                   one scan per GPU; here N scans run on N GPUs)
// Reads the .sens stream (SensReader), fuses every frame that has a valid pose into the hashed TSDF volume (poses come from
// the file: camera tracking / bundle adjustment is not part of this path), extracts the surface with marching cubes and
// writes `<base>_vh.ply` in the VCGLIB layout the `segment` stage reads (Server/config/scan_stages.json:33-37).
//
// Decode.  Default ("gpu"): the COMPRESSED payloads are uploaded and decoded in HBM — depth by the inflate kernel
// (csrc/inflate.cu), JPEG colour by the device JPEG decoder (csrc/jpeg_gpu.cu), which samples only the colour pixels the
// depth image needs — in chunks of a few hundred frames, two decoder threads (depth, colour) one chunk ahead of the
// integration, so no raw frame crosses PCIe or touches a host core.  "host" (SCN_FUSE_DECODE=host; automatic for PNG / raw
// colour): a host thread pool inflates / decodes one chunk ahead and the frames are handed over in pinned memory.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "scn_common.h"
#include "stage_pool.h"

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

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

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

void decode_chunk(const scn_sens* s, const scn_sens_info_t& in, const std::vector<int32_t>& lut, bool use_color, uint64_t f0, uint32_t n, Chunk& c, Pool& pool) {
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
      int r = scn_sens_frame_depth_u16(s, f0 + i, c.depth + (size_t)i * px);
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

// ---- one scene on one GPU ------------------------------------------------------------------------------------------------------
struct SceneJob {
  std::string sens_path, out_path;          // out_path empty = no mesh
  scn_tsdf_params params;                   // width / height / depth_shift are taken from the stream
  int device = 0;
  bool verbose = true;
  const char* decode_mode = nullptr;        // "gpu", "host" or null (automatic)
};

// The volume (16 GiB of voxel blocks with the CLI's defaults: cudaMalloc + first clear + cudaFree cost 0.15-0.4 s, as much as
// fusing a 5,578-frame scan; a reset clears only the used blocks in a few ms) is kept per device from one scene to the next - by
// the workers of scn_fuse_many and by consecutive scn_fuse_scene calls of a process alike; scn_release_cached_staging frees it.
struct VolSlot {
  int device = 0;
  scn_tsdf* vol = nullptr; scn_tsdf_params p;
  void release() { if (vol) scn_tsdf_destroy(vol); vol = nullptr; }
};
// The double-buffered frame arrays of the GPU-decode mode (3 GB at 1024-frame chunks with colour) come from a per-device pool
// like the decoders' staging: allocating and freeing them per scene cost up to 0.2 s (cudaFree synchronises the device).
struct FrameBufs {
  int device = 0;
  void* buf[4] = {nullptr, nullptr, nullptr, nullptr}; size_t cap[4] = {0, 0, 0, 0};       // depth x2, colour x2
  void* take(int i, size_t bytes) {
    if (bytes > cap[i]) { scn_device_free(buf[i]); buf[i] = scn_device_alloc(bytes); cap[i] = buf[i] ? bytes : 0; }
    return buf[i];
  }
  void release() { for (int i = 0; i < 4; ++i) { scn_device_free(buf[i]); buf[i] = nullptr; cap[i] = 0; } }
};
scn::StagePool<FrameBufs>& frame_pool() { static auto* p = new scn::StagePool<FrameBufs>(); return *p; }
scn::StagePool<VolSlot>& vol_pool() { static auto* p = new scn::StagePool<VolSlot>(); return *p; }

int fuse_scene(const SceneJob& job, scn_fuse_report_t* rep_out) {
  scn_fuse_report_t rep; memset(&rep, 0, sizeof(rep));
  rep.device = job.device;
  const double t_begin = now_s();
  if (scn_set_device(job.device)) return SCN_ERR_CUDA;
  std::thread warm([dev = job.device]() { scn_set_device(dev); scn_cuda_warmup(); });          // context creation overlaps reading the .sens file
  scn_sens* s = nullptr;
  const int open_rc = scn_sens_open(job.sens_path.c_str(), &s);
  const std::string open_err = open_rc ? scn_last_error() : "";
  warm.join();
  if (open_rc) return scn::fail(open_rc, "%s", open_err.c_str());
  scn_sens_info_t in; scn_sens_info(s, &in);
  scn_tsdf_params p = job.params;
  p.width = in.depth_width; p.height = in.depth_height; p.depth_shift = in.depth_shift;   // integrate at the stream's depth resolution
  if (p.batch_frames == 0) p.batch_frames = 32;
  const bool use_color = in.color_compression <= 2 && in.color_width > 0 && in.color_height > 0;   // raw, PNG, JPEG (sensorData.h:600-616)
  if (job.verbose)
    printf("fusing %s: %llu frames %ux%u, voxel %.4f m, truncation %.3f+%.3f*d\n", job.sens_path.c_str(), (unsigned long long)in.n_frames, in.depth_width,
           in.depth_height, p.voxel_size, p.trunc_base, p.trunc_scale);
  scn_tsdf* vol = nullptr;
  scn::StagePool<VolSlot>::Lease vs(vol_pool());
  if (vs->vol && !memcmp(&vs->p, &p, sizeof(p))) {                                     // same volume layout as the previous scene on this GPU
    vol = vs->vol;
    if (scn_tsdf_reset(vol)) { const std::string e = scn_last_error(); vs->release(); scn_sens_close(s); return scn::fail(SCN_ERR_CUDA, "%s", e.c_str()); }
    rep.volume_reused = 1;
  } else {
    vs->release();
    if (scn_tsdf_create(&p, job.device, &vol)) { const std::string e = scn_last_error(); scn_sens_close(s); return scn::fail(SCN_ERR_CUDA, "%s", e.c_str()); }
    vs->vol = vol; vs->p = p;
  }
  const size_t px = (size_t)in.depth_width * in.depth_height;
  std::vector<int32_t> lut; if (use_color) build_color_lut(in, lut);
  // decode mode
  const bool gpu_able = in.depth_compression == 0 || in.depth_compression == 1;
  const bool gpu_decode = gpu_able && !(job.decode_mode && !strcmp(job.decode_mode, "host"));
  if (job.verbose)
    printf("decode: %s\n", !gpu_decode ? "host thread pool" : (use_color && in.color_compression != 2) ? "GPU inflate (depth), host thread pool (raw / PNG colour)"
                                                                                      : "GPU inflate (depth) + GPU JPEG (colour), compressed payloads uploaded");
  rep.gpu_decode = gpu_decode ? 1 : 0;
  const double t0 = now_s();
  rep.setup_s = t0 - t_begin;
  int rc = 0; std::string err;
  auto fail_here = [&](const char* what) { if (!rc) { rc = 1; err = what && *what ? what : scn_last_error(); } };
  if (gpu_decode) {
    const char* ce = getenv("SCN_FUSE_CHUNK");
    const uint32_t CH = (uint32_t)std::max(16, std::min(4096, ce ? atoi(ce) : 1024));   // frames per decode chunk: enough streams in flight for the one-warp-per-frame decoders
    const uint64_t n_chunks = (in.n_frames + CH - 1) / CH;
    uint16_t* d_depth[2] = {nullptr, nullptr}; uint8_t* d_rgb[2] = {nullptr, nullptr}; int32_t* d_lut = nullptr;
    void* st_d = nullptr; void* st_c = nullptr;
    scn::StagePool<FrameBufs>::Lease fb(frame_pool());
    for (int b = 0; b < 2; ++b) {
      d_depth[b] = (uint16_t*)fb->take(b, (size_t)CH * px * 2);
      if (use_color) d_rgb[b] = (uint8_t*)fb->take(2 + b, (size_t)CH * px * 3);
      if (!d_depth[b] || (use_color && !d_rgb[b])) fail_here(nullptr);
    }
    if (use_color && !rc) { d_lut = (int32_t*)scn_device_alloc(px * 4); if (!d_lut || scn_memcpy_h2d(d_lut, lut.data(), px * 4, nullptr)) fail_here(nullptr); }
    if (!rc && (scn_stream_create(&st_d) || scn_stream_create(&st_c))) fail_here(nullptr);
    rep.buffers_s = now_s() - t0;
    // per chunk: the frames with a valid pose, compacted (a frame without a pose is never decoded: sensorData.h:382)
    struct ChunkPlan { std::vector<uint64_t> frames; std::vector<float> poses; };
    std::mutex m; std::condition_variable cv;
    std::vector<int> depth_done(n_chunks, 0), color_done(n_chunks, 0);            // 1 = ready, -1 = failed
    std::vector<ChunkPlan> plan(n_chunks);
    uint64_t consumed = 0;                                                          // chunks whose buffers are free again
    std::string dec_err; bool stop = false;
    for (uint64_t c = 0; c < n_chunks; ++c) {
      const uint64_t f0 = c * CH, f1 = std::min<uint64_t>(in.n_frames, f0 + CH);
      for (uint64_t f = f0; f < f1; ++f) {
        float T[16]; scn_sens_frame_meta(s, f, T, nullptr, nullptr, nullptr, nullptr);
        if (T[0] == -INFINITY) { ++rep.frames_skipped_pose; continue; }
        plan[c].frames.push_back(f); plan[c].poses.insert(plan[c].poses.end(), T, T + 16);
      }
    }
    std::unique_ptr<Pool> color_pool;
    if (use_color && in.color_compression != 2) color_pool.reset(new Pool(std::max(1u, std::min(32u, std::thread::hardware_concurrency())) - 1));
    uint8_t* h_rgb = nullptr;
    auto decoder = [&](bool is_color) {
      scn_set_device(job.device);
      std::vector<const uint8_t*> src; std::vector<uint64_t> len;
      for (uint64_t c = 0; c < n_chunks; ++c) {
        { std::unique_lock<std::mutex> l(m); cv.wait(l, [&]() { return stop || c < consumed + 2; }); if (stop) return; }
        const ChunkPlan& pl = plan[c];
        const uint32_t n = (uint32_t)pl.frames.size();
        int r = 0; std::string e;
        if (n) {
          src.resize(n); len.resize(n);
          for (uint32_t i = 0; i < n; ++i) {
            const uint8_t* cp = nullptr; const uint8_t* dp = nullptr; uint64_t cb = 0, db = 0;
            scn_sens_frame_payload(s, pl.frames[i], &cp, &dp);
            scn_sens_frame_meta(s, pl.frames[i], nullptr, nullptr, nullptr, &cb, &db);
            src[i] = is_color ? cp : dp; len[i] = is_color ? cb : db;
          }
          const double td0 = now_s();
          if (is_color && in.color_compression != 2) {
            // raw / PNG colour: host pool into a pinned chunk (registered to depth on the way), one upload
            if (!h_rgb) h_rgb = (uint8_t*)scn_host_alloc((size_t)CH * px * 3);
            if (!h_rgb) r = SCN_ERR_CUDA;
            std::atomic<uint32_t> nx{0}; std::atomic<int> prc{0}; std::mutex pem; std::string perr;
            auto work = [&]() {
              static thread_local std::vector<uint8_t> col;
              col.resize((size_t)in.color_width * in.color_height * 3);
              for (;;) {
                const uint32_t i = nx.fetch_add(1);
                if (i >= n || prc.load() || r) break;
                const int q = scn_sens_frame_color_rgb8(s, pl.frames[i], col.data());
                if (q) { std::lock_guard<std::mutex> l(pem); if (!prc.load()) { perr = scn_last_error(); prc.store(q); } break; }
                uint8_t* o = h_rgb + (size_t)i * px * 3;
                for (size_t pp = 0; pp < px; ++pp) { const int32_t qq = lut[pp]; if (qq >= 0) { o[3 * pp] = col[3 * (size_t)qq]; o[3 * pp + 1] = col[3 * (size_t)qq + 1]; o[3 * pp + 2] = col[3 * (size_t)qq + 2]; } else o[3 * pp] = o[3 * pp + 1] = o[3 * pp + 2] = 0; }
              }
            };
            if (!r) { color_pool->run(work); r = prc.load(); if (r) scn::fail(r, "%s", perr.c_str()); }
            if (!r) r = scn_memcpy_h2d(d_rgb[c & 1], h_rgb, (size_t)n * px * 3, st_c);
            if (!r) { std::lock_guard<std::mutex> l(m); rep.color_decode_s += now_s() - td0; }
          } else if (is_color) {
            uint32_t k = 0;
            r = scn_jpeg_decode_batch_device(src.data(), len.data(), n, in.color_width, in.color_height, d_lut, (uint32_t)px, d_rgb[c & 1], st_c, &k);
            double hs = 0, em = 0, cm = 0; scn_jpeg_last_timings(&hs, &em, &cm);
            if (!r) { std::lock_guard<std::mutex> l(m); rep.color_frames_on_device += k; rep.color_decode_s += now_s() - td0;
                      rep.color_host_s += hs; rep.color_entropy_s += em * 1e-3; rep.color_convert_s += cm * 1e-3; }
          } else if (in.depth_compression == 1) {
            r = scn_inflate_batch_device(src.data(), len.data(), n, (uint64_t)px * 2, d_depth[c & 1], st_d);
            double ps = 0, km = 0; scn_inflate_last_timings(&ps, &km, nullptr, nullptr);
            if (!r) { std::lock_guard<std::mutex> l(m); rep.depth_decode_s += now_s() - td0; rep.depth_pack_s += ps; rep.depth_kernel_s += km * 1e-3; }
          } else {
            for (uint32_t i = 0; i < n && !r; ++i) {
              if (len[i] < px * 2) { r = scn::fail(SCN_ERR_FORMAT, "frame %llu: invalid data", (unsigned long long)pl.frames[i]); break; }
              r = scn_memcpy_h2d(d_depth[c & 1] + (size_t)i * px, src[i], px * 2, st_d);
            }
          }
          if (r) e = scn_last_error();
        }
        { std::lock_guard<std::mutex> l(m); (is_color ? color_done : depth_done)[c] = r ? -1 : 1; if (r && dec_err.empty()) dec_err = e; }
        cv.notify_all();
        if (r) return;
      }
    };
    std::thread th_d, th_c;
    if (!rc) { th_d = std::thread(decoder, false); if (use_color) th_c = std::thread(decoder, true); }
    for (uint64_t c = 0; c < n_chunks && !rc; ++c) {
      const double tw0 = now_s();
      {
        std::unique_lock<std::mutex> l(m);
        cv.wait(l, [&]() { return depth_done[c] != 0 && (!use_color || color_done[c] != 0); });
        if (depth_done[c] < 0 || (use_color && color_done[c] < 0)) { fail_here(dec_err.c_str()); break; }
      }
      rep.decode_wait_s += now_s() - tw0;
      const ChunkPlan& pl = plan[c];
      if (!pl.frames.empty()) {
        const double ti0 = now_s();
        if (scn_tsdf_integrate_device(vol, (uint32_t)pl.frames.size(), d_depth[c & 1], use_color ? d_rgb[c & 1] : nullptr, pl.poses.data(), in.depth_intrinsic)) { fail_here(nullptr); break; }
        if (scn_tsdf_sync(vol)) { fail_here(nullptr); break; }                          // the buffers of this chunk may be overwritten now
        rep.integrate_s += now_s() - ti0;
      }
      { std::lock_guard<std::mutex> l(m); consumed = c + 1; }
      cv.notify_all();
    }
    const double tt0 = now_s();
    { std::lock_guard<std::mutex> l(m); stop = true; }
    cv.notify_all();
    if (th_d.joinable()) th_d.join();
    if (th_c.joinable()) th_c.join();
    scn_device_free(d_lut); scn_stream_destroy(st_d); scn_stream_destroy(st_c); scn_host_free(h_rgb);
    rep.teardown_s = now_s() - tt0;
  } else {
    // host decode is the bottleneck of this mode (one frame = ~2 ms inflate + ~1-3 ms JPEG on one core): up to 64 cores,
    // chunks of two frames per worker
    const unsigned threads = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    const uint32_t CH = std::max(32u, 2 * threads);
    Pool pool(threads - 1);
    Chunk ch[2];
    for (Chunk& c : ch) { c.depth = (uint16_t*)scn_host_alloc(CH * px * 2); c.rgb = use_color ? (uint8_t*)scn_host_alloc(CH * px * 3) : nullptr;
      if (!c.depth || (use_color && !c.rgb)) fail_here(nullptr); }
    uint64_t f = 0; int cur = 0;
    if (!rc && in.n_frames) decode_chunk(s, in, lut, use_color, 0, (uint32_t)std::min<uint64_t>(CH, in.n_frames), ch[0], pool);
    while (f < in.n_frames && !rc) {
      Chunk& c = ch[cur];
      if (c.rc) { fail_here(c.err.c_str()); break; }
      // the GPU consumes chunk `cur` asynchronously while the host decodes the next one into the other buffer
      if (scn_tsdf_integrate_batch(vol, c.n, c.depth, c.rgb, c.poses.data(), in.depth_intrinsic)) { fail_here(nullptr); break; }
      f += c.n;
      if (f < in.n_frames) decode_chunk(s, in, lut, use_color, f, (uint32_t)std::min<uint64_t>(CH, in.n_frames - f), ch[cur ^ 1], pool);
      if (scn_tsdf_sync(vol)) { fail_here(nullptr); break; }      // chunk `cur` may be overwritten next round
      cur ^= 1;
    }
    for (Chunk& c : ch) { scn_host_free(c.depth); scn_host_free(c.rgb); }
  }
  rep.fuse_s = now_s() - t0;
  if (!rc) {
    scn_tsdf_stats_t st; scn_tsdf_stats(vol, &st);
    rep.frames_integrated = st.frames_integrated; rep.frames_skipped = st.frames_skipped + rep.frames_skipped_pose;
    rep.blocks_allocated = st.blocks_allocated; rep.voxels_updated = st.voxels_updated;
    if (job.verbose)
      printf("integrated %llu frames (%llu skipped: invalid pose) in %.3f s = %.1f frames/s incl. decode; %llu blocks, %llu voxel updates\n",
             (unsigned long long)rep.frames_integrated, (unsigned long long)rep.frames_skipped, rep.fuse_s, rep.frames_integrated / std::max(rep.fuse_s, 1e-9),
             (unsigned long long)st.blocks_allocated, (unsigned long long)st.voxels_updated);
    size_t used = 0, total = 0;
    if (!scn_device_mem_info(&used, &total)) rep.device_bytes_in_use = used;
    if (!job.out_path.empty()) {
      const double tm0 = now_s();
      float* xyz = nullptr; uint8_t* rgb = nullptr; uint32_t* tri = nullptr; uint64_t nV = 0, nF = 0;
      if (scn_tsdf_extract_mesh(vol, &xyz, &rgb, &tri, &nV, &nF)) fail_here(nullptr);
      rep.mc_s = now_s() - tm0;
      if (!rc && scn_mesh_save_ply(job.out_path.c_str(), xyz, use_color ? rgb : nullptr, nV, tri, nF)) fail_here(nullptr);
      rep.ply_s = now_s() - tm0 - rep.mc_s;
      rep.mesh_vertices = nV; rep.mesh_faces = nF;
      if (!rc && job.verbose) printf("mesh written to %s with %llu vertices, %llu faces\n", job.out_path.c_str(), (unsigned long long)nV, (unsigned long long)nF);
      scn_free(xyz); scn_free(rgb); scn_free(tri);
    }
  }
  if (rc) vs->release();                                       // do not hand a volume of unknown state to the next scene
  scn_sens_close(s);
  rep.total_s = now_s() - t_begin;
  if (rep_out) *rep_out = rep;
  return rc ? scn::fail(SCN_ERR_FORMAT, "%s: %s", job.sens_path.c_str(), err.c_str()) : SCN_OK;
}

std::string default_out(const std::string& sens_path) { return sens_path.substr(0, sens_path.size() - 5) + "_vh.ply"; }

}  // namespace

extern "C" {

void scn_fuse_release_staging_() { frame_pool().trim(); vol_pool().trim(); }
size_t scn_fuse_report_sizeof(void) { return sizeof(scn_fuse_report_t); }

int scn_fuse_scene(const char* sens_path, const char* out_ply, const scn_tsdf_params* params, int device, const char* decode_mode, scn_fuse_report_t* report) {
  if (!sens_path || !params) return scn::fail(SCN_ERR_ARG, "null argument");
  SceneJob j; j.sens_path = sens_path; j.out_path = out_ply ? out_ply : ""; j.params = *params; j.device = device; j.verbose = false; j.decode_mode = decode_mode;
  return fuse_scene(j, report);
}

// One scene per GPU at a time: worker d takes the next scene of the list, fuses it on devices[d], repeats.  No data-path
// collective: the scenes are independent.
int scn_fuse_many(const char* const* sens_paths, const char* const* out_plys, uint32_t n_scenes, const scn_tsdf_params* params,
                  const int* devices, uint32_t n_devices, const char* decode_mode, scn_fuse_report_t* reports) {
  if (!sens_paths || !params || (!devices && n_devices) || !n_devices) return scn::fail(SCN_ERR_ARG, "null argument");
  std::atomic<uint32_t> next{0}; std::atomic<int> rc{0};
  std::mutex em; std::string first_err;
  auto worker = [&](int dev) {
    for (;;) {
      const uint32_t i = next.fetch_add(1);
      if (i >= n_scenes) break;
      SceneJob j; j.sens_path = sens_paths[i]; j.out_path = out_plys && out_plys[i] ? out_plys[i] : ""; j.params = *params; j.device = dev; j.verbose = false;
      j.decode_mode = decode_mode;
      scn_fuse_report_t r; memset(&r, 0, sizeof(r));
      const int e = fuse_scene(j, &r);
      r.status = e;
      if (reports) reports[i] = r;
      if (e) { std::lock_guard<std::mutex> l(em); if (!rc.load()) { first_err = scn_last_error(); rc.store(e); } }
    }
  };
  std::vector<std::thread> th;
  for (uint32_t d = 1; d < n_devices; ++d) th.emplace_back(worker, devices[d]);
  worker(devices[0]);
  for (auto& t : th) t.join();
  return rc.load() ? scn::fail(rc.load(), "%s", first_err.c_str()) : SCN_OK;
}

int scn_fuse_main(int argc, const char** argv) {
  std::vector<std::string> params, scenes; std::string out_path; int gpus = 0;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--gpus" && i + 1 < argc) { gpus = atoi(argv[++i]); continue; }
    if (a.size() > 5 && a.substr(a.size() - 5) == ".sens") scenes.push_back(a);
    else if (!scenes.empty()) out_path = a;
    else params.push_back(a);
  }
  if (scenes.empty()) {
    printf("Usage: fuse <params.txt> [<params2.txt>] <file.sens> [out.ply]\n       fuse --gpus N <params.txt> ... <a.sens> <b.sens> ...\n");
    return 255;
  }
  scn_tsdf_params p; scn_tsdf_default_params(&p);
  p.max_blocks = 1ull << 22; p.hash_slots = 1ull << 24;      // 16 GiB of voxel blocks unless the parameter file says otherwise (s_hashNumSDFBlocks)
  for (const std::string& f : params) if (scn_tsdf_params_from_file(f.c_str(), &p)) { fprintf(stderr, "%s\n", scn_last_error()); return 1; }
  p.batch_frames = 32;
  const char* mode = getenv("SCN_FUSE_DECODE");
  if (scenes.size() == 1 && gpus <= 1) {
    SceneJob j; j.sens_path = scenes[0]; j.out_path = out_path.empty() ? default_out(scenes[0]) : out_path; j.params = p; j.device = 0; j.verbose = true; j.decode_mode = mode;
    scn_fuse_report_t r;
    if (fuse_scene(j, &r)) { fprintf(stderr, "%s\n", scn_last_error()); return 1; }
    if (getenv("SCN_TIMING"))
      fprintf(stderr, "[timing] fuse %.3f s (waiting for the decoders %.3f s; depth decode %.3f s, colour decode %.3f s on their threads), marching cubes %.3f s, PLY %.3f s, total %.3f s, %.2f GB of HBM in use\n",
              r.fuse_s, r.decode_wait_s, r.depth_decode_s, r.color_decode_s, r.mc_s, r.ply_s, r.total_s, r.device_bytes_in_use / 1e9);
    return 0;
  }
  int ndev = scn_device_count();
  if (ndev <= 0) { fprintf(stderr, "%s\n", scn_last_error()); return 1; }
  if (gpus <= 0 || gpus > ndev) gpus = ndev;
  gpus = std::min<int>(gpus, (int)scenes.size());
  std::vector<const char*> sp, op; std::vector<std::string> outs; std::vector<int> devs;
  for (const std::string& s : scenes) outs.push_back(default_out(s));
  for (size_t i = 0; i < scenes.size(); ++i) { sp.push_back(scenes[i].c_str()); op.push_back(outs[i].c_str()); }
  for (int d = 0; d < gpus; ++d) devs.push_back(d);
  std::vector<scn_fuse_report_t> reps(scenes.size());
  const double t0 = now_s();
  const int rc = scn_fuse_many(sp.data(), op.data(), (uint32_t)scenes.size(), &p, devs.data(), (uint32_t)devs.size(), mode, reps.data());
  const double dt = now_s() - t0;
  uint64_t frames = 0;
  for (size_t i = 0; i < scenes.size(); ++i) {
    const scn_fuse_report_t& r = reps[i];
    printf("%s: GPU %d, %llu frames in %.3f s = %.1f frames/s incl. decode, mesh %llu vertices / %llu faces%s\n", scenes[i].c_str(), r.device,
           (unsigned long long)r.frames_integrated, r.fuse_s, r.frames_integrated / std::max(r.fuse_s, 1e-9), (unsigned long long)r.mesh_vertices,
           (unsigned long long)r.mesh_faces, r.status ? "  FAILED" : "");
    frames += r.frames_integrated;
  }
  printf("%zu scenes on %d GPUs: %llu frames in %.3f s wall = %.1f frames/s aggregate (file -> TSDF -> mesh)\n", scenes.size(), gpus, (unsigned long long)frames, dt, frames / std::max(dt, 1e-9));
  if (rc) { fprintf(stderr, "%s\n", scn_last_error()); return 1; }
  return 0;
}

}  // extern "C"

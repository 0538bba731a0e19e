// TEST INFRASTRUCTURE — not product code.
//
// C-callable shim around the UNMODIFIED reference ml::SensorData
// (/root/reference/SensReader/c++/src/sensorData.h, header-only).  Compiled by
// oracle/Makefile with -I pointing at the reference tree where it lies; nothing
// from the reference is copied into this repository.  Only
// oracle/_ref/libref_sens.so travels to the GPU box.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include "sensorData.h"

struct ref_sens { ml::SensorData sd; };

extern "C" {

// sensorData.h:855,1250 — whole file is loaded into RAM.
ref_sens* ref_sens_open(const char* path) {
  try {
    ref_sens* h = new ref_sens();
    h->sd.loadFromFile(std::string(path));
    return h;
  } catch (...) { return nullptr; }
}
void ref_sens_close(ref_sens* h) { delete h; }

// Header fields, sensorData.h:1669-1685.
void ref_sens_info(ref_sens* h, uint32_t* dims /*cw,ch,dw,dh*/, float* depth_shift,
                   int32_t* comp /*color,depth*/, uint64_t* n_frames, uint64_t* n_imu,
                   float* intr_extr /*4x16: colorI,colorE,depthI,depthE*/) {
  ml::SensorData& s = h->sd;
  dims[0] = s.m_colorWidth; dims[1] = s.m_colorHeight; dims[2] = s.m_depthWidth; dims[3] = s.m_depthHeight;
  *depth_shift = s.m_depthShift;
  comp[0] = (int)s.m_colorCompressionType; comp[1] = (int)s.m_depthCompressionType;
  *n_frames = s.m_frames.size(); *n_imu = s.m_IMUFrames.size();
  std::memcpy(intr_extr +  0, &s.m_calibrationColor.m_intrinsic, 64);
  std::memcpy(intr_extr + 16, &s.m_calibrationColor.m_extrinsic, 64);
  std::memcpy(intr_extr + 32, &s.m_calibrationDepth.m_intrinsic, 64);
  std::memcpy(intr_extr + 48, &s.m_calibrationDepth.m_extrinsic, 64);
}
uint64_t ref_sens_name(ref_sens* h, char* out, uint64_t cap) {
  const std::string& n = h->sd.m_sensorName;
  if (n.size() + 1 <= cap) { std::memcpy(out, n.c_str(), n.size() + 1); }
  return n.size();
}
void ref_sens_frame_meta(ref_sens* h, uint64_t i, float* cam2world16, uint64_t* ts_color,
                         uint64_t* ts_depth, uint64_t* color_bytes, uint64_t* depth_bytes) {
  const ml::SensorData::RGBDFrame& f = h->sd.m_frames[i];
  std::memcpy(cam2world16, &f.getCameraToWorld(), 64);
  *ts_color = f.getTimeStampColor(); *ts_depth = f.getTimeStampDepth();
  *color_bytes = f.getColorSizeBytes(); *depth_bytes = f.getDepthSizeBytes();
}
// sensorData.h:939-946 → :693-709 (stb zlib inflate) / :724-730 (raw).
int ref_sens_depth(ref_sens* h, uint64_t i, uint16_t* out) {
  try {
    unsigned short* d = h->sd.decompressDepthAlloc(i);
    if (!d) return -1;
    std::memcpy(out, d, (size_t)h->sd.m_depthWidth * h->sd.m_depthHeight * 2);
    std::free(d);
    return 0;
  } catch (...) { return -2; }
}
// sensorData.h:929-936 → :600-616 (stb JPEG/PNG decode, 3 channels).
int ref_sens_color(ref_sens* h, uint64_t i, uint8_t* out) {
  try {
    ml::vec3uc* c = h->sd.decompressColorAlloc(i);
    if (!c) return -1;
    std::memcpy(out, c, (size_t)h->sd.m_colorWidth * h->sd.m_colorHeight * 3);
    std::free(c);
    return 0;
  } catch (...) { return -2; }
}
// Reference writer: initDefault + addFrame + saveToFile (sensorData.h:888-929,1101-1109).
// colour is stored RAW (the reference cannot JPEG-encode without uplink, :591-593).
int ref_sens_write(const char* path, uint32_t cw, uint32_t ch, uint32_t dw, uint32_t dh,
                   const float* color_intr16, const float* depth_intr16, float depth_shift,
                   int depth_comp, uint64_t n_frames, const uint8_t* colors, const uint16_t* depths,
                   const float* poses16) {
  try {
    ml::SensorData sd;
    ml::mat4f ci, di;
    std::memcpy(&ci, color_intr16, 64); std::memcpy(&di, depth_intr16, 64);
    sd.initDefault(cw, ch, dw, dh, ml::SensorData::CalibrationData(ci), ml::SensorData::CalibrationData(di),
                   ml::SensorData::TYPE_RAW, (ml::SensorData::COMPRESSION_TYPE_DEPTH)depth_comp, depth_shift, "ref_shim");
    for (uint64_t i = 0; i < n_frames; ++i) {
      ml::mat4f p; std::memcpy(&p, poses16 + 16 * i, 64);
      sd.addFrame((const ml::vec3uc*)(colors + (size_t)i * cw * ch * 3), depths + (size_t)i * dw * dh, p, i * 33333, i * 33333);
    }
    sd.saveToFile(std::string(path));
    return 0;
  } catch (...) { return -1; }
}

}  // extern "C"

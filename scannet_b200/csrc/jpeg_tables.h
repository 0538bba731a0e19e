// JPEG Huffman lookup tables shared by the host decoder (jpeg.cpp) and the device decoder (jpeg_gpu.cu).
// Layout and construction follow the lookups of stb_image v2.08 (/root/reference/SensReader/c++/src/sensorData/stb_image.h:
// stbi__build_huffman :1590-1632, stbi__build_fast_ac :1636-1660) so that both decoders take the same path through the same
// bits: a 9-bit lookahead `fast` table, canonical maxcode/mincode/valptr for longer codes, and for AC tables `fast_ac`, which
// folds a short code and a small coefficient into one lookup.  Plain data: a table set is memcpy'd to the device as is.
#pragma once
#include <cstdint>
#include <cstring>

namespace scn_jpeg {

constexpr int kFastBits = 9;

// zig-zag order, padded so that a run past the end of a block lands on coefficient 63 (stb_image.h:1467-1481)
static const uint8_t kZigHost[64 + 15] = {0,1,8,16,9,2,3,10,17,24,32,25,18,11,4,5,12,19,26,33,40,48,41,34,27,20,13,6,7,14,21,28,35,42,49,56,57,50,43,36,
                                          29,22,15,23,30,37,44,51,58,59,52,45,38,31,39,46,53,60,61,54,47,55,62,63,
                                          63,63,63,63,63,63,63,63,63,63,63,63,63,63,63};

struct HuffTab {
  int32_t present;
  int32_t maxcode[18]; int32_t valptr[17]; int32_t mincode[17];
  uint8_t vals[256];
  uint16_t fast[1 << kFastBits];          // (symbol << 4) | length for codes of <= 9 bits, 0xFFFF = longer code
  int16_t fast_ac[1 << kFastBits];        // AC tables: (value << 8) | (run << 4) | (code+magnitude bits) when both fit in 9 bits, else 0

  void build_fast_ac() {
    for (int i = 0; i < (1 << kFastBits); ++i) {
      fast_ac[i] = 0;
      const uint16_t e = fast[i];
      if (e == 0xFFFF) continue;
      const int rs = e >> 4, len = e & 15, run = rs >> 4, mag = rs & 15;
      if (mag && len + mag <= kFastBits) {
        int k = ((i << len) & ((1 << kFastBits) - 1)) >> (kFastBits - mag);
        if (k < (1 << (mag - 1))) k -= (1 << mag) - 1;                       // extend
        if (k >= -128 && k <= 127) fast_ac[i] = (int16_t)((k * 256) + (run * 16) + (len + mag));
      }
    }
  }
  bool build(const uint8_t* counts, const uint8_t* symbols, int nsym) {
    int code = 0, k = 0;
    // validate before any table write (stb_image.h:1590-1615 rejects over-subscribed lengths before building its fast
    // table): an over-subscribed code would index far outside fast[]
    for (int l = 1, c = 0, tot = 0; l <= 16; ++l) {
      c += counts[l - 1]; tot += counts[l - 1];
      if (c > (1 << l) || tot > nsym || tot > 256) return false;
      c <<= 1;
    }
    for (int i = 0; i < (1 << kFastBits); ++i) fast[i] = 0xFFFF;
    memset(fast_ac, 0, sizeof(fast_ac));
    memset(vals, 0, sizeof(vals));
    for (int l = 1; l <= 16; ++l) {
      valptr[l] = k; mincode[l] = code;
      if (l <= kFastBits)
        for (int j = 0; j < counts[l - 1]; ++j) {
          const int c = (code + j) << (kFastBits - l);
          for (int f = 0; f < (1 << (kFastBits - l)); ++f) fast[c + f] = (uint16_t)((symbols[k + j] << 4) | l);
        }
      code += counts[l - 1]; k += counts[l - 1];
      if (code > (1 << l)) return false;
      maxcode[l] = counts[l - 1] ? code - 1 : -1;
      code <<= 1;
    }
    valptr[0] = 0; mincode[0] = 0; maxcode[0] = -1;
    maxcode[17] = 0x7FFFFFFF;
    memcpy(vals, symbols, (size_t)nsym);
    present = 1;
    return true;
  }
};

}  // namespace scn_jpeg

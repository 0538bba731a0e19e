// Device-wide exclusive prefix sum over uint32 (hand-written; three-phase reduce / scan / add,
// recursing on the block sums).  out[i] = sum(in[0..i)), and out[n] = total (out has n+1 slots).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace scn {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned warp_incl_scan(unsigned v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across a 256-thread CTA; returns the exclusive prefix, total in *total
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* total) {
  __shared__ unsigned s_w[kScanThreads / 32];
  __shared__ unsigned s_tot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned inc = warp_incl_scan(v, lane);
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < kScanThreads / 32 ? s_w[lane] : 0u;
    const unsigned wi = warp_incl_scan(w, lane);
    if (lane < kScanThreads / 32) s_w[lane] = wi - w;
    if (lane == kScanThreads / 32 - 1) s_tot = wi;
  }
  __syncthreads();
  const unsigned r = s_w[warp] + inc - v;
  if (total) *total = s_tot;
  __syncthreads();
  return r;
}

static __global__ void __launch_bounds__(kScanThreads)
k_scan_tiles(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned* __restrict__ tile_sums, size_t n) {
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  unsigned v[kScanItems], s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { v[i] = base + i < n ? in[base + i] : 0u; s += v[i]; }
  unsigned tot;
  unsigned ex = block_excl_scan(s, &tot);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

static __global__ void __launch_bounds__(kScanThreads)
k_scan_add(unsigned* __restrict__ out, const unsigned* __restrict__ tile_offs, size_t n, unsigned* total_slot) {
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  const unsigned add = tile_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) if (base + i < n) out[base + i] += add;
  if (total_slot && blockIdx.x == 0 && threadIdx.x == 0) *total_slot = tile_offs[gridDim.x];
}

static __global__ void k_scan_single_total(const unsigned* tile_sums, unsigned* total_slot) { *total_slot = tile_sums[0]; }

// scratch must hold at least scan_scratch_elems(n) uint32.  Returns number of kernels launched.
inline size_t scan_scratch_elems(size_t n) {
  size_t tot = 0;
  while (n > 1) { n = (n + kScanTile - 1) / kScanTile; tot += n + 1; if (n == 1) break; }
  return tot + 8;
}

inline int exclusive_scan_u32(const unsigned* in, unsigned* out /*n+1*/, size_t n, unsigned* scratch, cudaStream_t st) {
  if (n == 0) { cudaMemsetAsync(out, 0, 4, st); return 0; }
  const size_t tiles = (n + kScanTile - 1) / kScanTile;
  unsigned* sums = scratch;                 // tiles (+1) entries
  int launches = 1;
  k_scan_tiles<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, sums, n);
  if (tiles == 1) {
    k_scan_single_total<<<1, 1, 0, st>>>(sums, out + n);
    return launches + 1;
  }
  // scan the tile sums in place (sums[tiles] receives the grand total)
  launches += exclusive_scan_u32(sums, sums, tiles, scratch + tiles + 1, st);
  k_scan_add<<<(unsigned)tiles, kScanThreads, 0, st>>>(out, sums, n, out + n);
  return launches + 1;
}

}  // namespace scn

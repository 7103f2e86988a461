// Pieces of the top-2 merge shared by the matcher's translation unit and the peer exchange:
// the packed candidate key, the merge of key pairs, and the output of one merged row.
#pragma once
#include <cstdint>

namespace sod {

// A candidate as one signed 64-bit key (d2 << 32 | global row): signed order is the (distance, index)
// order of the merge (cv2's tie rule: lowest index).  No entry: kNoneKey.
constexpr long long kNoneKey = 0x7FFFFFFFFFFFFFFFll;

__device__ __forceinline__ longlong2 top2_pack_keys(int2 ci, uint2 cd) {
  longlong2 k = make_longlong2(kNoneKey, kNoneKey);
  if (ci.x >= 0) k.x = (static_cast<long long>(cd.x) << 32) | static_cast<uint32_t>(ci.x);
  if (ci.y >= 0) k.y = (static_cast<long long>(cd.y) << 32) | static_cast<uint32_t>(ci.y);
  return k;
}

// Fold one ascending key pair c (c.x <= c.y) into the running two smallest (k1 <= k2).
__device__ __forceinline__ void top2_fold_keys(long long& k1, long long& k2, longlong2 c) {
  if (c.x < k1) {
    k2 = min(k1, c.y);
    k1 = c.x;
  } else if (c.x < k2) {
    k2 = c.x;
  }
}

// Output of one merged row + the ratio test exactly as the reference evaluates it (main.py:81-82).
__device__ __forceinline__ void write_top2_row(int64_t row, int32_t i1, uint32_t d1, int32_t i2, uint32_t d2,
                                               int32_t* __restrict__ out_idx, uint32_t* __restrict__ out_d2,
                                               float* __restrict__ out_dist, uint8_t* __restrict__ out_pass,
                                               double ratio) {
  out_idx[row * 2 + 0] = i1;
  out_idx[row * 2 + 1] = i2;
  out_d2[row * 2 + 0] = d1;
  out_d2[row * 2 + 1] = d2;
  // OpenCV reports sqrt of the float32 squared distance; d2 < 2^24 converts exactly.
  const float f1 = (i1 >= 0) ? __fsqrt_rn(static_cast<float>(d1)) : __int_as_float(0x7f800000);
  const float f2 = (i2 >= 0) ? __fsqrt_rn(static_cast<float>(d2)) : __int_as_float(0x7f800000);
  if (out_dist) {
    out_dist[row * 2 + 0] = f1;
    out_dist[row * 2 + 1] = f2;
  }
  if (out_pass)
    out_pass[row] = (i2 >= 0 && static_cast<double>(f1) < ratio * static_cast<double>(f2)) ? 1 : 0;
}


// Merged keys of one row -> outputs (ratio test included).
__device__ __forceinline__ void write_top2_keys_row(int64_t row, long long k1, long long k2,
                                                    int32_t* __restrict__ out_idx, uint32_t* __restrict__ out_d2,
                                                    float* __restrict__ out_dist, uint8_t* __restrict__ out_pass,
                                                    double ratio) {
  const bool h1 = k1 != kNoneKey, h2 = k2 != kNoneKey;
  write_top2_row(row, h1 ? static_cast<int32_t>(k1 & 0xFFFFFFFFll) : -1,
                 h1 ? static_cast<uint32_t>(k1 >> 32) : 0xFFFFFFFFu,
                 h2 ? static_cast<int32_t>(k2 & 0xFFFFFFFFll) : -1,
                 h2 ? static_cast<uint32_t>(k2 >> 32) : 0xFFFFFFFFu, out_idx, out_d2, out_dist, out_pass, ratio);
}

}  // namespace sod

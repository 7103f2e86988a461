// K0-K3: descriptor packing, per-row norms, the tcgen05 2-NN matcher and the top-2 merge.
//
// Replaces cv2.BFMatcher().knnMatch(des_query, des, k=2) + the ratio loop of the reference
// (main.py:70-86).  |q-t|^2 = |q|^2 + |t|^2 - 2 q.t with u8 operands is exact in int32, so the
// contraction q.t runs on the 5th-gen tensor cores (tcgen05.mma kind::i8, u8 x u8 -> s32 in TMEM)
// and the top-2 selection happens in the TMEM epilogue without ever writing distances to HBM.
//
// Kernel anatomy (one persistent CTA per SM, 384 threads):
//   warp 0      TMA producer: query block (2 x 128 rows, double buffered) + database tiles
//               (128 rows = 16 KB, kStages-deep ring) + the 512 B slice of per-row constants cq.
//   warp 1      MMA issuer (one thread): per database tile 2 x 4 UTCIMMA (M=128,N=128,K=32) into a
//               double-buffered TMEM accumulator (2 halves x 2 buffers x 128 columns = 512 columns).
//   warp 2      TMEM allocator.
//   warps 4-11  epilogue: thread <-> query row (TMEM lane), sweeps the 128 columns of the tile and
//               keeps a running top-2 of the packed key  ((|t|^2 - 2 q.t) << 8) | column.
//
// Packed key.  cq[n] = (|t_n|^2 << 8) | (n & 127) is prepared once per database.  With
// Q = |q|^2 << 8,  key' = cq - 512*acc  equals  ((d2 << 8) | col) - Q  (mod 2^32).  d2 <= 128*255^2
// < 2^23 so (d2<<8|col) < 2^31 and Q < 2^31: key' never overflows int32 and the signed order of
// key' is the (d2, col) lexicographic order -> one IMAD per element, ties resolve to the lowest
// column.  Across tiles the column byte is cleared before a new tile is compared so that equal
// distances keep the EARLIER tile (lowest database index, as cv2 does).
#include <cudaTypedefs.h>

#include <climits>

#include "sod_common.cuh"
#include "sod_ptx.cuh"

#ifndef SOD_EXP
#define SOD_EXP 0  // kernel experiments (timing only, results invalid): 1 = no epilogue math, 2 = no TMEM loads
#endif

namespace sod {
namespace {

constexpr int kTileM = 128;                 // query rows per MMA (TMEM lanes)
constexpr int kHalves = 2;                  // query tiles per CTA sharing one database tile
constexpr int kBlockQ = kTileM * kHalves;   // 256 query rows per unit
constexpr int kTileN = SOD_TILE_ROWS;       // database rows per MMA
constexpr int kTileBytes = kTileN * SOD_DESC_DIM;  // 16 KB (A half-tile has the same size)
constexpr int kStages = 6;
constexpr int kCqSlots = kStages + 2;       // see the slot-reuse argument in the producer
constexpr int kChunk = 32;                  // accumulator columns per tcgen05.ld
constexpr int kCqTile = SOD_CQ_TILE_INTS;   // 128 packed keys + 4 per-chunk minima of |t|^2
constexpr int kCqTileBytes = kCqTile * 4;   // 528 B, a multiple of 16 for the bulk copy
constexpr int kInitKey = INT_MAX & ~0xFF;   // "no candidate yet", column byte cleared
constexpr int kEpiWarps = 8;
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int kTmemCols = 512;

constexpr int kOffA = 0;                                    // [2 buffers][2 halves][16 KB]
constexpr int kOffB = kOffA + 2 * kHalves * kTileBytes;     // [kStages][16 KB]
constexpr int kOffCq = kOffB + kStages * kTileBytes;        // [kCqSlots][128] int32
constexpr int kOffBar = kOffCq + kCqSlots * kCqTileBytes;
constexpr int kNumBars = 2 * kStages + 8 + kCqSlots;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;         // +1024: manual 1 KB alignment

struct MatchArgs {
  const int32_t* qn;   // [nq] |q|^2
  const int32_t* cq;   // [n_tiles][132] packed per-row constants + chunk minima
  uint32_t* part_d2;   // [n_seg][nq][2]
  int32_t* part_idx;   // [n_seg][nq][2]
  int nq;
  int n_tiles;
  int n_qblocks;
  int n_seg;
  int idx_base;
};

// One chunk of 32 accumulator columns -> running top-2, in three levels of increasing cost:
//  1. bound: every key of the chunk is >= (cmin - 2*max(acc)) << 8 where cmin is the smallest
//     |t|^2 of the chunk; if that already exceeds the thread's 2nd best nothing can change
//     (~0.5 ALU op per element: a 3-input max tree on the raw accumulators);
//  2. exact: compute the 32 packed keys and their minimum (1 IMAD + 0.5 min3 per element);
//  3. update: the plain running top-2 (3 ops per element) - after the first few thousand columns
//     of a sweep only a few percent of the chunks get here.
// All comparisons are exact, so pruning never changes the result.
__device__ __forceinline__ void top2_chunk(const uint32_t (&v)[32], const int4* __restrict__ cq4,
                                           int cmin, int& m1, int& m2, bool& touched) {
#if SOD_EXP == 1 || SOD_EXP == 2
  return;
#endif
  int amax = static_cast<int>(v[0]);
#pragma unroll
  for (int j = 1; j < 32; ++j) amax = max(amax, static_cast<int>(v[j]));
  // (cmin - 2*amax) is in d2 units relative to |q|^2; m2 >> 8 is the same quantity of the 2nd best.
  if (cmin - 2 * amax > (m2 >> 8)) return;
  int k[32];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int4 c = cq4[j];
    k[4 * j + 0] = c.x - 512 * static_cast<int>(v[4 * j + 0]);
    k[4 * j + 1] = c.y - 512 * static_cast<int>(v[4 * j + 1]);
    k[4 * j + 2] = c.z - 512 * static_cast<int>(v[4 * j + 2]);
    k[4 * j + 3] = c.w - 512 * static_cast<int>(v[4 * j + 3]);
  }
  int kmin = k[0];
#pragma unroll
  for (int j = 1; j < 32; ++j) kmin = min(kmin, k[j]);
  if (kmin >= m2) return;
  touched = true;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int t = max(m1, k[j]);
    m1 = min(m1, k[j]);
    m2 = min(m2, t);
  }
}

// tcgen05.wait::ld that also names the destination registers, so no use of them can be scheduled
// above the wait.
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]),
                 "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]),
                 "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]),
                 "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),
                 "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
match_top2_kernel(const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_db, const MatchArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1 KB alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bar0 = base + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto bar_afull = [&](int b) { return bar0 + 8u * (2 * kStages + b); };
  auto bar_aempty = [&](int b) { return bar0 + 8u * (2 * kStages + 2 + b); };
  auto bar_tfull = [&](int b) { return bar0 + 8u * (2 * kStages + 4 + b); };
  auto bar_tempty = [&](int b) { return bar0 + 8u * (2 * kStages + 6 + b); };
  auto bar_cqfull = [&](int s) { return bar0 + 8u * (2 * kStages + 8 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_db);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_afull(b), 1);
      mbar_init(bar_aempty(b), 1);
      mbar_init(bar_tfull(b), 1);
      mbar_init(bar_tempty(b), kEpiWarps);
    }
    for (int s = 0; s < kCqSlots; ++s) mbar_init(bar_cqfull(s), 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total_units = a.n_qblocks * a.n_seg;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t step = 0, ucount = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++ucount) {
        const int seg = u / a.n_qblocks;
        const int qb = u - seg * a.n_qblocks;
        const int t0 = static_cast<int>(static_cast<int64_t>(seg) * a.n_tiles / a.n_seg);
        const int t1 = static_cast<int>(static_cast<int64_t>(seg + 1) * a.n_tiles / a.n_seg);
        const uint32_t ab = ucount & 1u, aph = (ucount >> 1) & 1u;
        mbar_wait(bar_aempty(ab), aph ^ 1u);
        mbar_arrive_expect_tx(bar_afull(ab), kHalves * kTileBytes);
        for (int h = 0; h < kHalves; ++h)
          tma_load_2d(base + kOffA + (ab * kHalves + h) * kTileBytes, &tmap_q, bar_afull(ab), 0,
                      qb * kBlockQ + h * kTileM);
        for (int t = t0; t < t1; ++t, ++step) {
          const uint32_t s = step % kStages, ph = (step / kStages) & 1u;
          mbar_wait(bar_empty(s), ph ^ 1u);
          mbar_arrive_expect_tx(bar_full(s), kTileBytes);
          tma_load_2d(base + kOffB + s * kTileBytes, &tmap_db, bar_full(s), 0, t * kTileN);
          // cq slice rides in its own, deeper ring: stage s is released when the MMA that read it
          // retires, but the epilogue reads cq up to two accumulator buffers later.  The write of
          // step j happens after MMA(j-kStages) retired, hence after the epilogue finished step
          // j-kStages-2; steps j-kStages-1 .. j-1 may still be live -> kStages+2 slots suffice.
          const uint32_t slot = step % kCqSlots;
          mbar_arrive_expect_tx(bar_cqfull(slot), kCqTileBytes);
          bulk_load_1d(base + kOffCq + slot * kCqTileBytes, a.cq + static_cast<int64_t>(t) * kCqTile,
                       kCqTileBytes, bar_cqfull(slot));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_u8(kTileM, kTileN);
      uint32_t step = 0, ucount = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++ucount) {
        const int seg = u / a.n_qblocks;
        const int t0 = static_cast<int>(static_cast<int64_t>(seg) * a.n_tiles / a.n_seg);
        const int t1 = static_cast<int>(static_cast<int64_t>(seg + 1) * a.n_tiles / a.n_seg);
        const uint32_t ab = ucount & 1u, aph = (ucount >> 1) & 1u;
        mbar_wait(bar_afull(ab), aph);
        const uint64_t adesc0 = umma_desc_k128(base + kOffA + (ab * kHalves + 0) * kTileBytes);
        const uint64_t adesc1 = umma_desc_k128(base + kOffA + (ab * kHalves + 1) * kTileBytes);
        for (int t = t0; t < t1; ++t, ++step) {
          const uint32_t s = step % kStages, ph = (step / kStages) & 1u;
          const uint32_t acc = step & 1u, accph = (step >> 1) & 1u;
          mbar_wait(bar_tempty(acc), accph ^ 1u);
          mbar_wait(bar_full(s), ph);
          tc_fence_after();
          const uint64_t bdesc = umma_desc_k128(base + kOffB + s * kTileBytes);
          const uint32_t d0 = tmem_base + acc * (kHalves * kTileN);
#pragma unroll
          for (int k = 0; k < SOD_DESC_DIM / 32; ++k)  // K = 32 bytes per UTCIMMA: +2 x 16 B
            umma_i8(d0, adesc0 + 2 * k, bdesc + 2 * k, idesc, k > 0);
#pragma unroll
          for (int k = 0; k < SOD_DESC_DIM / 32; ++k)
            umma_i8(d0 + kTileN, adesc1 + 2 * k, bdesc + 2 * k, idesc, k > 0);
          umma_commit(bar_empty(s));    // database stage free once these MMAs retire
          umma_commit(bar_tfull(acc));  // accumulator ready for the epilogue
        }
        umma_commit(bar_aempty(ab));  // query block buffer free
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int e = warp - 4;
    const int h = e >> 2;       // which query half-tile
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    uint32_t step = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int seg = u / a.n_qblocks;
      const int qb = u - seg * a.n_qblocks;
      const int t0 = static_cast<int>(static_cast<int64_t>(seg) * a.n_tiles / a.n_seg);
      const int t1 = static_cast<int>(static_cast<int64_t>(seg + 1) * a.n_tiles / a.n_seg);
      const int row = qb * kBlockQ + h * kTileM + quad * 32 + lane;
      // Between tiles m1/m2 keep their column byte cleared: an equal distance in a later tile
      // then never displaces an earlier one (lowest database index wins, as in cv2).
      int m1 = kInitKey, m2 = kInitKey, i1 = -1, i2 = -1;
      for (int t = t0; t < t1; ++t, ++step) {
        const uint32_t acc = step & 1u, accph = (step >> 1) & 1u;
        const uint32_t slot = step % kCqSlots, cqph = (step / kCqSlots) & 1u;
        mbar_wait(bar_tfull(acc), accph);
        mbar_wait(bar_cqfull(slot), cqph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_sel + acc * (kHalves * kTileN) + h * kTileN;
        const int4* cq4 = reinterpret_cast<const int4*>(smem + kOffCq + slot * kCqTileBytes);
        const int4 cmin = cq4[kTileN / 4];  // per-chunk minima of |t|^2
        const int s1 = m1, s2 = m2;
        bool touched = false;
        uint32_t va[32], vb[32];
#if SOD_EXP == 2
        for (int j = 0; j < 32; ++j) va[j] = vb[j] = 0;
#define tmem_ld32(a, b)
#define tmem_ld_wait_on(a)
#endif
        tmem_ld32(taddr, va);
        tmem_ld_wait_on(va);
        tmem_ld32(taddr + kChunk, vb);
        top2_chunk(va, cq4, cmin.x, m1, m2, touched);
        tmem_ld_wait_on(vb);
        tmem_ld32(taddr + 2 * kChunk, va);
        top2_chunk(vb, cq4 + 8, cmin.y, m1, m2, touched);
        tmem_ld_wait_on(va);
        tmem_ld32(taddr + 3 * kChunk, vb);
        top2_chunk(va, cq4 + 16, cmin.z, m1, m2, touched);
        tmem_ld_wait_on(vb);
        // Every TMEM read of this accumulator buffer has landed: hand it back to the MMA warp.
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty(acc));
        top2_chunk(vb, cq4 + 24, cmin.w, m1, m2, touched);
        if (touched) {
          // Recover database indices of the entries that changed in this tile.
          const int tile_base = a.idx_base + t * kTileN;
          if (m1 != s1) {
            i2 = (m2 == s1) ? i1 : tile_base + (m2 & 0xFF);
            i1 = tile_base + (m1 & 0xFF);
          } else if (m2 != s2) {
            i2 = tile_base + (m2 & 0xFF);
          }
          m1 &= ~0xFF;
          m2 &= ~0xFF;
        }
      }
      if (row < a.nq) {
        const uint32_t Q = static_cast<uint32_t>(a.qn[row]) << 8;
        const int64_t o = (static_cast<int64_t>(seg) * a.nq + row) * 2;
        a.part_d2[o + 0] = (i1 >= 0) ? (static_cast<uint32_t>(m1) + Q) >> 8 : 0xFFFFFFFFu;
        a.part_d2[o + 1] = (i2 >= 0) ? (static_cast<uint32_t>(m2) + Q) >> 8 : 0xFFFFFFFFu;
        a.part_idx[o + 0] = i1;
        a.part_idx[o + 1] = i2;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// K1 (query side): one warp per descriptor row, 4 bytes per lane, __dp4a for the squares.
__global__ void row_sqnorm_kernel(const uint8_t* __restrict__ x, int64_t n_rows,
                                  int32_t* __restrict__ out) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const uint32_t w = reinterpret_cast<const uint32_t*>(x + row * SOD_DESC_DIM)[lane];
  uint32_t s = __dp4a(w, w, 0u);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = static_cast<int32_t>(s);
}

// K1 (database side): one CTA of 128 threads per tile of 128 rows, one row per thread (a full
// 128-byte line each).  Writes the tile's 128 packed keys (|t|^2 << 8 | column; INT_MAX for rows
// past the end, which can never win) and the minimum |t|^2 of each 32-row chunk.
__global__ void __launch_bounds__(kTileN)
db_prepare_kernel(const uint8_t* __restrict__ db, int64_t n_rows, int32_t* __restrict__ cq) {
  const int64_t tile = blockIdx.x;
  const int col = threadIdx.x;
  const int64_t row = tile * kTileN + col;
  int32_t key = INT_MAX, c = 0x3FFFFFFF;
  if (row < n_rows) {
    const uint4* p = reinterpret_cast<const uint4*>(db + row * SOD_DESC_DIM);
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SOD_DESC_DIM / 16; ++j) {
      const uint4 w = __ldg(p + j);
      s = __dp4a(w.x, w.x, s);
      s = __dp4a(w.y, w.y, s);
      s = __dp4a(w.z, w.z, s);
      s = __dp4a(w.w, w.w, s);
    }
    c = static_cast<int32_t>(s);
    key = static_cast<int32_t>((s << 8) | static_cast<uint32_t>(col));
  }
  int32_t* out = cq + tile * kCqTile;
  out[col] = key;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c = min(c, __shfl_xor_sync(0xffffffffu, c, o));
  if ((col & 31) == 0) out[kTileN + (col >> 5)] = c;
}

// K0: float32 -> u8 with an integrality check.
__global__ void pack_u8_kernel(const float4* __restrict__ src, int64_t n_vec4,
                               uchar4* __restrict__ dst, int32_t* __restrict__ nonint_flag) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_vec4) return;
  const float4 f = src[i];
  const float r[4] = {f.x, f.y, f.z, f.w};
  unsigned char b[4];
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float c = fminf(fmaxf(r[k], 0.f), 255.f);
    const float q = rintf(c);
    bad |= !(q == r[k]);
    b[k] = static_cast<unsigned char>(q);
  }
  dst[i] = make_uchar4(b[0], b[1], b[2], b[3]);
  if (bad) atomicOr(nonint_flag, 1);
}

// K3: merge candidate lists, exact ratio test.
__global__ void top2_merge_kernel(const int32_t* __restrict__ parts_idx,
                                  const uint32_t* __restrict__ parts_d2, int n_parts, int64_t nq,
                                  int32_t* __restrict__ out_idx, uint32_t* __restrict__ out_d2,
                                  float* __restrict__ out_dist, uint8_t* __restrict__ out_pass,
                                  double ratio) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= nq) return;
  uint32_t d1 = 0xFFFFFFFFu, d2 = 0xFFFFFFFFu;
  int32_t i1 = -1, i2 = -1;
  for (int p = 0; p < n_parts; ++p) {
    const int64_t o = (static_cast<int64_t>(p) * nq + row) * 2;
    const int2 ci = *reinterpret_cast<const int2*>(parts_idx + o);
    const uint2 cd = *reinterpret_cast<const uint2*>(parts_d2 + o);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int32_t ci_j = j ? ci.y : ci.x;
      const uint32_t cd_j = j ? cd.y : cd.x;
      if (ci_j < 0) continue;
      const bool lt1 = i1 < 0 || cd_j < d1 || (cd_j == d1 && ci_j < i1);
      const bool lt2 = i2 < 0 || cd_j < d2 || (cd_j == d2 && ci_j < i2);
      if (lt1) {
        d2 = d1; i2 = i1;
        d1 = cd_j; i1 = ci_j;
      } else if (lt2) {
        d2 = cd_j; i2 = ci_j;
      }
    }
  }
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

struct Plan {
  int n_tiles, n_qblocks, n_seg, grid;
};

// Split every query block's database sweep into n_seg contiguous segments so that
// n_qblocks*n_seg units fill the persistent grid evenly (cost = makespan in tile steps).
Plan make_plan(int64_t nq, int64_t ndb, int sms) {
  Plan p;
  p.n_tiles = static_cast<int>((ndb + kTileN - 1) / kTileN);
  p.n_qblocks = static_cast<int>((nq + kBlockQ - 1) / kBlockQ);
  p.n_seg = 1;
  if (p.n_tiles > 0 && p.n_qblocks > 0) {
    const int max_seg = p.n_tiles < 512 ? p.n_tiles : 512;
    int64_t best = -1;
    for (int s = 1; s <= max_seg; ++s) {
      const int64_t units = static_cast<int64_t>(p.n_qblocks) * s;
      const int64_t waves = (units + sms - 1) / sms;
      const int64_t cost = waves * ((p.n_tiles + s - 1) / s + 3);
      if (best < 0 || cost < best) {
        best = cost;
        p.n_seg = s;
      }
    }
  }
  const int64_t units = static_cast<int64_t>(p.n_qblocks) * p.n_seg;
  p.grid = static_cast<int>(units < sms ? units : sms);
  return p;
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

// [n_rows,128] u8 row-major, boxes of 128 rows x 128 B, 128-byte swizzle, OOB rows read as zero.
int make_desc_map(CUtensorMap* m, const uint8_t* ptr, int64_t n_rows) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SOD_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {SOD_DESC_DIM, static_cast<cuuint64_t>(n_rows)};
  const cuuint64_t strides[1] = {SOD_DESC_DIM};
  const cuuint32_t box[2] = {SOD_DESC_DIM, kTileN};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(ptr), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return SOD_ERR_CUDA;
  }
  return SOD_OK;
}

}  // namespace
}  // namespace sod

using namespace sod;

extern "C" {

int64_t sod_cq_ints(int64_t n_rows) {
  return n_rows <= 0 ? 0 : (n_rows + kTileN - 1) / kTileN * kCqTile;
}

int sod_pack_u8_from_f32(const float* src, int64_t n_rows, uint8_t* dst, int32_t* nonint_flag,
                         sod_stream_t stream) {
  SOD_CHECK_ARG(n_rows >= 0, "n_rows < 0");
  if (n_rows == 0) return SOD_OK;
  SOD_CHECK_ARG(src && dst && nonint_flag, "null pointer");
  const int64_t n4 = n_rows * (SOD_DESC_DIM / 4);
  const int threads = 256;
  pack_u8_kernel<<<static_cast<unsigned>((n4 + threads - 1) / threads), threads, 0,
                   static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(src), n4,
                                                        reinterpret_cast<uchar4*>(dst), nonint_flag);
  SOD_CHECK_LAUNCH("pack_u8_kernel");
  return SOD_OK;
}

int sod_db_prepare(const uint8_t* db, int64_t n_rows, int32_t* cq, sod_stream_t stream) {
  SOD_CHECK_ARG(n_rows >= 0 && n_rows < (int64_t(1) << 31) - kTileN, "n_rows out of range");
  if (n_rows == 0) return SOD_OK;
  SOD_CHECK_ARG(db && cq, "null pointer");
  SOD_CHECK_ARG((reinterpret_cast<uintptr_t>(db) & 15) == 0, "db must be 16-byte aligned");
  const int64_t tiles = (n_rows + kTileN - 1) / kTileN;
  db_prepare_kernel<<<static_cast<unsigned>(tiles), kTileN, 0, static_cast<cudaStream_t>(stream)>>>(
      db, n_rows, cq);
  SOD_CHECK_LAUNCH("db_prepare_kernel");
  return SOD_OK;
}

int sod_query_prepare(const uint8_t* q, int64_t n_rows, int32_t* qn, sod_stream_t stream) {
  SOD_CHECK_ARG(n_rows >= 0 && n_rows < (int64_t(1) << 31) - kBlockQ, "n_rows out of range");
  if (n_rows == 0) return SOD_OK;
  SOD_CHECK_ARG(q && qn, "null pointer");
  const int threads = 256;
  const int64_t blocks = (n_rows * 32 + threads - 1) / threads;
  row_sqnorm_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      q, n_rows, qn);
  SOD_CHECK_LAUNCH("row_sqnorm_kernel");
  return SOD_OK;
}

size_t sod_match_workspace_bytes(int64_t n_query, int64_t n_db) {
  if (n_query <= 0 || n_db <= 0) return 16;
  const int sms = device_sm_count();
  const Plan p = make_plan(n_query, n_db, sms > 0 ? sms : 148);
  return static_cast<size_t>(p.n_seg) * static_cast<size_t>(n_query) * 2 * 8 + 16;
}

int sod_match_top2(const uint8_t* q, const int32_t* qn, int64_t n_query, const uint8_t* db,
                   const int32_t* cq, int64_t n_db, int32_t db_index_base, int32_t* out_idx,
                   uint32_t* out_d2, void* workspace, size_t workspace_bytes, sod_stream_t stream) {
  SOD_CHECK_ARG(n_query >= 0 && n_db >= 0, "negative size");
  SOD_CHECK_ARG(n_query < (int64_t(1) << 31) - kBlockQ && n_db < (int64_t(1) << 31) - kTileN,
                "size out of range");
  SOD_CHECK_ARG(static_cast<int64_t>(db_index_base) + n_db < (int64_t(1) << 31),
                "db_index_base + n_db overflows int32");
  if (n_query == 0) return SOD_OK;
  SOD_CHECK_ARG(out_idx && out_d2, "null output pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int mthreads = 128;
  const unsigned mblocks = static_cast<unsigned>((n_query + mthreads - 1) / mthreads);
  if (n_db == 0) {
    top2_merge_kernel<<<mblocks, mthreads, 0, st>>>(nullptr, nullptr, 0, n_query, out_idx, out_d2,
                                                    nullptr, nullptr, 0.0);
    SOD_CHECK_LAUNCH("top2_merge_kernel");
    return SOD_OK;
  }
  SOD_CHECK_ARG(q && qn && db && cq && workspace, "null pointer");
  SOD_CHECK_ARG((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(db) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(cq) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "q, db, cq and workspace must be 16-byte aligned");
  const int sms = device_sm_count();
  if (sms <= 0) return SOD_ERR_CUDA;
  const Plan p = make_plan(n_query, n_db, sms);
  const size_t need = static_cast<size_t>(p.n_seg) * static_cast<size_t>(n_query) * 2 * 8;
  SOD_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);

  CUtensorMap map_q, map_db;
  int rc = make_desc_map(&map_q, q, n_query);
  if (rc != SOD_OK) return rc;
  rc = make_desc_map(&map_db, db, n_db);
  if (rc != SOD_OK) return rc;

  MatchArgs a;
  a.qn = qn;
  a.cq = cq;
  a.part_d2 = static_cast<uint32_t*>(workspace);
  a.part_idx = reinterpret_cast<int32_t*>(a.part_d2 + static_cast<size_t>(p.n_seg) * n_query * 2);
  a.nq = static_cast<int>(n_query);
  a.n_tiles = p.n_tiles;
  a.n_qblocks = p.n_qblocks;
  a.n_seg = p.n_seg;
  a.idx_base = db_index_base;

  static bool attr_set = false;
  if (!attr_set) {
    SOD_CHECK_CUDA(cudaFuncSetAttribute(match_top2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kSmemBytes));
    attr_set = true;
  }
  match_top2_kernel<<<p.grid, kThreads, kSmemBytes, st>>>(map_q, map_db, a);
  SOD_CHECK_LAUNCH("match_top2_kernel");
  top2_merge_kernel<<<mblocks, mthreads, 0, st>>>(a.part_idx, a.part_d2, p.n_seg, n_query, out_idx,
                                                  out_d2, nullptr, nullptr, 0.0);
  SOD_CHECK_LAUNCH("top2_merge_kernel");
  return SOD_OK;
}

int sod_top2_merge(const int32_t* parts_idx, const uint32_t* parts_d2, int32_t n_parts,
                   int64_t n_query, int32_t* out_idx, uint32_t* out_d2, float* out_dist,
                   uint8_t* out_pass, double ratio, sod_stream_t stream) {
  SOD_CHECK_ARG(n_parts >= 0 && n_query >= 0, "negative size");
  if (n_query == 0) return SOD_OK;
  SOD_CHECK_ARG(out_idx && out_d2, "null output pointer");
  SOD_CHECK_ARG(n_parts == 0 || (parts_idx && parts_d2), "null parts pointer");
  const int threads = 128;
  top2_merge_kernel<<<static_cast<unsigned>((n_query + threads - 1) / threads), threads, 0,
                      static_cast<cudaStream_t>(stream)>>>(parts_idx, parts_d2, n_parts, n_query,
                                                           out_idx, out_d2, out_dist, out_pass, ratio);
  SOD_CHECK_LAUNCH("top2_merge_kernel");
  return SOD_OK;
}

}  // extern "C"

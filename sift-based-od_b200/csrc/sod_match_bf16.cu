// 2-NN matching for descriptors that are NOT integer-valued 0..255 (BASELINE north star: "falls
// back to a bf16 path with a stated tolerance only for non-integer descriptors").  Replaces the same
// cv2.BFMatcher().knnMatch(des_query, des, k=2) call as sod_match.cu (reference main.py:72-73), but
// its distances are approximate:
//
//     d^2(q,t) = |q|^2 + |t|^2 - 2 q.t      norms in fp32 from the fp32 input (summed in fp64),
//                                           q.t on tcgen05 kind::f16 (bf16 x bf16 -> f32, UTCHMMA)
//
// Operands (built by sod_bf16_prepare):
//   split = 0   q -> bf16(q), t -> -2 * bf16(t)                                     K = 128
//   split = 1   x = hi + lo with hi = bf16(x), lo = bf16(x - hi); the kernel multiplies
//               [qh | qh | ql] by -2 * [th | tl | th], i.e. q.t ~ qh.th + qh.tl + ql.th   K = 384
// Stated tolerance on d^2, relative to |q|^2 + |t|^2:  split 0: 4e-3,  split 1: 2e-5  (tests hold the
// kernel to these); indices are those of the two smallest APPROXIMATE distances, ties -> lowest index.
//
// Kernel: one CTA per (128-query block, database segment).  Warp 0 = TMA producer (query block once,
// then 16 KB database K-blocks through a 6-stage ring, plus the tile's 128 fp32 |t|^2), warp 1 = MMA
// issuer (two 128x128 f32 accumulators in TMEM), warp 2 = TMEM allocator, warps 4-11 = epilogue: lane
// quadrant x column half; per 32-column chunk one FADD per element builds key = |t|^2 - 2 q.t, a
// min tree compares the chunk with the thread's 2nd best and only improving chunks are scanned.
// No pruning tables and no sorted database: this is the rarely used path, kept simple.
#include <cuda_bf16.h>

#include <cmath>

#include "sod_common.cuh"
#include "sod_ptx.cuh"
#include "sod_tma.cuh"

namespace sod {
namespace {

constexpr int kTile = 128;                     // query rows per CTA = database rows per tile
constexpr int kKBlock = 64;                    // bf16 elements per 128-byte swizzle row
constexpr int kBlockBytes = kTile * 128;       // one K-block of one tile: 16 KB
constexpr int kStages = 6;
constexpr int kTnSlots = 8;                    // see the reuse argument in the producer
constexpr int kTnBytes = kTile * 4;
constexpr int kMaxKBlocks = 6;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int kTmemCols = 256;                 // two 128-column f32 accumulators

constexpr int kOffB = 0;
constexpr int kOffTn = kOffB + kStages * kBlockBytes;
constexpr int kOffBar = kOffTn + kTnSlots * kTnBytes;
constexpr int kNumBars = 2 * kStages + 1 + 4 + kTnSlots;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kOffA = (kOffTmemPtr + 16 + 1023) / 1024 * 1024;  // [k_blocks][16 KB], sized at launch
constexpr int smem_bytes(int k_blocks) { return kOffA + k_blocks * kBlockBytes + 1024; }

struct Bf16Args {
  const float* qn;      // [nq]
  const float* dn;      // [n_tiles * 128], +inf on padding rows
  float* part_d2;       // [n_seg * 2][nq][2]
  int32_t* part_idx;    // [n_seg * 2][nq][2]
  int nq;
  int n_tiles;
  int n_seg;
  int k_blocks;         // 2 (plain) or 6 (split)
  int idx_base;
};

__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float min_tree32(const float* v) {
  float t[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) t[i] = fmin3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  t[10] = fminf(v[30], v[31]);
  const float a = fmin3(t[0], t[1], t[2]), b = fmin3(t[3], t[4], t[5]), c = fmin3(t[6], t[7], t[8]);
  return fmin3(fmin3(a, b, c), t[9], t[10]);
}

__global__ void __launch_bounds__(kThreads, 1)
match_top2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_q,
                       const __grid_constant__ CUtensorMap tmap_db, const Bf16Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar0 = base + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  const uint32_t bar_afull = bar0 + 8u * (2 * kStages);
  auto bar_tfull = [&](int acc) { return bar0 + 8u * (2 * kStages + 1 + acc); };
  auto bar_tempty = [&](int acc) { return bar0 + 8u * (2 * kStages + 3 + acc); };
  auto bar_tnfull = [&](int s) { return bar0 + 8u * (2 * kStages + 5 + s); };
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
    mbar_init(bar_afull, 1);
    for (int acc = 0; acc < 2; ++acc) {
      mbar_init(bar_tfull(acc), 1);
      mbar_init(bar_tempty(acc), kEpiWarps);
    }
    for (int s = 0; s < kTnSlots; ++s) mbar_init(bar_tnfull(s), 1);
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

  const int qb = blockIdx.x;
  const int seg = blockIdx.y;
  const int t0 = static_cast<int>(static_cast<int64_t>(seg) * a.n_tiles / a.n_seg);
  const int t1 = static_cast<int>(static_cast<int64_t>(seg + 1) * a.n_tiles / a.n_seg);
  const int kb_n = a.k_blocks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_afull, kb_n * kBlockBytes);
      for (int kb = 0; kb < kb_n; ++kb)
        tma_load_2d(base + kOffA + kb * kBlockBytes, &tmap_q, bar_afull, kb * kKBlock, qb * kTile);
      uint32_t bstep = 0;
      for (int t = t0; t < t1; ++t) {
        for (int kb = 0; kb < kb_n; ++kb, ++bstep) {
          const uint32_t s = bstep % kStages, ph = (bstep / kStages) & 1u;
          mbar_wait(bar_empty(s), ph ^ 1u);
          if (kb == 0) {
            // |t|^2 of tile j goes to slot j % 8.  The stage just acquired was freed by the MMAs of
            // tile >= j-3 (6 stages, >= 2 blocks per tile), which were issued after every epilogue
            // warp handed back the accumulator of tile j-5; so tiles j-5 .. j may be live: 6 slots.
            const uint32_t slot = static_cast<uint32_t>(t - t0) % kTnSlots;
            mbar_arrive_expect_tx(bar_tnfull(slot), kTnBytes);
            bulk_load_1d(base + kOffTn + slot * kTnBytes, a.dn + static_cast<int64_t>(t) * kTile, kTnBytes,
                         bar_tnfull(slot));
          }
          mbar_arrive_expect_tx(bar_full(s), kBlockBytes);
          tma_load_2d(base + kOffB + s * kBlockBytes, &tmap_db, bar_full(s), kb * kKBlock, t * kTile);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (convergent
    // warp, tcgen05 instructions from one elected lane)
    constexpr uint32_t idesc = umma_idesc_bf16(kTile, kTile);
    mbar_wait(bar_afull, 0);
    uint32_t bstep = 0, step = 0;
    for (int t = t0; t < t1; ++t, ++step) {
      const uint32_t acc = step & 1u;
      const uint32_t d = tmem_base + acc * kTile;
      if (step >= 2) mbar_wait(bar_tempty(acc), ((step - 2) >> 1) & 1u);
      for (int kb = 0; kb < kb_n; ++kb, ++bstep) {
        const uint32_t s = bstep % kStages, ph = (bstep / kStages) & 1u;
        mbar_wait(bar_full(s), ph);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k128(base + kOffA + kb * kBlockBytes);
        const uint64_t bdesc = umma_desc_k128(base + kOffB + s * kBlockBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kKBlock / 16; ++k)  // K = 16 bf16 = 32 B per UTCHMMA: +2 x 16 B
            umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(bar_empty(s));
          if (kb == kb_n - 1) umma_commit(bar_tfull(acc));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;            // TMEM lane quadrant this warp may read
    const int ch = (warp - 4) >> 2;       // which 64 columns of every tile
    const int row = qb * kTile + quad * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    float d1 = INFINITY, d2 = INFINITY;
    int i1 = -1, i2 = -1;
    uint32_t step = 0;
    for (int t = t0; t < t1; ++t, ++step) {
      const uint32_t acc = step & 1u, slot = step % kTnSlots;
      mbar_wait(bar_tnfull(slot), (step / kTnSlots) & 1u);
      mbar_wait(bar_tfull(acc), (step >> 1) & 1u);
      tc_fence_after();
      uint32_t v[64];
      tmem_ld64_wait(tmem_base + lane_sel + acc * kTile + ch * 64, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty(acc));
      const float4* tn4 = reinterpret_cast<const float4*>(smem + kOffTn + slot * kTnBytes) + ch * 16;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float k[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 n4 = tn4[c * 8 + j];
          k[4 * j + 0] = __uint_as_float(v[c * 32 + 4 * j + 0]) + n4.x;
          k[4 * j + 1] = __uint_as_float(v[c * 32 + 4 * j + 1]) + n4.y;
          k[4 * j + 2] = __uint_as_float(v[c * 32 + 4 * j + 2]) + n4.z;
          k[4 * j + 3] = __uint_as_float(v[c * 32 + 4 * j + 3]) + n4.w;
        }
        if (min_tree32(k) < d2) {
          const int col0 = t * kTile + ch * 64 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float kk = k[j];
            if (kk < d2) {
              if (kk < d1) {
                d2 = d1;
                i2 = i1;
                d1 = kk;
                i1 = col0 + j;
              } else {
                d2 = kk;
                i2 = col0 + j;
              }
            }
          }
        }
      }
    }
    if (row < a.nq) {
      const float qn = a.qn[row];
      const size_t o = (static_cast<size_t>(seg * 2 + ch) * a.nq + row) * 2;
      a.part_d2[o] = i1 >= 0 ? fmaxf(d1 + qn, 0.0f) : INFINITY;
      a.part_d2[o + 1] = i2 >= 0 ? fmaxf(d2 + qn, 0.0f) : INFINITY;
      a.part_idx[o] = i1 >= 0 ? a.idx_base + i1 : -1;
      a.part_idx[o + 1] = i2 >= 0 ? a.idx_base + i2 : -1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// One warp per output row: bf16 operand (optionally hi/lo split) and the fp32 squared norm.
__global__ void bf16_operand_kernel(const float* __restrict__ src, int64_t n_rows, int64_t n_out_rows,
                                    int side, int split, __nv_bfloat16* __restrict__ dst,
                                    float* __restrict__ norms, int32_t* __restrict__ nonfinite_flag) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_out_rows) return;
  const int cols = split ? 3 * SOD_DESC_DIM : SOD_DESC_DIM;
  __nv_bfloat16* out = dst + row * cols + lane * 4;
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < n_rows) x = reinterpret_cast<const float4*>(src + row * SOD_DESC_DIM)[lane];
  const float xs[4] = {x.x, x.y, x.z, x.w};
  const float scale = side ? -2.0f : 1.0f;  // exact in bf16: only the exponent changes
  double sq = 0.0;
  bool finite = true;
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    finite = finite && isfinite(xs[i]);
    sq += static_cast<double>(xs[i]) * static_cast<double>(xs[i]);
    hi[i] = __float2bfloat16_rn(xs[i]);
    lo[i] = __float2bfloat16_rn(xs[i] - __bfloat162float(hi[i]));
    hi[i] = __float2bfloat16_rn(scale * __bfloat162float(hi[i]));
    lo[i] = __float2bfloat16_rn(scale * __bfloat162float(lo[i]));
  }
  auto store4 = [](__nv_bfloat16* p, const __nv_bfloat16* v) {
    uint2 u;
    u.x = static_cast<uint32_t>(__bfloat16_as_ushort(v[0])) | static_cast<uint32_t>(__bfloat16_as_ushort(v[1])) << 16;
    u.y = static_cast<uint32_t>(__bfloat16_as_ushort(v[2])) | static_cast<uint32_t>(__bfloat16_as_ushort(v[3])) << 16;
    *reinterpret_cast<uint2*>(p) = u;
  };
  store4(out, hi);
  if (split) {  // query side [h | h | l], database side [h | l | h]
    store4(out + SOD_DESC_DIM, side ? lo : hi);
    store4(out + 2 * SOD_DESC_DIM, side ? hi : lo);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (__any_sync(0xffffffffu, !finite) && lane == 0 && nonfinite_flag) atomicOr(nonfinite_flag, 1);
  if (lane == 0) norms[row] = row < n_rows ? static_cast<float>(sq) : INFINITY;
}

// Merge candidate lists in (d^2, index) order and apply the ratio test the way the reference does:
// distances are float32 square roots, the comparison m.distance < ratio * n.distance runs in double
// (main.py:75-77).
__global__ void top2_merge_f32_kernel(const int32_t* __restrict__ parts_idx, const float* __restrict__ parts_d2,
                                      int n_parts, int64_t nq, int32_t* __restrict__ out_idx,
                                      float* __restrict__ out_d2, float* __restrict__ out_dist,
                                      uint8_t* __restrict__ out_pass, double ratio) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= nq) return;
  float d1 = INFINITY, d2 = INFINITY;
  int i1 = -1, i2 = -1;
  auto before = [](float da, int ia, float db, int ib) {  // (da, ia) < (db, ib); index -1 = empty, last
    if (ib < 0) return true;
    return da < db || (da == db && ia < ib);
  };
  for (int p = 0; p < n_parts; ++p)
    for (int k = 0; k < 2; ++k) {
      const size_t o = (static_cast<size_t>(p) * nq + row) * 2 + k;
      const int i = parts_idx[o];
      if (i < 0) continue;
      const float d = parts_d2[o];
      if (before(d, i, d1, i1)) {
        d2 = d1;
        i2 = i1;
        d1 = d;
        i1 = i;
      } else if (before(d, i, d2, i2)) {
        d2 = d;
        i2 = i;
      }
    }
  out_idx[row * 2] = i1;
  out_idx[row * 2 + 1] = i2;
  out_d2[row * 2] = d1;
  out_d2[row * 2 + 1] = d2;
  const float s1 = sqrtf(d1), s2 = sqrtf(d2);
  if (out_dist) {
    out_dist[row * 2] = s1;
    out_dist[row * 2 + 1] = s2;
  }
  if (out_pass) out_pass[row] = (i2 >= 0 && static_cast<double>(s1) < ratio * static_cast<double>(s2)) ? 1 : 0;
}

struct Bf16Plan {
  int n_qblocks, n_tiles, n_seg;
};
Bf16Plan make_bf16_plan(int64_t nq, int64_t n_db, int sms) {
  Bf16Plan p;
  p.n_qblocks = static_cast<int>((nq + kTile - 1) / kTile);
  p.n_tiles = static_cast<int>((n_db + kTile - 1) / kTile);
  p.n_seg = 1;  // split the database until there are two CTAs per SM or segments get short
  while (static_cast<int64_t>(p.n_qblocks) * p.n_seg < 2 * sms && p.n_tiles / (p.n_seg * 2) >= 8 && p.n_seg < 64)
    p.n_seg *= 2;
  return p;
}

}  // namespace
}  // namespace sod

using namespace sod;

extern "C" {

int64_t sod_bf16_operand_cols(int32_t split) { return split ? 3 * SOD_DESC_DIM : SOD_DESC_DIM; }

int64_t sod_bf16_db_rows(int64_t n_rows) { return n_rows <= 0 ? 0 : (n_rows + kTile - 1) / kTile * kTile; }

int sod_bf16_prepare(const float* src, int64_t n_rows, int32_t side, int32_t split, uint16_t* dst,
                     float* norms, int32_t* nonfinite_flag, sod_stream_t stream) {
  SOD_CHECK_ARG(n_rows >= 0, "negative size");
  SOD_CHECK_ARG(side == 0 || side == 1, "side must be 0 (query) or 1 (database)");
  if (n_rows == 0) return SOD_OK;
  SOD_CHECK_ARG(src && dst && norms, "null pointer");
  SOD_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                "src and dst must be 16-byte aligned");
  const int64_t n_out = side ? sod_bf16_db_rows(n_rows) : n_rows;
  const int threads = 256;
  const int64_t blocks = (n_out * 32 + threads - 1) / threads;
  bf16_operand_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      src, n_rows, n_out, side, split ? 1 : 0, reinterpret_cast<__nv_bfloat16*>(dst), norms, nonfinite_flag);
  SOD_CHECK_LAUNCH("bf16_operand_kernel");
  return SOD_OK;
}

size_t sod_match_bf16_workspace_bytes(int64_t n_query, int64_t n_db) {
  if (n_query <= 0 || n_db <= 0) return 16;
  const int sms = device_sm_count();
  const Bf16Plan p = make_bf16_plan(n_query, n_db, sms > 0 ? sms : 148);
  return static_cast<size_t>(p.n_seg) * 2 * static_cast<size_t>(n_query) * 2 * 8 + 16;
}

int sod_top2_merge_f32(const int32_t* parts_idx, const float* parts_d2, int32_t n_parts, int64_t n_query,
                       int32_t* out_idx, float* out_d2, float* out_dist, uint8_t* out_pass, double ratio,
                       sod_stream_t stream) {
  SOD_CHECK_ARG(n_parts >= 0 && n_query >= 0, "negative size");
  if (n_query == 0) return SOD_OK;
  SOD_CHECK_ARG(out_idx && out_d2, "null output pointer");
  SOD_CHECK_ARG(n_parts == 0 || (parts_idx && parts_d2), "null parts pointer");
  const int threads = 128;
  top2_merge_f32_kernel<<<static_cast<unsigned>((n_query + threads - 1) / threads), threads, 0,
                          static_cast<cudaStream_t>(stream)>>>(parts_idx, parts_d2, n_parts, n_query, out_idx,
                                                               out_d2, out_dist, out_pass, ratio);
  SOD_CHECK_LAUNCH("top2_merge_f32_kernel");
  return SOD_OK;
}

int sod_match_top2_bf16(const uint16_t* q_op, const float* qn, int64_t n_query, const uint16_t* db_op,
                        const float* dn, int64_t n_db, int32_t split, int32_t db_index_base, int32_t* out_idx,
                        float* out_d2, void* workspace, size_t workspace_bytes, sod_stream_t stream) {
  SOD_CHECK_ARG(n_query >= 0 && n_db >= 0, "negative size");
  SOD_CHECK_ARG(n_query < (int64_t(1) << 31) - kTile && n_db < (int64_t(1) << 31) - kTile, "size out of range");
  SOD_CHECK_ARG(static_cast<int64_t>(db_index_base) + n_db < (int64_t(1) << 31),
                "db_index_base + n_db overflows int32");
  if (n_query == 0) return SOD_OK;
  SOD_CHECK_ARG(out_idx && out_d2, "null output pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_db == 0)
    return sod_top2_merge_f32(nullptr, nullptr, 0, n_query, out_idx, out_d2, nullptr, nullptr, 0.0, stream);
  SOD_CHECK_ARG(q_op && qn && db_op && dn && workspace, "null pointer");
  SOD_CHECK_ARG((reinterpret_cast<uintptr_t>(q_op) & 15) == 0 && (reinterpret_cast<uintptr_t>(db_op) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dn) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "operands, dn and workspace must be 16-byte aligned");
  const int sms = device_sm_count();
  if (sms <= 0) return SOD_ERR_CUDA;
  const Bf16Plan p = make_bf16_plan(n_query, n_db, sms);
  SOD_CHECK_ARG(p.n_seg <= 65535, "too many segments");
  const size_t need = static_cast<size_t>(p.n_seg) * 2 * static_cast<size_t>(n_query) * 2 * 8;
  SOD_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  const int cols = static_cast<int>(sod_bf16_operand_cols(split));
  const int k_blocks = cols / kKBlock;

  CUtensorMap map_q, map_db;
  int rc = make_rowmajor_map(&map_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q_op, n_query, cols, kKBlock, kTile);
  if (rc != SOD_OK) return rc;
  rc = make_rowmajor_map(&map_db, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db_op, sod_bf16_db_rows(n_db), cols,
                         kKBlock, kTile);
  if (rc != SOD_OK) return rc;

  Bf16Args a;
  a.qn = qn;
  a.dn = dn;
  a.part_d2 = static_cast<float*>(workspace);
  a.part_idx = reinterpret_cast<int32_t*>(a.part_d2 + static_cast<size_t>(p.n_seg) * 2 * n_query * 2);
  a.nq = static_cast<int>(n_query);
  a.n_tiles = p.n_tiles;
  a.n_seg = p.n_seg;
  a.k_blocks = k_blocks;
  a.idx_base = db_index_base;

  SOD_CHECK_CUDA(cudaFuncSetAttribute(match_top2_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem_bytes(kMaxKBlocks)));  // per device, hence at every launch
  const dim3 grid(static_cast<unsigned>(p.n_qblocks), static_cast<unsigned>(p.n_seg));
  match_top2_bf16_kernel<<<grid, kThreads, smem_bytes(k_blocks), st>>>(map_q, map_db, a);
  SOD_CHECK_LAUNCH("match_top2_bf16_kernel");
  return sod_top2_merge_f32(a.part_idx, a.part_d2, p.n_seg * 2, n_query, out_idx, out_d2, nullptr, nullptr, 0.0,
                            stream);
}

}  // extern "C"

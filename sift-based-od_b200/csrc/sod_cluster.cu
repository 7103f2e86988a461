// Pose clustering after the affine stage: the O(V^2) neighbour tests of
// group_position (PostProcessing.py:14-37) and group_orientation (PostProcessing.py:39-63) as bit
// matrices, plus the connected components of each graph by a lock-free union-find.  The depth-first
// visiting order that fixes the reference's float summation order is recovered on the host from the
// bit rows (sod_b200/postprocess.py); the pair tests themselves use the reference's fp64 expressions.
#include "sod_common.cuh"

namespace sod {
namespace {

constexpr int kThreads = 256;

// Loads go to L2 (__ldcg): another SM's hook must not hide behind a stale L1 line.
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = __ldcg(parent + x);
  while (p != x) {  // path halving; parents only ever decrease, so stale reads are still ancestors
    const int g = __ldcg(parent + p);
    if (g != p) parent[x] = g;
    x = p;
    p = __ldcg(parent + x);
  }
  return x;
}

// Roots are always hooked under the smaller root: the final root of a component is its lowest index.
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    if (atomicCAS(parent + a, a, b) == a) return;
  }
}

__global__ void label_init_kernel(int* __restrict__ label, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) label[i] = i;
}

// Final labels.  No path compression here: a halving store that was computed from older parents could
// land after a thread has written its final label and replace it by a mere ancestor.  The walk only
// reads; the one store per thread writes a root, which is a valid parent for every concurrent reader.
__global__ void label_flatten_kernel(int* __restrict__ label, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x = i, p = __ldcg(label + x);
  while (p != x) {
    x = p;
    p = __ldcg(label + x);
  }
  label[i] = x;
}

// One CTA per bin i; its warps sweep the other bins 32 at a time (lane <-> j), __ballot_sync builds the
// adjacency word.  kind 0: position test, kind 1: angle test.
template <int kKind>
__global__ void __launch_bounds__(kThreads)
adjacency_kernel(const double* __restrict__ a0, const double* __restrict__ a1, const double* __restrict__ a2,
                 const double* __restrict__ a3, const int32_t* __restrict__ segment, int n, double limit,
                 uint32_t* __restrict__ adj, int* __restrict__ parent) {
  const int i = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int words = (n + 31) >> 5;
  const int seg_i = segment ? segment[i] : 0;
  const double xi = a0[i];
  double yi = 0.0, wi = 0.0, hi = 0.0;
  if (kKind == 0) {
    yi = a1[i];
    wi = a2[i];
    hi = a3[i];
  }
  for (int w = warp; w < words; w += kThreads / 32) {
    const int j = w * 32 + lane;
    bool edge = false;
    if (j < n && j != i && (!segment || segment[j] == seg_i)) {
      if (kKind == 0) {
        // abs(xa - xb) <= w_a*s_a/4 and abs(ya - yb) <= h_a*s_a/4 and the same against bin b
        const double dx = fabs(__dsub_rn(a0[j], xi)), dy = fabs(__dsub_rn(a1[j], yi));
        edge = dx <= a2[j] && dy <= a3[j] && dx <= wi && dy <= hi;
      } else {
        // abs(math.degrees(angle_a - angle_b)) <= 1; math.degrees(x) is x * (180 / pi) in double
        edge = fabs(__dmul_rn(__dsub_rn(a0[j], xi), 180.0 / 3.14159265358979323846)) <= limit;
      }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, edge);
    if (lane == 0) adj[static_cast<size_t>(i) * words + w] = word;
    if (edge && j < i) uf_union(parent, i, j);
  }
}

template <int kKind>
int run_adjacency(const double* a0, const double* a1, const double* a2, const double* a3, const int32_t* segment,
                  int64_t n, double limit, uint32_t* adj, int32_t* label, sod_stream_t stream) {
  SOD_CHECK_ARG(n >= 0 && n <= SOD_MAX_CLUSTER_BINS, "n must be in 0..%d", SOD_MAX_CLUSTER_BINS);
  if (n == 0) return SOD_OK;
  SOD_CHECK_ARG(a0 && adj && label && (kKind == 1 || (a1 && a2 && a3)), "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ni = static_cast<int>(n);
  const unsigned blocks = static_cast<unsigned>((ni + kThreads - 1) / kThreads);
  label_init_kernel<<<blocks, kThreads, 0, st>>>(label, ni);
  SOD_CHECK_LAUNCH("label_init_kernel");
  adjacency_kernel<kKind><<<static_cast<unsigned>(ni), kThreads, 0, st>>>(a0, a1, a2, a3, segment, ni, limit, adj,
                                                                         label);
  SOD_CHECK_LAUNCH("adjacency_kernel");
  label_flatten_kernel<<<blocks, kThreads, 0, st>>>(label, ni);
  SOD_CHECK_LAUNCH("label_flatten_kernel");
  return SOD_OK;
}

}  // namespace
}  // namespace sod

using namespace sod;

extern "C" {

int sod_pose_adjacency(const double* cx, const double* cy, const double* reach_x, const double* reach_y,
                       const int32_t* segment, int64_t n, uint32_t* adj, int32_t* label, sod_stream_t stream) {
  return run_adjacency<0>(cx, cy, reach_x, reach_y, segment, n, 0.0, adj, label, stream);
}

int sod_angle_adjacency(const double* angle, const int32_t* segment, int64_t n, double max_degrees, uint32_t* adj,
                        int32_t* label, sod_stream_t stream) {
  return run_adjacency<1>(angle, nullptr, nullptr, nullptr, segment, n, max_degrees, adj, label, stream);
}

}  // extern "C"

/* sod.h — C ABI of libsod_b200.so, the B200 (sm_100a) implementation of the SIFT object-detection
 * hot path of torn8to/sift-based-OD: 2-NN descriptor matching + ratio test, 4-D Hough pose-bin
 * voting, per-bin affine verification.
 *
 * The reference has no FFI: its boundary is the Python call surface.  Each entry
 * point below names the reference code it replaces; INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - all work is enqueued on `stream` (a cudaStream_t) and is asynchronous; no entry point
 *     synchronises, allocates or frees caller-visible memory;
 *   - return value: 0 = SOD_OK, negative = error; sod_last_error() gives a thread-local message;
 *   - descriptors are unsigned 8-bit, 128 per row, row-major (OpenCV SIFT descriptors are
 *     integer-valued 0..255 stored as float32; sod_pack_u8_from_f32 converts and verifies).
 */
#ifndef SOD_H_
#define SOD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOD_OK 0
#define SOD_ERR_INVALID_ARGUMENT (-1)
#define SOD_ERR_CUDA (-2)
#define SOD_ERR_UNSUPPORTED (-3)

#define SOD_DESC_DIM 128  /* bytes per descriptor row */
#define SOD_TILE_ROWS 128 /* database rows per MMA tile */
#define SOD_CQ_TILE_INTS 132 /* int32 per tile in the prepared `cq` array: 128 keys + 4 chunk minima */

typedef void* sod_stream_t; /* cudaStream_t */

/* ABI version (major*1000 + minor). */
int sod_version(void);
/* Message for the last error raised on the calling thread ("" if none). */
const char* sod_last_error(void);
/* Number of SMs of the current device (grid sizing; 148 on B200). Negative on error. */
int sod_device_sm_count(void);

/* Number of int32 elements of the prepared per-row constants `cq` for a database of n_rows:
 * ceil(n_rows / SOD_TILE_ROWS) * SOD_CQ_TILE_INTS. */
int64_t sod_cq_ints(int64_t n_rows);

/* float32 [n_rows,128] -> u8 [n_rows,128].  *nonint_flag (device int32, caller zeroes it) is set
 * to 1 if any value is not an integer in 0..255 (such a set needs the bf16 path, not built yet).
 * Replaces nothing in the reference; it is the packing step in front of main.py:71. */
int sod_pack_u8_from_f32(const float* src, int64_t n_rows, uint8_t* dst, int32_t* nonint_flag,
                         sod_stream_t stream);

/* K1 (database side).  Per tile of 128 rows, cq holds 128 packed keys
 * (sum_k db[i][k]^2 << 8) | (i % 128)  (INT32_MAX for rows past the end) followed by the minimum
 * sum of squares of each 32-row chunk (used to prune the epilogue); sod_cq_ints(n_rows) int32 in
 * total.  This is the per-row term of |q-t|^2 = |q|^2 + |t|^2 - 2 q.t that OpenCV recomputes per
 * pair inside cv::batchDistance (called from main.py:71). */
int sod_db_prepare(const uint8_t* db, int64_t n_rows, int32_t* cq, sod_stream_t stream);

/* K1 (query side).  qn[i] = sum_k q[i][k]^2. */
int sod_query_prepare(const uint8_t* q, int64_t n_rows, int32_t* qn, sod_stream_t stream);

/* Bytes of scratch sod_match_top2 needs for these sizes on the current device. */
size_t sod_match_workspace_bytes(int64_t n_query, int64_t n_db);

/* K2.  For every query row the two database rows with the smallest squared L2 distance, ascending,
 * ties -> lowest index: the result of cv2.BFMatcher().knnMatch(des_query, des, k=2)
 * (main.py:70-71) with exact integer distances.  out_idx[i][j] = db_index_base + row (or -1 when the
 * database has fewer than j+1 rows), out_d2[i][j] = squared distance (0xFFFFFFFF when idx = -1).
 * tcgen05 kind::i8 MMA, TMA-fed, top-2 selection fused in the TMEM epilogue. */
int sod_match_top2(const uint8_t* q, const int32_t* qn, int64_t n_query, const uint8_t* db,
                   const int32_t* cq, int64_t n_db, int32_t db_index_base, int32_t* out_idx,
                   uint32_t* out_d2, void* workspace, size_t workspace_bytes, sod_stream_t stream);

/* K3.  Merge n_parts candidate lists (parts_idx/parts_d2 are [n_parts][n_query][2], e.g. the
 * all-gathered shard-local results) into the global top-2 by (d2, idx) lexicographic order and
 * apply Lowe's ratio test exactly as the reference evaluates it (main.py:81-82):
 *   distance = sqrtf(d2) as float32 (OpenCV), pass = (double)dist1 < ratio * (double)dist2.
 * out_dist (float32 [n_query][2]) and out_pass (u8 [n_query]) may be NULL. */
int sod_top2_merge(const int32_t* parts_idx, const uint32_t* parts_d2, int32_t n_parts,
                   int64_t n_query, int32_t* out_idx, uint32_t* out_d2, float* out_dist,
                   uint8_t* out_pass, double ratio, sod_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SOD_H_ */

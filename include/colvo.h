/* colvo.h -- C ABI of the B200-native ColVO photometric-loss path (libcolvo_b200.so).
 *
 * Drop-in boundary for the data-parallel hot path of HNUicda/CoIVO ("ColVO"): the
 * per-pixel view-synthesis photometric loss that couples depth and pose
 * (/root/reference/README.md:7 "loss function constraints to couple depth and pose
 * estimation modes ... alignment of geometric projections between consecutive frames";
 * "LCC ... recalibrating the luminosity values of adjacent frames").
 *
 * The upstream repository ships no source, hence no FFI to mirror: every entry point below
 * cites the README sentence and the SURVEY.md section 8 row it implements, and
 * INTEGRATION.md shows the ctypes binding a maintainer of a PyTorch training loop adds.
 *
 * Conventions:
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every tensor is fp32, contiguous NCHW, in DEVICE memory unless the name says host;
 *   - the caller owns every buffer including the workspace; the library never allocates
 *     device memory, never retains a pointer, never synchronises the device;
 *   - all work is enqueued on the given cudaStream_t (passed as void*); calls are
 *     CUDA-graph capturable and re-entrant on disjoint buffers;
 *   - return value: 0 = success, >0 = a cudaError_t, <0 = a COLVO_E_* code;
 *     colvo_error_string() explains either.
 */
#ifndef COLVO_H_
#define COLVO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COLVO_VERSION 100          /* major*10000 + minor*100 + patch  (0.1.0) */
#define COLVO_MAX_SCALES 4
#define COLVO_MAX_SOURCES 2

/* flags */
#define COLVO_F_LCC 1u             /* apply the LCC brightness calibration (README.md:5,7)     */
#define COLVO_F_LCC_DETACH 2u      /* do not differentiate through (a, b)                       */
#define COLVO_F_SAVE_FOR_BWD 4u    /* forward also produces what colvo_photo_backward needs     */
#define COLVO_F_NO_SRC_GRAD 8u     /* backward: skip the scatter-add into grad_srcs             */
#define COLVO_F_PACKED_BF16 16u    /* tgt / srcs are RGBA-interleaved bf16 ([..,H,W,4], 8 B per pixel, A ignored);
                                      arithmetic stays fp32; requires COLVO_F_NO_SRC_GRAD in the backward
                                                                        */

#define COLVO_F_HOST_U8 32u        /* colvo_photo_step_host only: h_tgt / h_srcs are uint8 frames (what a video loader holds);
                                      they are copied as bytes and widened on the device, x = u8 * (1.0f / 255.0f)   */

#define COLVO_F_SCATTER_MERGE 64u   /* backward: warp-aggregated scatter into grad_srcs -- coincident bilinear taps of neighbouring
                                      pixels are summed inside the warp (shared-memory exchange, no shared atomics) before they go
                                      to global memory: two vector REDs per pixel, source and scale instead of four.  Same results;
                                      measured slower than the default scatter on B200, hence opt-in */

/* negative error codes */
#define COLVO_E_BAD_DESC (-1)
#define COLVO_E_WORKSPACE (-2)
#define COLVO_E_NULL_PTR (-3)
#define COLVO_E_MISALIGNED (-4)
#define COLVO_E_UNSUPPORTED (-5)
#define COLVO_E_TENSOR_MAP (-6)    /* cuTensorMapEncodeTiled refused the TMA descriptor of an internal buffer */

/* Problem descriptor.  h[k] = H >> k, w[k] = W >> k must hold (oracle A1/A3). */
typedef struct ColvoDesc {
  int32_t B, N, S, H, W;           /* triplets, sources (1..2), scales (1..4), full resolution  */
  int32_t h[COLVO_MAX_SCALES];
  int32_t w[COLVO_MAX_SCALES];
  float alpha;                     /* SSIM weight, 0.85                                          */
  float c1, c2;                    /* SSIM constants 1e-4, 9e-4                                  */
  float eps_proj;                  /* 1e-7 in iz = 1/(Z'+eps)                                    */
  float eps_lcc;                   /* 1e-6 on the LCC variance                                   */
  float eps_disp;                  /* 1e-7 on the mean inverse depth                             */
  float z_min;                     /* 1e-3                                                       */
  float smooth_weight;             /* 1e-3 (scaled by 2^-k per scale)                            */
  float geo_weight;                /* weight of the geometric-consistency term (needs src_depth) */
  uint32_t flags;
} ColvoDesc;

/* Fill a descriptor with the defaults of the path (oracle/photometric.py constants). */
int colvo_desc_init(ColvoDesc* d, int32_t B, int32_t N, int32_t S, int32_t H, int32_t W, uint32_t flags);

int colvo_version(void);
const char* colvo_error_string(int rc);

/* Scratch bytes colvo_photo_forward / colvo_photo_backward need (max of the two), and the
 * number of doubles in the `saved` buffer the forward hands to the backward. */
int colvo_workspace_bytes(const ColvoDesc* d, size_t* bytes);
int colvo_saved_doubles(const ColvoDesc* d, size_t* count);

/* Forward: SURVEY.md section 8(a) rows 0-10.
 *   tgt [B,3,H,W]  srcs [B,N,3,H,W]  depth[k] [B,1,h_k,w_k]  K [B,3,3]  T [B,N,4,4]
 *   (with COLVO_F_PACKED_BF16: tgt [B,H,W,4] bf16, srcs [B,N,H,W,4] bf16, 8-byte aligned)
 *   src_depth [B,N,1,H,W]     (nullable) depth maps of the source frames: with geo_weight != 0 adds the
 *                             geometric-consistency term (README.md:1,7)
 *   loss  [1]                 (out)
 *   ab    [B,N,S,2]           (out)  LCC gain/bias per warped frame and scale
 *   valid [B,N,S,H,W] u8      (out, nullable)  bit-exact projection validity
 *   sel   [B,S,H,W]   u8      (out; required with COLVO_F_SAVE_FOR_BWD, else nullable)
 *   saved [colvo_saved_doubles] (out; required with COLVO_F_SAVE_FOR_BWD, else nullable)
 */
int colvo_photo_forward(const ColvoDesc* d, const void* tgt, const void* srcs, const float* const* depth,
                        const float* K, const float* T, const float* src_depth, float* loss, float* ab,
                        uint8_t* valid, uint8_t* sel, double* saved, void* ws, size_t ws_bytes, void* stream);

/* The same forward with one more output (SURVEY.md section 8(f)-2 "also yields a soft occlusion mask"):
 *   occ [B,N,S,H,W] fp32  (out, nullable)  1 - diff of the geometric-consistency term where the projection is valid, 0
 *                         elsewhere (SC-Depth's weight mask: small where the re-projected depth disagrees with the
 *                         source frame's own depth map, i.e. at occlusions / moving tissue).  Needs src_depth and
 *                         geo_weight != 0 (COLVO_E_UNSUPPORTED otherwise); a constant, no gradient flows through it.
 */
int colvo_photo_forward_occ(const ColvoDesc* d, const void* tgt, const void* srcs, const float* const* depth,
                            const float* K, const float* T, const float* src_depth, float* loss, float* ab,
                            uint8_t* valid, uint8_t* sel, float* occ, double* saved, void* ws, size_t ws_bytes,
                            void* stream);

/* Backward: SURVEY.md section 8(a) row 11.  Inputs as in the forward plus its sel / saved.
 *   grad_loss  [1] device scalar (dL_total / dloss)
 *   grad_depth[k] [B,1,h_k,w_k]   (out, overwritten)
 *   grad_T     [B,N,4,4]          (out, overwritten; bottom row 0)
 *   grad_srcs  [B,N,3,H,W]        (out, overwritten; nullable with COLVO_F_NO_SRC_GRAD)
 *   grad_src_depth [B,N,1,H,W]    (out, overwritten; nullable)
 * No gradient is produced for K or tgt (oracle A14).
 */
int colvo_photo_backward(const ColvoDesc* d, const void* tgt, const void* srcs, const float* const* depth,
                         const float* K, const float* T, const float* src_depth, const float* grad_loss,
                         const uint8_t* sel, const double* saved, float* const* grad_depth, float* grad_T,
                         float* grad_srcs, float* grad_src_depth, void* ws, size_t ws_bytes, void* stream);

/* Inference-time warp + LCC consistency sweep over a frame sequence (BASELINE config 5;
 * README.md:29: depth maps are stitched along the trajectory -- this is the check that gates it).
 *   frames [F,3,H,W]  depth [F,1,H,W]  T [F-1,4,4] (T_{t->t+1})  K [3,3] (k_per_pair = 0) or [F-1,3,3]
 *   out    [F-1,4] = {mean pe over valid pixels, a, b, valid fraction}
 * flags: COLVO_F_LCC honoured.  Workspace: colvo_consistency_workspace_bytes.
 */
int colvo_consistency_workspace_bytes(int32_t F, int32_t H, int32_t W, size_t* bytes);
int colvo_consistency(int32_t F, int32_t H, int32_t W, uint32_t flags, const float* frames, const float* depth,
                      const float* T, const float* K, int32_t k_per_pair, float* out, void* ws, size_t ws_bytes,
                      void* stream);

/* End-to-end step on HOST buffers: H2D copies of the inputs (use pinned memory), forward,
 * backward with grad_loss = grad_scale, D2H copies of loss and gradients, all on `stream`.
 * `arena` is caller-owned DEVICE memory of colvo_step_host_arena_bytes bytes.
 * h_grad_srcs may be NULL with COLVO_F_NO_SRC_GRAD.  h_grad_depth == NULL keeps ALL gradients on the device (only
 * the loss is read back): a training step consumes them there; colvo_step_host_arena_grads gives their byte
 * offsets inside `arena` (grad_depth_off[S], grad_T [B,N,4,4], grad_srcs [B,N,3,H,W]).  The call does not synchronise: the host
 * outputs are complete once `stream` has been synchronised.  A caller that splits a batch into
 * chunks on several streams (to overlap H2D, compute and D2H) passes grad_scale = B_chunk / B and
 * combines the chunk losses with the same weights.
 * h_tgt [B,3,H,W] / h_srcs [B,N,3,H,W] are fp32, or uint8 with COLVO_F_HOST_U8 (a quarter of the image bytes cross
 * PCIe; the arena then also holds the byte staging area).
 */
int colvo_step_host_arena_bytes(const ColvoDesc* d, size_t* bytes);
int colvo_step_host_arena_grads(const ColvoDesc* d, size_t* grad_depth_off, size_t* grad_T_off, size_t* grad_srcs_off);
int colvo_photo_step_host(const ColvoDesc* d, const void* h_tgt, const void* h_srcs, const float* const* h_depth,
                          const float* h_K, const float* h_T, float* h_loss, float* const* h_grad_depth,
                          float* h_grad_T, float* h_grad_srcs, float grad_scale, void* arena, size_t arena_bytes,
                          void* stream);

/* Front-end: raw network outputs -> the tensors the loss takes.
 * Monodepth2's transformation_from_parameters / disp_to_depth (oracle/frontend.py).
 *   axisangle, translation [B,N,3]; bit n of invert_mask: source n's motion is given source->target and
 *   must be inverted (the frame t-1 convention); T, grad_T [B,N,4,4].
 *   disp_to_depth: depth = 1 / (1/max_depth + (1/min_depth - 1/max_depth) * disp), S tensors of counts[k]
 *   elements each (counts is a HOST array); the backward takes the forward's depth.
 */
int colvo_pose_from_axisangle(int32_t B, int32_t N, uint32_t invert_mask, const float* axisangle,
                              const float* translation, float* T, void* stream);
int colvo_pose_from_axisangle_backward(int32_t B, int32_t N, uint32_t invert_mask, const float* axisangle,
                                       const float* translation, const float* grad_T, float* grad_axisangle,
                                       float* grad_translation, void* stream);
int colvo_disp_to_depth(int32_t S, const int64_t* counts, const float* const* disp, float* const* depth,
                        float min_depth, float max_depth, void* stream);
int colvo_disp_to_depth_backward(int32_t S, const int64_t* counts, const float* const* depth,
                                 const float* const* grad_depth, float* const* grad_disp, float min_depth,
                                 float max_depth, void* stream);

/* Profiling aid (bench.py's roofline leg): bracket the NEXT launch of one kernel with two
 * caller-owned cudaEvent_t on the launching stream.  One-shot, process-wide; pass
 * which = 0 to clear.  Not for use under CUDA-graph capture.
 *   which: 1 = k_photo_fwd, 2 = k_photo_bwd, 3 = k_warp_stats, 4 = k_consistency_pe (first pass of a sweep) */
#define COLVO_K_PHOTO_FWD 1
#define COLVO_K_PHOTO_BWD 2
#define COLVO_K_WARP_STATS 3
#define COLVO_K_CONSISTENCY_PE 4
int colvo_debug_time_kernel(int which, void* ev_start, void* ev_stop);

#ifdef __cplusplus
}
#endif
#endif /* COLVO_H_ */

/*
 * ezflow_b200 — C ABI of the B200-native similarity / cost-volume hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch
 * types.  Every pointer is a DEVICE pointer unless it says "host".  Every entry
 * point returns 0 on success and a negative ezb_status otherwise; it never aborts,
 * never synchronises the device and never allocates device memory (callers pass
 * outputs and workspaces).  The calling thread's current CUDA device must be the
 * one that owns the pointers; work is enqueued on `stream` (a cudaStream_t).
 * The library is re-entrant: no mutable global state besides a thread-local error
 * string, so one thread per device (nn.DataParallel, reference
 * ezflow/engine/eval.py:306) is safe.
 *
 * Each function cites the reference interface (neu-vi/ezflow v0.2.5, paths relative
 * to the reference checkout) whose arithmetic it replaces.
 */
#ifndef EZFLOW_B200_H
#define EZFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EZB_VERSION 100 /* major*10000 + minor*100 + patch */

typedef enum {
  EZB_OK = 0,
  EZB_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, ...) -> ValueError   */
  EZB_ERR_UNSUPPORTED = -2, /* valid in the reference but not built here -> NotImplementedError     */
  EZB_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed -> RuntimeError                   */
  EZB_ERR_WORKSPACE = -4    /* workspace too small -> RuntimeError                                  */
} ezb_status;

/* storage type of the correlation pyramid (the arithmetic is always fp16 x fp16 -> fp32 on tcgen05) */
typedef enum { EZB_F32 = 0, EZB_F16 = 1 } ezb_dtype;

#define EZB_MAX_LEVELS 8

int ezb_version(void);
/* copies the calling thread's last error message (NUL-terminated) into buf; returns its length */
int ezb_last_error(char* buf, size_t n);

/* ---------------------------------------------------------------------------------------------
 * RAFT all-pairs correlation pyramid (K1 + K2)
 *   replaces MutliScalePairwise4DCorr.__init__ / .corr
 *   ezflow/similarity/correlation/pairwise.py:26-41, :78-88
 *
 * Pyramid storage (the library's choice, "subtile-sequence"): ONE buffer per call holding
 * all levels.  h[l] = H >> l, w[l] = W >> l (floor, F.avg_pool2d default); at most 4 levels.  Level l of
 * a query's correlation map is cut into 8x8-pixel subtiles; the subtiles of all levels form one sequence
 *   s = sub_base[l] + (y/8) * ceil(w[l]/8) + x/8
 * and 4 consecutive subtiles are one 256-element record (= one GEMM tile of 256 accumulator columns):
 *   tile = s / 4,   column = (s % 4) * 64 + (y % 8) * 8 + x % 8.
 * Queries (q = h1*W + w1) are grouped 32 at a time.  Each (pair b, tile t, query group g) owns a
 * contiguous *group block* of `group_bytes` at  ((b * n_tiles + t) * n_groups + g) * group_bytes;
 * ezb_corr_pyramid_record_offset() gives the byte offset of a record column inside a block.
 * Pixels outside a level are stored as zeros.  Stored value * deq_scale[b] = reference value.
 * ------------------------------------------------------------------------------------------- */
int ezb_corr_pyramid_layout(int H, int W, int L, int dtype,
                            int* h /*[L] host*/, int* w /*[L] host*/, int* sub_base /*[L+1] host*/,
                            int* n_tiles, int* n_groups,
                            int64_t* group_bytes, int64_t* pair_bytes /* all host, may be null */);

/* (tile, column) of every pixel of one level, row-major (h[level] * w[level] entries each, host) */
int ezb_corr_pyramid_level_map(int H, int W, int L, int level, int32_t* tile, int32_t* column);

/* byte offset inside a group block of record column `column` (0..255) for query `lane` (0..31) of the
 * group; -1 on bad arguments */
int64_t ezb_corr_pyramid_record_offset(int dtype, int lane, int column);

/* bytes of scratch ezb_corr_pyramid_fwd needs (fp16 operand tiles, pooled fmap2 levels, scales) */
int ezb_corr_pyramid_workspace_bytes(int B, int C, int H, int W, int L, size_t* bytes /*host*/);

int ezb_corr_pyramid_fwd(const float* fmap1, const float* fmap2, /* (B,C,H,W) fp32 contiguous */
                         int B, int C, int H, int W, int L, int dtype,
                         void* pyramid /* B * pair_bytes, 128-byte aligned, written */,
                         float* deq_scale /*[B] device, written*/,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Same as ezb_corr_pyramid_fwd; additionally records two CUDA events (cudaEvent_t, may be null) on
 * `stream` immediately before and after the tcgen05 GEMM kernel so a caller can time that launch alone
 * (bench.py's roofline leg).  The prep kernels (absmax / convert) run before ev_gemm_start. */
int ezb_corr_pyramid_fwd_ex(const float* fmap1, const float* fmap2, int B, int C, int H, int W, int L, int dtype,
                            void* pyramid, float* deq_scale, void* workspace, size_t workspace_bytes,
                            void* stream, void* ev_gemm_start, void* ev_gemm_end);

/* thin cudaEvent helpers so ctypes callers need no second CUDA binding */
int ezb_event_create(void** ev /*host out*/);
int ezb_event_destroy(void* ev);
int ezb_event_elapsed_ms(void* ev_start, void* ev_end, float* ms /*host out*/); /* synchronises on ev_end */

/* ---------------------------------------------------------------------------------------------
 * Pyramid lookup (K3) and its adjoint w.r.t. the pyramid (K3b)
 *   replaces MutliScalePairwise4DCorr.__call__ + bilinear_sampler
 *   ezflow/similarity/correlation/pairwise.py:43-69, ezflow/utils/resampling.py:56-85
 *   out (B, L*(2r+1)^2, H, W) fp32; channel l*(2r+1)^2 + i*(2r+1) + j samples level l at
 *   (x/2^l + i - r, y/2^l + j - r) -- the slow window index offsets x.
 * ------------------------------------------------------------------------------------------- */
int ezb_corr_lookup_fwd(const void* pyramid, const float* deq_scale /*[B]*/,
                        const float* coords /* (B,2,H,W): ch0 = x, ch1 = y */,
                        int B, int H, int W, int L, int radius, int dtype,
                        float* out, void* stream);

/* Lookup fused with the 1x1 convolution that consumes it in RAFT's update block:
 *   replaces  corr = corr_fn(coords); cor = F.relu(self.convc1(corr))
 *   ezflow/similarity/correlation/pairwise.py:43-69 + ezflow/decoder/iterative/recurrent_lookup.py:153,182,204
 *   out (B, Cout, H, W) = act( conv1x1(lookup(coords), weight (Cout, L*(2r+1)^2), bias) ),  act = ReLU when relu != 0.
 * The lookup features are produced straight into the A operand of a tcgen05 GEMM (fp16 operands, fp32 accumulation,
 * <= 1e-3 relative) and never written to HBM.  fp16 pyramids, L <= 4, radius <= 4, Cout <= 256; anything else returns
 * EZB_ERR_UNSUPPORTED and the caller runs the two separate ops.
 * The weights are packed once (fp16, K-major, one 96-slot K block per level, scaled by a power of two) into a buffer of
 * ezb_lookup_conv_packed_weight_bytes(); wscale[0] receives the factor that undoes the scaling. */
int ezb_lookup_conv_packed_weight_bytes(int L, size_t* bytes /*host*/);
int ezb_lookup_conv_pack_weights(const float* weight /* (Cout, L*(2r+1)^2) */, int Cout, int L, int radius,
                                 void* packed /* 128-byte aligned, written */, float* wscale /*[1] device, written*/, void* stream);
int ezb_corr_lookup_conv1x1_fwd(const void* pyramid, const float* deq_scale, const float* coords, int B, int H, int W, int L,
                                int radius, int dtype, const void* packed_weight, const float* wscale,
                                const float* bias /*[Cout] or NULL*/, int Cout, int relu, float* out, void* stream);

/* Gradient levels are separate fp32 buffers, row-major per query: element (b,q,y,x) of level l at
 * (b*N+q)*qstride[l] + y*pitch[l] + x with pitch[l] = w[l] and qstride[l] = h[l]*w[l] rounded up to 4 (a 16-byte
 * aligned row per query, so TMA can read a level as a 2-D tensor) (ezb_corr_grad_layout).
 * The lookup adjoint accumulates into them (+=). */
int ezb_corr_grad_layout(int H, int W, int L, int* h, int* w, int* pitch, int64_t* qstride /*[L] host*/);
int ezb_corr_lookup_bwd(const float* dout /* (B, L*(2r+1)^2, H, W) */, const float* coords,
                        int B, int H, int W, int L, int radius,
                        float* const* dlvl_ptrs /*[L] host*/, void* stream);
/* The same for ALL lookups of a step at once (RAFT: 12 per step, models/raft.py:114-129): dlvl += sum_k adjoint(dout_k, coords_k).
 * Per (query, level) the window gradients of the lookups that fall inside a 16 x 20 tile centred on the window of lookup 0
 * (pass the most converged iterate first) are summed in shared memory and each DRAM line is read-modify-written once instead
 * of once per lookup; windows outside the tile (and radius > 4 or L > 4 altogether) take the per-lookup path, so any
 * coordinates and any order are valid. */
int ezb_corr_lookup_bwd_multi(const float* const* dout_ptrs /*[n] host array of device pointers*/,
                              const float* const* coords_ptrs /*[n] host*/, int n_lookups, int B, int H, int W, int L, int radius,
                              float* const* dlvl_ptrs /*[L] host*/, void* stream);
/* Adjoint of ezb_corr_lookup_fwd w.r.t. coords (autograd of `coords / 2**i`, bilinear_sampler's normalisation and
 * grid_sample's grid argument, pairwise.py:50-67, utils/resampling.py:71-83): dcoords (B,2,H,W) fp32 is overwritten.
 * RAFT detaches coords (models/raft.py:115); the entry point exists for callers that do not. */
int ezb_corr_lookup_bwd_coords(const void* pyramid, const float* deq_scale /*[B]*/, const float* coords,
                               const float* dout /* (B, L*(2r+1)^2, H, W) */, int B, int H, int W, int L, int radius,
                               int dtype, float* dcoords, void* stream);

/* Adjoint of the pyramid build (K2b + K1b): dfmap1, dfmap2 (B,C,H,W) fp32 from the gradient levels.
 * With a workspace of ezb_corr_pyramid_bwd_workspace_bytes() (non-zero when C <= 256, L <= 4 and H*W % 4 == 0) both
 * GEMMs run on tcgen05 (kind::tf32) over all levels at once and the gradient levels are left untouched; otherwise
 * (workspace == NULL) the levels are folded into level 0 in place (dlvl[0] is overwritten with the total) and fp32
 * CUDA-core GEMMs run. */
int ezb_corr_pyramid_bwd_workspace_bytes(int B, int C, int H, int W, int L, size_t* bytes /*host*/);
int ezb_corr_pyramid_bwd(float* const* dlvl_ptrs /*[L] host*/, const float* fmap1, const float* fmap2,
                         int B, int C, int H, int W, int L, float* dfmap1, float* dfmap2,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Support of the gradient pyramid.  The reference's autograd materialises the full gradient of the 4-D volume
 * (pairwise.py:31-41 under .backward(): (B*N, 1, H, W) per level), although a query's gradient is non-zero only inside the
 * windows its lookups read (pairwise.py:50-63: (2r+1)^2 taps per level and lookup, 12 lookups per RAFT step,
 * models/raft.py:114-129) — about 1 % of a 1080p pyramid.
 *   ezb_corr_grad_ranges        ranges[b][T][l] = {t0, t1} (int32 pairs, l < 4; ezb_corr_grad_ranges_count int32 in all):
 *                               the 256-target tiles [t0, t1) of level l that a window of any query of the 256-query tile T
 *                               can touch, over the coords of all `n_lookups` lookups of the step ((B,2,H,W) each).
 *   ezb_corr_pyramid_bwd_ex     ezb_corr_pyramid_bwd that skips everything outside `ranges` (NULL: dense).  Only with the
 *                               tensor-core adjoint (ezb_corr_grad_ranges_supported() == 1), else EZB_ERR_UNSUPPORTED.
 *   ezb_corr_grad_rezero        zeroes the ranges again: gradient levels that were all-zero before the step's
 *                               ezb_corr_lookup_bwd calls are all-zero afterwards, without a memset of the whole pyramid. */
int ezb_corr_grad_ranges_supported(int C, int H, int W, int L);
int ezb_corr_grad_ranges_count(int B, int H, int W, int64_t* n_int32 /*host*/);
int ezb_corr_grad_ranges(const float* const* coords_ptrs /*[n_lookups] host array of device pointers*/, int n_lookups,
                         int B, int H, int W, int L, int radius, int32_t* ranges /*device, written*/, void* stream);
int ezb_corr_pyramid_bwd_ex(float* const* dlvl_ptrs /*[L] host*/, const float* fmap1, const float* fmap2,
                            int B, int C, int H, int W, int L, float* dfmap1, float* dfmap2,
                            void* workspace, size_t workspace_bytes, const int32_t* ranges /*device or NULL*/, void* stream);
int ezb_corr_grad_rezero(float* const* dlvl_ptrs /*[L] host*/, const int32_t* ranges /*device*/, int B, int H, int W, int L,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Local displacement-window correlation (K4 / K4b), also CorrelationLayer and the engine of K6
 *   replaces iter_spatial_correlation_sample  ezflow/similarity/correlation/sampler.py:29-129
 *   out[b,pi,pj,y,x] = act( scale * sum_c in1p[b,c,y*sh,x*sw] * in2p[b,c,y*sh+pi*dph-mdh, x*sw+pj*dpw-mdw] )
 *   in*p = inputs zero-padded by (padh,padw); md = dp*(P-1)/2; out (B,Ph,Pw,oh,ow),
 *   oh = ceil((H+2*padh)/sh), ow likewise.  act = leaky_relu(slope) (slope 1 = identity).
 *   `out_batch_stride` (elements) lets K6 write straight into a slice of the concatenated volume.
 *   With a workspace of ezb_local_corr_workspace_bytes() (and C <= 256 per group) the
 *   contraction runs on tcgen05 tensor cores with fp16 operands / fp32 accumulation (<= 1e-3 relative,
 *   the north_star tolerance for cost volumes); with workspace == NULL the exact fp32 CUDA-core kernel
 *   runs (any C).
 * ------------------------------------------------------------------------------------------- */
int ezb_local_corr_workspace_bytes(int BG /* batch * groups */, int C /* per group */, int H, int W, int padh, int padw,
                                   size_t* bytes /*host*/);

int ezb_local_corr_fwd(const float* in1, const float* in2, int B, int C, int H, int W,
                       int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw,
                       float scale, float slope, float* out, int64_t out_batch_stride,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Warp-fused forward (section 8f-2): out = act(scale * corr(in1, warp(in2, flow))) where warp is ezb_warp_fwd's
 * (ezflow/utils/warp.py:5-42) — PWC-Net's decoder warps features2 by the up-sampled flow right before every
 * correlation (ezflow/decoder/pyramid.py:138-141) and applies LeakyReLU(0.1) right after it (:98-102).  The warp is
 * evaluated while the target operand is converted to fp16, so the warped tensor never exists in HBM.  Tensor-core
 * path only: EZB_ERR_UNSUPPORTED when workspace == NULL or C > 256 (callers then run the two kernels). */
int ezb_warp_local_corr_fwd(const float* in1, const float* in2, const float* flow /* (B,2,H,W) */, int B, int C, int H, int W,
                            int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw,
                            float scale, float slope, float* out, int64_t out_batch_stride,
                            void* workspace, size_t workspace_bytes, void* stream);

int ezb_local_corr_bwd(const float* in1, const float* in2, const float* dout, int B, int C, int H, int W,
                       int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw,
                       float scale, int64_t dout_batch_stride,
                       float* din1, float* din2, void* stream);

/* Tensor-core adjoints (K4b on tcgen05): with a workspace of ezb_local_corr_bwd_workspace_bytes(), the regular case —
 * stride 1, no padding, square odd window <= 25, one isotropic patch dilation <= 8, C <= 256 — runs as two banded GEMMs
 * (fp16 operands, fp32 accumulation, <= 1e-3 relative); anything else, or workspace == NULL, takes ezb_local_corr_bwd's exact
 * fp32 kernels. */
int ezb_local_corr_bwd_workspace_bytes(int BG, int C, int H, int W, size_t* bytes /*host*/);
int ezb_local_corr_bwd_ex(const float* in1, const float* in2, const float* dout, int B, int C, int H, int W,
                          int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw,
                          float scale, int64_t dout_batch_stride,
                          float* din1, float* din2, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DCVNet Matryoshka dilated cost volume (K6): all dilations of one scale in one launch
 *   replaces MatryoshkaDilatedCostVolume.forward  ezflow/similarity/learnable_cost.py:306-325
 *   x1,x2 (B, G*Cg, H, W) viewed (B*G, Cg, H, W); out[b, d*G+g, pi, pj, y, x]; P = 2*max_disp+1,
 *   stride s, dilation_patch = dil[d]; optional leaky_relu(0.1) fused (slope).
 * ------------------------------------------------------------------------------------------- */
int ezb_matryoshka_fwd(const float* x1, const float* x2, int B, int G, int Cg, int H, int W,
                       int P, int stride, const int* dil /*[D] host*/, int D, float slope,
                       float* out, int64_t out_batch_stride, void* workspace, size_t workspace_bytes, void* stream);

/* Adjoint of ezb_matryoshka_fwd over ALL dilations in one call (autograd of learnable_cost.py:306-325): dx1, dx2
 * (B, G*Cg, H, W) from dout (B, D*G, P, P, oh, ow).  workspace: ezb_local_corr_bwd_workspace_bytes(B*G, Cg, H, W).
 * use_tensor_cores != 0: dilations the tcgen05 adjoint covers accumulate straight into the outputs; the rest (and all of them
 * when 0) go through the fp32 kernels, which need G == 1 here (EZB_ERR_UNSUPPORTED otherwise: call ezb_local_corr_bwd per
 * dilation and group). */
int ezb_matryoshka_bwd(const float* x1, const float* x2, const float* dout, int B, int G, int Cg, int H, int W,
                       int P, int stride, const int* dil /*[D] host*/, int D, int64_t dout_batch_stride,
                       float* dx1, float* dx2, void* workspace, size_t workspace_bytes, int use_tensor_cores, void* stream);

/* x / (||x||_2 over channels + eps)   learnable_cost.py:426-428 */
int ezb_l2_normalize_fwd(const float* x, int B, int C, int HW, float eps, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Feature warping (K5)   replaces warp  ezflow/utils/warp.py:5-42
 *   out = grid_sample(x, pixel_grid + flow, bilinear, zeros, align_corners) * binarised validity mask
 * ------------------------------------------------------------------------------------------- */
int ezb_warp_fwd(const float* x, const float* flow, int B, int C, int H, int W, float* out, void* stream);
int ezb_warp_bwd(const float* x, const float* flow, const float* dout, int B, int C, int H, int W,
                 float* dx, float* dflow, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Convex flow upsampling (K7)   replaces convex_upsample_flow  ezflow/utils/resampling.py:112-141
 * (called after every RAFT iteration, models/raft.py:127, and by DCVNet,
 * decoder/dilated_flow_stack_filter.py:219)
 *   flow (N,C,H,W) fp32, C in 1..4; mask_logits (N, 9*s*s, H, W) fp32 viewed (N,1,9,s,s,H,W);
 *   out (N,C,s*H,s*W) = sum_k softmax_k(mask_logits) * unfold3x3(flow)       (zero padded)
 * Backward: dflow (N,C,H,W) and / or dmask_logits (same shape as mask_logits); either may be NULL.
 * dflow needs a workspace of ezb_convex_upsample_bwd_workspace_bytes (per-tap partial sums that a
 * second pass gathers, so the result does not depend on atomics ordering).
 * ------------------------------------------------------------------------------------------- */
int ezb_convex_upsample_fwd(const float* flow, const float* mask_logits, int N, int C, int H, int W,
                            int out_stride, float* out, void* stream);
int ezb_convex_upsample_bwd_workspace_bytes(int N, int C, int H, int W, size_t* bytes);
int ezb_convex_upsample_bwd(const float* flow, const float* mask_logits, const float* dout, int N, int C,
                            int H, int W, int out_stride, float* dflow, float* dmask_logits,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * the remaining models' volume assemblies and the memory-free lookup
 *
 * VCN per-channel shifted products    replaces VCN._corr_fn  ezflow/models/vcn.py:162-206
 *   out (B, C, 2*mdx+1, 2*mdy+1, H, W):  out[b,c,i,j,y,x] = leaky_relu(f1[b,c,y,x] * f2[b,c,y+j-mdy,x+i-mdx], slope),
 *   0 where the shifted pixel is outside (mdx = max_disp, mdy = max_disp // factorization).
 * DICL displacement-stacked volume     replaces the assembly in LearnableMatchingCost.forward
 *   ezflow/similarity/learnable_cost.py:157-231 (everything before self.matching_net):
 *   out (B*U*V, 2C, H, W), image (b,i,j) = [x[b] ; y[b] shifted by (j-max_v, i-max_u)] on the overlap, zero elsewhere and
 *   (remove_warp_hole) where the shifted y sums to zero over the channels.  `mask` (B*U*V, H, W) bytes, optional, receives
 *   the keep-mask the backward needs.
 * On-demand correlation lookup: the output of ezb_corr_lookup_fwd (pairwise.py:43-69) computed from the feature maps
 *   without the all-pairs volume (exact fp32); prepare() lays fmap1 and the pooled fmap2 levels out in `workspace`.
 * ------------------------------------------------------------------------------------------- */
int ezb_vcn_corr_fwd(const float* f1, const float* f2, int B, int C, int H, int W, int max_disp_x, int max_disp_y,
                     float slope, float* out, void* stream);
int ezb_vcn_corr_bwd(const float* f1, const float* f2, const float* dout, int B, int C, int H, int W, int max_disp_x,
                     int max_disp_y, float slope, float* df1, float* df2, void* stream);
int ezb_dicl_volume_fwd(const float* x, const float* y, int B, int C, int H, int W, int max_u, int max_v,
                        int remove_warp_hole, float* out, unsigned char* mask /* may be NULL */, void* stream);
int ezb_dicl_volume_bwd(const float* dout, const unsigned char* mask, int B, int C, int H, int W, int max_u, int max_v,
                        float* dx, float* dy, void* stream);
int ezb_corr_ondemand_workspace_bytes(int B, int C, int H, int W, int L, size_t* bytes /*host*/);
int ezb_corr_ondemand_prepare(const float* fmap1, const float* fmap2, int B, int C, int H, int W, int L, void* workspace,
                              size_t workspace_bytes, void* stream);
int ezb_corr_ondemand_lookup(const void* workspace, const float* coords, int B, int C, int H, int W, int L, int radius,
                             float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EZFLOW_B200_H */

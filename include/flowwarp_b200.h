/*
 * flowwarp_b200.h — C-ABI of the B200-native optical-flow backward warp + mask-weighted blend.
 *
 * This is the drop-in boundary for ONE hot path of lzhangbj/deep_video_interpolation_extrapolation:
 *
 *   reference interface replaced                                   file:line (under /root/reference)
 *   ------------------------------------------------------------   ---------------------------------
 *   FlowWrapper.forward(x, flow)   base grid - flow, grid_sample    utils/net_utils.py:93-114
 *   warp(frame, flow, opt, floww, mask)       T frames, gated flow  utils/net_utils.py:116-121
 *   warp_back(frame, flowback, opt, floww, mask)                    utils/net_utils.py:124-129
 *   inline bidirectional warp (border padding)                      nets/OpticalUnet.py:7-15,123-139
 *   mask weighting of the two warps                                 nets/OpticalUnet.py:141-146
 *   autograd of all of the above (ATen grid_sampler_2d_backward)    torch: ATen/native/GridSampler.h:43-83
 *
 * The reference is pure Python over torch; it has no FFI of its own.  The binding a maintainer
 * adds is the ctypes stub in deep_video_interpolation_extrapolation_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers + sizes + strides; no torch types.  All tensors are fp32.  Strides are in ELEMENTS.
 *     The innermost (W) stride of every tensor is 1.
 *   - logical batch is (N, T): T = opt.vid_length frames that share one launch.  A tensor that is
 *     shared by all T frames (the single source frame of warp()) passes a T-stride of 0.
 *   - every function launches asynchronously on `stream` (a cudaStream_t passed as void*), allocates
 *     no device memory, never throws and never aborts.  Process-wide state is limited to three mutex-guarded caches: the A/B
 *     environment knobs (FWB_KERNELS, FWB_TILE_*: read once, fwb_reload_env() re-reads), the per-(kernel, device) record
 *     of the dynamic shared-memory attribute already requested from the CUDA runtime, and the texture-object descriptors
 *     of dense source tensors (fwb_release_cache() destroys them).
 *   - limits: N*T <= 65535 (grid.z), H, W <= 32767, every in-plane offset (C*channel stride + (H+2)*row stride) < 2^31
 *     for sources, outputs and gradients; violations return FWB_E_SHAPE / FWB_E_RANGE.
 *   - return value: 0 = ok; > 0 = a cudaError_t from a launch; < 0 = FWB_E_* argument error.
 *   - sampling arithmetic (bit-exact contract, see DESIGN.md "Coordinate arithmetic"):
 *        bx[j]   = linspace(-1,1,W)[j]  (CPU torch.linspace bit pattern; -1 when W == 1)
 *        f       = gate ? flow*gate : flow                  (one fp32 rounding)
 *        gx      = sign < 0 ? bx - f : bx + f
 *        ix      = align_corners ? ((gx+1)/2)*(W-1) : fma(gx+1, W, -1)/2
 *        border  : ix = min(W-1, max(ix, 0));  non-finite or out-of-int-range -> -100
 *                  (a NaN flow under border padding samples at coordinate 0 in the forward, as ATen's forward does; the backward
 *                  keeps those taps, whereas ATen's backward sends the NaN coordinate to -100 and returns zero gradients there:
 *                  a known deviation for NaN flows only)
 *        x0      = floor(ix); taps (x0,y0) (x0+1,y0) (x0,y0+1) (x0+1,y0+1); valid bit per tap = in image
 *        out[c]  = sum_d  blend_d * ( v_nw*nw + v_ne*ne + v_sw*sw + v_se*se )   in that order
 */
#ifndef FLOWWARP_B200_H
#define FLOWWARP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FWB_VERSION 0x00010000 /* major.minor.patch = 1.0.0 */
#define FWB_MAX_GROUPS 4       /* channel groups that share flow/mask (RGB, seg, ...) */

/* padding_mode — torch F.grid_sample padding_mode (utils/net_utils.py:113 uses zeros,
 * nets/OpticalUnet.py:135,139 uses border) */
#define FWB_PAD_ZEROS 0
#define FWB_PAD_BORDER 1

/* flags */
#define FWB_FLAG_DETERMINISTIC 1u /* grad_src must be bit-exact run to run (kernel 3 as an owner gather, fwb_owner.cuh) */
#define FWB_FLAG_ATOMIC_SRC 2u    /* grad_src by global atomics (ATen-style scatter; non-deterministic; for A/B runs) */
#define FWB_FLAG_FUSED_BWD 4u     /* fwb_warp_blend_backward_flow also produces grad_src (kernels 2+3 fused: shared-memory
                                   * fixed-point tiles + vector reductions, non-deterministic); fwb_warp_blend_backward_src
                                   * then returns at once.  Ignored with FWB_FLAG_DETERMINISTIC / FWB_FLAG_ATOMIC_SRC. */
#define FWB_FLAG_GRAD_SRC_ZEROED 8u /* with FWB_FLAG_FUSED_BWD: every grad_src plane is already zero on entry (the caller
                                   * zeroed it earlier, e.g. on a side stream while the forward ran), so the library skips
                                   * its own memsets.  The fused backward ACCUMULATES into grad_src. */

/* argument errors (negative return values) */
#define FWB_E_NULL -1      /* a required pointer is NULL */
#define FWB_E_SHAPE -2     /* N,T,H,W,C out of range (empty spatial dims are an error, as in torch) */
#define FWB_E_DIRS -3      /* n_dirs not in {1,2} */
#define FWB_E_GROUPS -4    /* n_groups not in 1..FWB_MAX_GROUPS */
#define FWB_E_MODE -5      /* unknown padding_mode / align_corners / sign */
#define FWB_E_ALIGN -6     /* a pointer is not 4-byte aligned */
#define FWB_E_WORKSPACE -7 /* workspace missing or too small */
#define FWB_E_RANGE -8     /* H or W above 32767 (segment tables are int16) */

/* One warp direction: its flow, optional flow-gating mask (utils/net_utils.py:118 `flow*mask`)
 * and optional blend weight (nets/OpticalUnet.py:141-146). */
typedef struct fwb_dir {
  const float* flow; /* [N,2,T,H,W]: channel 0 horizontal, 1 vertical, normalised units */
  int64_t flow_sn, flow_sc, flow_st, flow_sh;
  const float* gate; /* optional [N,T,H,W] */
  int64_t gate_sn, gate_st, gate_sh;
  const float* blend; /* optional [N,T,H,W] */
  int64_t blend_sn, blend_st, blend_sh;
  float sign; /* -1: grid = base - flow (warp, forward flow); +1: grid = base + flow (warp_back) */
  int32_t _pad;
} fwb_dir;

/* One channel group: per-direction source planes and the output planes. */
typedef struct fwb_group {
  int32_t C;
  int32_t _pad;
  const float* src[2]; /* [N,T,C,H,W] (T-stride 0 when one frame feeds all T) */
  int64_t src_sn[2], src_st[2], src_sc[2], src_sh[2];
  float* out; /* [N,T,C,H,W] */
  int64_t out_sn, out_st, out_sc, out_sh;
} fwb_group;

typedef struct fwb_problem {
  int32_t N, T, H, W;
  int32_t n_dirs;   /* 1 or 2 */
  int32_t n_groups; /* 1..FWB_MAX_GROUPS */
  int32_t padding_mode;
  int32_t align_corners;
  uint32_t flags;
  int32_t _pad;
  fwb_dir dir[2];
  fwb_group grp[FWB_MAX_GROUPS];
} fwb_problem;

/* Gradient buffers for the backward entry points.  Any output pointer may be NULL (= not needed). */
typedef struct fwb_grads {
  const float* grad_out[FWB_MAX_GROUPS]; /* [N,T,C,H,W], strides below */
  int64_t go_sn[FWB_MAX_GROUPS], go_st[FWB_MAX_GROUPS], go_sc[FWB_MAX_GROUPS], go_sh[FWB_MAX_GROUPS];
  float* grad_src[FWB_MAX_GROUPS][2]; /* same logical shape as src; T-stride 0 => summed over T */
  int64_t gs_sn[FWB_MAX_GROUPS][2], gs_st[FWB_MAX_GROUPS][2], gs_sc[FWB_MAX_GROUPS][2],
      gs_sh[FWB_MAX_GROUPS][2];
  float* grad_flow[2]; /* [N,2,T,H,W] */
  int64_t gf_sn[2], gf_sc[2], gf_st[2], gf_sh[2];
  float* grad_gate[2]; /* [N,T,H,W] */
  int64_t gg_sn[2], gg_st[2], gg_sh[2];
  float* grad_blend[2]; /* [N,T,H,W] */
  int64_t gb_sn[2], gb_st[2], gb_sh[2];
} fwb_grads;

/* The mask blend that follows the warp in `refine` (utils/net_utils.py:141-143):
 *     out[n,t,c] = input[n,t,c] * mask[n,t] + noise[n,c] * (1 - mask[n,t])
 * with the `cat([noise_bg, zeros(20)])` of `:134-136` folded in: only the first Cn <= C channels have a noise plane,
 * the others blend against zero.  All strides in elements, W-stride 1.  Forward reads input/mask/noise and writes
 * out; backward reads grad_out (+ input, mask, noise) and writes whichever of grad_input / grad_mask / grad_noise
 * is non-NULL.  Both are single streaming passes, deterministic (no atomics). */
typedef struct fwb_blend {
  int32_t N, T, C, Cn, H, W;
  const float* input; /* [N,T,C,H,W] */
  int64_t in_sn, in_st, in_sc, in_sh;
  const float* mask; /* [N,T,H,W] */
  int64_t m_sn, m_st, m_sh;
  const float* noise; /* [N,Cn,H,W]; NULL <=> Cn = 0 */
  int64_t nz_sn, nz_sc, nz_sh;
  float* out; /* [N,T,C,H,W] (forward) */
  int64_t out_sn, out_st, out_sc, out_sh;
  const float* grad_out; /* [N,T,C,H,W] (backward) */
  int64_t go_sn, go_st, go_sc, go_sh;
  float* grad_input; /* [N,T,C,H,W] or NULL */
  int64_t gi_sn, gi_st, gi_sc, gi_sh;
  float* grad_mask; /* [N,T,H,W] or NULL */
  int64_t gm_sn, gm_st, gm_sh;
  float* grad_noise; /* [N,Cn,H,W] or NULL */
  int64_t gn_sn, gn_sc, gn_sh;
} fwb_blend;

/* Compact segmentation format: the K-class map is given as uint8 LABELS [N,T,H,W] instead of the
 * K-channel one-hot float tensor the reference's loader builds (folder.py:193-200) and warps with the RGB frame's flow
 * (nets/VAE_S.py:134-135, nets/InterNet.py:15-18).  out[N,T,K,H,W] is bit-identical to the dense op applied to
 * one_hot(labels); labels >= K belong to no class.  There is no gradient w.r.t. the labels; the backward produces the
 * label warp's contribution to grad_flow / grad_gate / grad_blend, overwriting them (accumulate = 0) or adding to what
 * the RGB backward already wrote there (accumulate = 1, same stream).  `dir` as in fwb_problem. */
typedef struct fwb_label_problem {
  int32_t N, T, H, W;
  int32_t n_dirs; /* 1 or 2 */
  int32_t K;      /* classes = output channels, 1..256 */
  int32_t padding_mode;
  int32_t align_corners;
  fwb_dir dir[2];
  const uint8_t* labels[2]; /* per direction, [N,T,H,W] uint8 (T-stride 0 when one map feeds all T) */
  int64_t lab_sn[2], lab_st[2], lab_sh[2];
  float* out; /* [N,T,K,H,W] (forward) */
  int64_t out_sn, out_st, out_sc, out_sh;
  const float* grad_out; /* [N,T,K,H,W] (backward) */
  int64_t go_sn, go_st, go_sc, go_sh;
  float* grad_flow[2]; /* [N,2,T,H,W] or NULL */
  int64_t gf_sn[2], gf_sc[2], gf_st[2], gf_sh[2];
  float* grad_gate[2]; /* [N,T,H,W] or NULL */
  int64_t gg_sn[2], gg_st[2], gg_sh[2];
  float* grad_blend[2]; /* [N,T,H,W] or NULL */
  int64_t gb_sn[2], gb_st[2], gb_sh[2];
  int32_t accumulate;
  int32_t _pad;
} fwb_label_problem;

/* A strided fp32 view [N,T,C,H,W] (element strides, W-stride 1) for the loss entry points: a frame slice flow[:, :, i] of the
 * reference's [N,2,T,H,W] layout is a view with sc = the T*H*W plane stride and st = H*W: no copy. */
typedef struct fwb_view {
  float* ptr; /* NULL: tensor absent */
  int64_t sn, st, sc, sh;
} fwb_view;

/* Library version (FWB_VERSION of the build). */
int32_t fwb_version(void);

/* Human-readable text for a return code of this library (static storage). */
const char* fwb_strerror(int32_t code);

/* Re-read the A/B environment knobs (FWB_KERNELS, FWB_TILE_*).  They are otherwise read once per process.  Test / A-B hook:
 * do not call it while another thread is inside a launch function. */
void fwb_reload_env(void);

/* Destroy the cached texture objects (the forward / backward read dense sources through the texture units; the descriptors
 * are created once per (pointer, extent, pitch, device) and kept).  Call only when no kernel of this library is in flight. */
void fwb_release_cache(void);

/* Kernel 1 — fused forward: flow->coordinate, floor/fraction, validity, 4-tap bilinear gather
 * over every channel of every group, for 1 or 2 directions, and the blend-weighted sum.
 * Replaces utils/net_utils.py:93-121 (+124-129) and nets/OpticalUnet.py:123-146. */
int32_t fwb_warp_blend_forward(const fwb_problem* p, void* stream);

/* The same forward, which additionally sets every grad_src plane named in `g` to zero (only g->grad_src and its
 * strides are read).  torch zero-fills the gradient of grid_sample's input before its atomicAdd scatter
 * (ATen grid_sampler_2d_backward, reached from utils/net_utils.py:113 by autograd); here the forward's tiles do it
 * with spare store bandwidth, so the fused backward can be called with FWB_FLAG_GRAD_SRC_ZEROED and needs no memset
 * pass.  Planes the kernel cannot clear in place (unaligned / odd strides) are cleared by memsets on `stream`. */
int32_t fwb_warp_blend_forward_zero(const fwb_problem* p, const fwb_grads* g, void* stream);

/* `refine`'s mask blend (utils/net_utils.py:141-143), see fwb_blend above. */
int32_t fwb_mask_blend_forward(const fwb_blend* b, void* stream);
int32_t fwb_mask_blend_backward(const fwb_blend* b, void* stream);

/* Label-map warp (+ gate) (+ blend), see fwb_label_problem above. */
int32_t fwb_label_warp_blend_forward(const fwb_label_problem* p, void* stream);
int32_t fwb_label_warp_blend_backward(const fwb_label_problem* p, void* stream);

/* Debug / parity: integer sample indices and validity bits of direction `d`.
 * x0,y0: int32 [N,T,H,W] contiguous; valid: uint8 [N,T,H,W], bit0 nw, bit1 ne, bit2 sw, bit3 se.
 * Same device function as the forward kernel. */
int32_t fwb_sample_indices(const fwb_problem* p, int32_t d, int32_t* x0, int32_t* y0,
                           uint8_t* valid, float* ix, float* iy, void* stream);

/* Bytes of scratch the backward entry points need for this problem (segment tables of the
 * owner-gather kernel); 0 is a valid answer. */
size_t fwb_workspace_bytes(const fwb_problem* p);

/* Kernel 2 — gradient w.r.t. flow, gate and blend weight, as a gather (no atomics).
 * Replaces autograd of grid_sample's grid input + the mul/sub/transposes around it
 * (utils/net_utils.py:109-113,118).  Also fills the per-segment tap bounding boxes in
 * `workspace` that kernel 3 consumes (so call it before fwb_warp_blend_backward_src). */
int32_t fwb_warp_blend_backward_flow(const fwb_problem* p, const fwb_grads* g, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* Kernel 3 — gradient w.r.t. the sources (image / segmentation planes).
 * Replaces the atomicAdd scatter of ATen grid_sampler_2d_backward. */
int32_t fwb_warp_blend_backward_src(const fwb_problem* p, const fwb_grads* g, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* The flow-regularisation losses that consume the warp.  Their Python source is deleted from the reference;
 * the formulas are read from the bytecode in __pycache__/losses.cpython-36.pyc and use utils/net_utils.py:243-248
 * (gradientx / gradienty) and the FlowWrapper held as self.flowwarp (runners/VAEer.py:53):
 *   TrainingLoss._flowgradloss (pyc line 413):  flow*128, image*256,  w = exp(-mean_c |grad(image)|),
 *       loss = mean |gradientx(flow) * wx| + mean |gradienty(flow) * wy|, summed over the T frames of the views / T
 *   TrainingLoss._flowconsist (pyc line 481):   mean(mask * |a - b|) with a = flowwarp(...) (this library's warp) per term
 * partials: scratch of fwb_loss_partials_bytes(N, T, H) bytes; loss: one device float.  Deterministic (fixed-order sums). */
size_t fwb_loss_partials_bytes(int32_t N, int32_t T, int32_t H);
int32_t fwb_flowgrad_loss_forward(const fwb_view* flow /* C = 2 */, const fwb_view* image, int32_t N, int32_t T, int32_t C,
                                  int32_t H, int32_t W, void* partials, float* loss, void* stream);
/* gradient w.r.t. flow (the image is data); grad_loss: one device float */
int32_t fwb_flowgrad_loss_backward(const fwb_view* flow, const fwb_view* image, int32_t N, int32_t T, int32_t C, int32_t H,
                                   int32_t W, const float* grad_loss, const fwb_view* grad_flow, void* stream);
/* loss = sum_t mean_{n,c,h,w}( mask * |a - b| ) (mask [N,T,1,H,W] view or NULL) */
int32_t fwb_masked_abs_forward(const fwb_view* a, const fwb_view* b, const fwb_view* mask, int32_t N, int32_t T, int32_t C,
                              int32_t H, int32_t W, void* partials, float* loss, void* stream);
int32_t fwb_masked_abs_backward(const fwb_view* a, const fwb_view* b, const fwb_view* mask, int32_t N, int32_t T, int32_t C,
                               int32_t H, int32_t W, const float* grad_loss, const fwb_view* grad_a, const fwb_view* grad_b,
                               const fwb_view* grad_mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWWARP_B200_H */

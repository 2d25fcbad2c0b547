// fwb_label.cuh — the warp (+ gate) (+ blend) of a K-class segmentation map given as uint8 LABELS instead of the
// K-channel one-hot float tensor the reference builds in its loader (folder.py:193-200) and warps with the same
// flow as the RGB frame (nets/VAE_S.py:134-135, nets/InterNet.py:15-18).  SURVEY.md §8(f) row 4.
//
// Result: bit-identical to warping one_hot(labels) with the dense kernels.  Why: a one-hot tap value is exactly 0 or 1,
// so ATen's accumulation  r = v_nw*w_nw; r = fma(v_ne, w_ne, r); r = fma(v_sw, w_sw, r); r = fma(v_se, w_se, r)
// degenerates to "add the weights of the taps whose label is c, in the order nw, ne, sw, se" (fma(1, w, r) is one rounding
// of r + w, fma(0, w, r) is r) — which is what label_value() does.  What changes is the traffic: the source read is
// 1 byte per tap instead of 4 K bytes, there is no gradient w.r.t. the source (labels are not differentiable; the reference
// re-one-hots by argmax and lets the gradient flow through RGB only, runners/ExtraTrainer.py:254-310), and the coordinate
// gradient needs grad_out only at the (at most) 4 classes under the 4 taps:
//     gix = sum_c g_c (uy (v_ne,c - v_nw,c) + ty (v_se,c - v_sw,c)) = uy (g[l_ne] - g[l_nw]) + ty (g[l_se] - g[l_sw])
// Algorithmic bytes per pixel (2 directions): forward 2*(1*~1) + 24 + 4K  vs  12K + 24 dense.
//
// One thread per output pixel, lanes along W (all K output planes are written with full 128-byte lines); labels come
// through the read-only path (a [H,W] uint8 map is 1/80 of the one-hot tensor and stays in L1/L2).
#pragma once
#include "fwb_coords.cuh"

namespace fwb {

struct LabelP {
  Geo geo;
  DirP dir[2];
  int K;
  const uint8_t* lab[2];
  long long lab_sn[2], lab_st[2];
  int lab_sh[2];
  float* out;
  long long out_sn, out_st;
  int out_sc, out_sh;
  const float* go;
  long long go_sn, go_st;
  int go_sc, go_sh;
  float* grad_flow[2];
  long long gf_sn[2], gf_sc[2], gf_st[2], gf_sh[2];
  float* grad_gate[2];
  long long gg_sn[2], gg_st[2], gg_sh[2];
  float* grad_blend[2];
  long long gb_sn[2], gb_st[2], gb_sh[2];
  int accumulate;
};

constexpr int LB_THREADS = 256;  // 32 x 8 pixels per CTA
constexpr int LB_NOLABEL = 0xffff;

// labels under the 4 taps (LB_NOLABEL where the tap is outside the image: it contributes 0, like zeros padding)
__device__ __forceinline__ void tap_labels(const uint8_t* __restrict__ lab, int sh, const Tap& k, int* l) {
  const uint8_t* p = lab + (long long)k.y0 * sh + k.x0;
  l[0] = (k.valid & 1u) ? (int)__ldg(p) : LB_NOLABEL;
  l[1] = (k.valid & 2u) ? (int)__ldg(p + 1) : LB_NOLABEL;
  l[2] = (k.valid & 4u) ? (int)__ldg(p + sh) : LB_NOLABEL;
  l[3] = (k.valid & 8u) ? (int)__ldg(p + sh + 1) : LB_NOLABEL;
}

// value of class c of the warped one-hot map: the dense `bilinear()` with tap values in {0, 1}
__device__ __forceinline__ float label_value(const int* l, const float* w, int c) {
  float r = (l[0] == c) ? w[0] : 0.0f;  // __fmul_rn(1, w) == w, __fmul_rn(0, w) == 0
  r = (l[1] == c) ? __fadd_rn(r, w[1]) : r;
  r = (l[2] == c) ? __fadd_rn(r, w[2]) : r;
  r = (l[3] == c) ? __fadd_rn(r, w[3]) : r;
  return r;
}

template <int NDIRS>
__global__ void __launch_bounds__(LB_THREADS) label_fwd_kernel(const __grid_constant__ LabelP P) {
  const Geo& G = P.geo;
  const int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (j >= G.W || i >= G.H) return;
  const int n = blockIdx.z / G.T, t = blockIdx.z - n * G.T;
  int l[NDIRS][4];
  float w[NDIRS][4], bl[NDIRS];
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    compute_tap(G, P.dir[d], n, t, i, j, k);
    w[d][0] = __fmul_rn(k.ux, k.uy);
    w[d][1] = __fmul_rn(k.tx, k.uy);
    w[d][2] = __fmul_rn(k.ux, k.ty);
    w[d][3] = __fmul_rn(k.tx, k.ty);
    bl[d] = k.blend;
    has_bl[d] = P.dir[d].blend != nullptr;
    tap_labels(P.lab[d] + n * P.lab_sn[d] + t * P.lab_st[d], P.lab_sh[d], k, l[d]);
  }
  float* op = P.out + n * P.out_sn + t * P.out_st + (long long)i * P.out_sh + j;
#pragma unroll 4
  for (int c = 0; c < P.K; ++c) {
    float r = 0.0f;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      float a = label_value(l[d], w[d], c);
      if (has_bl[d]) a = __fmul_rn(a, bl[d]);
      r = (d == 0) ? a : __fadd_rn(r, a);
    }
    __stcs(op + (long long)c * P.out_sc, r);
  }
}

// coordinate gradient of the label warp -> grad_flow / grad_gate / grad_blend (written, or added when P.accumulate:
// the same thread owns the same cells in the RGB backward that ran before on the stream, so += needs no atomics)
template <int NDIRS>
__global__ void __launch_bounds__(LB_THREADS) label_bwd_kernel(const __grid_constant__ LabelP P) {
  const Geo& G = P.geo;
  const int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (j >= G.W || i >= G.H) return;
  const int n = blockIdx.z / G.T, t = blockIdx.z - n * G.T;
  const float* gp = P.go + n * P.go_sn + t * P.go_st + (long long)i * P.go_sh + j;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    compute_tap(G, P.dir[d], n, t, i, j, k);
    int l[4];
    tap_labels(P.lab[d] + n * P.lab_sn[d] + t * P.lab_st[d], P.lab_sh[d], k, l);
    float g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q] = (l[q] < P.K) ? __ldg(gp + (long long)l[q] * P.go_sc) : 0.0f;  // labels >= K: no class
    const bool has_bl = P.dir[d].blend != nullptr;
    // grad_blend = sum_c g_c * warped_c = sum over taps of w_tap * g[l_tap]
    const float gbl = fmaf(g[3], k.tx * k.ty, fmaf(g[2], k.ux * k.ty, fmaf(g[1], k.tx * k.uy, g[0] * (k.ux * k.uy))));
    const float s = has_bl ? k.blend : 1.0f;
    const float gix = s * fmaf(k.ty, g[3] - g[2], k.uy * (g[1] - g[0]));
    const float giy = s * fmaf(k.tx, g[3] - g[1], k.ux * (g[2] - g[0]));
    float gfx = k.mx * gix, gfy = k.my * giy;
    if (P.dir[d].sign < 0.0f) {
      gfx = -gfx;
      gfy = -gfy;
    }
    const bool gated = P.dir[d].gate != nullptr, acc = P.accumulate != 0;
    if (P.grad_gate[d] && gated) {
      float* o = P.grad_gate[d] + n * P.gg_sn[d] + t * P.gg_st[d] + (long long)i * P.gg_sh[d] + j;
      const float v = __fadd_rn(__fmul_rn(gfx, k.fx), __fmul_rn(gfy, k.fy));
      *o = acc ? *o + v : v;
    }
    if (P.grad_flow[d]) {
      float* o = P.grad_flow[d] + n * P.gf_sn[d] + t * P.gf_st[d] + (long long)i * P.gf_sh[d] + j;
      const float vx = gated ? gfx * k.gate : gfx, vy = gated ? gfy * k.gate : gfy;
      o[0] = acc ? o[0] + vx : vx;
      o[P.gf_sc[d]] = acc ? o[P.gf_sc[d]] + vy : vy;
    }
    if (P.grad_blend[d] && has_bl) {
      float* o = P.grad_blend[d] + n * P.gb_sn[d] + t * P.gb_st[d] + (long long)i * P.gb_sh[d] + j;
      *o = acc ? *o + gbl : gbl;
    }
  }
}

}  // namespace fwb

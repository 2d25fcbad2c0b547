// fwb_generic.cuh — the generic (global-memory gather) per-pixel bodies of kernels 1 and 2.
// Used (a) by the plain one-thread-per-pixel kernels that serve tensors the staged kernels cannot take
// (pointers / strides not 16-byte aligned) and (b) as the in-kernel path of a staged CTA whose source
// footprint does not fit in shared memory (wild flows).
#pragma once
#include "fwb_coords.cuh"

namespace fwb {

__device__ __forceinline__ float ldg_if(const float* p, bool ok) { return ok ? __ldg(p) : 0.0f; }

// 4-tap bilinear value; accumulation order nw, ne, sw, se (torch:_decomp/decompositions.py:4515-4537)
__device__ __forceinline__ float bilinear(const float* __restrict__ s, int o, int sh, unsigned v, float wnw,
                                          float wne, float wsw, float wse) {
  const float a = ldg_if(s + o, v & 1u), b = ldg_if(s + o + 1, v & 2u);
  const float c = ldg_if(s + o + sh, v & 4u), d = ldg_if(s + o + sh + 1, v & 8u);
  float r = __fmul_rn(a, wnw);
  r = __fmaf_rn(b, wne, r);
  r = __fmaf_rn(c, wsw, r);
  r = __fmaf_rn(d, wse, r);
  return r;
}

// Kernel 1 body for one output pixel: all directions, all channel groups.
template <int NDIRS>
__device__ __forceinline__ void fwd_generic_pixel(const Params& P, int n, int t, int i, int j) {
  const Geo& G = P.geo;
  float w[NDIRS][4], bl[NDIRS];
  int x0[NDIRS], y0[NDIRS];
  unsigned v[NDIRS];
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    compute_tap(G, P.dir[d], n, t, i, j, k);
    w[d][0] = __fmul_rn(k.ux, k.uy);
    w[d][1] = __fmul_rn(k.tx, k.uy);
    w[d][2] = __fmul_rn(k.ux, k.ty);
    w[d][3] = __fmul_rn(k.tx, k.ty);
    x0[d] = k.x0;
    y0[d] = k.y0;
    v[d] = k.valid;
    bl[d] = k.blend;
    has_bl[d] = P.dir[d].blend != nullptr;
  }
  for (int g = 0; g < G.n_groups; ++g) {
    const GroupP& R = P.grp[g];
    const float* s[NDIRS];
    int o[NDIRS];
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      s[d] = R.src[d] + n * R.src_sn[d] + t * R.src_st[d];
      o[d] = y0[d] * R.src_sh[d] + x0[d];
    }
    float* out = R.out + n * R.out_sn + t * R.out_st + (long long)i * R.out_sh + j;
#pragma unroll 4
    for (int c = 0; c < R.C; ++c) {
      float r = 0.0f;
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        float a = bilinear(s[d] + (long long)c * R.src_sc[d], o[d], R.src_sh[d], v[d], w[d][0], w[d][1],
                           w[d][2], w[d][3]);
        if (has_bl[d]) a = __fmul_rn(a, bl[d]);
        r = (d == 0) ? a : __fadd_rn(r, a);
      }
      __stcs(out + (long long)c * R.out_sc, r);
    }
  }
}

// Kernel 2 epilogue for one (pixel, direction): coordinate gradient -> grad_flow / grad_gate / grad_blend.
__device__ __forceinline__ void bwdflow_store(const Params& P, const GradP& Q, int d, int n, int t, int i, int j,
                                              const Tap& k, float gix, float giy, float gbl) {
  float gfx = k.mx * gix, gfy = k.my * giy;
  if (P.dir[d].sign < 0.0f) {
    gfx = -gfx;
    gfy = -gfy;
  }
  const bool gated = P.dir[d].gate != nullptr;
  if (Q.grad_gate[d] && gated)
    Q.grad_gate[d][n * Q.gg_sn[d] + t * Q.gg_st[d] + (long long)i * Q.gg_sh[d] + j] =
        __fadd_rn(__fmul_rn(gfx, k.fx), __fmul_rn(gfy, k.fy));
  if (Q.grad_flow[d]) {
    float* o = Q.grad_flow[d] + n * Q.gf_sn[d] + t * Q.gf_st[d] + (long long)i * Q.gf_sh[d] + j;
    o[0] = gated ? gfx * k.gate : gfx;
    o[Q.gf_sc[d]] = gated ? gfy * k.gate : gfy;
  }
  if (Q.grad_blend[d] && P.dir[d].blend != nullptr)
    Q.grad_blend[d][n * Q.gb_sn[d] + t * Q.gb_st[d] + (long long)i * Q.gb_sh[d] + j] = gbl;
}

// Kernel 2 body for one output pixel.
//   gix = sum_c gw_c * [ uy*(v_ne - v_nw) + ty*(v_se - v_sw) ]
//   giy = sum_c gw_c * [ ux*(v_sw - v_nw) + tx*(v_se - v_ne) ]       (OOB tap value = 0)
// which is ATen's grid_sampler_2d_backward accumulation with the common factors pulled out.
template <int NDIRS>
__device__ __forceinline__ void bwdflow_generic_pixel(const Params& P, const GradP& Q, int n, int t, int i, int j) {
  const Geo& G = P.geo;
  Tap k[NDIRS];
  float gix[NDIRS], giy[NDIRS], gbl[NDIRS];
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    compute_tap(G, P.dir[d], n, t, i, j, k[d]);
    gix[d] = giy[d] = gbl[d] = 0.0f;
    has_bl[d] = P.dir[d].blend != nullptr;
  }
  for (int g = 0; g < G.n_groups; ++g) {
    const GroupP& R = P.grp[g];
    if (!Q.grad_out[g]) continue;
    const float* go = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)i * Q.go_sh[g] + j;
    const float* s[NDIRS];
    int o[NDIRS];
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      s[d] = R.src[d] + n * R.src_sn[d] + t * R.src_st[d];
      o[d] = k[d].y0 * R.src_sh[d] + k[d].x0;
    }
#pragma unroll 2
    for (int c = 0; c < R.C; ++c) {
      const float gout = __ldg(go + (long long)c * Q.go_sc[g]);
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const float* sp = s[d] + (long long)c * R.src_sc[d] + o[d];
        const int sh = R.src_sh[d];
        const unsigned v = k[d].valid;
        const float a = ldg_if(sp, v & 1u), b = ldg_if(sp + 1, v & 2u);
        const float cc = ldg_if(sp + sh, v & 4u), dd = ldg_if(sp + sh + 1, v & 8u);
        float gw = gout;
        if (has_bl[d]) {
          const float top = fmaf(b, k[d].tx, a * k[d].ux), bot = fmaf(dd, k[d].tx, cc * k[d].ux);
          gbl[d] = fmaf(gout, fmaf(bot, k[d].ty, top * k[d].uy), gbl[d]);
          gw = gout * k[d].blend;
        }
        gix[d] = fmaf(gw, fmaf(k[d].ty, dd - cc, k[d].uy * (b - a)), gix[d]);
        giy[d] = fmaf(gw, fmaf(k[d].tx, dd - b, k[d].ux * (cc - a)), giy[d]);
      }
    }
  }
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) bwdflow_store(P, Q, d, n, t, i, j, k[d], gix[d], giy[d], gbl[d]);
}

// flattened channel index (over all groups) -> group, channel
__device__ __forceinline__ void chan_lookup(const Params& P, int cf, int& g, int& c) {
  g = 0;
  c = cf;
  while (g + 1 < P.geo.n_groups && c >= P.grp[g].C) {
    c -= P.grp[g].C;
    ++g;
  }
}

// Kernel 1 body for one output pixel, channels first, first+step, ... of the flattened channel list (a warp
// serves one pixel with lane = first, step = 32).
template <int NDIRS>
__device__ __forceinline__ void fwd_generic_pixel_strided(const Params& P, int n, int t, int i, int j, int first,
                                                          int step) {
  const Geo& G = P.geo;
  float w[NDIRS][4], bl[NDIRS];
  int x0[NDIRS], y0[NDIRS];
  unsigned v[NDIRS];
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    compute_tap(G, P.dir[d], n, t, i, j, k);
    w[d][0] = __fmul_rn(k.ux, k.uy);
    w[d][1] = __fmul_rn(k.tx, k.uy);
    w[d][2] = __fmul_rn(k.ux, k.ty);
    w[d][3] = __fmul_rn(k.tx, k.ty);
    x0[d] = k.x0;
    y0[d] = k.y0;
    v[d] = k.valid;
    bl[d] = k.blend;
    has_bl[d] = P.dir[d].blend != nullptr;
  }
  int Ctot = 0;
  for (int g = 0; g < G.n_groups; ++g) Ctot += P.grp[g].C;
  for (int cf = first; cf < Ctot; cf += step) {
    int g, c;
    chan_lookup(P, cf, g, c);
    const GroupP& R = P.grp[g];
    float r = 0.0f;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const float* s = R.src[d] + n * R.src_sn[d] + t * R.src_st[d] + (long long)c * R.src_sc[d];
      float a = bilinear(s, y0[d] * R.src_sh[d] + x0[d], R.src_sh[d], v[d], w[d][0], w[d][1], w[d][2], w[d][3]);
      if (has_bl[d]) a = __fmul_rn(a, bl[d]);
      r = (d == 0) ? a : __fadd_rn(r, a);
    }
    __stcs(R.out + n * R.out_sn + t * R.out_st + (long long)c * R.out_sc + (long long)i * R.out_sh + j, r);
  }
}

// Kernel 2 body for one output pixel served by a whole warp: lane l takes channels l, l+32, ...; the partial
// sums are combined in a fixed butterfly order (deterministic).
template <int NDIRS>
__device__ __forceinline__ void bwdflow_generic_pixel_warp(const Params& P, const GradP& Q, int n, int t, int i,
                                                           int j) {
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31;
  Tap k[NDIRS];
  float gix[NDIRS], giy[NDIRS], gbl[NDIRS];
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    compute_tap(G, P.dir[d], n, t, i, j, k[d]);
    gix[d] = giy[d] = gbl[d] = 0.0f;
    has_bl[d] = P.dir[d].blend != nullptr;
  }
  int Ctot = 0;
  for (int g = 0; g < G.n_groups; ++g) Ctot += P.grp[g].C;
  for (int cf = lane; cf < Ctot; cf += 32) {
    int g, c;
    chan_lookup(P, cf, g, c);
    if (!Q.grad_out[g]) continue;
    const GroupP& R = P.grp[g];
    const float gout = __ldg(Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)c * Q.go_sc[g] +
                             (long long)i * Q.go_sh[g] + j);
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const int sh = R.src_sh[d];
      const float* sp = R.src[d] + n * R.src_sn[d] + t * R.src_st[d] + (long long)c * R.src_sc[d] + k[d].y0 * sh + k[d].x0;
      const unsigned v = k[d].valid;
      const float a = ldg_if(sp, v & 1u), b = ldg_if(sp + 1, v & 2u);
      const float cc = ldg_if(sp + sh, v & 4u), dd = ldg_if(sp + sh + 1, v & 8u);
      float gw = gout;
      if (has_bl[d]) {
        const float top = fmaf(b, k[d].tx, a * k[d].ux), bot = fmaf(dd, k[d].tx, cc * k[d].ux);
        gbl[d] = fmaf(gout, fmaf(bot, k[d].ty, top * k[d].uy), gbl[d]);
        gw = gout * k[d].blend;
      }
      gix[d] = fmaf(gw, fmaf(k[d].ty, dd - cc, k[d].uy * (b - a)), gix[d]);
      giy[d] = fmaf(gw, fmaf(k[d].tx, dd - b, k[d].ux * (cc - a)), giy[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gix[d] += __shfl_xor_sync(0xffffffffu, gix[d], o);
      giy[d] += __shfl_xor_sync(0xffffffffu, giy[d], o);
      gbl[d] += __shfl_xor_sync(0xffffffffu, gbl[d], o);
    }
    if (lane == 0) bwdflow_store(P, Q, d, n, t, i, j, k[d], gix[d], giy[d], gbl[d]);
  }
}

// Slow-pixel bodies of the staged kernels: the taps were computed once per pixel into shared memory (`k`).
// Kernel 1: one (pixel, flattened channel) item.
template <int NDIRS>
__device__ __forceinline__ void fwd_slow_item(const Params& P, int n, int t, int i, int j, const Tap* k, int cf) {
  int g, c;
  chan_lookup(P, cf, g, c);
  const GroupP& R = P.grp[g];
  float r = 0.0f;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    const float* s = R.src[d] + n * R.src_sn[d] + t * R.src_st[d] + (long long)c * R.src_sc[d];
    float a = bilinear(s, k[d].y0 * R.src_sh[d] + k[d].x0, R.src_sh[d], k[d].valid, __fmul_rn(k[d].ux, k[d].uy),
                       __fmul_rn(k[d].tx, k[d].uy), __fmul_rn(k[d].ux, k[d].ty), __fmul_rn(k[d].tx, k[d].ty));
    if (P.dir[d].blend != nullptr) a = __fmul_rn(a, k[d].blend);
    r = (d == 0) ? a : __fadd_rn(r, a);
  }
  __stcs(R.out + n * R.out_sn + t * R.out_st + (long long)c * R.out_sc + (long long)i * R.out_sh + j, r);
}

// Kernel 2: one pixel per warp, lane l takes channels l, l+32, ...; fixed butterfly order (deterministic).
template <int NDIRS>
__device__ __forceinline__ void bwdflow_slow_warp(const Params& P, const GradP& Q, int n, int t, int i, int j, const Tap* k) {
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31;
  float gix[NDIRS], giy[NDIRS], gbl[NDIRS];
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    gix[d] = giy[d] = gbl[d] = 0.0f;
    has_bl[d] = P.dir[d].blend != nullptr;
  }
  int Ctot = 0;
  for (int g = 0; g < G.n_groups; ++g) Ctot += P.grp[g].C;
  for (int cf = lane; cf < Ctot; cf += 32) {
    int g, c;
    chan_lookup(P, cf, g, c);
    if (!Q.grad_out[g]) continue;
    const GroupP& R = P.grp[g];
    const float gout = __ldg(Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)c * Q.go_sc[g] +
                             (long long)i * Q.go_sh[g] + j);
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const int sh = R.src_sh[d];
      const float* sp = R.src[d] + n * R.src_sn[d] + t * R.src_st[d] + (long long)c * R.src_sc[d] + k[d].y0 * sh + k[d].x0;
      const unsigned v = k[d].valid;
      const float a = ldg_if(sp, v & 1u), b = ldg_if(sp + 1, v & 2u);
      const float cc = ldg_if(sp + sh, v & 4u), dd = ldg_if(sp + sh + 1, v & 8u);
      float gw = gout;
      if (has_bl[d]) {
        const float top = fmaf(b, k[d].tx, a * k[d].ux), bot = fmaf(dd, k[d].tx, cc * k[d].ux);
        gbl[d] = fmaf(gout, fmaf(bot, k[d].ty, top * k[d].uy), gbl[d]);
        gw = gout * k[d].blend;
      }
      gix[d] = fmaf(gw, fmaf(k[d].ty, dd - cc, k[d].uy * (b - a)), gix[d]);
      giy[d] = fmaf(gw, fmaf(k[d].tx, dd - b, k[d].ux * (cc - a)), giy[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gix[d] += __shfl_xor_sync(0xffffffffu, gix[d], o);
      giy[d] += __shfl_xor_sync(0xffffffffu, giy[d], o);
      gbl[d] += __shfl_xor_sync(0xffffffffu, gbl[d], o);
    }
    if (lane == 0) bwdflow_store(P, Q, d, n, t, i, j, k[d], gix[d], giy[d], gbl[d]);
  }
}

}  // namespace fwb

// fwb_blend.cuh — the mask blend that follows the warp in `refine` (utils/net_utils.py:141-143):
//     out[n,t,c] = input[n,t,c] * mask[n,t] + noise[n,c] * (1 - mask[n,t])
// torch runs it as four pointwise kernels per frame (mul, rsub, mul, add: each reads and writes a full [N,C,H,W]
// tensor) inside a Python loop over t, plus the `cat` of the 3-channel noise with 20 zero planes (`:134-136`).
// Here it is one pass: a thread owns 4 consecutive pixels (one float4 per plane) of one (n, t) and walks the
// channels, so the mask is read once and every input / output plane exactly once — pure HBM streaming
// (algorithmic bytes per pixel and frame: 4C in + 4 mask + 4C out, noise amortised over T).
// Channels c >= Cn have no noise plane (the reference concatenates zeros there): out = input * mask + 0 * (1 - mask),
// which is input * mask bit for bit (x + 0 == x, and the product 0 * (1 - m) is +0 for the finite masks of the path).
// Rounding follows torch's op sequence exactly (no FMA contraction): a = in*m; b = 1-m; c = nz*b; out = a+c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flowwarp_b200.h"

namespace fwb {

struct BlendP {
  int N, T, C, Cn, H, W;
  const float* in;
  long long in_sn, in_st, in_sc, in_sh;
  const float* m;
  long long m_sn, m_st, m_sh;
  const float* nz;
  long long nz_sn, nz_sc, nz_sh;
  float* out;
  long long out_sn, out_st, out_sc, out_sh;
  const float* go;
  long long go_sn, go_st, go_sc, go_sh;
  float* gi;
  long long gi_sn, gi_st, gi_sc, gi_sh;
  float* gm;
  long long gm_sn, gm_st, gm_sh;
  float* gn;
  long long gn_sn, gn_sc, gn_sh;
};

template <int V>
struct Vec;
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ static Vec ld(const float* p) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(p));
    return Vec{{a.x, a.y, a.z, a.w}};
  }
  __device__ __forceinline__ static Vec ld_keep(const float* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    return Vec{{a.x, a.y, a.z, a.w}};
  }
  __device__ __forceinline__ void st(float* p) const { __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3])); }
};
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ static Vec ld(const float* p) { return Vec{{__ldcs(p)}}; }
  __device__ __forceinline__ static Vec ld_keep(const float* p) { return Vec{{__ldg(p)}}; }
  __device__ __forceinline__ void st(float* p) const { __stcs(p, v[0]); }
};

// forward: grid (ceil(W/V/128), H, N*T), 128 threads
template <int V>
__global__ void __launch_bounds__(128) mask_blend_fwd_kernel(const __grid_constant__ BlendP B) {
  const int j = (blockIdx.x * 128 + threadIdx.x) * V, i = blockIdx.y;
  if (j >= B.W) return;
  const int n = blockIdx.z / B.T, t = blockIdx.z - n * B.T;
  const Vec<V> m = Vec<V>::ld(B.m + n * B.m_sn + t * B.m_st + i * B.m_sh + j);
  Vec<V> om;
#pragma unroll
  for (int k = 0; k < V; ++k) om.v[k] = __fsub_rn(1.0f, m.v[k]);
  const float* ip = B.in + n * B.in_sn + t * B.in_st + i * B.in_sh + j;
  float* op = B.out + n * B.out_sn + t * B.out_st + i * B.out_sh + j;
  const float* zp = B.nz ? B.nz + n * B.nz_sn + i * B.nz_sh + j : nullptr;
#pragma unroll 4
  for (int c = 0; c < B.C; ++c) {
    const Vec<V> x = Vec<V>::ld(ip + c * B.in_sc);
    Vec<V> r;
    if (c < B.Cn) {
      const Vec<V> z = Vec<V>::ld_keep(zp + c * B.nz_sc);  // re-read by the other T-1 frames: keep it cached
#pragma unroll
      for (int k = 0; k < V; ++k) r.v[k] = __fadd_rn(__fmul_rn(x.v[k], m.v[k]), __fmul_rn(z.v[k], om.v[k]));
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) r.v[k] = __fadd_rn(__fmul_rn(x.v[k], m.v[k]), __fmul_rn(0.0f, om.v[k]));
    }
    r.st(op + c * B.out_sc);
  }
}

// backward, kernel A: grid (ceil(W/V/128), H, N*T), 128 threads; a thread owns its pixels of one (n, t) for every c:
//   grad_input[n,t,c] = g * mask[n,t]
//   grad_mask[n,t]    = sum_c g * (input[n,t,c] - noise[n,c])      (register accumulator, fixed order: deterministic)
//   grad_noise[n,c]   = g * (1 - mask[n,0])                        only when T == 1 (no sum over frames needed)
template <int V>
__global__ void __launch_bounds__(128, 8) mask_blend_bwd_kernel(const __grid_constant__ BlendP B) {
  const int j = (blockIdx.x * 128 + threadIdx.x) * V, i = blockIdx.y;
  if (j >= B.W) return;
  const int n = blockIdx.z / B.T, t = blockIdx.z - n * B.T;
  const Vec<V> m = Vec<V>::ld_keep(B.m + n * B.m_sn + t * B.m_st + i * B.m_sh + j);
  Vec<V> gm;
#pragma unroll
  for (int k = 0; k < V; ++k) gm.v[k] = 0.f;
  const bool want_gm = B.gm != nullptr, want_gi = B.gi != nullptr, want_gn = B.gn != nullptr && B.T == 1;
  const float* gp = B.go + n * B.go_sn + t * B.go_st + i * B.go_sh + j;
  const float* ip = B.in + n * B.in_sn + t * B.in_st + i * B.in_sh + j;
  float* gip = want_gi ? B.gi + n * B.gi_sn + t * B.gi_st + i * B.gi_sh + j : nullptr;
#pragma unroll 2
  for (int c = 0; c < B.C; ++c) {
    const Vec<V> g = Vec<V>::ld(gp + c * B.go_sc);
    if (want_gi) {
      Vec<V> r;
#pragma unroll
      for (int k = 0; k < V; ++k) r.v[k] = __fmul_rn(g.v[k], m.v[k]);
      r.st(gip + c * B.gi_sc);
    }
    if (want_gm) {
      const Vec<V> x = Vec<V>::ld(ip + c * B.in_sc);
      if (c < B.Cn) {
        const Vec<V> z = Vec<V>::ld_keep(B.nz + n * B.nz_sn + c * B.nz_sc + i * B.nz_sh + j);
#pragma unroll
        for (int k = 0; k < V; ++k) gm.v[k] = fmaf(g.v[k], x.v[k] - z.v[k], gm.v[k]);
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) gm.v[k] = fmaf(g.v[k], x.v[k], gm.v[k]);
      }
    }
    if (want_gn && c < B.Cn) {
      Vec<V> r;
#pragma unroll
      for (int k = 0; k < V; ++k) r.v[k] = g.v[k] * (1.0f - m.v[k]);
      r.st(B.gn + n * B.gn_sn + c * B.gn_sc + i * B.gn_sh + j);
    }
  }
  if (want_gm) gm.st(B.gm + n * B.gm_sn + t * B.gm_st + i * B.gm_sh + j);
}

// backward, kernel B (T > 1 only): grad_noise[n,c] = sum_t g[n,t,c] * (1 - mask[n,t]) for the Cn noise channels;
// grid (ceil(W/V/128), H, N*Cn).  Re-reads Cn of the C grad_out planes (3 of 23 on the path) and the masks.
template <int V>
__global__ void __launch_bounds__(128, 8) mask_blend_gnoise_kernel(const __grid_constant__ BlendP B) {
  const int j = (blockIdx.x * 128 + threadIdx.x) * V, i = blockIdx.y;
  if (j >= B.W) return;
  const int n = blockIdx.z / B.Cn, c = blockIdx.z - n * B.Cn;
  Vec<V> acc;
#pragma unroll
  for (int k = 0; k < V; ++k) acc.v[k] = 0.f;
  for (int t = 0; t < B.T; ++t) {
    const Vec<V> g = Vec<V>::ld_keep(B.go + n * B.go_sn + t * B.go_st + c * B.go_sc + i * B.go_sh + j);
    const Vec<V> m = Vec<V>::ld_keep(B.m + n * B.m_sn + t * B.m_st + i * B.m_sh + j);
#pragma unroll
    for (int k = 0; k < V; ++k) acc.v[k] = fmaf(g.v[k], 1.0f - m.v[k], acc.v[k]);
  }
  acc.st(B.gn + n * B.gn_sn + c * B.gn_sc + i * B.gn_sh + j);
}

static inline bool blend_vec_ok(const fwb_blend* b, bool bwd) {
  if (b->W & 3) return false;
  auto al = [](const void* p, std::initializer_list<int64_t> st) {
    if ((uintptr_t)p & 15u) return false;
    for (int64_t s : st)
      if (s & 3) return false;
    return true;
  };
  if (!al(b->input, {b->in_sn, b->in_st, b->in_sc, b->in_sh}) || !al(b->mask, {b->m_sn, b->m_st, b->m_sh})) return false;
  if (b->noise && !al(b->noise, {b->nz_sn, b->nz_sc, b->nz_sh})) return false;
  if (!bwd) return al(b->out, {b->out_sn, b->out_st, b->out_sc, b->out_sh});
  if (!al(b->grad_out, {b->go_sn, b->go_st, b->go_sc, b->go_sh})) return false;
  if (b->grad_input && !al(b->grad_input, {b->gi_sn, b->gi_st, b->gi_sc, b->gi_sh})) return false;
  if (b->grad_mask && !al(b->grad_mask, {b->gm_sn, b->gm_st, b->gm_sh})) return false;
  if (b->grad_noise && !al(b->grad_noise, {b->gn_sn, b->gn_sc, b->gn_sh})) return false;
  return true;
}

static inline int blend_validate(const fwb_blend* b, bool bwd) {
  if (!b) return FWB_E_NULL;
  if (b->N < 0 || b->T < 1 || b->C < 1 || b->H < 1 || b->W < 1 || b->Cn < 0 || b->Cn > b->C) return FWB_E_SHAPE;
  if (b->H > 65535 || (long long)b->N * b->T > 65535 || (long long)b->N * b->Cn > 65535) return FWB_E_RANGE;
  if (!b->input || !b->mask || (b->Cn > 0 && !b->noise)) return FWB_E_NULL;
  if (!bwd && !b->out) return FWB_E_NULL;
  if (bwd && !b->grad_out) return FWB_E_NULL;
  const void* ps[] = {b->input, b->mask, b->noise, b->out, b->grad_out, b->grad_input, b->grad_mask, b->grad_noise};
  for (const void* p : ps)
    if ((uintptr_t)p & 3u) return FWB_E_ALIGN;
  return 0;
}

static inline void blend_params(const fwb_blend* b, BlendP& B) {
  B.N = b->N, B.T = b->T, B.C = b->C, B.Cn = b->noise ? b->Cn : 0, B.H = b->H, B.W = b->W;
  B.in = b->input, B.in_sn = b->in_sn, B.in_st = b->in_st, B.in_sc = b->in_sc, B.in_sh = b->in_sh;
  B.m = b->mask, B.m_sn = b->m_sn, B.m_st = b->m_st, B.m_sh = b->m_sh;
  B.nz = b->noise, B.nz_sn = b->nz_sn, B.nz_sc = b->nz_sc, B.nz_sh = b->nz_sh;
  B.out = b->out, B.out_sn = b->out_sn, B.out_st = b->out_st, B.out_sc = b->out_sc, B.out_sh = b->out_sh;
  B.go = b->grad_out, B.go_sn = b->go_sn, B.go_st = b->go_st, B.go_sc = b->go_sc, B.go_sh = b->go_sh;
  B.gi = b->grad_input, B.gi_sn = b->gi_sn, B.gi_st = b->gi_st, B.gi_sc = b->gi_sc, B.gi_sh = b->gi_sh;
  B.gm = b->grad_mask, B.gm_sn = b->gm_sn, B.gm_st = b->gm_st, B.gm_sh = b->gm_sh;
  B.gn = b->grad_noise, B.gn_sn = b->gn_sn, B.gn_sc = b->gn_sc, B.gn_sh = b->gn_sh;
}

}  // namespace fwb

// fwb_tex.cuh — the 4-tap gather through the TEXTURE units.
//
// tex2Dgather (SASS TLD4) returns the 2x2 quad (x0..x0+1, y0..y0+1) of one fp32 plane in ONE instruction, the exact stored
// values (no filtering), so a (pixel, direction, channel) costs one texture instruction instead of four shared-memory or
// global loads, needs no staging of the source footprint and leaves the LSU / shared-memory pipe to the stores (forward) or
// to the scatter (backward).  Measured on the config-2 flows (tools/mb_tex.cu): 0.195 ms for the whole forward against 0.263
// for 4 x LDG and 0.26 for the shared-memory tile kernel; the rate does not depend on the warp's patch shape (TEX-pipe bound).
//
// A source tensor [N, T|1, C, H, W] that is dense above its rows (row pitch a multiple of 32 bytes, planes back to back) is
// described as pitch-linear 2-D textures of W x (nb * Tn * C * H) texels, one object per block of nb clips (a pitch-linear
// texture holds at most 65000 rows); plane (n, t, c) starts at row ((n % nb) * Tn + t) * C * H + c * H.  Taps outside the
// image would read the neighbouring plane (or the clamped edge): they are masked to zero with the validity bits, which is
// also what zeros padding needs.  Texture objects are created by the host side once per (pointer, shape) and cached.
#pragma once
#include "fwb_coords.cuh"
#include "fwb_generic.cuh"

namespace fwb {

constexpr int TX_MAXBLK = 8;  // texture objects (blocks of clips) per source tensor

struct TexSrc {
  unsigned long long tex[TX_MAXBLK];
  int nb;      // clips per texture object
  int rows_n;  // texture rows per clip   (Tn * C * H)
  int rows_t;  // texture rows per frame  (C * H), 0 when one frame serves all T
};
struct TexP {
  TexSrc s[FWB_MAX_GROUPS][2];
};

// the quad of taps of one plane: (x0,y0) (x0+1,y0) (x0,y0+1) (x0+1,y0+1); fx1 = x0 + 1, fy1 = first row of the plane + y0 + 1
// (texel centres sit at +0.5: the footprint of coordinate x0 + 1 is exactly the texels x0 and x0 + 1)
__device__ __forceinline__ void tex_quad(unsigned long long tx, float fx1, float fy1, unsigned v, float& a, float& b, float& c, float& d) {
  const float4 q = tex2Dgather<float4>((cudaTextureObject_t)tx, fx1, fy1, 0);
  a = (v & 1u) ? q.w : 0.0f;
  b = (v & 2u) ? q.z : 0.0f;
  c = (v & 4u) ? q.x : 0.0f;
  d = (v & 8u) ? q.y : 0.0f;
}

// Optional side job of the forward: zero-fill of the grad_src planes the fused backward will accumulate into (see
// fwb_warp_blend_forward_zero).  Every CTA clears the block of its own output coordinates in every plane.
struct ZeroP {
  float* gs[FWB_MAX_GROUPS][2];
  long long sn[FWB_MAX_GROUPS][2], st[FWB_MAX_GROUPS][2];
  int sc[FWB_MAX_GROUPS][2];
  int sh[2];  // row stride per direction (all groups agree, host-checked)
  int on;
};
__device__ __forceinline__ float* zero_plane(const ZeroP& Z, const Geo& G, int g, int d, int n, int t, int c) {
  float* p = Z.gs[g][d];
  if (!p || (t != 0 && Z.st[g][d] == 0)) return nullptr;  // a source shared by all T frames is cleared by the t == 0 tiles
  return p + n * Z.sn[g][d] + t * Z.st[g][d] + (long long)c * Z.sc[g][d];
}

constexpr int TXF_THREADS = 256;
#ifndef TXF_ZERO_FIRST
#define TXF_ZERO_FIRST 1  // zero-fill side job before (1) or after (0) the channel loop
#endif
constexpr int TXF_CB = 4;  // channels per batch of texture fetches
constexpr int TXF_TW = 32, TXF_TH = 8;  // CTA tile: 4 x 2 warp patches of 8 x 4 pixels

// ---------------------------------------------------------------------------------------------
// Kernel 1 on the texture path: one thread per output pixel, all directions, all channel groups.
// Arithmetic identical to fwd_generic_pixel (nw, ne, sw, se; * blend; sum of the directions).
// ---------------------------------------------------------------------------------------------
template <int NDIRS>
__global__ void __launch_bounds__(TXF_THREADS) fwd_tex_kernel(const __grid_constant__ Params P, const __grid_constant__ TexP X,
                                                              const __grid_constant__ ZeroP Z, int Ctot) {
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * TXF_TW + (warp & 3) * 8 + (lane & 7);
  const int i = blockIdx.y * TXF_TH + (warp >> 2) * 4 + (lane >> 3);
  int n, t;
  if (G.T == 1) {
    n = blockIdx.z;
    t = 0;
  } else {
    n = blockIdx.z / G.T;
    t = blockIdx.z - n * G.T;
  }
  const bool in = j < G.W && i < G.H;
#if TXF_ZERO_FIRST
  if (Z.on) {  // 64 float4 per plane and tile: a quarter of the CTA per plane, the planes of a (group, direction) round robin
    const int f4 = threadIdx.x & 63, zi = blockIdx.y * TXF_TH + (f4 >> 3), zj = blockIdx.x * TXF_TW + (f4 & 7) * 4;
    if (zi < G.H && zj < G.W) {
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int g = 0; g < G.n_groups; ++g) {
        const int C = P.grp[g].C;
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) {
          float* zp = Z.gs[g][d];
          if (!zp || (t != 0 && Z.st[g][d] == 0)) continue;  // a source shared by all T frames is cleared by the t == 0 tiles
          zp += n * Z.sn[g][d] + t * Z.st[g][d] + (long long)zi * Z.sh[d] + zj;
          const long long sc = Z.sc[g][d];
          for (int c = threadIdx.x >> 6; c < C; c += TXF_THREADS / 64) *reinterpret_cast<float4*>(zp + c * sc) = z4;
        }
      }
    }
  }
#endif
  float w[NDIRS][4], bl[NDIRS], fx1[NDIRS], fy1[NDIRS];
  unsigned v[NDIRS];
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    compute_tap(G, P.dir[d], n, t, min(i, G.H - 1), min(j, G.W - 1), k);  // ragged tiles: clamped address, store masked
    w[d][0] = __fmul_rn(k.ux, k.uy);
    w[d][1] = __fmul_rn(k.tx, k.uy);
    w[d][2] = __fmul_rn(k.ux, k.ty);
    w[d][3] = __fmul_rn(k.tx, k.ty);
    v[d] = k.valid;
    bl[d] = k.blend;
    has_bl[d] = P.dir[d].blend != nullptr;
    // a pixel without any valid tap fetches texel (0, 0) of the slab (any address will do, the values are masked)
    fx1[d] = k.valid ? (float)(k.x0 + 1) : 0.0f;
    fy1[d] = k.valid ? (float)(k.y0 + 1) : 0.0f;
  }
  for (int g = 0; g < G.n_groups; ++g) {
    const GroupP& R = P.grp[g];
    unsigned long long tx[NDIRS];
    float row[NDIRS];
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const TexSrc& S = X.s[g][d];
      const int blk = n / S.nb;
      tx[d] = S.tex[blk];
      row[d] = (float)((n - blk * S.nb) * S.rows_n + t * S.rows_t) + fy1[d];
    }
    float* out = R.out + n * R.out_sn + t * R.out_st + (long long)i * R.out_sh + j;
    const float fH = (float)G.H;
    // channels in batches of TXF_CB: all texture fetches of a batch are issued before the first result is used
    for (int c0 = 0; c0 < R.C; c0 += TXF_CB) {
      float4 q[TXF_CB][NDIRS];
#pragma unroll
      for (int u = 0; u < TXF_CB; ++u)
#pragma unroll
        for (int d = 0; d < NDIRS; ++d)  // rows past the last channel stay inside the slab or clamp: fetched, never used
          q[u][d] = tex2Dgather<float4>((cudaTextureObject_t)tx[d], fx1[d], row[d] + (float)u * fH, 0);
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) row[d] += (float)TXF_CB * fH;  // integers below 2^24: exact
#pragma unroll
      for (int u = 0; u < TXF_CB; ++u) {
        float r = 0.0f;
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) {
          const float a = (v[d] & 1u) ? q[u][d].w : 0.0f, b = (v[d] & 2u) ? q[u][d].z : 0.0f;
          const float cc = (v[d] & 4u) ? q[u][d].x : 0.0f, dd = (v[d] & 8u) ? q[u][d].y : 0.0f;
          float s = __fmul_rn(a, w[d][0]);
          s = __fmaf_rn(b, w[d][1], s);
          s = __fmaf_rn(cc, w[d][2], s);
          s = __fmaf_rn(dd, w[d][3], s);
          if (has_bl[d]) s = __fmul_rn(s, bl[d]);
          r = (d == 0) ? s : __fadd_rn(r, s);
        }
        if (in && c0 + u < R.C) __stcs(out + (long long)(c0 + u) * R.out_sc, r);
      }
    }
  }
#if !TXF_ZERO_FIRST
  if (Z.on) {  // 64 float4 per plane and tile: a quarter of the CTA per plane, planes round robin
    const int f4 = threadIdx.x & 63, zi = blockIdx.y * TXF_TH + (f4 >> 3), zj = blockIdx.x * TXF_TW + (f4 & 7) * 4;
    if (zi < G.H && zj < G.W) {
      for (int pl = threadIdx.x >> 6; pl < Ctot * NDIRS; pl += TXF_THREADS / 64) {
        const int cf = pl / NDIRS, d = pl - cf * NDIRS;
        int g, c;
        chan_lookup(P, cf, g, c);
        float* zp = zero_plane(Z, G, g, d, n, t, c);
        if (zp) *reinterpret_cast<float4*>(zp + (long long)zi * Z.sh[d] + zj) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
#endif
}

// ---------------------------------------------------------------------------------------------
// Kernel 2 on the texture path: coordinate gradient -> grad_flow / grad_gate / grad_blend, one thread per output pixel.
//   gix = sum_c gw_c * [ uy*(v_ne - v_nw) + ty*(v_se - v_sw) ],  giy = sum_c gw_c * [ ux*(v_sw - v_nw) + tx*(v_se - v_ne) ]
// (ATen grid_sampler_2d_backward with the common factors pulled out; taps outside the image count as zero), the same
// arithmetic as bwdflow_generic_pixel.  Channels in batches: every texture fetch and grad_out load of a batch is issued
// before the first result is used.
// ---------------------------------------------------------------------------------------------
template <int NDIRS>
__global__ void __launch_bounds__(TXF_THREADS, 3) bwd_flow_tex_kernel(const __grid_constant__ Params P, const __grid_constant__ GradP Q,
                                                                   const __grid_constant__ TexP X) {
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * TXF_TW + (warp & 3) * 8 + (lane & 7);
  const int i = blockIdx.y * TXF_TH + (warp >> 2) * 4 + (lane >> 3);
  int n, t;
  if (G.T == 1) {
    n = blockIdx.z;
    t = 0;
  } else {
    n = blockIdx.z / G.T;
    t = blockIdx.z - n * G.T;
  }
  const bool in = j < G.W && i < G.H;
  const int ic = min(i, G.H - 1), jc = min(j, G.W - 1);  // ragged tiles: clamped address, stores masked
  // only what the channel loop needs stays in registers; the gradient multipliers, the raw flow and the gate are recomputed for
  // the final store (once per pixel)
  float tx[NDIRS], ty[NDIRS], ux[NDIRS], uy[NDIRS], bl[NDIRS];
  unsigned vld[NDIRS];
  float gix[NDIRS], giy[NDIRS], gbl[NDIRS], fx1[NDIRS], fy1[NDIRS];
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    compute_tap(G, P.dir[d], n, t, ic, jc, k);
    tx[d] = k.tx, ty[d] = k.ty, ux[d] = k.ux, uy[d] = k.uy, bl[d] = k.blend, vld[d] = k.valid;
    gix[d] = giy[d] = gbl[d] = 0.0f;
    has_bl[d] = P.dir[d].blend != nullptr;
    fx1[d] = k.valid ? (float)(k.x0 + 1) : 0.0f;
    fy1[d] = k.valid ? (float)(k.y0 + 1) : 0.0f;
  }
  const float fH = (float)G.H;
  for (int g = 0; g < G.n_groups; ++g) {
    const GroupP& R = P.grp[g];
    if (!Q.grad_out[g]) continue;
    const float* go = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)ic * Q.go_sh[g] + jc;
    unsigned long long th[NDIRS];
    float row[NDIRS];
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const TexSrc& S = X.s[g][d];
      const int blk = n / S.nb;
      th[d] = S.tex[blk];
      row[d] = (float)((n - blk * S.nb) * S.rows_n + t * S.rows_t) + fy1[d];
    }
    for (int c0 = 0; c0 < R.C; c0 += TXF_CB) {
      float4 q[TXF_CB][NDIRS];
      float gv[TXF_CB];
#pragma unroll
      for (int u = 0; u < TXF_CB; ++u) {
        gv[u] = (c0 + u < R.C) ? __ldcs(go + (long long)(c0 + u) * Q.go_sc[g]) : 0.0f;  // a channel past the end contributes 0
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) q[u][d] = tex2Dgather<float4>((cudaTextureObject_t)th[d], fx1[d], row[d] + (float)u * fH, 0);
      }
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) row[d] += (float)TXF_CB * fH;
#pragma unroll
      for (int u = 0; u < TXF_CB; ++u) {
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) {
          const unsigned v = (c0 + u < R.C) ? vld[d] : 0u;  // (the fetched texels of a channel past the end are not finite-safe)
          const float a = (v & 1u) ? q[u][d].w : 0.0f, b = (v & 2u) ? q[u][d].z : 0.0f;
          const float cc = (v & 4u) ? q[u][d].x : 0.0f, dd = (v & 8u) ? q[u][d].y : 0.0f;
          float gw = gv[u];
          if (has_bl[d]) {
            const float top = fmaf(b, tx[d], a * ux[d]), bot = fmaf(dd, tx[d], cc * ux[d]);
            gbl[d] = fmaf(gv[u], fmaf(bot, ty[d], top * uy[d]), gbl[d]);
            gw = gv[u] * bl[d];
          }
          gix[d] = fmaf(gw, fmaf(ty[d], dd - cc, uy[d] * (b - a)), gix[d]);
          giy[d] = fmaf(gw, fmaf(tx[d], dd - b, ux[d] * (cc - a)), giy[d]);
        }
      }
    }
  }
  if (in) {
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      Tap k;
      compute_tap(G, P.dir[d], n, t, i, j, k);
      bwdflow_store(P, Q, d, n, t, i, j, k, gix[d], giy[d], gbl[d]);
    }
  }
}

}  // namespace fwb

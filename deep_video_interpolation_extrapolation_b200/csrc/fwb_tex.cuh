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
#ifndef TXF_ZSTREAM
#define TXF_ZSTREAM 1  // zero-fill stores with the evict-first hint (st.global.cs): the zeros are not read again before the backward
#endif
#if TXF_ZSTREAM
#define TXF_ZST(p, v) __stcs(p, v)
#else
#define TXF_ZST(p, v) (*(p) = (v))
#endif
constexpr int TXF_CB = 4;  // channels per batch of texture fetches
constexpr int TXF_TW = 32, TXF_TH = 8;  // CTA tile: 4 x 2 warp patches of 8 x 4 pixels

// One batch of NB channels of the forward: every texture fetch of the batch is issued before the first result is used; the
// output pointer and the texture rows advance by one plane per channel (no per-channel index arithmetic).
template <int NDIRS>
struct TxfLoop {
  unsigned long long tx[NDIRS];
  float row[NDIRS];  // texture row of the quad's lower row in the current channel (integers below 2^24: exact)
  float* out;
  int out_sc;
};
template <int NDIRS, int NB, bool MASKED>
__device__ __forceinline__ void txf_batch(TxfLoop<NDIRS>& L, const float (&fx1)[NDIRS], float fH, const float (&w)[NDIRS][4],
                                          const float (&bl)[NDIRS], const unsigned (&m)[NDIRS][4], bool in) {
  float4 q[NB][NDIRS];
#pragma unroll
  for (int u = 0; u < NB; ++u)
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) q[u][d] = tex2Dgather<float4>((cudaTextureObject_t)L.tx[d], fx1[d], __fmaf_rn((float)u, fH, L.row[d]), 0);
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) L.row[d] = __fmaf_rn((float)NB, fH, L.row[d]);
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    float r = 0.0f;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      // TLD4 component order: w = (x0,y0) z = (x0+1,y0) x = (x0,y0+1) y = (x0+1,y0+1)
      float a = q[u][d].w, b = q[u][d].z, cc = q[u][d].x, dd = q[u][d].y;
      if (MASKED) {
        a = __uint_as_float(__float_as_uint(a) & m[d][0]);
        b = __uint_as_float(__float_as_uint(b) & m[d][1]);
        cc = __uint_as_float(__float_as_uint(cc) & m[d][2]);
        dd = __uint_as_float(__float_as_uint(dd) & m[d][3]);
      }
      float s = __fmul_rn(a, w[d][0]);
      s = __fmaf_rn(b, w[d][1], s);
      s = __fmaf_rn(cc, w[d][2], s);
      s = __fmaf_rn(dd, w[d][3], s);
      s = __fmul_rn(s, bl[d]);
      r = (d == 0) ? s : __fadd_rn(r, s);
    }
    if (in) __stcs(L.out, r);
    L.out += L.out_sc;
  }
}

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
          for (int c = threadIdx.x >> 6; c < C; c += TXF_THREADS / 64) TXF_ZST(reinterpret_cast<float4*>(zp + c * sc), z4);
        }
      }
    }
  }
#endif
  float w[NDIRS][4], bl[NDIRS], fx1[NDIRS], fy1[NDIRS];
  unsigned v[NDIRS];
  bool part = false;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    compute_tap(G, P.dir[d], n, t, min(i, G.H - 1), min(j, G.W - 1), k);  // ragged tiles: clamped address, store masked
    w[d][0] = __fmul_rn(k.ux, k.uy);
    w[d][1] = __fmul_rn(k.tx, k.uy);
    w[d][2] = __fmul_rn(k.ux, k.ty);
    w[d][3] = __fmul_rn(k.tx, k.ty);
    v[d] = k.valid;
    bl[d] = k.blend;  // 1.0 without a blend weight: s * 1.0 == s bit for bit
    part |= k.valid != 15u;
    // a pixel without any valid tap fetches texel (0, 0) of the slab (any address will do, the values are masked)
    fx1[d] = k.valid ? (float)(k.x0 + 1) : 0.0f;
    fy1[d] = k.valid ? (float)(k.y0 + 1) : 0.0f;
  }
  // a warp whose 32 x NDIRS quads lie inside the image (the common case) runs the channel loop without the validity selects
  const bool masked = __any_sync(0xffffffffu, part);
  const float fH = (float)G.H;
  for (int g = 0; g < G.n_groups; ++g) {
    const GroupP& R = P.grp[g];
    TxfLoop<NDIRS> L;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const TexSrc& S = X.s[g][d];
      const int blk = n / S.nb;
      L.tx[d] = S.tex[blk];
      L.row[d] = (float)((n - blk * S.nb) * S.rows_n + t * S.rows_t) + fy1[d];
    }
    L.out = R.out + n * R.out_sn + t * R.out_st + (long long)min(i, G.H - 1) * R.out_sh + min(j, G.W - 1);
    L.out_sc = R.out_sc;
    const int C = R.C;
    if (!masked) {
      unsigned m[NDIRS][4];
      int c = 0;
#pragma unroll 1
      for (; c + TXF_CB <= C; c += TXF_CB) txf_batch<NDIRS, TXF_CB, false>(L, fx1, fH, w, bl, m, in);
      switch (C - c) {
        case 1: txf_batch<NDIRS, 1, false>(L, fx1, fH, w, bl, m, in); break;
        case 2: txf_batch<NDIRS, 2, false>(L, fx1, fH, w, bl, m, in); break;
        case 3: txf_batch<NDIRS, 3, false>(L, fx1, fH, w, bl, m, in); break;
        default: break;
      }
    } else {
      unsigned m[NDIRS][4];  // all-ones / zero per tap: a tap outside the image counts as the value 0 (AND on the fetched bits)
#pragma unroll
      for (int d = 0; d < NDIRS; ++d)
#pragma unroll
        for (int q = 0; q < 4; ++q) m[d][q] = (v[d] >> q) & 1u ? 0xffffffffu : 0u;
      int c = 0;
#pragma unroll 1
      for (; c + TXF_CB <= C; c += TXF_CB) txf_batch<NDIRS, TXF_CB, true>(L, fx1, fH, w, bl, m, in);
      switch (C - c) {
        case 1: txf_batch<NDIRS, 1, true>(L, fx1, fH, w, bl, m, in); break;
        case 2: txf_batch<NDIRS, 2, true>(L, fx1, fH, w, bl, m, in); break;
        case 3: txf_batch<NDIRS, 3, true>(L, fx1, fH, w, bl, m, in); break;
        default: break;
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
        if (zp) TXF_ZST(reinterpret_cast<float4*>(zp + (long long)zi * Z.sh[d] + zj), make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
  }
#endif
}

// One batch of NB channels of kernel 2: all texture fetches and grad_out loads of the batch are issued before the first use.
template <int NDIRS>
struct TxbLoop {
  unsigned long long th[NDIRS];
  float row[NDIRS];
  const float* go;  // grad_out of this pixel in the current channel
  int go_sc;
};
template <int NDIRS, int NB, bool MASKED>
__device__ __forceinline__ void txb_batch(TxbLoop<NDIRS>& L, const float (&fx1)[NDIRS], float fH, float (&A)[NDIRS][4],
                                          const unsigned (&m)[NDIRS][4]) {
  float4 q[NB][NDIRS];
  float gv[NB];
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    gv[u] = __ldcs(L.go);
    L.go += L.go_sc;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) q[u][d] = tex2Dgather<float4>((cudaTextureObject_t)L.th[d], fx1[d], __fmaf_rn((float)u, fH, L.row[d]), 0);
  }
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) L.row[d] = __fmaf_rn((float)NB, fH, L.row[d]);
#pragma unroll
  for (int u = 0; u < NB; ++u)
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      // TLD4 component order: w = (x0,y0) z = (x0+1,y0) x = (x0,y0+1) y = (x0+1,y0+1)
      float a = q[u][d].w, b = q[u][d].z, cc = q[u][d].x, dd = q[u][d].y;
      if (MASKED) {
        a = __uint_as_float(__float_as_uint(a) & m[d][0]);
        b = __uint_as_float(__float_as_uint(b) & m[d][1]);
        cc = __uint_as_float(__float_as_uint(cc) & m[d][2]);
        dd = __uint_as_float(__float_as_uint(dd) & m[d][3]);
      }
      A[d][0] = fmaf(gv[u], a, A[d][0]);
      A[d][1] = fmaf(gv[u], b, A[d][1]);
      A[d][2] = fmaf(gv[u], cc, A[d][2]);
      A[d][3] = fmaf(gv[u], dd, A[d][3]);
    }
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
  // Only the four sums  A_k = sum_c grad_out_c * tap_k,c  per direction are accumulated in the channel loop (one TLD4 and four
  // FFMA per channel and direction); the bilinear weights do not depend on the channel and are applied once at the end:
  //   gix = bl [ty (D - C) + uy (B - A)],  giy = bl [tx (D - B) + ux (C - A)],  gblend = uy (ux A + tx B) + ty (ux C + tx D)
  // (ATen grid_sampler_2d_backward with the common factors pulled out; taps outside the image count as zero).  The taps are
  // recomputed for the final store, so nothing but the fetch coordinates and the sums lives across the loop.
  float fx1[NDIRS], fy1[NDIRS], A[NDIRS][4];
  unsigned m[NDIRS][4];
  bool part = false;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    compute_tap(G, P.dir[d], n, t, ic, jc, k);
    fx1[d] = k.valid ? (float)(k.x0 + 1) : 0.0f;
    fy1[d] = k.valid ? (float)(k.y0 + 1) : 0.0f;
    part |= k.valid != 15u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      m[d][q] = (k.valid >> q) & 1u ? 0xffffffffu : 0u;
      A[d][q] = 0.0f;
    }
  }
  const bool masked = __any_sync(0xffffffffu, part);
  const float fH = (float)G.H;
  for (int g = 0; g < G.n_groups; ++g) {
    const GroupP& R = P.grp[g];
    if (!Q.grad_out[g]) continue;
    TxbLoop<NDIRS> L;
    L.go = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)ic * Q.go_sh[g] + jc;
    L.go_sc = Q.go_sc[g];
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const TexSrc& S = X.s[g][d];
      const int blk = n / S.nb;
      L.th[d] = S.tex[blk];
      L.row[d] = (float)((n - blk * S.nb) * S.rows_n + t * S.rows_t) + fy1[d];
    }
    const int C = R.C;
    int c = 0;
    if (!masked) {
#pragma unroll 1
      for (; c + TXF_CB <= C; c += TXF_CB) txb_batch<NDIRS, TXF_CB, false>(L, fx1, fH, A, m);
      switch (C - c) {
        case 1: txb_batch<NDIRS, 1, false>(L, fx1, fH, A, m); break;
        case 2: txb_batch<NDIRS, 2, false>(L, fx1, fH, A, m); break;
        case 3: txb_batch<NDIRS, 3, false>(L, fx1, fH, A, m); break;
        default: break;
      }
    } else {
#pragma unroll 1
      for (; c + TXF_CB <= C; c += TXF_CB) txb_batch<NDIRS, TXF_CB, true>(L, fx1, fH, A, m);
      switch (C - c) {
        case 1: txb_batch<NDIRS, 1, true>(L, fx1, fH, A, m); break;
        case 2: txb_batch<NDIRS, 2, true>(L, fx1, fH, A, m); break;
        case 3: txb_batch<NDIRS, 3, true>(L, fx1, fH, A, m); break;
        default: break;
      }
    }
  }
  if (in) {
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      Tap k;
      compute_tap(G, P.dir[d], n, t, i, j, k);
      const float a = A[d][0], b = A[d][1], cc = A[d][2], dd = A[d][3];
      float gbl = 0.0f, sc = 1.0f;
      if (P.dir[d].blend != nullptr) {
        const float top = fmaf(b, k.tx, a * k.ux), bot = fmaf(dd, k.tx, cc * k.ux);
        gbl = fmaf(bot, k.ty, top * k.uy);
        sc = k.blend;
      }
      const float gix = sc * fmaf(k.ty, dd - cc, k.uy * (b - a));
      const float giy = sc * fmaf(k.tx, dd - b, k.ux * (cc - a));
      bwdflow_store(P, Q, d, n, t, i, j, k, gix, giy, gbl);
    }
  }
}

}  // namespace fwb

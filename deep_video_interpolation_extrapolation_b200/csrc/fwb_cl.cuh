// fwb_cl.cuh — fused backward (kernels 2+3) with a CHANNEL-PER-LANE scatter.
//
// The scatter of grad_out * weight into grad_src (ATen grid_sampler_2d_backward's atomicAdd, reached from
// utils/net_utils.py:113) was bound by same-bank serialisation when a warp instruction carried 32 PIXELS of one channel:
// neighbouring pixels hit the same / neighbouring source cells (3.8 wavefronts per ATOMS).  Here a
// warp instruction carries the <= 32 CHANNELS of ONE (pixel, direction): lane c adds into plane c of a shared-memory copy of
// the tile's source footprint, the planes are an odd number of words apart, so the 23 lanes of the headline op always hit
// 23 different banks: one wavefront per ATOMS, and one descriptor (4 weights + 2 offsets, broadcast loads) serves all
// channels of the item.  The coordinate gradient (kernel 2) needs the opposite mapping (a thread sums over the channels of
// its pixel) and runs on the TEXTURE units (fwb_tex.cuh), so the CTA is warp-specialised:
//
//     warps 0-7   (thread = pixel)     taps, footprint tables, item descriptors | kernel 2: TLD4 quads x grad_out -> grad_flow ...
//     warps 8-15  (lane = channel)     grad_out tile -> shared memory, scale vote | scatter (ATOMS.ADD.S32) | flush (RED.F32)
//
// so that texture fetches and shared-memory traffic are in flight at the same time.  (Both pass through the one L1TEX data stage
// of the SM, which is what bounds the kernel: 62.8 % LSU + 18.8 % TEX wavefronts, profiles/r2_ncu_full_summary.txt; making one
// role faster makes the other slower, profiles/r2_cl_timeline.txt.)  A CTA owns a 32x8 tile of output pixels
// of one (n, t); 2 CTAs per SM.  The two directions are scattered one after the other into the same accumulator planes.
// Fixed point as in fwb_tile.cuh: scale per channel and tile 2^(20 - exponent(max|grad_out| * max|blend|)),
// float -> int by fma(x, y, 1.5 * 2^23); non-finite grad_out channels and items whose taps are far from the rest of the tile
// (border-clamped outliers, or a direction whose footprint does not fit) go to global memory with float atomics.
#pragma once
#include "fwb_coords.cuh"
#include "fwb_generic.cuh"
#include "fwb_tex.cuh"
#include "fwb_util.cuh"

namespace fwb {

constexpr int CL_TW = 32, CL_TH = 8, CL_NPIX = CL_TW * CL_TH;
constexpr int CL_THREADS = 2 * CL_NPIX;  // 8 pixel warps + 8 channel warps
constexpr int CL_R = 24;                 // an item displaced farther than this from the tile's anchor is SLOW
constexpr int CL_ROWS = 64;              // source-row window of a direction (CL_TH + 2 CL_R + 2 = 58 used), 2 rows per lane
constexpr int CL_ZPAD = 4;               // dummy cells in front of every plane: the target of items without a tap
constexpr int CL_SLOTS = 6;              // footprint cells per thread of a flush half (128 threads): <= 768 cells per direction
constexpr int CL_MAXC = 31;              // channels (lanes 0 .. C-1; lane C counts the taps per footprint cell)
constexpr int CL_GS = CL_NPIX + 4;       // words between the channels of the staged grad_out tile: 16-byte aligned rows whose
                                         // 4-word groups of 8 consecutive channels cover the 32 banks (LDS.128 by lane = channel:
                                         // 3 wavefronts for 24 lanes x 4 pixels, the minimum)
#ifndef CL_CB_N
#define CL_CB_N 1
#endif
#ifndef CL_COMBINE
#define CL_COMBINE 0  // 1: pre-add the contributions of consecutive items to common footprint cells before the shared atomics
                      // (A/B: correct, 0.540 against 0.521 ms on config 2 - the uniform branches and the serial pending state cost
                      // more than the ~25 % of the ATOMS they save at |grad flow| 0.6)
#endif
#ifndef CL_PATCH
#define CL_PATCH 2  // the pixels of a warp: 2 = a 16 x 2 patch of the tile (0.515 ms on config 2), 1 = 8 x 4 (0.521), 0 = a 32 x 1 row (0.535)
#endif
#ifndef CL_DBG
#define CL_DBG 0  // 2: the channel role returns after staging grad_out (timing experiment: the pixel role alone)
#endif
#ifndef CL_REGS_K2
#define CL_REGS_K2 0  // setmaxnreg of the pixel-role warps (0: leave the launch value)
#define CL_REGS_K3 0
#endif
constexpr int CL_CB = CL_CB_N;                 // channels per batch of texture fetches in kernel 2

// -DCL_PROF: phase timeline of the kernel (clock64 deltas of the first thread of each role, summed over the CTAs; read back and
// cleared by fwb_debug_cl_prof, tools/cl_prof.py).  Slots 0.. pixel role, 16.. channel role.
#ifdef CL_PROF
__device__ unsigned long long cl_prof_acc[64];
#define CL_T0() long long cl_tprev = clock64()
#define CL_T(idx)                                                                  \
  do {                                                                             \
    if ((threadIdx.x & 255) == 0) {                                                \
      const long long now_ = clock64();                                            \
      atomicAdd(&cl_prof_acc[idx], (unsigned long long)(now_ - cl_tprev));         \
      cl_tprev = now_;                                                             \
    }                                                                              \
  } while (0)
#else
#define CL_T0()
#define CL_T(idx)
#endif

struct ClChan {  // per flattened channel (groups that have a grad_out)
  float* gs[2];  // grad_src plane of (n, t, c) per direction, or NULL
  int g, c;
  int pad_[2];
};

struct ClGrp {  // per group that has a grad_out: where the flush finds its grad_src planes
  float* gsb[2];  // grad_src of (n, t, channel 0) per direction, or NULL
  int gsc[2];     // plane stride
  int C, cf0;     // channels, first flattened channel
};

struct ClTab {  // footprint of one direction: row r is source row ybase + r, columns [rowx, rowx + len) at word rowbase of a plane
  int xlo[CL_ROWS], xhi[CL_ROWS];
  int rowx[CL_ROWS], rowbase[CL_ROWS], rowoff[CL_ROWS];
  unsigned akey;
  int cells, ybase, ok;
};

__device__ __forceinline__ void cl_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(CL_NPIX) : "memory"); }
__device__ __forceinline__ void cl_red_s32(unsigned a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_red_s32_4(unsigned a, int v) { asm volatile("red.shared.add.s32 [%0+4], %1;" ::"r"(a), "r"(v) : "memory"); }

// one warp: segment lengths and placement of one direction's rows.  A pixel recorded [x0, x0+1] on its nw row only; its sw / se
// taps are the same columns one row down, so the segment of row r is own[r] U own[r-1].
__device__ __forceinline__ void cl_tab_scan(ClTab& tb, int max_cells) {
  const int lane = threadIdx.x & 31;
  int olo[2], ohi[2], xs[2], ln[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    olo[h] = tb.xlo[2 * lane + h];
    ohi[h] = tb.xhi[2 * lane + h];
  }
  int plo = __shfl_up_sync(0xffffffffu, olo[1], 1), phi = __shfl_up_sync(0xffffffffu, ohi[1], 1);
  if (lane == 0) plo = 0x7fffffff, phi = -0x7fffffff;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int lo = min(olo[h], h == 0 ? plo : olo[0]), hi = max(ohi[h], h == 0 ? phi : ohi[0]);
    const bool has = lo <= hi;
    xs[h] = has ? lo : 0;
    ln[h] = has ? hi - lo + 1 : 0;
  }
  const int mine = ln[0] + ln[1];
  int v = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  int run = v - mine;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = 2 * lane + h;
    tb.rowx[r] = xs[h];
    tb.rowoff[r] = run;
    tb.rowbase[r] = CL_ZPAD + run;
    run += ln[h];
  }
  if (lane == 31) {
    tb.cells = run;
    tb.ok = run <= max_cells;
  }
}

// one SLOW item (pixel, direction) or the items of non-finite channels: exact float atomics to global memory, lane = channel
template <int NDIRS>
__device__ __forceinline__ void cl_exact_item(const Params& P, const GradP& Q, const ClChan* chan, const float* gos, int Cn, int n,
                                              int t, int i, int j, int pp, int d, bool lane_on) {
  const int lane = threadIdx.x & 31;
  Tap k;
  compute_tap(P.geo, P.dir[d], n, t, i, j, k);
  if (lane < Cn && lane_on) {
    const float g = gos[lane * CL_GS + pp];
    scatter_atomic_px(Q, chan[lane].g, d, n, t, chan[lane].c, k, P.dir[d].blend ? g * k.blend : g);
  }
}

// flush of `np` planes (every second plane of a group): thread <-> footprint cells htid + s * 128, s < NSL (consecutive lanes =
// consecutive cells of a source row: coalesced RED.F32).  raw + cmb = sum of the fixed-point contributions of the cell (see the
// scatter), * Sinv = float.  Cells outside the image / without a tap add an exact 0 at a clamped address; only the last slot
// (ragged: cells past the end of the footprint) is predicated.
template <int NSL>
__device__ __forceinline__ void cl_flush_planes(float* gsp, int gstep, int np, const float* sinv, unsigned ap, unsigned ap_step,
                                                const int* goff, const int* cmb, bool last_on) {
  // (opaque copies: keep the strides in registers instead of re-deriving them from the kernel parameters in the loop)
  asm volatile("" : "+r"(gstep));
  asm volatile("" : "+r"(ap_step));
  asm volatile("" : "+l"(gsp));
  int off[NSL];  // element offset of the cell in the current plane, relative to gsp (host-checked: fits in 32 bits)
#pragma unroll
  for (int s = 0; s < NSL; ++s) off[s] = goff[s];
#pragma unroll 1
  for (int q = 0; q < np; ++q) {
    const float Sinv = sinv[2 * q];
#pragma unroll
    for (int s = 0; s < NSL; ++s) {
      if (s < NSL - 1 || last_on) {
        int raw;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(raw) : "r"(ap + (unsigned)(s * (CL_NPIX / 2) * 4)));
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gsp + off[s]), "f"((float)(raw + cmb[s]) * Sinv) : "memory");
      }
      off[s] += gstep;
    }
    ap += ap_step;
  }
}

// everything channel independent about one pixel: the arithmetic of compute_tap (fwb_coords.cuh), bit for bit
template <int NDIRS>
struct ClTaps {
  float tx[NDIRS], ty[NDIRS], ux[NDIRS], uy[NDIRS], bl[NDIRS];
  int x0[NDIRS], y0[NDIRS];
  unsigned vld[NDIRS];
  unsigned clip;  // 2 bits per direction: border padding clipped ix / iy (the coordinate gradient is then zero)
};
template <int NDIRS, bool ALIGN, bool BORDER>
__device__ __forceinline__ void cl_taps(const Params& P, int n, int t, int ic, int jc, bool inimg, ClTaps<NDIRS>& k) {
  const Geo& G = P.geo;
  float lfx[NDIRS], lfy[NDIRS], lgt[NDIRS], lbl[NDIRS];  // all global loads first: one exposed memory latency
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    const DirP& D = P.dir[d];
    const DirAt at = dir_at(D, n, t);
    const int of = ic * (int)D.flow_sh + jc;
    lfx[d] = __ldg(at.flow + of);
    lfy[d] = __ldg(at.flow + D.flow_sc + of);
    lgt[d] = at.gate ? __ldg(at.gate + (ic * (int)D.gate_sh + jc)) : 1.0f;
    lbl[d] = at.blend ? __ldg(at.blend + (ic * (int)D.blend_sh + jc)) : 1.0f;
  }
  const float bx = base_coord(jc, G.W, G.stepx), by = base_coord(ic, G.H, G.stepy);
  const float fW = (float)G.W, fW1 = (float)(G.W - 1), fH = (float)G.H, fH1 = (float)(G.H - 1);
  k.clip = 0u;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    const DirP& D = P.dir[d];
    float fx = lfx[d], fy = lfy[d];
    if (D.gate != nullptr) {
      fx = __fmul_rn(fx, lgt[d]);
      fy = __fmul_rn(fy, lgt[d]);
    }
    k.bl[d] = lbl[d];
    const float gx = __fmaf_rn(D.sign, fx, bx), gy = __fmaf_rn(D.sign, fy, by);
    const float ix = source_index_fast<ALIGN, BORDER>(gx, fW, fW1), iy = source_index_fast<ALIGN, BORDER>(gy, fH, fH1);
    if (BORDER) {  // clipped <=> the unclipped coordinate was <= 0 or >= size-1
      const float cx_ = ALIGN ? __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), fW1) : __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.0f), fW, -1.0f), 0.5f);
      const float cy_ = ALIGN ? __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), fH1) : __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.0f), fH, -1.0f), 0.5f);
      k.clip |= ((unsigned)(cx_ <= 0.f || cx_ >= fW1) | ((unsigned)(cy_ <= 0.f || cy_ >= fH1) << 1)) << (2 * d);
    }
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    k.x0[d] = (int)fx0;
    k.y0[d] = (int)fy0;
    k.tx[d] = __fsub_rn(ix, fx0);
    k.ty[d] = __fsub_rn(iy, fy0);
    k.ux[d] = __fsub_rn(__fadd_rn(fx0, 1.0f), ix);
    k.uy[d] = __fsub_rn(__fadd_rn(fy0, 1.0f), iy);
    const bool xin0 = (unsigned)k.x0[d] < (unsigned)G.W, xin1 = (unsigned)(k.x0[d] + 1) < (unsigned)G.W;
    const bool yin0 = (unsigned)k.y0[d] < (unsigned)G.H, yin1 = (unsigned)(k.y0[d] + 1) < (unsigned)G.H;
    k.vld[d] = inimg ? ((unsigned)(xin0 && yin0) | ((unsigned)(xin1 && yin0) << 1) | ((unsigned)(xin0 && yin1) << 2) |
                        ((unsigned)(xin1 && yin1) << 3))
                     : 0u;
  }
}


// kernel 2's channel loop of one group: A_k += grad_out_c * tap_k,c for the quads of channel c in every direction.  The fetches
// of channel c + 1 travel while those of channel c are summed (2 .. 4 TLD4 in flight per thread).
template <int NDIRS>
struct ClK2 {
  unsigned long long th[NDIRS];
  float fx1[NDIRS], row[NDIRS];  // fetch coordinates; row advances by H per channel (integers below 2^24: exact)
  float fH;
  const float* gq;  // staged grad_out of this pixel, channel stride CL_GS
};
template <int NDIRS>
__device__ __forceinline__ void cl_k2_fetch(ClK2<NDIRS>& L, float4 (&q)[NDIRS]) {
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    q[d] = tex2Dgather<float4>((cudaTextureObject_t)L.th[d], L.fx1[d], L.row[d], 0);
    L.row[d] += L.fH;
  }
}
template <int NDIRS, bool MASKED>
__device__ __forceinline__ void cl_k2_acc(const float4 (&q)[NDIRS], float gv, float (&A)[NDIRS][4], const unsigned (&m)[NDIRS][4]) {
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    // TLD4 component order: w = (x0,y0) z = (x0+1,y0) x = (x0,y0+1) y = (x0+1,y0+1)
    float a = q[d].w, b = q[d].z, c = q[d].x, e = q[d].y;
    if (MASKED) {  // a tap outside the image counts as the value 0
      a = __uint_as_float(__float_as_uint(a) & m[d][0]);
      b = __uint_as_float(__float_as_uint(b) & m[d][1]);
      c = __uint_as_float(__float_as_uint(c) & m[d][2]);
      e = __uint_as_float(__float_as_uint(e) & m[d][3]);
    }
    A[d][0] = fmaf(gv, a, A[d][0]);
    A[d][1] = fmaf(gv, b, A[d][1]);
    A[d][2] = fmaf(gv, c, A[d][2]);
    A[d][3] = fmaf(gv, e, A[d][3]);
  }
}
template <int NDIRS, bool MASKED>
__device__ __forceinline__ void cl_k2_group(ClK2<NDIRS>& L, int C, float (&A)[NDIRS][4], const unsigned (&m)[NDIRS][4]) {
  // (a rolling window of 4 channels, 8 TLD4 in flight per thread, was measured: kernel 2's phase shrinks from 38 K to 30 K
  // cycles per tile, the channel role's phases grow by as much - the two roles share the L1TEX data path - and the kernel
  // takes 0.567 instead of 0.532 ms)
  float4 qa[NDIRS], qb[NDIRS];
  const float* gq = L.gq;
  cl_k2_fetch<NDIRS>(L, qa);
  int c = 0;
#pragma unroll 1
  for (; c + 2 < C; c += 2) {
    cl_k2_fetch<NDIRS>(L, qb);
    cl_k2_acc<NDIRS, MASKED>(qa, gq[0], A, m);
    cl_k2_fetch<NDIRS>(L, qa);
    cl_k2_acc<NDIRS, MASKED>(qb, gq[CL_GS], A, m);
    gq += 2 * CL_GS;
  }
  if (C - c == 2) {
    cl_k2_fetch<NDIRS>(L, qb);
    cl_k2_acc<NDIRS, MASKED>(qa, gq[0], A, m);
    cl_k2_acc<NDIRS, MASKED>(qb, gq[CL_GS], A, m);
  } else {
    cl_k2_acc<NDIRS, MASKED>(qa, gq[0], A, m);
  }
  L.gq += C * CL_GS;
}

struct ClSlow {  // a SLOW item with its taps (the first CL_SLOWCAP of a tile; the rest are scattered by their owner threads)
  float w[4];   // bilinear weights * blend
  int goff;     // y0 * row stride + x0 inside a grad_src plane
  int key;      // pixel << 8 | direction << 4 | validity bits
};
constexpr int CL_SLOWCAP = 64;
constexpr unsigned CL_OWNER_MARK = 4u;  // oo[].y of an item with zero weights (its ATOMS, if any, land in the dummy cells in front of a plane)

// grad_src[tap] += w_tap * g for the 4 taps of one item in one plane (ATen grid_sampler_2d_backward's atomicAdd scatter); the
// (x0, x0+1) pair of a row goes out as one 8-byte vector reduction when both taps are inside the image and the pair is aligned
__device__ __forceinline__ void cl_scatter_exact(float* gs, int sh, unsigned vld, const float* w, float g) {
  const int off[4] = {0, 1, sh, sh + 1};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float* p0 = gs + off[2 * r];
    const unsigned vv = (vld >> (2 * r)) & 3u;
    const unsigned a16 = (unsigned)(reinterpret_cast<uintptr_t>(p0) & 15u);
    if (vv == 3u && (a16 & 7u) == 0u) {
      red_add_v2(p0, w[2 * r] * g, w[2 * r + 1] * g);
    } else if (vv == 3u && a16 == 4u) {
      // the pair sits in the middle of a 16-byte quad of its row (rows are 16-byte aligned: host-checked for this kernel): one
      // 16-byte reduction whose outer lanes add an exact 0 - the L2 reduction rate is per operation, not per byte
      red_add_v4(p0 - 1, make_float4(0.0f, w[2 * r] * g, w[2 * r + 1] * g, 0.0f));
    } else {
#pragma unroll
      for (int q = 2 * r; q < 2 * r + 2; ++q)
        if (vld & (1u << q)) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gs + off[q]), "f"(w[q] * g) : "memory");
    }
  }
}

// one SLOW item scattered by the thread that owns its pixel: grad_src[tap] += w_tap * grad_out for every channel, straight to global
// memory (exact float reductions; grad_out from the staged tile)
template <int NDIRS>
__device__ __forceinline__ void cl_owner_scatter(const Params& P, const GradP& Q, const ClTaps<NDIRS>& k, int d, int n, int t,
                                                 const float* gos, int pp) {
  const Geo& G = P.geo;
  const float b = P.dir[d].blend ? k.bl[d] : 1.0f;
  const float wl = b * k.ux[d], wr = b * k.tx[d];
  const float w[4] = {wl * k.uy[d], wr * k.uy[d], wl * k.ty[d], wr * k.ty[d]};
  const float* gv = gos + pp;
  for (int g = 0; g < G.n_groups; ++g) {
    if (!Q.grad_out[g]) continue;
    const int C = P.grp[g].C;
    float* gs = Q.grad_src[g][d];
    if (gs) {
      const int sh = Q.gs_sh[g][d];
      gs += n * Q.gs_sn[g][d] + t * Q.gs_st[g][d] + (long long)k.y0[d] * sh + k.x0[d];
      const long long gsc = Q.gs_sc[g][d];
      for (int c = 0; c < C; ++c) cl_scatter_exact(gs + c * gsc, sh, k.vld[d], w, gv[c * CL_GS]);
    }
    gv += C * CL_GS;
  }
}

template <int NDIRS, bool ALIGN, bool BORDER>
__global__ void __launch_bounds__(CL_THREADS, 2) bwd_cl_kernel(const __grid_constant__ Params P, const __grid_constant__ GradP Q,
                                                               const __grid_constant__ TexP X, int acc_words, int go16) {
  extern __shared__ float4 cl_smem4[];
  __shared__ ClTab tab[NDIRS];
  __shared__ __align__(16) ClChan chan[CL_MAXC];
  __shared__ __align__(16) ClGrp grp_s[FWB_MAX_GROUPS];
  __shared__ int ngrp_s;
  __shared__ __align__(16) float4 wq[NDIRS * CL_NPIX];  // item (p, d) at d * 256 + p: bilinear weights * blend
  __shared__ __align__(8) uint2 oo[NDIRS * CL_NPIX];    //                           byte offsets of the nw / sw taps inside a plane
  __shared__ __align__(8) ClSlow slowtap[CL_SLOWCAP];
  __shared__ int nslow_s, nfar_s;
  __shared__ unsigned amax_s[32];
  __shared__ float sinv_s[32];
  __shared__ __align__(16) float zeros_s[32];
  __shared__ unsigned blmax_s;
  __shared__ unsigned gsmask_s[2];  // per direction: channels (bits) that have a grad_src plane

  const Geo& G = P.geo;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pw = warp & 7, pp = (pw << 5) | lane;  // the pixel this thread owns
  // a pixel warp is one 32 x 1 row of the tile: coalesced grad_out staging, conflict-free reads of the staged tile in kernel 2,
  // x-consecutive items in the scatter (the texture rate does not depend on the patch shape)
#if CL_PATCH == 2
  const int j = blockIdx.x * CL_TW + (pw & 1) * 16 + (lane & 15);
  const int i = blockIdx.y * CL_TH + (pw >> 1) * 2 + (lane >> 4);
#elif CL_PATCH
  const int j = blockIdx.x * CL_TW + (pw & 3) * 8 + (lane & 7);
  const int i = blockIdx.y * CL_TH + (pw >> 2) * 4 + (lane >> 3);
#else
  const int j = blockIdx.x * CL_TW + lane;
  const int i = blockIdx.y * CL_TH + pw;
#endif
  const int i0 = blockIdx.y * CL_TH;
  int n, t;
  if (G.T == 1) {
    n = blockIdx.z;
    t = 0;
  } else {
    n = blockIdx.z / G.T;
    t = blockIdx.z - n * G.T;
  }
  const bool inimg = j < G.W && i < G.H;
  const int ic = min(i, G.H - 1), jc = min(j, G.W - 1);  // ragged tiles: clamped address, results masked
  float* const gos = reinterpret_cast<float*>(cl_smem4);
  int Cn = 0;  // flattened channels of the groups that have a grad_out (host-checked: 1 .. CL_MAXC)
  for (int g = 0; g < G.n_groups; ++g)
    if (Q.grad_out[g]) Cn += P.grp[g].C;

  if (warp < 8) {
    // ==================================================================== pixel role
#if CL_REGS_K2
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CL_REGS_K2));
#endif
    CL_T0();
    ClTaps<NDIRS> k;
    cl_taps<NDIRS, ALIGN, BORDER>(P, n, t, ic, jc, inimg, k);
#ifdef CL_DELAY
    __nanosleep(CL_DELAY);  // experiment: is the kernel bound by the pixel role's latency chain (time grows with the delay) or by throughput?
#endif
    CL_T(0);
    unsigned slowbits = 0u;
    {
      // ---- footprint tables and item descriptors of the scatter (thread = pixel; barrier 1 = the 8 pixel warps)
      for (int q = tid; q < NDIRS * CL_ROWS; q += CL_NPIX) {
        ClTab& T = tab[(NDIRS > 1 && q >= CL_ROWS) ? 1 : 0];
        const int r = q >= CL_ROWS ? q - CL_ROWS : q;
        T.xlo[r] = 0x7fffffff;
        T.xhi[r] = -0x7fffffff;
      }
      if (tid < NDIRS) tab[tid].akey = 0xffffffffu;
      if (tid == 0) {
        blmax_s = 0u;
        nslow_s = 0;
        nfar_s = 0;
      }
      cl_bar(1);
      float blm = 0.f;
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        if (inimg) blm = __uint_as_float(max(__float_as_uint(blm), __float_as_uint(P.dir[d].blend ? fabsf(k.bl[d]) : 1.0f)));
        // anchor vote: the displacement (x0 - j, y0 - i) of the tile's least displaced item, as one 32-bit key
        const int dx = min(max(k.x0[d] - j, -1024), 1023), dy = min(max(k.y0[d] - i, -1024), 1023);
        const unsigned mag = (unsigned)min(abs(dx) + abs(dy), 1023);
        const unsigned key = k.vld[d] ? ((mag << 22) | (((unsigned)dx & 0x7ffu) << 11) | ((unsigned)dy & 0x7ffu)) : 0xffffffffu;
        const unsigned best = __reduce_min_sync(0xffffffffu, key);
        if (lane == 0 && best != 0xffffffffu) atomicMin(&tab[d].akey, best);
      }
      {
        const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(blm));  // NaN wins (bit patterns of non-negative floats)
        if (lane == 0 && mb != 0u) atomicMax(&blmax_s, mb);
      }
      cl_bar(1);
      bool fast[NDIRS], slow[NDIRS];
      int rr[NDIRS];
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const unsigned ak = tab[d].akey;  // ~0 when no item of the tile has a tap
        const int adx = ((int)(ak << 10)) >> 21, ady = ((int)(ak << 21)) >> 21;
        const bool far = abs(k.x0[d] - j - adx) > CL_R || abs(k.y0[d] - i - ady) > CL_R;
        fast[d] = k.vld[d] != 0u && !far;
        slow[d] = k.vld[d] != 0u && far;
        rr[d] = k.y0[d] - (i0 + ady - CL_R);  // 0 .. CL_TH + 2 CL_R - 1 for fast items
        if (tid == 0) tab[d].ybase = i0 + ady - CL_R;
        if (fast[d]) {
          atomicMin(&tab[d].xlo[rr[d]], k.x0[d]);
          atomicMax(&tab[d].xhi[rr[d]], k.x0[d] + 1);
        }
        const unsigned fm = __ballot_sync(0xffffffffu, slow[d]);
        if (lane == 0 && fm) atomicAdd(&nfar_s, __popc(fm));
      }
      cl_bar(1);
      if (warp == 0) {
        const int cmax = min(acc_words / (Cn + 1) - CL_ZPAD - 1, CL_SLOTS * (CL_NPIX / 2));
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) cl_tab_scan(tab[d], cmax);
      }
      cl_bar(1);
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const ClTab& T = tab[d];
        const float b = P.dir[d].blend ? k.bl[d] : 1.0f;
        const float wl = b * k.ux[d], wr = b * k.tx[d];
        float4 w = make_float4(wl * k.uy[d], wr * k.uy[d], wl * k.ty[d], wr * k.ty[d]);
        // taps outside the image (zeros padding) add nothing: their footprint cells only ever receive zeros
        if (!(k.vld[d] & 1u)) w.x = 0.f;
        if (!(k.vld[d] & 2u)) w.y = 0.f;
        if (!(k.vld[d] & 4u)) w.z = 0.f;
        if (!(k.vld[d] & 8u)) w.w = 0.f;
        uint2 o = make_uint2(0u, 0u);
        if (fast[d] && T.ok) {
          o.x = 4u * (unsigned)(T.rowbase[rr[d]] + (k.x0[d] - T.rowx[rr[d]]));
          o.y = 4u * (unsigned)(T.rowbase[rr[d] + 1] + (k.x0[d] - T.rowx[rr[d] + 1]));
        } else {
          if (slow[d] || fast[d]) {
            // SLOW: a few outliers per tile go to the channel role with their taps (lanes = channels); when there are many
            // (large displacements, a footprint that does not fit) every thread scatters its own items after kernel 2
            const int slot = (slow[d] && nfar_s <= CL_SLOWCAP) ? atomicAdd(&nslow_s, 1) : CL_SLOWCAP;
            if (slot < CL_SLOWCAP) {
              int sh = 0;
              for (int g = G.n_groups - 1; g >= 0; --g)
                if (Q.grad_out[g] && Q.grad_src[g][d]) sh = Q.gs_sh[g][d];
              ClSlow e;
              e.w[0] = w.x, e.w[1] = w.y, e.w[2] = w.z, e.w[3] = w.w;
              e.goff = k.y0[d] * sh + k.x0[d];
              e.key = (pp << 8) | (d << 4) | (int)k.vld[d];
              slowtap[slot] = e;
            } else {
              slowbits |= 1u << d;
              // with two directions the items of the second one are scattered by the channel-role thread of the same pixel
              // (that role idles when footprints do not fit, while this one still has kernel 2 to do): marked in the descriptor
              if (NDIRS == 2 && d == 1) o.y = CL_OWNER_MARK;
            }
          }
          w = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        wq[d * CL_NPIX + pp] = w;
        oo[d * CL_NPIX + pp] = o;
      }
    }
    CL_T(1);
    __syncthreads();  // descriptors and tables -> channel role; the staged grad_out tile -> kernel 2
    CL_T(2);
    // ---- kernel 2 on the texture units
    // gix = sum_c gw_c [uy (b_c - a_c) + ty (d_c - c_c)] etc. (ATen grid_sampler_2d_backward): the weights do not depend on the
    // channel, so only the four sums  A_k = sum_c grad_out_c * tap_k,c  are accumulated per (pixel, direction) - one TLD4 and
    // four FFMA per channel - and the weights are applied once at the end.  The fetches run one batch ahead of the sums.
    float fx1[NDIRS], fy1[NDIRS];
    bool part = false;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      fx1[d] = k.vld[d] ? (float)(k.x0[d] + 1) : 0.0f;
      fy1[d] = k.vld[d] ? (float)(k.y0[d] + 1) : 0.0f;
      part |= k.vld[d] != 15u;
    }
    const bool masked = __any_sync(0xffffffffu, part);  // some tap of this warp is outside the image (borders, ragged tiles)
    float A[NDIRS][4];
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) A[d][0] = A[d][1] = A[d][2] = A[d][3] = 0.0f;
    {
      ClK2<NDIRS> L;
      L.fH = (float)G.H;
      L.gq = gos + pp;
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) L.fx1[d] = fx1[d];
      unsigned m[NDIRS][4];
      if (masked) {
#pragma unroll
        for (int d = 0; d < NDIRS; ++d)
#pragma unroll
          for (int q = 0; q < 4; ++q) m[d][q] = (k.vld[d] >> q) & 1u ? 0xffffffffu : 0u;
      }
      for (int g = 0; g < G.n_groups; ++g) {
        if (!Q.grad_out[g]) continue;
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) {
          const TexSrc& S = X.s[g][d];
          const int blk = n / S.nb;
          L.th[d] = S.tex[blk];
          L.row[d] = (float)((n - blk * S.nb) * S.rows_n + t * S.rows_t) + fy1[d];
        }
        if (masked)
          cl_k2_group<NDIRS, true>(L, P.grp[g].C, A, m);
        else
          cl_k2_group<NDIRS, false>(L, P.grp[g].C, A, m);
      }
    }
    CL_T(3);
    // the taps are recomputed for the epilogue (flow / mask reloads hit the L1): nothing but the fetch coordinates and the four
    // sums per direction lives across the channel loop, which keeps it free of spills and rematerialised address arithmetic
    {
      int ic2 = ic, jc2 = jc;
      asm volatile("" : "+r"(ic2), "+r"(jc2));
      cl_taps<NDIRS, ALIGN, BORDER>(P, n, t, ic2, jc2, inimg, k);
    }
    if (inimg) {
      // coordinate gradient -> grad_flow / grad_gate / grad_blend.  The multipliers are d(ix)/d(gx) = W/2 or (W-1)/2, zero where
      // border padding clipped the coordinate; the raw flow and the gate are reloaded only when there is a gate.
      const float mxc = ALIGN ? __fmul_rn((float)(G.W - 1), 0.5f) : __fmul_rn((float)G.W, 0.5f);
      const float myc = ALIGN ? __fmul_rn((float)(G.H - 1), 0.5f) : __fmul_rn((float)G.H, 0.5f);
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const float a = A[d][0], b = A[d][1], cc = A[d][2], dd = A[d][3];
        float gbl = 0.0f, sc = 1.0f;
        if (P.dir[d].blend != nullptr) {
          const float top = fmaf(b, k.tx[d], a * k.ux[d]), bot = fmaf(dd, k.tx[d], cc * k.ux[d]);
          gbl = fmaf(bot, k.ty[d], top * k.uy[d]);
          sc = k.bl[d];
        }
        const float gix = sc * fmaf(k.ty[d], dd - cc, k.uy[d] * (b - a));
        const float giy = sc * fmaf(k.tx[d], dd - b, k.ux[d] * (cc - a));
        const unsigned cb = k.clip >> (2 * d);
        Tap kk;
        kk.mx = (cb & 1u) ? 0.f : mxc;
        kk.my = (cb & 2u) ? 0.f : myc;
        kk.fx = kk.fy = 0.f;
        kk.gate = 1.f;
        const DirP& D = P.dir[d];
        if (D.gate != nullptr) {
          const DirAt at = dir_at(D, n, t);
          const int of = i * (int)D.flow_sh + j;
          kk.fx = __ldg(at.flow + of);
          kk.fy = __ldg(at.flow + D.flow_sc + of);
          kk.gate = __ldg(at.gate + (i * (int)D.gate_sh + j));
        }
        bwdflow_store(P, Q, d, n, t, i, j, kk, gix, giy, gbl);
      }
    }
    // ---- SLOW items (taps far from the rest of the tile, or a direction whose footprint does not fit): this thread owns the
    // taps, so it adds its contributions straight to global memory (exact float reductions; grad_out from the staged tile)
    CL_T(4);
    if (slowbits) {
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        if (!((slowbits >> d) & 1u)) continue;
        if (NDIRS == 2 && d == 1) continue;  // (handed to the channel-role thread of this pixel, see CL_OWNER_MARK)
        cl_owner_scatter<NDIRS>(P, Q, k, d, n, t, gos, pp);
      }
    }
    CL_T(5);
    return;
  }

  // ====================================================================== channel role: kernel 3 (scatter into grad_src)
#if CL_REGS_K3
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CL_REGS_K3));
#endif
  const int ctid = tid - CL_NPIX;
  CL_T0();
  // ---- the grad_out tile travels to shared memory while the taps are computed: 16-byte cp.async when the planes allow it (go16:
  // W % 4 == 0, 16-byte aligned pointers and strides - a thread copies 4 consecutive pixels of a patch row, channels 4 apart),
  // else 4-byte copies (any plane stride).  Out-of-image pixels of ragged tiles stage the value of a clamped address: their
  // items carry zero weights.
  if (go16) {
    // chunk q = 4 consecutive pixels: patch pw_ (8 x 4 pixels, pp order), row r_ of the patch, half h_ of the row
    const int q = ctid & 63, cg = ctid >> 6;
    const int pw_ = q >> 3;
#if CL_PATCH == 1  // 8 x 4 patches: row r_ of the patch, half h_ of the row
    const int r_ = (q >> 1) & 3, h_ = q & 1;
    const int cj = min(blockIdx.x * CL_TW + (pw_ & 3) * 8 + h_ * 4, G.W - 4), ci = min(blockIdx.y * CL_TH + (pw_ >> 2) * 4 + r_, G.H - 1);
    const unsigned gd0 = (unsigned)__cvta_generic_to_shared(gos + (pw_ << 5) + r_ * 8 + h_ * 4);
#elif CL_PATCH == 2  // 16 x 2 patches: row r_ of the patch, quarter h_ of the row
    const int r_ = (q >> 2) & 1, h_ = q & 3;
    const int cj = min(blockIdx.x * CL_TW + (pw_ & 1) * 16 + h_ * 4, G.W - 4), ci = min(blockIdx.y * CL_TH + (pw_ >> 1) * 2 + r_, G.H - 1);
    const unsigned gd0 = (unsigned)__cvta_generic_to_shared(gos + (pw_ << 5) + r_ * 16 + h_ * 4);
#else  // 32 x 1 rows: eighth h_ of the row
    const int h_ = q & 7;
    const int cj = min(blockIdx.x * CL_TW + h_ * 4, G.W - 4), ci = min(blockIdx.y * CL_TH + pw_, G.H - 1);
    const unsigned gd0 = (unsigned)__cvta_generic_to_shared(gos + (pw_ << 5) + h_ * 4);
#endif
    int cf0 = 0;  // flattened index of the group's channel 0
    for (int g = 0; g < G.n_groups; ++g) {
      if (!Q.grad_out[g]) continue;
      const int C = P.grp[g].C;
      const long long sc = Q.go_sc[g];
      const float* gp = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)ci * Q.go_sh[g] + cj + cg * sc;
      unsigned gd = gd0 + 4u * CL_GS * (unsigned)(cf0 + cg);
      for (int c = cg; c < C; c += 4) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gd), "l"(gp) : "memory");
        gp += 4 * sc;
        gd += 16u * CL_GS;
      }
      cf0 += C;
    }
    cp_async_commit();
  } else {
    unsigned gd = (unsigned)__cvta_generic_to_shared(gos + pp);
    for (int g = 0; g < G.n_groups; ++g) {
      if (!Q.grad_out[g]) continue;
      const float* gp = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)ic * Q.go_sh[g] + jc;
      const int C = P.grp[g].C;
      const long long sc = Q.go_sc[g];
#pragma unroll 4
      for (int c = 0; c < C; ++c) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(gd), "l"(gp) : "memory");
        gp += sc;
        gd += 4u * CL_GS;
      }
    }
    cp_async_commit();
  }
  if (warp == 8) {  // channel table over the groups that have a grad_out (the others contribute nothing)
    int base = 0, ng = 0;
    unsigned m0 = 0u, m1 = 0u;
    for (int g = 0; g < G.n_groups; ++g) {
      if (!Q.grad_out[g]) continue;
      const int C = P.grp[g].C;
      ClChan e;
      e.g = g;
      e.c = lane;
      e.pad_[0] = e.pad_[1] = 0;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        float* gs = (d < NDIRS && lane < C) ? Q.grad_src[g][d] : nullptr;
        e.gs[d] = gs ? gs + n * Q.gs_sn[g][d] + t * Q.gs_st[g][d] + (long long)lane * Q.gs_sc[g][d] : nullptr;
      }
      if (lane < C) chan[base + lane] = e;
      if (lane == 0) {
        ClGrp gr;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          gr.gsb[d] = e.gs[d];  // lane 0: channel 0 of the group
          gr.gsc[d] = d < NDIRS ? Q.gs_sc[g][d] : 0;
        }
        gr.C = C;
        gr.cf0 = base;
        grp_s[ng] = gr;
      }
      ++ng;
      m0 |= __ballot_sync(0xffffffffu, e.gs[0] != nullptr) << base;
      m1 |= __ballot_sync(0xffffffffu, e.gs[1] != nullptr) << base;
      base += C;
    }
    amax_s[lane] = 0u;
    zeros_s[lane] = 0.0f;
    if (lane == 0) {
      gsmask_s[0] = m0;
      gsmask_s[1] = m1;
      ngrp_s = ng;
    }
  }
  int gsh[NDIRS];  // row stride of the grad_src planes per direction (all groups agree, host-checked)
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    gsh[d] = 0;
    for (int g = G.n_groups - 1; g >= 0; --g)
      if (Q.grad_out[g] && Q.grad_src[g][d]) gsh[d] = Q.gs_sh[g][d];
  }
  CL_T(16);
  cp_async_wait<0>();
  CL_T(17);
  __syncthreads();
  CL_T(18);
  if (CL_DBG & 2) return;  // timing experiment: the pixel role alone


  int* const acc = reinterpret_cast<int*>(gos + ((Cn * CL_GS + 3) & ~3));
  const unsigned acc_s = (unsigned)__cvta_generic_to_shared(acc);
  constexpr float MAGIC = 12582912.0f;  // 1.5 * 2^23: fma(x, y, MAGIC) holds round-to-nearest(x*y) in its low mantissa bits
  constexpr int MAGIC_BITS = 0x4B400000;
  const float* gl = gos + (lane < Cn ? lane : 0) * CL_GS + (pw << 5);
  if (lane < Cn) {  // max |grad_out| of this lane's channel over the warp's 32 pixels; NaN wins (integer compare of the magnitudes)
    unsigned m = 0u;
#pragma unroll
    for (int u = 0; u < 32; u += 4) {
      const float4 g4 = *reinterpret_cast<const float4*>(gl + u);
      m = max(max(m, __float_as_uint(g4.x) & 0x7fffffffu), __float_as_uint(g4.y) & 0x7fffffffu);
      m = max(max(m, __float_as_uint(g4.z) & 0x7fffffffu), __float_as_uint(g4.w) & 0x7fffffffu);
    }
    if (m != 0u) atomicMax(&amax_s[lane], m);
  }
  CL_T(19);
  float S = 0.f, magic = 0.f;
  bool finite = true, any_nonfinite = false, scaled = false;
#pragma unroll 1
  for (int d = 0; d <= NDIRS; ++d) {
    // (one more round than there are directions: the scale of the channels is needed even when no direction scatters)
    const bool live = d < NDIRS && tab[d < NDIRS ? d : 0].ok && tab[d < NDIRS ? d : 0].cells != 0 && gsmask_s[d < NDIRS ? d : 0] != 0u;
    if (!live && (scaled || d < NDIRS)) continue;  // (uniform)
    const ClTab& T = tab[d < NDIRS ? d : 0];
    const unsigned gsm = live ? gsmask_s[d] : 0u;
    const int PS = (CL_ZPAD + T.cells) | 1;
    if (live) {  // clear the planes (and the tap counter plane)
      const int n4 = ((Cn + 1) * PS + 3) >> 2;
      for (int q = ctid; q < n4; q += CL_NPIX) asm volatile("st.shared.v4.s32 [%0], {%1,%1,%1,%1};" ::"r"(acc_s + 16u * (unsigned)q), "r"(0) : "memory");
    }
    CL_T(20 + 8 * d);
    cl_bar(2);  // the planes are clear; (first round) the maxima of all eight warps are in
    CL_T(21 + 8 * d);
    if (!scaled) {  // scale of this lane's channel
      scaled = true;
      if (lane < Cn) {
        const unsigned ab = amax_s[lane], bb = blmax_s;
        const float prod = __uint_as_float(ab) * __uint_as_float(bb);
        const unsigned pb = __float_as_uint(prod);
        finite = ab < 0x7f800000u && bb < 0x7f800000u && pb < 0x7f800000u;
        const int sexp = min(252, max(2, 274 - (int)(pb >> 23)));  // biased exponent of 2^(20 - exponent(amax))
        if (finite && pb != 0u) S = __uint_as_float((unsigned)sexp << 23);
        if (warp == 8) sinv_s[lane] = (finite && pb != 0u) ? __uint_as_float((unsigned)(254 - sexp) << 23) : 0.f;
      }
      any_nonfinite = __any_sync(0xffffffffu, !finite);
      if (lane >= Cn || !finite) gl = zeros_s;  // the tap counter lane and non-finite channels put in exact zeros
      // lane Cn counts the taps per cell: fma(0, w, denorm_min) has the bit pattern 1
      magic = lane == Cn ? __int_as_float(1) : MAGIC;
    }
    if (!live) continue;
    if (lane <= Cn) {
      // every tap adds the raw bits of fma(g * S, w, MAGIC) = MAGIC_BITS + k (k = the rounded fixed-point product); a cell hit
      // by cnt taps then holds cnt * MAGIC_BITS + sum k (mod 2^32): the flush takes cnt * MAGIC_BITS off again
      const float Sd = ((gsm >> lane) & 1u) ? S : 0.f;
      const unsigned la = acc_s + 4u * (unsigned)(lane * PS);
      const float4* wd = wq + d * CL_NPIX + (pw << 5);
      const uint2* od = oo + d * CL_NPIX + (pw << 5);
      // Consecutive items of a warp (x-neighbours of a patch row) mostly land on the same or on the next footprint cells: their
      // contributions to a common cell are added in registers first (raw fixed-point bits: integer adds, exact; the tap counter lane
      // adds its 1s the same way), so a run of k neighbours costs 2 k + 2 ATOMS instead of 4 k.  The comparisons are on the item's
      // offsets, i.e. warp-uniform.  PA/PB: pending cell pair of the upper / lower tap row, P[0..3] their pending sums.
#if CL_COMBINE
      unsigned PA = 0u, PB = 0u;
      int P0 = 0, P1 = 0, P2 = 0, P3 = 0;
      bool pend = false;
#endif
#pragma unroll 1
      for (int u = 0; u < 32; u += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gl + u);  // grad_out of this lane's channel at 4 items
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float gsv = gv[v] * Sd;
          const float4 w = wd[u + v];
          const uint2 o = od[u + v];
          const unsigned a0 = la + o.x, a1 = la + o.y;
          const int p0 = __float_as_int(fmaf(gsv, w.x, magic)), p1 = __float_as_int(fmaf(gsv, w.y, magic));
          const int p2 = __float_as_int(fmaf(gsv, w.z, magic)), p3 = __float_as_int(fmaf(gsv, w.w, magic));
#if CL_COMBINE
          if (pend && a0 == PA + 4u && a1 == PB + 4u) {  // one cell to the right: the left column of the pending pair is complete
            cl_red_s32(PA, P0);
            cl_red_s32(PB, P2);
            P0 = P1 + p0, P2 = P3 + p2, P1 = p1, P3 = p3;
            PA = a0, PB = a1;
          } else if (pend && a0 == PA && a1 == PB) {  // the same four cells
            P0 += p0, P1 += p1, P2 += p2, P3 += p3;
          } else {
            if (pend) {
              cl_red_s32(PA, P0);
              cl_red_s32_4(PA, P1);
              cl_red_s32(PB, P2);
              cl_red_s32_4(PB, P3);
            }
            P0 = p0, P1 = p1, P2 = p2, P3 = p3;
            PA = a0, PB = a1;
            pend = true;
          }
#else
          cl_red_s32(a0, p0);
          cl_red_s32_4(a0, p1);
          cl_red_s32(a1, p2);
          cl_red_s32_4(a1, p3);
#endif
        }
      }
#if CL_COMBINE
      if (pend) {
        cl_red_s32(PA, P0);
        cl_red_s32_4(PA, P1);
        cl_red_s32(PB, P2);
        cl_red_s32_4(PB, P3);
      }
#endif
    }
    CL_T(22 + 8 * d);
    cl_bar(2);
    CL_T(23 + 8 * d);
    // flush: two halves of 128 threads, half h takes the planes cf = h, h + 2, ...
    const int half = ctid >> 7, htid = ctid & 127;
    const int nsl = (T.cells + CL_NPIX / 2 - 1) / (CL_NPIX / 2);
    int goff[CL_SLOTS], cmb[CL_SLOTS];
    bool last_on = false;
#pragma unroll
    for (int s = 0; s < CL_SLOTS; ++s) {
      goff[s] = 0;
      cmb[s] = 0;
      if (s < nsl) {
        const int kk = htid + s * (CL_NPIX / 2);
        if (kk < T.cells) {
          int r = 0;  // last row with rowoff[r] <= kk
#pragma unroll
          for (int step = CL_ROWS / 2; step > 0; step >>= 1)
            if (T.rowoff[r + step] <= kk) r += step;
          const int y = T.ybase + r, col = T.rowx[r] + (kk - T.rowoff[r]);
          const int cnt = acc[Cn * PS + CL_ZPAD + kk];
          // a cell outside the image only ever received zeros (weights masked): it adds 0.0 at the clamped address
          goff[s] = min(max(y, 0), G.H - 1) * gsh[d] + min(max(col, 0), G.W - 1);
          cmb[s] = -cnt * MAGIC_BITS;
          if (s == nsl - 1) last_on = true;
        }
      }
    }
    CL_T(24 + 8 * d);
    const unsigned ps_b = 4u * (unsigned)PS;
    const unsigned ap0 = acc_s + 4u * (unsigned)(CL_ZPAD + htid);
    const int ngrp = ngrp_s;
    for (int gq_ = 0; gq_ < ngrp; ++gq_) {
      const ClGrp& GR = grp_s[gq_];
      float* gsb = GR.gsb[d];
      if (!gsb) continue;
      const int C = GR.C, cf = GR.cf0, gsc = GR.gsc[d];
      const int c0 = (half ^ cf) & 1;  // first channel of this group whose flattened index has this half's parity
      const int np = (C - c0 + 1) >> 1;
      float* gsp = gsb + (long long)c0 * gsc;
      const unsigned ap = ap0 + ps_b * (unsigned)(cf + c0);
      const float* si = sinv_s + cf + c0;
      switch (nsl) {
        case 1: cl_flush_planes<1>(gsp, 2 * gsc, np, si, ap, 2 * ps_b, goff, cmb, last_on); break;
        case 2: cl_flush_planes<2>(gsp, 2 * gsc, np, si, ap, 2 * ps_b, goff, cmb, last_on); break;
        case 3: cl_flush_planes<3>(gsp, 2 * gsc, np, si, ap, 2 * ps_b, goff, cmb, last_on); break;
        case 4: cl_flush_planes<4>(gsp, 2 * gsc, np, si, ap, 2 * ps_b, goff, cmb, last_on); break;
        case 5: cl_flush_planes<5>(gsp, 2 * gsc, np, si, ap, 2 * ps_b, goff, cmb, last_on); break;
        default: cl_flush_planes<6>(gsp, 2 * gsc, np, si, ap, 2 * ps_b, goff, cmb, last_on); break;
      }
    }
    CL_T(25 + 8 * d);
    cl_bar(2);  // the planes are reused by the next direction; sinv_s is read by every warp
    CL_T(26 + 8 * d);
  }
  // the cached SLOW items: exact float reductions, lane = channel
  {
    const int nslow = min(nslow_s, CL_SLOWCAP);
    for (int q = pw; q < nslow; q += 8) {
      const ClSlow& e = slowtap[q];
      const int sp = e.key >> 8, sd = (e.key >> 4) & 1;
      if (lane < Cn) {
        float* gs = chan[lane].gs[sd];
        if (gs) cl_scatter_exact(gs + e.goff, gsh[sd], (unsigned)e.key & 15u, e.w, gos[lane * CL_GS + sp]);
      }
    }
  }
  CL_T(44);
  // the owner-scattered items of the second direction: this thread's pixel is pp as well; taps recomputed (flow / mask reloads hit the L1)
  if (NDIRS == 2) {
    const uint2 om = oo[CL_NPIX + pp];
    if (__any_sync(0xffffffffu, om.x == 0u && om.y == CL_OWNER_MARK)) {
      ClTaps<NDIRS> k;
      cl_taps<NDIRS, ALIGN, BORDER>(P, n, t, ic, jc, inimg, k);
      if (om.x == 0u && om.y == CL_OWNER_MARK) cl_owner_scatter<NDIRS>(P, Q, k, 1, n, t, gos, pp);
    }
  }
  // every (fast) item of a non-finite channel: exact float atomics, lane = channel
  if (any_nonfinite) {
    for (int u = 0; u < 32; ++u) {
      const int sp = (pw << 5) | u;
#if CL_PATCH == 2
      const int sj = blockIdx.x * CL_TW + (pw & 1) * 16 + (u & 15), si = blockIdx.y * CL_TH + (pw >> 1) * 2 + (u >> 4);
#elif CL_PATCH
      const int sj = blockIdx.x * CL_TW + (pw & 3) * 8 + (u & 7), si = blockIdx.y * CL_TH + (pw >> 2) * 4 + (u >> 3);
#else
      const int sj = blockIdx.x * CL_TW + u, si = blockIdx.y * CL_TH + pw;
#endif
      for (int d = 0; d < NDIRS; ++d)
        if (oo[d * CL_NPIX + sp].x != 0u) cl_exact_item<NDIRS>(P, Q, chan, gos, Cn, n, t, si, sj, sp, d, !finite);
    }
  }
}

}  // namespace fwb

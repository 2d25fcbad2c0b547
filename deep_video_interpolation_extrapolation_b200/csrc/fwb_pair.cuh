// fwb_pair.cuh — kernels 1 and 2 on shared-memory tiles of CHANNEL PAIRS.
//
// Both kernels are bound by instruction issue once the gather no longer thrashes L1 (ncu: the planar staged
// version ran at 43-56 % issue utilisation for 170 M warp instructions against the generic kernel's 172 M).
// What cuts the instruction count is the layout of the staged tile: two channel planes are INTERLEAVED in
// shared memory (cell = float2 {c, c+1}), so that
//   * one LDS.64 fetches a tap for two channels,
//   * one FFMA2 (Blackwell's packed fp32 FMA, `fma.rn.f32x2`) accumulates both — the two channels share the
//     bilinear weight, and each half is an IEEE fp32 fma, bit-identical to the scalar chain,
// i.e. 4 LDS.64 + 4 FFMA2 per (pixel, direction, channel PAIR) instead of 8 LDS + 8 FFMA.
// The interleave happens on the way in, in two hops.  (1) cp.async copies a thread's 16-byte pieces of planes
// c and c+1 into a thread-private planar scratch ring in shared memory, PR_NS - 1 pairs ahead of the gather
// (the copies are in flight while earlier pairs are computed; no registers are held).  (2) Once its own copies
// of the next pair have landed (cp.async.wait_group — no barrier needed, the scratch is private), the thread
// reads them back (2 LDS.128) and writes them interleaved (2 STS.128 {c0,d0,c1,d1} {c2,d2,c3,d3}) into the
// gather buffer of the next pair.  One __syncthreads per channel pair.
// A CTA owns a 32x16 tile of output pixels of one (n, t); 8 warps; warp w owns rows w and w+8, lane = column.
#pragma once
#include "fwb_coords.cuh"
#include "fwb_generic.cuh"
#include "fwb_stage.cuh"

namespace fwb {

constexpr int PR_TW = 32, PR_TH = 16;
constexpr int PR_THREADS = 256;
constexpr int PR_PPT = 2;    // pixels per thread
constexpr int PR_SLOTS = 2;  // 16-byte pieces per thread, plane and direction -> <= 512 float4 per plane
constexpr int PR_MAXSLOW = 32;  // slow pixels a tile may have before the whole tile goes generic
constexpr int PR_NS = 3;     // planar scratch stages: copies run PR_NS - 1 channel pairs ahead of the gather

typedef unsigned long long f32x2;  // two packed floats in a 64-bit register pair

__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 as2(float2 v) { return pk2(v.x, v.y); }

// ---------------------------------------------------------------------------------------------
// prologue shared by both kernels
// ---------------------------------------------------------------------------------------------
struct PairPix {  // per (pixel of this thread, direction)
  float tx, ty, ux, uy, bl;
  int o0, o1;  // cell offsets, inside a staged pair plane, of the nw and sw taps (direction slot included)
};

struct PairLoad {  // the pieces this thread loads for every plane of a direction
  int ycol[PR_SLOTS];   // (y << 16) | col of the first float (0 when the piece is outside the image)
  int bytes[PR_SLOTS];  // bytes inside the image: 16, 0 (zero fill), 4/8/12 (right edge), -1 = no piece
};

template <int NDIRS>
struct PairCtx {
  int n, t, j;
  int irow[PR_PPT];
  bool inimg[PR_PPT];  // pixel is inside the image
  bool act[PR_PPT];    // ... and served by the staged loop (not slow)
  PairPix px[PR_PPT][NDIRS];
  PairLoad ld[NDIRS];
  int slot[NDIRS];  // cells per direction slot (ZPAD included)
  int cells;        // cells per pair plane
  int pieces;       // 16-byte pieces per plane over all directions
  int ok;
};

__device__ __forceinline__ void pair_assign(const StageTab& tb, int H, int W, PairLoad& ld) {
  const int span = tb.ymax + 1 - tb.ymin;
#pragma unroll
  for (int s = 0; s < PR_SLOTS; ++s) {
    const int k = threadIdx.x + s * PR_THREADS;
    ld.ycol[s] = 0;
    ld.bytes[s] = -1;
    if (k < tb.total4) {
      int lo = 0, hi = span - 1;  // first r with rowoff4[r+1] > k
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tb.rowoff4[mid + 1] > k)
          hi = mid;
        else
          lo = mid + 1;
      }
      const int y = tb.ymin + lo, col = tb.rowx[lo] + 4 * (k - tb.rowoff4[lo]);
      const bool in = y >= 0 && y < H && col >= 0 && col < W;
      ld.ycol[s] = in ? ((y << 16) | col) : 0;
      ld.bytes[s] = in ? 4 * min(4, W - col) : 0;
    }
  }
}

template <int NDIRS, bool ALIGN, bool BORDER>
__device__ __forceinline__ void pair_prologue(const Params& P, StageTab* tb, StageSlow& slow, Tap (*slowtap)[NDIRS], float* smem,
                                              int smem_floats, int n_stages, PairCtx<NDIRS>& cx) {
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cx.j = blockIdx.x * PR_TW + lane;
  cx.n = blockIdx.z / G.T;
  cx.t = blockIdx.z - cx.n * G.T;
#pragma unroll
  for (int q = 0; q < PR_PPT; ++q) {
    cx.irow[q] = blockIdx.y * PR_TH + warp + 8 * q;
    cx.inimg[q] = cx.j < G.W && cx.irow[q] < G.H;
  }
  int x0[PR_PPT][NDIRS], y0[PR_PPT][NDIRS];
  bool has[PR_PPT][NDIRS];
  stage_tab_init(tb, NDIRS);
  if (threadIdx.x == 0) slow.n = 0;
  __syncthreads();
#pragma unroll
  const float bx = base_coord(cx.j, G.W, G.stepx);
  for (int d = 0; d < NDIRS; ++d) {
    const DirAt at = dir_at(P.dir[d], cx.n, cx.t);
#pragma unroll
    for (int q = 0; q < PR_PPT; ++q) {
      FastTap k;
      k.valid = 0u;
      k.x0 = k.y0 = 0;
      k.ux = k.uy = k.tx = k.ty = 0.f;
      k.blend = 0.f;
      if (cx.inimg[q]) compute_tap_fast<ALIGN, BORDER>(G, P.dir[d], at, bx, cx.irow[q], cx.j, k);
      has[q][d] = k.valid != 0u;
      PairPix& px = cx.px[q][d];
      px.tx = k.tx;
      px.ty = k.ty;
      px.ux = k.ux;
      px.uy = k.uy;
      px.bl = k.blend;
      x0[q][d] = k.x0;
      y0[q][d] = k.y0;
      stage_anchor_vote(tb[d], has[q][d], k.x0 - cx.j, k.y0 - cx.irow[q]);
    }
  }
  __syncthreads();
  // slow pixels: some direction's taps are far from where the rest of the tile samples
#pragma unroll
  for (int q = 0; q < PR_PPT; ++q) {
    bool far = false;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) far |= has[q][d] && !stage_inlier(tb[d], x0[q][d] - cx.j, y0[q][d] - cx.irow[q]);
    cx.act[q] = cx.inimg[q] && !far;
    if (far) {
      const int slot = atomicAdd(&slow.n, 1);
      if (slot < PR_MAXSLOW) slow.pix[slot] = (unsigned short)(((warp + 8 * q) << 5) | lane);
    }
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      has[q][d] = has[q][d] && cx.act[q];
      stage_tab_add(tb[d], has[q][d], x0[q][d], y0[q][d]);
    }
  }
  __syncthreads();
  if (warp < NDIRS) stage_tab_scan(tb[warp]);
  __syncthreads();
  cx.ok = slow.n <= PR_MAXSLOW;
  cx.cells = 0;
  cx.pieces = 0;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    cx.ok &= tb[d].ok && tb[d].total4 <= PR_SLOTS * PR_THREADS;
    cx.slot[d] = ST_ZPAD + 4 * tb[d].total4;
    cx.cells += cx.slot[d];
    cx.pieces += tb[d].ok ? tb[d].total4 : 0;
  }
  // 2 gather buffers x cells x float2  +  PR_NS scratch stages x 2 planes x pieces x float4
  if (4 * cx.cells + n_stages * 8 * cx.pieces > smem_floats) cx.ok = 0;
  if (!cx.ok) return;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    const int doff = d == 0 ? 0 : cx.slot[0];
#pragma unroll
    for (int q = 0; q < PR_PPT; ++q) {
      stage_offsets(tb[d], has[q][d], x0[q][d], y0[q][d], cx.px[q][d].o0, cx.px[q][d].o1);
      cx.px[q][d].o0 += doff;
      cx.px[q][d].o1 += doff;
    }
    pair_assign(tb[d], G.H, G.W, cx.ld[d]);
  }
  // taps of the slow pixels, once per pixel (full version: kernel 2 also needs the multipliers and the raw flow)
  if (threadIdx.x < slow.n) {
    const int pix = slow.pix[threadIdx.x];
#pragma unroll
    for (int d = 0; d < NDIRS; ++d)
      compute_tap(G, P.dir[d], cx.n, cx.t, blockIdx.y * PR_TH + (pix >> 5), blockIdx.x * PR_TW + (pix & 31), slowtap[threadIdx.x][d]);
  }
  // zero pad of every direction slot of both buffers
  for (int k = threadIdx.x; k < 2 * NDIRS * ST_ZPAD; k += PR_THREADS) {
    const int z = k % ST_ZPAD, sl = k / ST_ZPAD;
    const int d = sl % NDIRS, b = sl / NDIRS;
    reinterpret_cast<float2*>(smem)[b * cx.cells + (d == 0 ? 0 : cx.slot[0]) + z] = make_float2(0.f, 0.f);
  }
}

// The producer side: walks the channel pairs of all (non-skipped) groups in order and loads this thread's
// pieces of planes c and c+1 into registers.  The image width is a multiple of 4 here (host-checked), so a piece
// is either completely inside the image or completely outside (zero).
template <int NDIRS>
struct PairStream {
  int g, c, C;               // next pair to load: channels c, c+1 of group g
  const float* pa[NDIRS];    // plane c of group g at (n, t)
  int sc[NDIRS];             // channel stride
  int goff[NDIRS][PR_SLOTS]; // element offset of this thread's pieces inside a plane

  __device__ __forceinline__ bool valid(const Params& P) const { return g < P.geo.n_groups; }
  __device__ __forceinline__ void enter(const Params& P, const PairCtx<NDIRS>& cx, unsigned skip_mask) {
    while (g < P.geo.n_groups && ((skip_mask >> g) & 1u)) ++g;
    if (g >= P.geo.n_groups) return;
    const GroupP& R = P.grp[g];
    c = 0;
    C = R.C;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      pa[d] = R.src[d] + cx.n * R.src_sn[d] + cx.t * R.src_st[d];
      sc[d] = R.src_sc[d];
#pragma unroll
      for (int s = 0; s < PR_SLOTS; ++s) goff[d][s] = (cx.ld[d].ycol[s] >> 16) * R.src_sh[d] + (cx.ld[d].ycol[s] & 0xffff);
    }
  }
  __device__ __forceinline__ void advance(const Params& P, const PairCtx<NDIRS>& cx, unsigned skip_mask) {
    c += 2;
    if (c < C) {
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) pa[d] += 2 * (long long)sc[d];
    } else {
      ++g;
      enter(P, cx, skip_mask);
    }
  }
  // cp.async this thread's pieces of planes c, c+1 into scratch stage `st` (shared-space byte address of the
  // stage); always commits a group, also when there is nothing left to copy
  __device__ __forceinline__ void issue(const Params& P, const PairCtx<NDIRS>& cx, unsigned st) {
    if (valid(P)) {
      const bool two = c + 1 < C;
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const float* pb = pa[d] + (two ? sc[d] : 0);  // no second channel: any finite data will do, it is not stored
#pragma unroll
        for (int s = 0; s < PR_SLOTS; ++s)
          if (cx.ld[d].bytes[s] >= 0) {
            const unsigned dst = st + 16u * (unsigned)((d == 0 ? 0 : cx.slot[0] / 4 - 1) + threadIdx.x + s * PR_THREADS);
            cp_async16(dst, pa[d] + goff[d][s], cx.ld[d].bytes[s]);
            cp_async16(dst + 16u * (unsigned)cx.pieces, pb + goff[d][s], cx.ld[d].bytes[s]);
          }
      }
    }
    cp_async_commit();
  }
};

// hop 2: this thread's pieces of one pair, scratch stage -> interleaved gather buffer
template <int NDIRS>
__device__ __forceinline__ void pair_transpose(const PairCtx<NDIRS>& cx, const float4* st, float2* buf) {
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    float4* dst = reinterpret_cast<float4*>(buf + (d == 0 ? 0 : cx.slot[0]) + ST_ZPAD);
    const float4* src = st + (d == 0 ? 0 : cx.slot[0] / 4 - 1);
#pragma unroll
    for (int s = 0; s < PR_SLOTS; ++s)
      if (cx.ld[d].bytes[s] >= 0) {
        const int k = threadIdx.x + s * PR_THREADS;
        const float4 a = src[k], b = src[k + cx.pieces];
        dst[2 * k] = make_float4(a.x, b.x, a.y, b.y);
        dst[2 * k + 1] = make_float4(a.z, b.z, a.w, b.w);
      }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel 1 (channel pairs): fused forward warp (+gate) (+blend), NDIRS directions, all channel groups.
// ---------------------------------------------------------------------------------------------
template <int NDIRS, bool ALIGN, bool BORDER>
__global__ void __launch_bounds__(PR_THREADS, 3) fwd_pair_kernel(const __grid_constant__ Params P, int smem_floats) {
  extern __shared__ float4 pr_smem4[];
  float* const smem = reinterpret_cast<float*>(pr_smem4);
  __shared__ StageTab tb[NDIRS];
  __shared__ StageSlow slow;
  __shared__ Tap slowtap[PR_MAXSLOW][NDIRS];
  PairCtx<NDIRS> cx;
  pair_prologue<NDIRS, ALIGN, BORDER>(P, tb, slow, slowtap, smem, smem_floats, PR_NS, cx);
  const int n = cx.n, t = cx.t, j = cx.j;

  if (!cx.ok) {  // wild flow: the tile's source footprint does not fit -> gather from global memory
#pragma unroll
    for (int q = 0; q < PR_PPT; ++q)
      if (cx.inimg[q]) fwd_generic_pixel<NDIRS>(P, n, t, cx.irow[q], j);
    return;
  }
  f32x2 w[PR_PPT][NDIRS][4], bl[PR_PPT][NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d)
#pragma unroll
    for (int q = 0; q < PR_PPT; ++q) {
      const PairPix& px = cx.px[q][d];
      const float w0 = __fmul_rn(px.ux, px.uy), w1 = __fmul_rn(px.tx, px.uy);
      const float w2 = __fmul_rn(px.ux, px.ty), w3 = __fmul_rn(px.tx, px.ty);
      w[q][d][0] = pk2(w0, w0);
      w[q][d][1] = pk2(w1, w1);
      w[q][d][2] = pk2(w2, w2);
      w[q][d][3] = pk2(w3, w3);
      bl[q][d] = pk2(px.bl, px.bl);  // 1.0f when the direction has no blend weight: exact
    }
  float2* const cells = reinterpret_cast<float2*>(smem);
  const f32x2 zero2 = pk2(0.f, 0.f), one2 = pk2(1.f, 1.f);

  // scratch ring behind the two gather buffers
  const float4* const scr = reinterpret_cast<const float4*>(cells + 2 * cx.cells);
  const unsigned scr_s = (unsigned)__cvta_generic_to_shared(scr);
  const unsigned stage_bytes = 32u * (unsigned)cx.pieces;
  PairStream<NDIRS> ps;
  ps.g = 0;
  ps.enter(P, cx, 0u);
  ps.issue(P, cx, scr_s);  // pair 0
  if (ps.valid(P)) ps.advance(P, cx, 0u);
  ps.issue(P, cx, scr_s + stage_bytes);  // pair 1
  // slow pixels, while the first copies are in flight: one (pixel, channel) item per thread
  __syncthreads();  // slowtap
  {
    int Ctot = 0;
    for (int g = 0; g < P.geo.n_groups; ++g) Ctot += P.grp[g].C;
    for (int it = threadIdx.x; it < slow.n * Ctot; it += PR_THREADS) {
      const int sidx = it / Ctot, pix = slow.pix[sidx];
      fwd_slow_item<NDIRS>(P, n, t, blockIdx.y * PR_TH + (pix >> 5), blockIdx.x * PR_TW + (pix & 31), slowtap[sidx], it - sidx * Ctot);
    }
  }
  cp_async_wait<1>();
  pair_transpose<NDIRS>(cx, scr, cells);
  __syncthreads();
  int b = 0, st_next = 1;  // st_next: scratch stage that holds the pair after the one being gathered
  for (int g = 0; g < P.geo.n_groups; ++g) {  // the consumer walks the same order, one pair behind the producer
    const GroupP& R = P.grp[g];
    float* op = R.out + n * R.out_sn + t * R.out_st + j;
    const int osc = R.out_sc;
    long long orow[PR_PPT];
#pragma unroll
    for (int q = 0; q < PR_PPT; ++q) orow[q] = (long long)cx.irow[q] * R.out_sh;
    for (int c = 0; c < R.C; c += 2) {
      {  // copies of the pair two ahead
        if (ps.valid(P)) ps.advance(P, cx, 0u);
        const int st2 = st_next == PR_NS - 1 ? 0 : st_next + 1;
        ps.issue(P, cx, scr_s + (unsigned)st2 * stage_bytes);
      }
      // ---- gather pair (g, c) from buffer b
      const float2* sp = cells + b * cx.cells;
      const bool two = c + 1 < R.C;
#pragma unroll
      for (int q = 0; q < PR_PPT; ++q) {
        f32x2 r = zero2;
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) {
          const float2* s0 = sp + cx.px[q][d].o0;
          const float2* s1 = sp + cx.px[q][d].o1;
          f32x2 a = ffma2(as2(s0[0]), w[q][d][0], zero2);
          a = ffma2(as2(s0[1]), w[q][d][1], a);
          a = ffma2(as2(s1[0]), w[q][d][2], a);
          a = ffma2(as2(s1[1]), w[q][d][3], a);
          a = ffma2(a, bl[q][d], zero2);
          r = (d == 0) ? a : ffma2(a, one2, r);
        }
        float r0, r1;
        upk2(r, r0, r1);
        float* o = op + orow[q];
        st_cs_if(o, r0, cx.act[q]);
        st_cs_if(o + osc, r1, cx.act[q] && two);
      }
      cp_async_wait<1>();  // this thread's copies of the next pair have landed
      pair_transpose<NDIRS>(cx, scr + (size_t)st_next * 2 * cx.pieces, cells + (b ^ 1) * cx.cells);
      st_next = st_next == PR_NS - 1 ? 0 : st_next + 1;
      __syncthreads();
      b ^= 1;
      op += 2 * (long long)osc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2 (channel pairs): gradient w.r.t. flow / gate / blend weight.  Same staging as kernel 1; grad_out is
// read once, coalesced, straight from global memory (every thread needs exactly its own pixels).
//   gix = sum_c gw_c * [ uy*(v_ne - v_nw) + ty*(v_se - v_sw) ]
//   giy = sum_c gw_c * [ ux*(v_sw - v_nw) + tx*(v_se - v_ne) ]       (OOB tap value = 0)
// ---------------------------------------------------------------------------------------------
template <int NDIRS, bool ALIGN, bool BORDER>
__global__ void __launch_bounds__(PR_THREADS, 2) bwd_flow_pair_kernel(const __grid_constant__ Params P,
                                                                     const __grid_constant__ GradP Q, int smem_floats) {
  extern __shared__ float4 pr_smem4[];
  float* const smem = reinterpret_cast<float*>(pr_smem4);
  __shared__ StageTab tb[NDIRS];
  __shared__ StageSlow slow;
  __shared__ Tap slowtap[PR_MAXSLOW][NDIRS];
  const Geo& G = P.geo;
  const int warp = threadIdx.x >> 5;
  PairCtx<NDIRS> cx;
  pair_prologue<NDIRS, ALIGN, BORDER>(P, tb, slow, slowtap, smem, smem_floats, PR_NS, cx);
  const int n = cx.n, t = cx.t, j = cx.j;

  if (!cx.ok) {
#pragma unroll
    for (int q = 0; q < PR_PPT; ++q)
      if (cx.inimg[q]) bwdflow_generic_pixel<NDIRS>(P, Q, n, t, cx.irow[q], j);
    return;
  }
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) has_bl[d] = P.dir[d].blend != nullptr;
  unsigned skip = 0u;  // groups without grad_out contribute nothing
#pragma unroll
  for (int g = 0; g < FWB_MAX_GROUPS; ++g) skip |= (Q.grad_out[g] == nullptr ? 1u : 0u) << g;

  f32x2 tx2[PR_PPT][NDIRS], ty2[PR_PPT][NDIRS], ux2[PR_PPT][NDIRS], uy2[PR_PPT][NDIRS];
  float gix[PR_PPT][NDIRS], giy[PR_PPT][NDIRS], gbl[PR_PPT][NDIRS];
#pragma unroll
  for (int q = 0; q < PR_PPT; ++q)
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const PairPix& px = cx.px[q][d];
      tx2[q][d] = pk2(px.tx, px.tx);
      ty2[q][d] = pk2(px.ty, px.ty);
      ux2[q][d] = pk2(px.ux, px.ux);
      uy2[q][d] = pk2(px.uy, px.uy);
      gix[q][d] = giy[q][d] = gbl[q][d] = 0.f;
    }
  float2* const cells = reinterpret_cast<float2*>(smem);
  const f32x2 zero2 = pk2(0.f, 0.f), mone2 = pk2(-1.f, -1.f);

  const float4* const scr = reinterpret_cast<const float4*>(cells + 2 * cx.cells);
  const unsigned scr_s = (unsigned)__cvta_generic_to_shared(scr);
  const unsigned stage_bytes = 32u * (unsigned)cx.pieces;
  PairStream<NDIRS> ps;
  ps.g = 0;
  ps.enter(P, cx, skip);
  ps.issue(P, cx, scr_s);  // pair 0
  if (ps.valid(P)) ps.advance(P, cx, skip);
  ps.issue(P, cx, scr_s + stage_bytes);  // pair 1
  // slow pixels, while the first copies are in flight: one warp per pixel, lanes over the channels, warp-reduced
  __syncthreads();  // slowtap
  for (int s = warp; s < slow.n; s += PR_THREADS / 32) {
    const int pix = slow.pix[s];
    bwdflow_slow_warp<NDIRS>(P, Q, n, t, blockIdx.y * PR_TH + (pix >> 5), blockIdx.x * PR_TW + (pix & 31), slowtap[s]);
  }
  cp_async_wait<1>();
  pair_transpose<NDIRS>(cx, scr, cells);
  __syncthreads();
  int b = 0, st_next = 1;
  for (int g = 0; g < P.geo.n_groups; ++g) {
    if ((skip >> g) & 1u) continue;
    const GroupP& R = P.grp[g];
    const float* gp = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + j;
    const int gsc = Q.go_sc[g];
    long long grow[PR_PPT];
#pragma unroll
    for (int q = 0; q < PR_PPT; ++q) grow[q] = (long long)cx.irow[q] * Q.go_sh[g];
    for (int c = 0; c < R.C; c += 2) {
      {  // copies of the pair two ahead
        if (ps.valid(P)) ps.advance(P, cx, skip);
        const int st2 = st_next == PR_NS - 1 ? 0 : st_next + 1;
        ps.issue(P, cx, scr_s + (unsigned)st2 * stage_bytes);
      }
      const bool two = c + 1 < R.C;
      float go0[PR_PPT], go1[PR_PPT];
#pragma unroll
      for (int q = 0; q < PR_PPT; ++q) {
        go0[q] = cx.act[q] ? __ldcs(gp + grow[q]) : 0.f;
        go1[q] = (cx.act[q] && two) ? __ldcs(gp + grow[q] + gsc) : 0.f;
      }
      const float2* sp = cells + b * cx.cells;
#pragma unroll
      for (int q = 0; q < PR_PPT; ++q) {
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) {
          const PairPix& px = cx.px[q][d];
          const float2* s0 = sp + px.o0;
          const float2* s1 = sp + px.o1;
          const f32x2 a = as2(s0[0]), bq = as2(s0[1]), c_ = as2(s1[0]), dd = as2(s1[1]);
          float gw0 = go0[q], gw1 = go1[q];
          if (has_bl[d]) {
            const f32x2 top = ffma2(bq, tx2[q][d], ffma2(a, ux2[q][d], zero2));
            const f32x2 bot = ffma2(dd, tx2[q][d], ffma2(c_, ux2[q][d], zero2));
            const f32x2 val = ffma2(bot, ty2[q][d], ffma2(top, uy2[q][d], zero2));
            float v0, v1;
            upk2(val, v0, v1);
            gbl[q][d] = fmaf(go1[q], v1, fmaf(go0[q], v0, gbl[q][d]));
            gw0 *= px.bl;
            gw1 *= px.bl;
          }
          const f32x2 dx_top = ffma2(a, mone2, bq), dx_bot = ffma2(c_, mone2, dd);  // b - a, dd - cc
          const f32x2 dy_rgt = ffma2(bq, mone2, dd), dy_lft = ffma2(a, mone2, c_);  // dd - b, cc - a
          const f32x2 ex = ffma2(ty2[q][d], dx_bot, ffma2(uy2[q][d], dx_top, zero2));
          const f32x2 ey = ffma2(tx2[q][d], dy_rgt, ffma2(ux2[q][d], dy_lft, zero2));
          float ex0, ex1, ey0, ey1;
          upk2(ex, ex0, ex1);
          upk2(ey, ey0, ey1);
          gix[q][d] = fmaf(gw1, ex1, fmaf(gw0, ex0, gix[q][d]));
          giy[q][d] = fmaf(gw1, ey1, fmaf(gw0, ey0, giy[q][d]));
        }
      }
      cp_async_wait<1>();
      pair_transpose<NDIRS>(cx, scr + (size_t)st_next * 2 * cx.pieces, cells + (b ^ 1) * cx.cells);
      st_next = st_next == PR_NS - 1 ? 0 : st_next + 1;
      __syncthreads();
      b ^= 1;
      gp += 2 * (long long)gsc;
    }
  }

#pragma unroll
  for (int q = 0; q < PR_PPT; ++q) {
    if (!cx.act[q]) continue;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      Tap k;
      compute_tap(G, P.dir[d], n, t, cx.irow[q], j, k);  // mx, my, fx, fy, gate (cheaper to recompute than to hold)
      bwdflow_store(P, Q, d, n, t, cx.irow[q], j, k, gix[q][d], giy[q][d], gbl[q][d]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernels 2 + 3 FUSED (the default, non-deterministic backward): gradient w.r.t. flow / gate / blend AND
// w.r.t. the sources in one pass over grad_out.
//
// The image-gradient scatter of a tile lands exactly on the cells the tile gathers from — its row-segment
// footprint — so it is accumulated in a shared-memory copy of that footprint and only the footprint (1.3-1.4
// cells per pixel, not 4 taps per pixel) goes to global memory, as 16-byte vector reductions
// (red.global.add.v4.f32) into the zero-initialised grad_src.  B200 has no native fp32 shared-memory atomic
// (atomicAdd(float*) on shared compiles to a CAS spin loop, 2-3x slower than ATOMS.ADD), so the tile accumulates
// in int32 FIXED POINT: per channel pair the CTA takes amax = max |grad_out * blend| over its pixels and scales
// by 2^(20 - exponent(amax)); a contribution w * g (|w| <= 1) is then below 2^21, so the 512 pixels of a tile
// cannot overflow int32 whatever the flow does, and the rounding error is 2^-21 of amax per contribution.
// A pair whose amax is not finite takes plain global float atomics, so inf / NaN gradients still propagate.
// The order of the global reductions varies from run to run: this is the NON-DETERMINISTIC mode; the
// deterministic mode is kernel 2 followed by the owner gather (fwb_owner.cuh / fwb_csr.cuh).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// global-atomic scatter of one pixel's 4 taps of one direction for channels c (and c+1)
__device__ __forceinline__ void scatter_atomic_px(const GradP& Q, int g, int d, int n, int t, int c, bool two, const Tap& k, float gw0,
                                                  float gw1) {
  float* gs = Q.grad_src[g][d];
  if (!gs) return;
  const int sh = Q.gs_sh[g][d];
  gs += n * Q.gs_sn[g][d] + t * Q.gs_st[g][d] + (long long)c * Q.gs_sc[g][d] + (long long)k.y0 * sh + k.x0;
  const float w[4] = {k.ux * k.uy, k.tx * k.uy, k.ux * k.ty, k.tx * k.ty};
  const int off[4] = {0, 1, sh, sh + 1};
  // the (x0, x0+1) pair of a row goes out as ONE 8-byte vector reduction when both taps are inside the image and the pair
  // is 8-byte aligned (half of the pixels when the strides are even): the L2 atomic units see 25 % fewer operations
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float* p0 = gs + off[2 * r];
    const unsigned vv = (k.valid >> (2 * r)) & 3u;
    if (vv == 3u && (reinterpret_cast<uintptr_t>(p0) & 7u) == 0u) {
      red_add_v2(p0, w[2 * r] * gw0, w[2 * r + 1] * gw0);
      if (two && (Q.gs_sc[g][d] & 1) == 0) {
        red_add_v2(p0 + Q.gs_sc[g][d], w[2 * r] * gw1, w[2 * r + 1] * gw1);
      } else if (two) {
        atomicAdd(p0 + Q.gs_sc[g][d], w[2 * r] * gw1);
        atomicAdd(p0 + Q.gs_sc[g][d] + 1, w[2 * r + 1] * gw1);
      }
    } else {
#pragma unroll
      for (int q = 2 * r; q < 2 * r + 2; ++q)
        if (k.valid & (1u << q)) {
          atomicAdd(gs + off[q], w[q] * gw0);
          if (two) atomicAdd(gs + off[q] + Q.gs_sc[g][d], w[q] * gw1);
        }
    }
  }
}

// one pixel, all channels of all groups: kernel 2's generic body + global-atomic scatter (tiles that do not fit)
template <int NDIRS>
__device__ __forceinline__ void bwd_fused_generic_pixel(const Params& P, const GradP& Q, int n, int t, int i, int j) {
  bwdflow_generic_pixel<NDIRS>(P, Q, n, t, i, j);
  Tap k[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) compute_tap(P.geo, P.dir[d], n, t, i, j, k[d]);
  for (int g = 0; g < P.geo.n_groups; ++g) {
    if (!Q.grad_out[g]) continue;
    const float* go = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)i * Q.go_sh[g] + j;
    for (int c = 0; c < P.grp[g].C; ++c) {
      const float gout = __ldg(go + (long long)c * Q.go_sc[g]);
#pragma unroll
      for (int d = 0; d < NDIRS; ++d)
        scatter_atomic_px(Q, g, d, n, t, c, false, k[d], P.dir[d].blend ? gout * k[d].blend : gout, 0.f);
    }
  }
}

template <int NDIRS, bool ALIGN, bool BORDER>
__global__ void __launch_bounds__(PR_THREADS, 2) bwd_fused_pair_kernel(const __grid_constant__ Params P,
                                                                      const __grid_constant__ GradP Q, int smem_floats) {
  extern __shared__ float4 pr_smem4[];
  float* const smem = reinterpret_cast<float*>(pr_smem4);
  __shared__ StageTab tb[NDIRS];
  __shared__ StageSlow slow;
  __shared__ Tap slowtap[PR_MAXSLOW][NDIRS];
  __shared__ unsigned amax_s[2];
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PairCtx<NDIRS> cx;
  // memory: 1 gather buffer + 1 accumulator (cells x 8 B each) + 2 scratch stages
  pair_prologue<NDIRS, ALIGN, BORDER>(P, tb, slow, slowtap, smem, smem_floats, 2, cx);
  const int n = cx.n, t = cx.t, j = cx.j;

  if (!cx.ok) {
#pragma unroll
    for (int q = 0; q < PR_PPT; ++q)
      if (cx.inimg[q]) bwd_fused_generic_pixel<NDIRS>(P, Q, n, t, cx.irow[q], j);
    return;
  }
  bool has_bl[NDIRS];
  float blmax[PR_PPT];
#pragma unroll
  for (int q = 0; q < PR_PPT; ++q) blmax[q] = 0.f;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    has_bl[d] = P.dir[d].blend != nullptr;
#pragma unroll
    for (int q = 0; q < PR_PPT; ++q) blmax[q] = fmaxf(blmax[q], has_bl[d] ? fabsf(cx.px[q][d].bl) : 1.0f);
  }
  unsigned skip = 0u;  // groups without grad_out contribute nothing
#pragma unroll
  for (int g = 0; g < FWB_MAX_GROUPS; ++g) skip |= (Q.grad_out[g] == nullptr ? 1u : 0u) << g;

  f32x2 tx2[PR_PPT][NDIRS], ty2[PR_PPT][NDIRS], ux2[PR_PPT][NDIRS], uy2[PR_PPT][NDIRS];
  float gix[PR_PPT][NDIRS], giy[PR_PPT][NDIRS], gbl[PR_PPT][NDIRS];
#pragma unroll
  for (int q = 0; q < PR_PPT; ++q)
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const PairPix& px = cx.px[q][d];
      tx2[q][d] = pk2(px.tx, px.tx);
      ty2[q][d] = pk2(px.ty, px.ty);
      ux2[q][d] = pk2(px.ux, px.ux);
      uy2[q][d] = pk2(px.uy, px.uy);
      gix[q][d] = giy[q][d] = gbl[q][d] = 0.f;
    }
  float2* const cells = reinterpret_cast<float2*>(smem);                 // the gather buffer
  int2* const acc = reinterpret_cast<int2*>(cells + cx.cells);           // the fixed-point accumulator
  const float4* const scr = reinterpret_cast<const float4*>(cells + 2 * cx.cells);
  const unsigned scr_s = (unsigned)__cvta_generic_to_shared(scr);
  const unsigned stage_bytes = 32u * (unsigned)cx.pieces;
  const f32x2 zero2 = pk2(0.f, 0.f), mone2 = pk2(-1.f, -1.f);
  for (int k = threadIdx.x; k < cx.cells; k += PR_THREADS) acc[k] = make_int2(0, 0);
  if (threadIdx.x < 2) amax_s[threadIdx.x] = 0u;

  PairStream<NDIRS> ps;
  ps.g = 0;
  ps.enter(P, cx, skip);
  ps.issue(P, cx, scr_s);  // pair 0
  __syncthreads();         // slowtap, acc, amax_s
  // slow pixels, while the first copies are in flight: one warp per pixel, lanes over the channels
  for (int s = warp; s < slow.n; s += PR_THREADS / 32) {
    const int pix = slow.pix[s];
    const int si = blockIdx.y * PR_TH + (pix >> 5), sj = blockIdx.x * PR_TW + (pix & 31);
    bwdflow_slow_warp<NDIRS>(P, Q, n, t, si, sj, slowtap[s]);
    int Ctot = 0;
    for (int g = 0; g < G.n_groups; ++g) Ctot += P.grp[g].C;
    for (int cf = lane; cf < Ctot; cf += 32) {
      int g, c;
      chan_lookup(P, cf, g, c);
      if (!Q.grad_out[g]) continue;
      const float gout = __ldg(Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)c * Q.go_sc[g] + (long long)si * Q.go_sh[g] + sj);
#pragma unroll
      for (int d = 0; d < NDIRS; ++d)
        scatter_atomic_px(Q, g, d, n, t, c, false, slowtap[s][d], has_bl[d] ? gout * slowtap[s][d].blend : gout, 0.f);
    }
  }
  // grad_out of the first pair and its amax
  float go0[PR_PPT], go1[PR_PPT];
#pragma unroll
  for (int q = 0; q < PR_PPT; ++q) go0[q] = go1[q] = 0.f;
  {
    int g0 = 0;
    while (g0 < G.n_groups && ((skip >> g0) & 1u)) ++g0;
    if (g0 < G.n_groups) {
      const float* gp = Q.grad_out[g0] + n * Q.go_sn[g0] + t * Q.go_st[g0] + j;
      float m = 0.f;
#pragma unroll
      for (int q = 0; q < PR_PPT; ++q) {
        const float* gq = gp + (long long)cx.irow[q] * Q.go_sh[g0];
        go0[q] = cx.act[q] ? __ldcs(gq) : 0.f;
        go1[q] = (cx.act[q] && P.grp[g0].C > 1) ? __ldcs(gq + Q.go_sc[g0]) : 0.f;
        const float a0 = fabsf(go0[q]), a1 = fabsf(go1[q]);
        // NaN must win the max: compare on the bit patterns (non-negative floats order like unsigned ints)
        m = __uint_as_float(max(__float_as_uint(m), max(__float_as_uint(a0 * blmax[q]), __float_as_uint(a1 * blmax[q]))));
      }
      unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
      if (lane == 0) atomicMax(&amax_s[0], mb);
    }
  }
  cp_async_wait<0>();
  pair_transpose<NDIRS>(cx, scr, cells);
  __syncthreads();
  int st_next = 1, pi = 0;  // scratch stage of the next pair; parity of the pair being processed
  for (int g = 0; g < G.n_groups; ++g) {
    if ((skip >> g) & 1u) continue;
    const GroupP& R = P.grp[g];
    const float* gp = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + j;
    const int gsc = Q.go_sc[g];
    bool want[NDIRS];
    bool any_want = false;
    float* gsp[NDIRS];             // grad_src plane of the current pair's first channel
    long long gsc2[NDIRS];         // 2 channel strides
    int gsoff[NDIRS][PR_SLOTS];    // element offset of this thread's pieces inside a grad_src plane
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      want[d] = Q.grad_src[g][d] != nullptr;
      any_want |= want[d];
      gsp[d] = want[d] ? Q.grad_src[g][d] + n * Q.gs_sn[g][d] + t * Q.gs_st[g][d] : nullptr;
      gsc2[d] = 2ll * Q.gs_sc[g][d];
#pragma unroll
      for (int sl = 0; sl < PR_SLOTS; ++sl) gsoff[d][sl] = (cx.ld[d].ycol[sl] >> 16) * Q.gs_sh[g][d] + (cx.ld[d].ycol[sl] & 0xffff);
    }
    for (int c = 0; c < R.C; c += 2) {
      // ---- copies of the next pair; its grad_out and amax
      if (ps.valid(P)) ps.advance(P, cx, skip);
      ps.issue(P, cx, scr_s + (unsigned)st_next * stage_bytes);
      float gn0[PR_PPT], gn1[PR_PPT];
#pragma unroll
      for (int q = 0; q < PR_PPT; ++q) gn0[q] = gn1[q] = 0.f;
      if (ps.valid(P)) {
        const int g2 = ps.g;
        const float* gp2 = Q.grad_out[g2] + n * Q.go_sn[g2] + t * Q.go_st[g2] + (long long)ps.c * Q.go_sc[g2] + j;
#pragma unroll
        for (int q = 0; q < PR_PPT; ++q) {
          const float* gq = gp2 + (long long)cx.irow[q] * Q.go_sh[g2];
          gn0[q] = cx.act[q] ? __ldcs(gq) : 0.f;
          gn1[q] = (cx.act[q] && ps.c + 1 < ps.C) ? __ldcs(gq + Q.go_sc[g2]) : 0.f;
        }
      }
      // ---- scale of this pair
      const unsigned ab = amax_s[pi];
      const bool finite = ab < 0x7f800000u;
      const int sexp = min(252, max(2, 275 - 21 + 20 - (int)(ab >> 23)));  // biased exponent of 2^(20 - exponent(amax))
      const float S = __uint_as_float((unsigned)sexp << 23), Sinv = __uint_as_float((unsigned)(254 - sexp) << 23);
      const bool two = c + 1 < R.C;
      // ---- gather (kernel 2) + scatter into the accumulator
#pragma unroll
      for (int q = 0; q < PR_PPT; ++q) {
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) {
          const PairPix& px = cx.px[q][d];
          const float2* s0 = cells + px.o0;
          const float2* s1 = cells + px.o1;
          const f32x2 a = as2(s0[0]), bq = as2(s0[1]), c_ = as2(s1[0]), dd = as2(s1[1]);
          float gw0 = go0[q], gw1 = go1[q];
          if (has_bl[d]) {
            const f32x2 top = ffma2(bq, tx2[q][d], ffma2(a, ux2[q][d], zero2));
            const f32x2 bot = ffma2(dd, tx2[q][d], ffma2(c_, ux2[q][d], zero2));
            const f32x2 val = ffma2(bot, ty2[q][d], ffma2(top, uy2[q][d], zero2));
            float v0, v1;
            upk2(val, v0, v1);
            gbl[q][d] = fmaf(go1[q], v1, fmaf(go0[q], v0, gbl[q][d]));
            gw0 *= px.bl;
            gw1 *= px.bl;
          }
          const f32x2 dx_top = ffma2(a, mone2, bq), dx_bot = ffma2(c_, mone2, dd);  // b - a, dd - cc
          const f32x2 dy_rgt = ffma2(bq, mone2, dd), dy_lft = ffma2(a, mone2, c_);  // dd - b, cc - a
          const f32x2 ex = ffma2(ty2[q][d], dx_bot, ffma2(uy2[q][d], dx_top, zero2));
          const f32x2 ey = ffma2(tx2[q][d], dy_rgt, ffma2(ux2[q][d], dy_lft, zero2));
          float ex0, ex1, ey0, ey1;
          upk2(ex, ex0, ex1);
          upk2(ey, ey0, ey1);
          gix[q][d] = fmaf(gw1, ex1, fmaf(gw0, ex0, gix[q][d]));
          giy[q][d] = fmaf(gw1, ey1, fmaf(gw0, ey0, giy[q][d]));
          if (want[d] && ab != 0u) {
            if (finite) {
              const f32x2 sg = pk2(gw0 * S, gw1 * S);
              const float w0 = px.ux * px.uy, w1 = px.tx * px.uy, w2 = px.ux * px.ty, w3 = px.tx * px.ty;
              int* a0 = reinterpret_cast<int*>(acc + px.o0);
              int* a1 = reinterpret_cast<int*>(acc + px.o1);
              float v0, v1;
              upk2(ffma2(sg, pk2(w0, w0), zero2), v0, v1);
              atomicAdd(a0, __float2int_rn(v0));
              atomicAdd(a0 + 1, __float2int_rn(v1));
              upk2(ffma2(sg, pk2(w1, w1), zero2), v0, v1);
              atomicAdd(a0 + 2, __float2int_rn(v0));
              atomicAdd(a0 + 3, __float2int_rn(v1));
              upk2(ffma2(sg, pk2(w2, w2), zero2), v0, v1);
              atomicAdd(a1, __float2int_rn(v0));
              atomicAdd(a1 + 1, __float2int_rn(v1));
              upk2(ffma2(sg, pk2(w3, w3), zero2), v0, v1);
              atomicAdd(a1 + 2, __float2int_rn(v0));
              atomicAdd(a1 + 3, __float2int_rn(v1));
            } else if (cx.act[q]) {  // inf / NaN in grad_out: exact float atomics straight to global memory
              Tap k;
              compute_tap(G, P.dir[d], n, t, cx.irow[q], j, k);
              scatter_atomic_px(Q, g, d, n, t, c, two, k, gw0, gw1);
            }
          }
        }
      }
      {  // amax of the next pair (its grad_out has had the whole gather to arrive)
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < PR_PPT; ++q) {
          const float a0 = fabsf(gn0[q]) * blmax[q], a1 = fabsf(gn1[q]) * blmax[q];
          // NaN must win the max: compare on the bit patterns (non-negative floats order like unsigned ints)
          m = __uint_as_float(max(__float_as_uint(m), max(__float_as_uint(a0), __float_as_uint(a1))));
        }
        const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
        if (lane == 0 && mb != 0u) atomicMax(&amax_s[pi ^ 1], mb);
      }
      cp_async_wait<0>();  // this thread's copies of the next pair have landed
      __syncthreads();     // (A) everyone is done with the gather buffer and with the scatter
      pair_transpose<NDIRS>(cx, scr + (size_t)st_next * 2 * cx.pieces, cells);
      st_next ^= 1;
      // ---- flush the accumulator: this thread's in-image pieces -> red.v4 into grad_src, then zero them
      if (any_want && ab != 0u && finite) {
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) {
          if (!want[d]) continue;
          int4* ap = reinterpret_cast<int4*>(acc + (d == 0 ? 0 : cx.slot[0]) + ST_ZPAD);
#pragma unroll
          for (int sl = 0; sl < PR_SLOTS; ++sl)
            if (cx.ld[d].bytes[sl] == 16) {
              const int k = threadIdx.x + sl * PR_THREADS;
              const int4 u = ap[2 * k], v = ap[2 * k + 1];  // {c0,d0,c1,d1} {c2,d2,c3,d3}
              const int nz0 = u.x | u.z | v.x | v.z, nz1 = u.y | u.w | v.y | v.w;
              if ((nz0 | nz1) != 0) {
                float* dst = gsp[d] + gsoff[d][sl];
                if (nz0 != 0)
                  red_add_v4(dst, make_float4((float)u.x * Sinv, (float)u.z * Sinv, (float)v.x * Sinv, (float)v.z * Sinv));
                if (two && nz1 != 0)
                  red_add_v4(dst + (gsc2[d] >> 1), make_float4((float)u.y * Sinv, (float)u.w * Sinv, (float)v.y * Sinv, (float)v.w * Sinv));
                ap[2 * k] = make_int4(0, 0, 0, 0);
                ap[2 * k + 1] = make_int4(0, 0, 0, 0);
              }
            }
        }
      }
#pragma unroll
      for (int d = 0; d < NDIRS; ++d)
        if (want[d]) gsp[d] += gsc2[d];
      if (threadIdx.x == 0) amax_s[pi] = 0u;
      __syncthreads();  // (B) gather buffer holds the next pair; accumulator is clean
      pi ^= 1;
#pragma unroll
      for (int q = 0; q < PR_PPT; ++q) {
        go0[q] = gn0[q];
        go1[q] = gn1[q];
      }
    }
  }

#pragma unroll
  for (int q = 0; q < PR_PPT; ++q) {
    if (!cx.act[q]) continue;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      Tap k;
      compute_tap(G, P.dir[d], n, t, cx.irow[q], j, k);  // mx, my, fx, fy, gate (cheaper to recompute than to hold)
      bwdflow_store(P, Q, d, n, t, cx.irow[q], j, k, gix[q][d], giy[q][d], gbl[q][d]);
    }
  }
}

}  // namespace fwb

// fwb_stage.cuh — shared-memory staging of source tiles, and kernels 1 and 2 built on it.
//
// The per-pixel gather of the warp (4 taps x C channels x 2 directions = 184 loads per pixel in the headline
// config) is bound by L1 wavefronts when it goes to global memory: with a rough flow the 32 taps of one warp
// instruction land on 7-11 different 128 B lines (ncu: l1tex 79 % busy, 0.30 ms for 0.10 ms of HBM traffic).
// Here a CTA owns a 32x32 tile of output pixels and
//   1. computes every pixel's taps once (channel independent),
//   2. builds, per direction, the ROW-SEGMENT TABLE of the source cells those taps touch: for every source row
//      the [xlo, xhi] column range, 16 B aligned — the deformed image of the tile, not its bounding box (a
//      bounding box over-reads 3.5x with the headline flows, the row segments 1.5-1.7x),
//   3. streams those segments for CC channel planes at a time into a compact shared-memory slot with 16-byte
//      cp.async (zero-filled outside the image, so the zeros padding needs no predicate in the inner loop),
//      double buffered: chunk k+1 is in flight while chunk k is gathered,
//   4. gathers the 4 taps from shared memory (1-2 wavefronts per warp instruction instead of 7-11).
// Which cp.async a thread issues is also channel independent: it is computed once and kept in registers.
// A tile whose segments do not fit (wild flows) takes the generic global-memory path of the same kernel.
#pragma once
#include "fwb_coords.cuh"
#include "fwb_generic.cuh"

namespace fwb {

constexpr int ST_TW = 32, ST_TH = 32;  // output tile of one CTA
constexpr int ST_THREADS = 256;        // 8 warps; warp w owns rows w, w+8, w+16, w+24 of the tile, lane = column
constexpr int ST_PPT = 4;              // pixels per thread
constexpr int ST_ROWS = 128;           // source rows a direction may span (row table is indexed y & 127)
constexpr int ST_SLOTS = 3;            // cp.async per thread per plane and direction -> <= 768 float4 per plane
constexpr int ST_ZPAD = 4;             // zero cells in front of every plane slot (target of tap-less pixels)
constexpr int ST_CCMAX = 4;            // channel planes per chunk (upper bound)
constexpr int ST_R = 24;               // a pixel whose displacement is farther than this from the tile's anchor is SLOW
constexpr int ST_MAXSLOW = 128;        // slow pixels a tile may have before the whole tile goes generic

struct StageTab {  // per direction, in shared memory
  int xlo[ST_ROWS], xhi[ST_ROWS];  // column range per source row (indexed y & 63)
  int rowoff4[ST_ROWS + 1];        // exclusive prefix of the segment lengths, float4 units, row r = y - ymin
  int rowx[ST_ROWS];               // 4-aligned first column of row r
  int ymin, ymax;
  int total4;  // float4 per plane
  int ok;      // stageable
  unsigned long long akey;  // anchor vote: (|dx|+|dy|) << 40 | dx << 20 | dy of the least displaced pixel
};

struct StageSlow {  // pixels of the tile that are served from global memory (border-clamped / wild ones)
  int n;
  unsigned short pix[ST_MAXSLOW];  // (row in tile << 5) | column in tile
};

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const float* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
// predicated streaming store (no branch around it)
__device__ __forceinline__ void st_cs_if(float* p, float v, bool pred) {
  asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp st.global.cs.f32 [%0], %1;\n}\n" ::"l"(p), "f"(v), "r"((int)pred)
               : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- step 2: row-segment table of one direction.  Called by all threads; `has[q]` = pixel q of this thread has
// at least one in-image tap (then x0 in [-1, W-1], y0 in [-1, H-1]).
__device__ __forceinline__ void stage_tab_init(StageTab* tb, int ndirs) {
  for (int k = threadIdx.x; k < ndirs * ST_ROWS; k += blockDim.x) {
    tb[k / ST_ROWS].xlo[k % ST_ROWS] = 0x7fffffff;
    tb[k / ST_ROWS].xhi[k % ST_ROWS] = -0x7fffffff;
  }
  if (threadIdx.x < ndirs) {
    tb[threadIdx.x].ymin = 0x7fffffff;
    tb[threadIdx.x].ymax = -0x7fffffff;
    tb[threadIdx.x].akey = ~0ull;
  }
}

// The ANCHOR of a direction is the displacement (x0 - j, y0 - i) of the tile's least displaced pixel: robust
// against the few pixels that border clamping or a wild flow throws far away (they have large displacements).
__device__ __forceinline__ void stage_anchor_vote(StageTab& tb, bool has, int dx, int dy) {
  const unsigned mag = (unsigned)min(abs(dx) + abs(dy), 0xfffff);
  const unsigned key = has ? ((mag << 5) | (threadIdx.x & 31u)) : 0xffffffffu;
  const unsigned best = __reduce_min_sync(0xffffffffu, key);
  if (has && key == best)
    atomicMin(&tb.akey, ((unsigned long long)mag << 40) | ((unsigned long long)((unsigned)dx & 0xfffffu) << 20) |
                            (unsigned long long)((unsigned)dy & 0xfffffu));
}
__device__ __forceinline__ bool stage_inlier(const StageTab& tb, int dx, int dy) {
  const unsigned long long k = tb.akey;
  const int adx = ((int)((unsigned)(k >> 20) << 12)) >> 12, ady = ((int)((unsigned)k << 12)) >> 12;  // sign-extend 20 bits
  return abs(dx - adx) <= ST_R && abs(dy - ady) <= ST_R;
}

__device__ __forceinline__ void stage_tab_add(StageTab& tb, bool has, int x0, int y0) {
  int lo = has ? y0 : 0x7fffffff, hi = has ? y0 + 1 : -0x7fffffff;  // rows y0 and y0 + 1
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(&tb.ymin, lo);
    atomicMax(&tb.ymax, hi);
  }
  if (has) {
    atomicMin(&tb.xlo[y0 & (ST_ROWS - 1)], x0);
    atomicMax(&tb.xhi[y0 & (ST_ROWS - 1)], x0 + 1);
    atomicMin(&tb.xlo[(y0 + 1) & (ST_ROWS - 1)], x0);
    atomicMax(&tb.xhi[(y0 + 1) & (ST_ROWS - 1)], x0 + 1);
  }
}

// one warp per table: prefix sum of the segment lengths.  After this rowoff4/rowx are indexed by r = y - ymin.
// (xlo/xhi = first/last column needed in a row, ymin/ymax = first/last row.)
__device__ __forceinline__ void stage_tab_scan(StageTab& tb) {
  const int lane = threadIdx.x & 31;
  const int ymin = tb.ymin, ymax = tb.ymax;
  if (ymin > ymax) {  // no pixel of the tile has a tap: nothing to stage, every pixel reads the zero pad
    for (int r = lane; r < ST_ROWS; r += 32) tb.rowoff4[r] = 0;
    if (lane == 0) {
      tb.rowoff4[ST_ROWS] = 0;
      tb.total4 = 0;
      tb.ok = 1;
      tb.ymin = 0;
      tb.ymax = 0;
    }
    return;
  }
  const int span = ymax + 1 - ymin;  // rows ymin .. ymax
  if (span > ST_ROWS) {
    if (lane == 0) tb.ok = 0;
    return;
  }
  int run = 0;
  constexpr int NH = ST_ROWS / 32;
  int xs[NH], ln[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const int r = h * 32 + lane;
    xs[h] = 0;
    ln[h] = 0;
    if (r < span) {
      const int y = ymin + r;
      const int lo = tb.xlo[y & (ST_ROWS - 1)], hi = tb.xhi[y & (ST_ROWS - 1)];
      if (lo <= hi) {
        xs[h] = (lo >> 2) << 2;  // arithmetic shift: -1 -> -4
        ln[h] = (((hi + 4) >> 2) << 2) - xs[h];
        ln[h] >>= 2;
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    int v = ln[h];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    const int r = h * 32 + lane;
    tb.rowoff4[r] = run + v - ln[h];
    tb.rowx[r] = xs[h];
    run += __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) {
    tb.rowoff4[ST_ROWS] = run;
    tb.total4 = run;
    tb.ok = run <= ST_SLOTS * ST_THREADS;
  }
}

// offset (floats, from the start of the direction's plane slot) of the nw tap and of the sw tap
__device__ __forceinline__ void stage_offsets(const StageTab& tb, bool has, int x0, int y0, int& o0, int& o1) {
  o0 = o1 = 0;
  if (has) {
    const int r = y0 - tb.ymin;
    o0 = ST_ZPAD + tb.rowoff4[r] * 4 + (x0 - tb.rowx[r]);
    o1 = ST_ZPAD + tb.rowoff4[r + 1] * 4 + (x0 - tb.rowx[r + 1]);
  }
}

struct StageLoad {  // the cp.async this thread issues for every plane of a direction
  int ycol[ST_SLOTS];   // (y << 16) | col of the first float of slot s (0 when the cell is outside the image)
  int bytes[ST_SLOTS];  // bytes that come from memory: 0 = zero fill; -1 = this thread has no slot s
};

__device__ __forceinline__ void stage_assign(const StageTab& tb, int H, int W, StageLoad& ld) {
  const int span = tb.ymax + 1 - tb.ymin;
#pragma unroll
  for (int s = 0; s < ST_SLOTS; ++s) {
    const int k = threadIdx.x + s * ST_THREADS;
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

struct StagePlan {  // CTA-uniform
  int ok;           // both directions stageable and at least one plane per chunk fits
  int cc;           // channel planes per chunk
  int slot[2];      // floats per plane slot of each direction (ZPAD included, multiple of 4)
  int plane;        // floats per staged plane (= sum of the slots)
};

__device__ __forceinline__ StagePlan stage_plan(const StageTab* tb, int ndirs, int smem_floats) {
  StagePlan pl;
  pl.ok = 1;
  pl.plane = 0;
  pl.slot[0] = pl.slot[1] = 0;
  for (int d = 0; d < ndirs; ++d) {
    pl.ok &= tb[d].ok;
    pl.slot[d] = ST_ZPAD + 4 * tb[d].total4;
    pl.plane += pl.slot[d];
  }
  pl.cc = pl.ok ? min(ST_CCMAX, smem_floats / (2 * pl.plane)) : 0;
  if (pl.cc < 1) pl.ok = 0;
  return pl;
}

// ---------------------------------------------------------------------------------------------
// Common prologue of the staged kernels: taps, anchor vote, slow-pixel list, row-segment tables, plan,
// per-pixel slot offsets and per-thread load assignment.
// ---------------------------------------------------------------------------------------------
struct StagePix {  // per (pixel of this thread, direction)
  float tx, ty, ux, uy, bl;
  int o0, o1;  // float offsets, inside a staged plane, of the nw and sw taps (direction slot offset included)
};

template <int NDIRS>
struct StageCtx {
  int n, t, j;
  int irow[ST_PPT];
  bool inimg[ST_PPT];  // pixel is inside the image
  bool act[ST_PPT];    // ... and served by the staged loop (not slow)
  StagePix px[ST_PPT][NDIRS];
  StageLoad ld[NDIRS];
  StagePlan pl;
};

template <int NDIRS>
__device__ __forceinline__ void stage_prologue(const Params& P, StageTab* tb, StageSlow& slow, float* smem,
                                               int smem_floats, StageCtx<NDIRS>& cx) {
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cx.j = blockIdx.x * ST_TW + lane;
  cx.n = blockIdx.z / G.T;
  cx.t = blockIdx.z - cx.n * G.T;
#pragma unroll
  for (int q = 0; q < ST_PPT; ++q) {
    cx.irow[q] = blockIdx.y * ST_TH + warp + 8 * q;
    cx.inimg[q] = cx.j < G.W && cx.irow[q] < G.H;
  }
  int x0[ST_PPT][NDIRS], y0[ST_PPT][NDIRS];
  bool has[ST_PPT][NDIRS];
  stage_tab_init(tb, NDIRS);
  if (threadIdx.x == 0) slow.n = 0;
  __syncthreads();
#pragma unroll
  for (int d = 0; d < NDIRS; ++d)
#pragma unroll
    for (int q = 0; q < ST_PPT; ++q) {
      Tap k;
      k.valid = 0u;
      k.x0 = k.y0 = 0;
      k.ux = k.uy = k.tx = k.ty = 0.f;
      k.blend = 0.f;
      if (cx.inimg[q]) compute_tap(G, P.dir[d], cx.n, cx.t, cx.irow[q], cx.j, k);
      has[q][d] = k.valid != 0u;
      StagePix& px = cx.px[q][d];
      px.tx = k.tx;
      px.ty = k.ty;
      px.ux = k.ux;
      px.uy = k.uy;
      px.bl = k.blend;
      x0[q][d] = k.x0;
      y0[q][d] = k.y0;
      stage_anchor_vote(tb[d], has[q][d], k.x0 - cx.j, k.y0 - cx.irow[q]);
    }
  __syncthreads();
  // slow pixels: some direction's taps are far from where the rest of the tile samples
#pragma unroll
  for (int q = 0; q < ST_PPT; ++q) {
    bool far = false;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d)
      far |= has[q][d] && !stage_inlier(tb[d], x0[q][d] - cx.j, y0[q][d] - cx.irow[q]);
    cx.act[q] = cx.inimg[q] && !far;
    if (far) {
      const int slot = atomicAdd(&slow.n, 1);
      if (slot < ST_MAXSLOW) slow.pix[slot] = (unsigned short)(((warp + 8 * q) << 5) | lane);
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
  cx.pl = stage_plan(tb, NDIRS, smem_floats);
  if (slow.n > ST_MAXSLOW) cx.pl.ok = 0;
  if (!cx.pl.ok) return;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    const int doff = d == 0 ? 0 : cx.pl.slot[0];
#pragma unroll
    for (int q = 0; q < ST_PPT; ++q) {
      stage_offsets(tb[d], has[q][d], x0[q][d], y0[q][d], cx.px[q][d].o0, cx.px[q][d].o1);
      cx.px[q][d].o0 += doff;
      cx.px[q][d].o1 += doff;
    }
    stage_assign(tb[d], G.H, G.W, cx.ld[d]);
  }
  const int nz = 2 * cx.pl.cc * NDIRS * ST_ZPAD;  // zero pad of every plane slot of both stages
  for (int k = threadIdx.x; k < nz; k += ST_THREADS) {
    const int z = k % ST_ZPAD, sl = k / ST_ZPAD;
    const int d = sl % NDIRS, p = sl / NDIRS;  // p = stage * cc + plane
    smem[p * cx.pl.plane + (d == 0 ? 0 : cx.pl.slot[0]) + z] = 0.f;
  }
}

// The producer side of the pipeline: walks the channel planes of all groups in order, CC at a time (a chunk
// never straddles two groups), and issues this thread's cp.async for them.
template <int NDIRS>
struct StageProducer {
  int g, c;                 // next plane to issue
  const float* ptr[NDIRS];  // its plane pointer per direction
  int goff[NDIRS][ST_SLOTS];  // element offset of this thread's pieces inside a plane (depends on the group's row stride)
  unsigned sdst[NDIRS][ST_SLOTS];  // shared-space byte address of the pieces inside plane 0 of stage 0

  __device__ __forceinline__ void enter_group(const Params& P, const StageCtx<NDIRS>& cx) {
    const GroupP& R = P.grp[g];
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      ptr[d] = R.src[d] + cx.n * R.src_sn[d] + cx.t * R.src_st[d];
#pragma unroll
      for (int s = 0; s < ST_SLOTS; ++s) goff[d][s] = (cx.ld[d].ycol[s] >> 16) * R.src_sh[d] + (cx.ld[d].ycol[s] & 0xffff);
    }
  }
  __device__ __forceinline__ void init(const Params& P, const StageCtx<NDIRS>& cx, unsigned smem_base) {
    g = 0;
    c = 0;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d)
#pragma unroll
      for (int s = 0; s < ST_SLOTS; ++s)
        sdst[d][s] = smem_base + 4u * (unsigned)((d == 0 ? 0 : cx.pl.slot[0]) + ST_ZPAD + 4 * (threadIdx.x + s * ST_THREADS));
    enter_group(P, cx);
  }
  // issue the next chunk into stage `stage`; returns the number of planes (0 = no more planes)
  __device__ __forceinline__ int issue(const Params& P, const StageCtx<NDIRS>& cx, int stage, unsigned skip_mask) {
    while (g < P.geo.n_groups && (c >= P.grp[g].C || ((skip_mask >> g) & 1u))) {
      ++g;
      c = 0;
      if (g < P.geo.n_groups) enter_group(P, cx);
    }
    if (g >= P.geo.n_groups) return 0;
    const GroupP& R = P.grp[g];
    const int nch = min(cx.pl.cc, R.C - c);
    unsigned pb = 4u * (unsigned)(stage * cx.pl.cc * cx.pl.plane);
    for (int u = 0; u < nch; ++u) {
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const float* pp = ptr[d] + (long long)(c + u) * R.src_sc[d];
#pragma unroll
        for (int s = 0; s < ST_SLOTS; ++s)
          if (cx.ld[d].bytes[s] >= 0) cp_async16(sdst[d][s] + pb, pp + goff[d][s], cx.ld[d].bytes[s]);
      }
      pb += 4u * (unsigned)cx.pl.plane;
    }
    c += nch;
    cp_async_commit();
    return nch;
  }
};

// ---------------------------------------------------------------------------------------------
// Kernel 1 (staged): fused forward warp (+gate) (+blend), NDIRS directions, all channel groups.
// ---------------------------------------------------------------------------------------------
template <int NDIRS>
__global__ void __launch_bounds__(ST_THREADS, 2) fwd_staged_kernel(const __grid_constant__ Params P, int smem_floats) {
  extern __shared__ float4 st_smem4[];
  float* const smem = reinterpret_cast<float*>(st_smem4);
  __shared__ StageTab tb[NDIRS];
  __shared__ StageSlow slow;
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  StageCtx<NDIRS> cx;
  stage_prologue<NDIRS>(P, tb, slow, smem, smem_floats, cx);
  const int n = cx.n, t = cx.t, j = cx.j;

  if (!cx.pl.ok) {  // wild flow: the tile's source footprint does not fit -> gather from global memory
#pragma unroll
    for (int q = 0; q < ST_PPT; ++q)
      if (cx.inimg[q]) fwd_generic_pixel<NDIRS>(P, n, t, cx.irow[q], j);
    return;
  }
  const int cc = cx.pl.cc, plane = cx.pl.plane;
  float w[ST_PPT][NDIRS][4];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
#pragma unroll
    for (int q = 0; q < ST_PPT; ++q) {
      const StagePix& px = cx.px[q][d];
      w[q][d][0] = __fmul_rn(px.ux, px.uy);
      w[q][d][1] = __fmul_rn(px.tx, px.uy);
      w[q][d][2] = __fmul_rn(px.ux, px.ty);
      w[q][d][3] = __fmul_rn(px.tx, px.ty);
    }
  }
  StageProducer<NDIRS> prod;
  prod.init(P, cx, (unsigned)__cvta_generic_to_shared(smem));

  int stage = 0;
  int nch = prod.issue(P, cx, 0, 0u);
  for (int g = 0; g < G.n_groups; ++g) {  // consumer walks the same plane order
    const GroupP& R = P.grp[g];
    float* op = R.out + n * R.out_sn + t * R.out_st + j;
    int orow[ST_PPT];
#pragma unroll
    for (int q = 0; q < ST_PPT; ++q) orow[q] = cx.irow[q] * R.out_sh;
    for (int c0 = 0; c0 < R.C; c0 += cc) {
      const int nxt = prod.issue(P, cx, stage ^ 1, 0u);
      if (nxt)
        cp_async_wait<1>();
      else
        cp_async_wait<0>();
      __syncthreads();
      const float* sp = smem + stage * cc * plane;
      for (int u = 0; u < nch; ++u) {
#pragma unroll
        for (int q = 0; q < ST_PPT; ++q) {
          float r = 0.f;
#pragma unroll
          for (int d = 0; d < NDIRS; ++d) {
            const float* s0 = sp + cx.px[q][d].o0;
            const float* s1 = sp + cx.px[q][d].o1;
            float a = __fmul_rn(s0[0], w[q][d][0]);
            a = __fmaf_rn(s0[1], w[q][d][1], a);
            a = __fmaf_rn(s1[0], w[q][d][2], a);
            a = __fmaf_rn(s1[1], w[q][d][3], a);
            a = __fmul_rn(a, cx.px[q][d].bl);  // 1.0f when the direction has no blend weight: exact
            r = (d == 0) ? a : __fadd_rn(r, a);
          }
          st_cs_if(op + orow[q], r, cx.act[q]);
        }
        sp += plane;
        op += R.out_sc;
      }
      __syncthreads();
      stage ^= 1;
      nch = nxt;
    }
  }
  // slow pixels: one warp per pixel, lanes stride over the channels
  for (int s = warp; s < slow.n; s += ST_THREADS / 32) {
    const int pix = slow.pix[s];
    fwd_generic_pixel_strided<NDIRS>(P, n, t, blockIdx.y * ST_TH + (pix >> 5), blockIdx.x * ST_TW + (pix & 31), lane, 32);
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2 (staged): gradient w.r.t. flow / gate / blend weight.  Same staging as kernel 1; grad_out is read
// once, coalesced, straight from global memory (every thread needs exactly its own pixels).
// ---------------------------------------------------------------------------------------------
template <int NDIRS>
__global__ void __launch_bounds__(ST_THREADS, 2) bwd_flow_staged_kernel(const __grid_constant__ Params P,
                                                                       const __grid_constant__ GradP Q,
                                                                       int smem_floats) {
  extern __shared__ float4 st_smem4[];
  float* const smem = reinterpret_cast<float*>(st_smem4);
  __shared__ StageTab tb[NDIRS];
  __shared__ StageSlow slow;
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  StageCtx<NDIRS> cx;
  stage_prologue<NDIRS>(P, tb, slow, smem, smem_floats, cx);
  const int n = cx.n, t = cx.t, j = cx.j;

  if (!cx.pl.ok) {
#pragma unroll
    for (int q = 0; q < ST_PPT; ++q)
      if (cx.inimg[q]) bwdflow_generic_pixel<NDIRS>(P, Q, n, t, cx.irow[q], j);
    return;
  }
  const int cc = cx.pl.cc, plane = cx.pl.plane;
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) has_bl[d] = P.dir[d].blend != nullptr;
  unsigned skip = 0u;  // groups without grad_out contribute nothing
#pragma unroll
  for (int g = 0; g < FWB_MAX_GROUPS; ++g) skip |= (Q.grad_out[g] == nullptr ? 1u : 0u) << g;

  float gix[ST_PPT][NDIRS], giy[ST_PPT][NDIRS], gbl[ST_PPT][NDIRS];
#pragma unroll
  for (int q = 0; q < ST_PPT; ++q)
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) gix[q][d] = giy[q][d] = gbl[q][d] = 0.f;

  StageProducer<NDIRS> prod;
  prod.init(P, cx, (unsigned)__cvta_generic_to_shared(smem));
  int stage = 0;
  int nch = prod.issue(P, cx, 0, skip);
  for (int g = 0; g < G.n_groups; ++g) {
    if ((skip >> g) & 1u) continue;
    const GroupP& R = P.grp[g];
    const float* gp = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + j;
    int grow[ST_PPT];
#pragma unroll
    for (int q = 0; q < ST_PPT; ++q) grow[q] = cx.irow[q] * Q.go_sh[g];
    for (int c0 = 0; c0 < R.C; c0 += cc) {
      // this chunk's grad_out values (coalesced), issued before the wait so that they overlap it
      float go[ST_CCMAX][ST_PPT];
#pragma unroll
      for (int u = 0; u < ST_CCMAX; ++u)
#pragma unroll
        for (int q = 0; q < ST_PPT; ++q)
          go[u][q] = (u < nch && cx.act[q]) ? __ldcs(gp + (long long)(c0 + u) * Q.go_sc[g] + grow[q]) : 0.f;
      const int nxt = prod.issue(P, cx, stage ^ 1, skip);
      if (nxt)
        cp_async_wait<1>();
      else
        cp_async_wait<0>();
      __syncthreads();
      const float* sp = smem + stage * cc * plane;
#pragma unroll
      for (int u = 0; u < ST_CCMAX; ++u) {
        if (u < nch) {
#pragma unroll
          for (int q = 0; q < ST_PPT; ++q) {
            const float gout = go[u][q];
#pragma unroll
            for (int d = 0; d < NDIRS; ++d) {
              const StagePix& px = cx.px[q][d];
              const float* s0 = sp + px.o0;
              const float* s1 = sp + px.o1;
              const float a = s0[0], b = s0[1], c_ = s1[0], dd = s1[1];
              float gw = gout;
              if (has_bl[d]) {
                const float top = fmaf(b, px.tx, a * px.ux), bot = fmaf(dd, px.tx, c_ * px.ux);
                gbl[q][d] = fmaf(gout, fmaf(bot, px.ty, top * px.uy), gbl[q][d]);
                gw = gout * px.bl;
              }
              gix[q][d] = fmaf(gw, fmaf(px.ty, dd - c_, px.uy * (b - a)), gix[q][d]);
              giy[q][d] = fmaf(gw, fmaf(px.tx, dd - b, px.ux * (c_ - a)), giy[q][d]);
            }
          }
          sp += plane;
        }
      }
      __syncthreads();
      stage ^= 1;
      nch = nxt;
    }
  }

#pragma unroll
  for (int q = 0; q < ST_PPT; ++q) {
    if (!cx.act[q]) continue;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      Tap k;
      compute_tap(G, P.dir[d], n, t, cx.irow[q], j, k);  // mx, my, fx, fy, gate (cheaper to recompute than to hold)
      bwdflow_store(P, Q, d, n, t, cx.irow[q], j, k, gix[q][d], giy[q][d], gbl[q][d]);
    }
  }
  // slow pixels: one warp per pixel, lanes stride over the channels, warp-reduced
  for (int s = warp; s < slow.n; s += ST_THREADS / 32) {
    const int pix = slow.pix[s];
    bwdflow_generic_pixel_warp<NDIRS>(P, Q, n, t, blockIdx.y * ST_TH + (pix >> 5), blockIdx.x * ST_TW + (pix & 31));
  }
}

}  // namespace fwb

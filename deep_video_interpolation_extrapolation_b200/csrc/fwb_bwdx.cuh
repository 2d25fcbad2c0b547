// fwb_bwdx.cuh — kernel 3 of the default backward: the source-gradient scatter as a SORTED shared-memory gather.
//
// The default backward on dense sources is two launches that load two different pipes of the SM:
//   * kernel 2 (bwd_flow_tex_kernel, fwb_tex.cuh): coordinate gradient -> grad_flow / grad_gate / grad_blend, one TLD4 per
//     (pixel, direction, channel) on the TEXTURE units, no shared memory;
//   * kernel 3 (bwd_src_sorted_kernel, this file): grad_src on the SHARED-MEMORY pipe.  It needs grad_out, the flows and the
//     masks, but not the source planes.
// Which cell of a tile's footprint a tap adds to does not depend on the channel, so the tile's taps (4 per pixel and direction,
// 4096 for a 32x16 tile and two directions) are counting-sorted by cell ONCE per tile; every thread then owns 16 consecutive
// records (pixel, weight) of the sorted list in registers.  Per channel it reads the 16 grad_out*blend values of its records
// from a shared copy of the tile (LDS, duplicates broadcast), sums each run of equal cells in a register and writes the run's
// total to a float accumulator with a plain store.  A run that continues in the next lane is handed over with one shuffle (a
// segmented warp scan when a run covers three or more lanes); the part of a run that lies in the previous warp goes to a
// side slot that the flush adds.  No atomics, every cell has exactly one writer.  This replaces the 4 integer shared-memory
// atomics per tap of the tile kernel (fwb_tile.cuh: 3.8 wavefronts each, no broadcast, ~60 % of that kernel's shared-memory
// time) by one load per tap and ~0.35 stores, needs no fixed-point scale vote, and propagates inf / NaN like the reference's
// float atomicAdd does.  (Fused with kernel 2 in one kernel the register budget does not hold both halves: 229 registers
// wanted, 128 available at 2 CTAs/SM -> spills and rematerialised address arithmetic, 1.07 ms against 0.65 ms for the
// integer-atomic kernel; hence two launches.)
// The accumulators (row-segment footprint of the tile, double buffered over the channels) are flushed into grad_src with
// red.global.add.v4.f32 while the next channel is processed; grad_src must be zero on entry.
// Reference: ATen grid_sampler_2d_backward (atomicAdd scatter of w*gOut), reached from utils/net_utils.py:113 and
// nets/OpticalUnet.py:132-139 by autograd; mask weighting nets/OpticalUnet.py:141-146.
#pragma once
#include "fwb_tile.cuh"

namespace fwb {

#ifndef BX_MINCTA
#define BX_MINCTA 2  // resident CTAs per SM the kernel is compiled for
#endif
constexpr int BX_THREADS = 256;

struct BxChan {  // per flattened channel, in shared memory
  float* gs[2];  // grad_src plane per direction, may be NULL
  const float* go;
  int g, c;
};

__device__ __forceinline__ void bx_sts(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
// one record of the sorted tap list, written as PTX blocks so that the unpacked addresses never leave the block (the compiler
// would otherwise keep 2 x 16 loop-invariant addresses in registers) and nothing turns into a branch:
//   bx_rec_load : g = shared[gb + (pk & 0xffff)]
//   bx_rec_end  : if (endm & BIT) { shared[ab + (pk >> 16)] = acc; acc = 0; }
__device__ __forceinline__ float bx_rec_load(unsigned pk, unsigned gb) {
  float v;
  asm volatile("{\n .reg .u32 a;\n and.b32 a, %1, 0xffff;\n add.u32 a, a, %2;\n ld.shared.f32 %0, [a];\n}\n" : "=f"(v) : "r"(pk), "r"(gb));
  return v;
}
template <unsigned BIT>
__device__ __forceinline__ void bx_rec_end(unsigned pk, unsigned ab, unsigned endm, float& acc) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .u32 a, t;\n and.b32 t, %3, %4;\n setp.ne.u32 p, t, 0;\n shr.u32 a, %1, 16;\n add.u32 a, a, %2;\n"
      " @p st.shared.f32 [a], %0;\n selp.f32 %0, 0f00000000, %0, p;\n}\n"
      : "+f"(acc)
      : "r"(pk), "r"(ab), "r"(endm), "n"(BIT)
      : "memory");
}
// Byte offset of pixel q of this thread, direction d, inside a grad_out tile buffer: the tile row-major (32 pixels = 32 banks
// per row) with a row skew of 11 words.  The records a warp reads in one instruction are 16 apart in the sorted list: their
// pixels run along two or three neighbouring output rows over similar columns, which this layout spreads over distinct banks
// (the 8x4 patch order of the threads would fold a row of 32 pixels onto 8 banks: measured 3.8 wavefronts per load).
template <int PPT>
__device__ __forceinline__ unsigned bx_gpos(int d, int q) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = (warp & 3) * 8 + (lane & 7), y = ((warp >> 2) + 2 * q) * 4 + (lane >> 3);  // position inside the tile
  return 4u * (unsigned)(d * (BX_THREADS * PPT) + y * 32 + ((x + 11 * y) & 31));
}
template <int R, int NREC>
struct BxRun {
  static __device__ __forceinline__ void go(const unsigned* pk, const float* wv, const float* gv, unsigned ab, unsigned endm, float& acc) {
    acc = fmaf(gv[R], wv[R], acc);
    bx_rec_end<(1u << R)>(pk[R], ab, endm, acc);
    BxRun<R + 1, NREC>::go(pk, wv, gv, ab, endm, acc);
  }
};
template <int NREC>
struct BxRun<NREC, NREC> {
  static __device__ __forceinline__ void go(const unsigned*, const float*, const float*, unsigned, unsigned, float&) {}
};

template <int NDIRS, bool ALIGN, bool BORDER, int PPT, int MINCTA = BX_MINCTA>
__global__ void __launch_bounds__(BX_THREADS, MINCTA) bwd_src_sorted_kernel(const __grid_constant__ Params P, const __grid_constant__ GradP Q,
                                                                              int smem_floats) {
  constexpr int NTHR = BX_THREADS, SLOTS = TL_SLOTS;
  constexpr int NREC = PPT * NDIRS * 4;     // records (taps) per thread
  constexpr int NPIX = NTHR * PPT;          // pixels per tile
  constexpr int GBUF_F = 2 * NDIRS * NPIX;  // floats of the two grad_out*blend tiles (double buffered over the channels)
  extern __shared__ float4 tl_smem4[];
  float* const smem = reinterpret_cast<float*>(tl_smem4);
  __shared__ StageSlow slow;
  __shared__ Tap slowtap[TL_MAXSLOW][NDIRS];
  __shared__ __align__(16) BxChan tab[TL_MAXCH];
  __shared__ int nchan_s, g0_s;
  __shared__ unsigned wsum[NTHR / 32];
  __shared__ unsigned sidecell[NTHR / 32];     // accumulator byte offset of the run that crosses the boundary behind warp w
  __shared__ float side[2][NTHR / 32];         // its partial sum inside warp w, per accumulator parity
  __shared__ unsigned maxcnt_s;
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int n, t, j, total, stage_f;
  int irow[PPT];
  bool act[PPT];
  float bl[PPT][NDIRS];
  unsigned info[SLOTS], pd[SLOTS];
  // records of this thread in the sorted tap list: packed (accumulator byte offset << 16 | byte offset of the pixel's
  // grad_out*blend inside a tile buffer) and bilinear weight
  unsigned pk[NREC];
  float wv[NREC];
  unsigned endm;  // bit k: a run of equal cells ends at record k (its sum is stored there)
  bool head_pending, tail_open, transparent, chain_warp;
  unsigned fixm[SLOTS];  // flush: up to two 6-bit entries (warp + 1) | float index << 4: piece s also receives that side slot
  // grad_out of the first channel: issued before the prologue so that its latency hides behind it
  float ego[PPT];
  {
    int g0e = 0;
    while (g0e < G.n_groups - 1 && !Q.grad_out[g0e]) ++g0e;
    const int nt = blockIdx.z, ne = G.T == 1 ? nt : nt / G.T, te = nt - ne * G.T;
    const int je = blockIdx.x * TL_TW + (warp & 3) * 8 + (lane & 7);
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int ie = blockIdx.y * (8 * PPT) + ((warp >> 2) + 2 * q) * 4 + (lane >> 3);
      ego[q] = (Q.grad_out[g0e] && je < G.W && ie < G.H)
                   ? __ldcs(Q.grad_out[g0e] + ne * Q.go_sn[g0e] + te * Q.go_st[g0e] + (long long)ie * Q.go_sh[g0e] + je)
                   : 0.f;
    }
  }
  {
    TileCtx<NDIRS, PPT, SLOTS> cx;
    // loop layout: two accumulators + the grad_out tiles; sort scratch: counters + the record list
    const int budget = min(min(smem_floats - GBUF_F, 2 * (smem_floats - 2 * NTHR * NREC)), 2 * 4096);  // prefix scan: <= 4096 cells
    tile_prologue<NDIRS, ALIGN, BORDER, PPT, SLOTS, NTHR>(P, reinterpret_cast<TileTab*>(smem), slow, slowtap, budget, 2, cx);
    n = cx.n, t = cx.t, j = cx.j;
    if (!cx.ok) {
#pragma unroll
      for (int q = 0; q < PPT; ++q)
        if (cx.inimg[q]) bwd_src_generic_pixel<NDIRS>(P, Q, n, t, cx.irow[q], j);
      return;
    }
    total = cx.total;
    stage_f = cx.stage_f;
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      irow[q] = cx.irow[q];
      act[q] = cx.act[q];
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const TilePix& px = cx.px[q][d];
        bl[q][d] = px.bl;
        // this pixel's 4 records: cell (float index inside an accumulator) and weight; a tap outside the image (or any tap
        // of a slow / out-of-tile pixel) goes to one of 16 dummy cells in front of the accumulator with weight zero
        const int cell[4] = {px.o0, px.o0 + 1, px.o1, px.o1 + 1};
        const float w4[4] = {__fmul_rn(px.ux, px.uy), __fmul_rn(px.tx, px.uy), __fmul_rn(px.ux, px.ty), __fmul_rn(px.tx, px.ty)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool on = (px.vld >> k) & 1u;
          const int r = (q * NDIRS + d) * 4 + k;
          pk[r] = (unsigned)(on ? cell[k] : (int)(threadIdx.x & (TL_ZPAD - 1)));  // cell index for now
          wv[r] = on ? w4[k] : 0.0f;
        }
      }
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      info[s] = cx.info[s];
      pd[s] = 4u * (unsigned)cx.pdst[s];
    }
  }
  // channel table over the groups that have a grad_out (the others contribute nothing)
  if (threadIdx.x < 32) {
    int base = 0, g0 = -1;
    for (int g = 0; g < G.n_groups; ++g) {
      if (!Q.grad_out[g]) continue;
      if (g0 < 0) g0 = g;
      const GroupP& R = P.grp[g];
      for (int c = lane; c < R.C; c += 32) {
        BxChan e;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          float* gs = (d < NDIRS) ? Q.grad_src[g][d] : nullptr;
          e.gs[d] = gs ? gs + n * Q.gs_sn[g][d] + t * Q.gs_st[g][d] + (long long)c * Q.gs_sc[g][d] : nullptr;
        }
        e.go = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)c * Q.go_sc[g];
        e.g = g;
        e.c = c;
        tab[base + c] = e;
      }
      base += R.C;
    }
    if (lane == 0) {
      nchan_s = base;
      g0_s = g0 < 0 ? 0 : g0;
    }
  }
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) has_bl[d] = P.dir[d].blend != nullptr;
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
  __syncthreads();  // the tables in dynamic shared memory are dead from here on; tab / slowtap are visible

  // ------------------------------------------------------------------ counting sort of the tile's taps by cell
  {
    unsigned* const cnt = reinterpret_cast<unsigned*>(smem);  // [stage_f] (stage_f is a multiple of 4)
    unsigned* const rec_pk = cnt + stage_f;                   // [NTHR * NREC]
    float* const rec_w = reinterpret_cast<float*>(rec_pk + NTHR * NREC);
    const int n4 = stage_f >> 2;
    for (int k = threadIdx.x; k < n4; k += NTHR) reinterpret_cast<uint4*>(cnt)[k] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) maxcnt_s = 0u;
    __syncthreads();
    unsigned rank[NREC];
#pragma unroll
    for (int r = 0; r < NREC; ++r) rank[r] = atomicAdd(&cnt[pk[r]], 1u);
    __syncthreads();
    {  // exclusive prefix sum of the counters, in place: per thread `per` uint4, warp scan, warp totals
      const int per = (n4 + NTHR - 1) / NTHR;  // <= 4 (stage_f <= 4096)
      uint4 v[4];
      unsigned s = 0u, mx = 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = threadIdx.x * per + u;
        v[u] = (u < per && idx < n4) ? reinterpret_cast<uint4*>(cnt)[idx] : make_uint4(0u, 0u, 0u, 0u);
        s += v[u].x + v[u].y + v[u].z + v[u].w;
        if (idx >= TL_ZPAD / 4) mx = max(mx, max(max(v[u].x, v[u].y), max(v[u].z, v[u].w)));  // the dummy cells do not count
      }
      mx = __reduce_max_sync(0xffffffffu, mx);
      if (lane == 0 && mx > 0u) atomicMax(&maxcnt_s, mx);
      unsigned incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      unsigned run = incl - s;
      for (int w = 0; w < warp; ++w) run += wsum[w];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = threadIdx.x * per + u;
        uint4 o;
        o.x = run, run += v[u].x;
        o.y = run, run += v[u].y;
        o.z = run, run += v[u].z;
        o.w = run, run += v[u].w;
        if (u < per && idx < n4) reinterpret_cast<uint4*>(cnt)[idx] = o;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NREC; ++r) {
      const int qd = r >> 2;  // (q * NDIRS + d)
      const int q = qd / NDIRS, d = qd - q * NDIRS;
      const unsigned slot = cnt[pk[r]] + rank[r];
      const unsigned pixb = bx_gpos<PPT>(d, q);  // byte offset of this pixel inside a grad_out tile buffer
      rec_pk[slot] = ((4u * pk[r]) << 16) | pixb;
      rec_w[slot] = wv[r];
    }
    __syncthreads();
    const int r0 = threadIdx.x * NREC;
#pragma unroll
    for (int r = 0; r < NREC; r += 4) {
      const uint4 a = *reinterpret_cast<const uint4*>(rec_pk + r0 + r);
      const float4 b = *reinterpret_cast<const float4*>(rec_w + r0 + r);
      pk[r] = a.x, pk[r + 1] = a.y, pk[r + 2] = a.z, pk[r + 3] = a.w;
      wv[r] = b.x, wv[r + 1] = b.y, wv[r + 2] = b.z, wv[r + 3] = b.w;
    }
    const unsigned none = 0xffffu;
    const unsigned prevc = threadIdx.x > 0 ? (rec_pk[r0 - 1] >> 16) : none;
    const unsigned nextc = threadIdx.x < NTHR - 1 ? (rec_pk[r0 + NREC] >> 16) : none;
    endm = 0u;
#pragma unroll
    for (int r = 0; r < NREC; ++r) {
      const unsigned c0 = pk[r] >> 16, c1 = r + 1 < NREC ? (pk[r + 1] >> 16) : nextc;
      endm |= (c0 != c1 ? 1u : 0u) << r;
    }
    const unsigned dummy_b = 4u * TL_ZPAD;  // accumulator byte offsets below this are the dummy cells: never stored
    const bool head_open = (pk[0] >> 16) == prevc && (pk[0] >> 16) >= dummy_b;
    tail_open = !((endm >> (NREC - 1)) & 1u) && (pk[NREC - 1] >> 16) >= dummy_b;
    const bool whole = (endm & ((1u << (NREC - 1)) - 1u)) == 0u;  // one run fills the chunk
    transparent = whole && head_open && tail_open;                 // ... and continues on both sides
    head_pending = head_open && !(whole && tail_open);             // the open head run ends inside this chunk
    chain_warp = __any_sync(0xffffffffu, transparent);
    if (lane == 31) sidecell[warp] = tail_open ? (pk[NREC - 1] >> 16) : none;
  }
  __syncthreads();  // sidecell / maxcnt_s visible
  if (maxcnt_s > 16u * NREC) {  // a cell with hundreds of taps (the tile folds onto a few cells): a run must not cross two warp boundaries
#pragma unroll
    for (int q = 0; q < PPT; ++q)
      if (j < G.W && irow[q] < G.H) bwd_src_generic_pixel<NDIRS>(P, Q, n, t, irow[q], j);
    return;
  }
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    // at most two warp boundaries can be crossed inside one 4-cell piece (three would need a cell with > 16 * NREC taps)
    fixm[s] = 0u;
    for (int w = NTHR / 32 - 2; w >= 0; --w) {
      const unsigned off = sidecell[w] - pd[s];  // wraps when below the piece
      if (sidecell[w] != 0xffffu && off < 16u) fixm[s] = (fixm[s] << 6) | (unsigned)(w + 1) | ((off >> 2) << 4);
    }
  }
  __syncthreads();  // the sort scratch is dead: accumulators and grad_out tiles take its place

  const int Cn = nchan_s, g0 = g0_s;
  int goff[SLOTS];  // element offset of piece tid + s*256 in its grad_src plane, or PIECE_NONE
  unsigned tsel[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int d = piece_dir(info[s]);
    tsel[s] = 8u * (unsigned)d;
    const bool none = threadIdx.x + s * NTHR >= total, zero = piece_zero(info[s]);
    goff[s] = (none || zero || Q.grad_src[g0][d] == nullptr) ? PIECE_NONE : piece_y(info[s]) * Q.gs_sh[g0][d] + piece_col(info[s]);
  }
  int gooff[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) gooff[q] = irow[q] * Q.go_sh[g0] + j;
  const unsigned stage_b = 4u * (unsigned)stage_f;
  const unsigned acc_s = smem_s;                     // two accumulators, stage layout
  const unsigned gb_s = smem_s + 2u * stage_b;       // two grad_out*blend tile buffers of NDIRS * NPIX floats
  constexpr unsigned GB_B = 4u * NDIRS * NPIX;
  {  // every cell of both accumulators starts from zero: a cell that has records is overwritten for every channel, one without
     // keeps its zero (the flush skips it)
    const int n4 = stage_f >> 1;  // two accumulators of stage_f / 4 float4 each
    for (int k = threadIdx.x; k < n4; k += NTHR) tl_smem4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);
  const unsigned head_a = pk[0] >> 16;  // accumulator byte offset of the open head run
  const unsigned side_s = (unsigned)__cvta_generic_to_shared(&side[0][0]);

  auto load_go = [&](int cf, float* go) {
    const bool live = cf < Cn;
    const float* gp = tab[live ? cf : 0].go;
#pragma unroll
    for (int q = 0; q < PPT; ++q) go[q] = (live && act[q]) ? __ldcs(gp + gooff[q]) : 0.f;
  };
  // grad_out * blend of this thread's pixels into tile buffer `buf` (what the records of all threads read)
  auto put_g = [&](const float* go, unsigned buf) {
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
      for (int d = 0; d < NDIRS; ++d)
        bx_sts(gb_s + buf * GB_B + bx_gpos<PPT>(d, q), has_bl[d] ? go[q] * bl[q][d] : go[q]);
  };
  // flush this thread's pieces of the accumulator of parity `pr` (channel cf) into grad_src
  auto flush = [&](int cf, unsigned pr) {
    const unsigned te = tab_s + (unsigned)sizeof(BxChan) * (unsigned)cf;  // &tab[cf].gs[0]
    const unsigned ab = acc_s + pr * stage_b;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if (goff[s] < 0) continue;
      float4 u;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(u.x), "=f"(u.y), "=f"(u.z), "=f"(u.w) : "r"(ab + pd[s]));
      if (fixm[s]) {  // cells whose run started in the previous warp: add that warp's part (rare)
        unsigned f = fixm[s];
#pragma unroll 1
        do {
          const float sv = side[pr][(f & 15u) - 1u];
          const unsigned l = (f >> 4) & 3u;
          u.x += l == 0u ? sv : 0.f, u.y += l == 1u ? sv : 0.f, u.z += l == 2u ? sv : 0.f, u.w += l == 3u ? sv : 0.f;
          f >>= 6;
        } while (f);
      }
      if (u.x == 0.f && u.y == 0.f && u.z == 0.f && u.w == 0.f) continue;
      float* gs = reinterpret_cast<float*>(tl_lds64(te + tsel[s]));
      if (gs) red_add_v4(gs + goff[s], u);
    }
  };

  float go[PPT], gn[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) go[q] = (Cn > 0 && act[q]) ? ego[q] : 0.f;  // channel 0 (loaded before the prologue)
  put_g(go, 0u);
  // slow pixels (far from the rest of the tile): global float atomics, one (pixel, channel) item per thread
  {
    const int nslow = slow.n;
    for (int it = threadIdx.x; it < nslow * Cn; it += NTHR) {
      const int cf = it / nslow, sidx = it - cf * nslow, pix = slow.pix[sidx];  // lanes = different pixels: no same-address
      const int si = blockIdx.y * (8 * PPT) + (pix >> 5), sj = blockIdx.x * TL_TW + (pix & 31);
      const BxChan& tc = tab[cf];
      const float gout = __ldg(tc.go + si * Q.go_sh[g0] + sj);
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const Tap& k = slowtap[sidx][d];
        scatter_atomic_px(Q, tc.g, d, n, t, tc.c, k, has_bl[d] ? gout * k.blend : gout);
      }
    }
  }
  unsigned par = 0u;  // cf & 1
#pragma unroll 1
  for (int cf = 0; cf < Cn; ++cf) {
    __syncthreads();  // grad_out tile of channel cf complete; accumulator and side slots of cf-1 complete; flush of cf-2 complete
    load_go(cf + 1, gn);
    if (cf > 0) flush(cf - 1, par ^ 1u);
    // ---- kernel 3: this thread's records of the sorted tap list
    {
      const unsigned gb = gb_s + par * GB_B, ab = acc_s + par * stage_b;
      float gv[NREC];
#pragma unroll
      for (int r = 0; r < NREC; ++r) gv[r] = bx_rec_load(pk[r], gb);  // all loads first: one exposed latency
      // every run's sum is stored where the run ends (runs of dummy cells land in the pad in front of the accumulator);
      // afterwards `acc` is the partial sum of the run that continues in the next thread (0 when the last run closed)
      float acc = 0.f;
      BxRun<0, NREC>::go(pk, wv, gv, ab, endm, acc);
      float v = tail_open ? acc : 0.f;
      if (chain_warp) {  // a run covers whole lanes: segmented scan (a transparent lane passes what it receives on)
        bool f = transparent;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float pv = __shfl_up_sync(0xffffffffu, v, o);
          const bool pf = __shfl_up_sync(0xffffffffu, f ? 1 : 0, o) != 0;
          if (lane >= o && f) {
            v += pv;
            f = pf;
          }
        }
      }
      const float carry = __shfl_up_sync(0xffffffffu, v, 1);
      // the open head run ended inside this chunk and was stored without the previous lane's part: add it (this thread is the
      // only writer of that cell; the previous WARP's part arrives through the side slot at flush time)
      if (head_pending && lane != 0) bx_sts(ab + head_a, tl_lds(ab + head_a) + carry);
      if (lane == 31 && tail_open) bx_sts(side_s + 4u * (par * (NTHR / 32) + (unsigned)warp), v);
    }
    put_g(gn, par ^ 1u);
#pragma unroll
    for (int q = 0; q < PPT; ++q) go[q] = gn[q];
    par ^= 1u;
  }
  __syncthreads();
  if (Cn > 0) flush(Cn - 1, par ^ 1u);
}

}  // namespace fwb

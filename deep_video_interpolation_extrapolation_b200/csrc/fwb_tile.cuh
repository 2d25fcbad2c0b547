// fwb_tile.cuh — the default kernels: forward (kernel 1) and fused backward (kernels 2+3) on PLANAR
// shared-memory tiles, one channel plane per pipeline stage.
//
// A CTA owns a 32x16 tile of output pixels of one (n, t): 8 warps, warp w owns rows w and w+8, lane = column.
// The prologue computes everything that does not depend on the channel — taps, bilinear weights, the tile's
// row-segment footprint for both directions, which 16-byte pieces of that footprint each
// thread copies, the offsets of each pixel's taps inside the staged footprint — and keeps it in registers.
// The channel loop is then a cp.async ring (TL_FD / TL_BD stages, ONE __syncthreads per channel):
//     wait own copies of channel c | barrier | issue copies of channel c+D-1 | gather channel c from smem
// Per (pixel, direction, channel) the forward costs 4 LDS.32 + 4 FFMA; the footprint pieces outside the image are
// zero-filled by cp.async (src-size 0), so the gather needs no validity test: an out-of-image tap reads 0.
// The backward adds, per channel, the scatter of grad_out*weight into an int32 FIXED-POINT shared-memory copy
// of the same footprint (ATOMS.ADD is native, fp32 shared atomics are CAS loops on sm_100), double buffered so
// that the flush of channel c-1 (int -> float, red.global.add.v4.f32 on the pieces the thread owns) overlaps
// the scatter of channel c.  Scale per channel: 2^(20 - exponent(max |grad_out*blend| over the tile)).
// Pixels whose taps are far from the rest of the tile (border-clamped outliers) are SLOW: served from global
// memory.  A tile whose footprint does not fit goes to the generic per-pixel path.
#pragma once
#include "fwb_coords.cuh"
#include "fwb_generic.cuh"
#include "fwb_util.cuh"
#include "fwb_tex.cuh"

namespace fwb {

#ifndef TL_FWD_MINCTA
#define TL_FWD_MINCTA 3  // resident CTAs per SM the forward is compiled for (4 = 64 registers: spills, measured below)
#endif
constexpr int TL_TW = 32, TL_TH = 16;
constexpr int TL_THREADS = 256;
constexpr int TL_PPT = 2;       // pixels per thread
constexpr int TL_SLOTS = 3;     // 16-byte pieces per thread and channel (both directions together): <= 768 pieces
constexpr int TL_MAXCH = 64;    // flattened channels per launch (table in shared memory)
constexpr int TL_MAXSLOW = 32;  // slow pixels a tile may have before the whole tile goes generic
constexpr int TL_FD = 4;        // forward ring depth
constexpr int TL_BD = 4;        // backward ring depth

constexpr int TL_R = 24;                       // a pixel displaced farther than this from the tile's anchor is SLOW
constexpr int TL_ROWS = TL_TH + 2 * TL_R + 2;  // <= 66 source rows an inlier tap can touch: window [i0 + ady - R, ...)
constexpr int TL_ROWS_P = 96;                  // padded: 3 rows per lane in the scan
constexpr int TL_ZPAD = 16;                    // zero cells in front of every stage (target of tap-less pixels)
constexpr int TL_SK4 = 3;                      // row skew in pieces: cell (x, row r) sits at offset == x + 12 r (mod 32)
constexpr bool TL_FILL_GAPS = false;           // A/B: fill the row placement gaps with real pieces (measured slower: 0.323 / 0.728 ms vs 0.304 / 0.680)
#ifndef TL_BWD_MINCTA
#define TL_BWD_MINCTA 2  // resident CTAs per SM the scatter backward is compiled for (3 = 85 registers: measured below)
#endif
#ifndef TL_BWDX_MINCTA
#define TL_BWDX_MINCTA 2  // resident CTAs per SM the texture-path scatter backward is compiled for
#endif
#ifndef TL_SMOOTH_LEN4
#define TL_SMOOTH_LEN4 10  // a tile whose widest row segment has at most this many 16-byte pieces counts as near-rigid
#endif
constexpr int TL_SK4_SMOOTH = 2;                // row skew (pieces) of near-rigid tiles
constexpr int TL_AL4 = 8;                      // row placement period in 4-cell pieces: a staged cell of image column x sits at cell offset == x (mod 16)

struct TilePix {  // per (pixel of this thread, direction)
  float tx, ty, ux, uy, bl;
  unsigned clip;  // bit 0 / 1: border padding clipped ix / iy (the coordinate gradient is then zero)
  int o0, o1;  // float offsets inside a stage of the nw and sw taps
  float fx1, fy1;  // x0 + 1, y0 + 1 (texture path: fwb_tex.cuh); 0 when the pixel has no tap in this tile
  unsigned vld;    // validity bits of the 4 taps (0 for slow / out-of-image pixels)
};

// Row-segment table of one direction.  Row r is source row ybase + r.  The segment of row r, 4-aligned columns
// [rowx, rowx + 4*len4), is stored at float offset rowbase[r] of a stage with rowbase[r] == rowx[r] (mod 32):
// the shared-memory bank of a staged cell is its image column mod 32, so the 32 taps of a warp instruction
// (32 consecutive output columns -> nearly consecutive source columns, whatever rows they fall on) are
// bank-conflict free unless the flow stretches the row beyond 32 columns or folds it.
struct TileTab {
  int xlo[TL_ROWS_P], xhi[TL_ROWS_P];
  int rowx[TL_ROWS_P];
  int rowbase[TL_ROWS_P];
  int rowoff4[129];  // exclusive prefix of the piece counts; entries above TL_ROWS_P hold INT_MAX (binary search)
  unsigned akey32;
  int pad0;
  int total4;   // 16-byte pieces of this direction
  int alloc4;   // float4 allocated (padding included), multiple of 8
  int ybase;
  int pad;
};

template <int NDIRS, int PPT, int SLOTS>
struct TileCtx {
  int n, t, j;
  int irow[PPT];
  bool inimg[PPT], act[PPT];
  TilePix px[PPT][NDIRS];
  unsigned info[SLOTS];  // piece tid + s*256: dir << 31 | y << 16 | zero-fill << 15 | col
  int pdst[SLOTS];       // float offset of that piece inside a stage
  int total;                // pieces per channel, both directions
  int stage_f;              // floats per stage (zero pad + both directions' slots)
  int ok;
};

__device__ __forceinline__ int piece_dir(unsigned info) { return (int)(info >> 31); }
__device__ __forceinline__ int piece_y(unsigned info) { return (int)((info >> 16) & 0x7fffu); }
__device__ __forceinline__ int piece_col(unsigned info) { return (int)(info & 0x7fffu); }
__device__ __forceinline__ bool piece_zero(unsigned info) { return (info & 0x8000u) != 0u; }

// one warp per direction: segment lengths, placement, prefix sums.  A pixel recorded [x0, x0+1] on its nw row only;
// its sw / se taps are the same columns one row down, so the segment of row r is own[r] U own[r-1].
__device__ __forceinline__ void tile_tab_scan(TileTab& tb, int slot_start4) {
  const int lane = threadIdx.x & 31;
  int xs[3], ln[3], ph[3];
  int olo[3], ohi[3];
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    olo[h] = tb.xlo[3 * lane + h];
    ohi[h] = tb.xhi[3 * lane + h];
  }
  int plo = __shfl_up_sync(0xffffffffu, olo[2], 1), phi = __shfl_up_sync(0xffffffffu, ohi[2], 1);
  if (lane == 0) plo = 0x7fffffff, phi = -0x7fffffff;
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    const int lo = min(olo[h], h == 0 ? plo : olo[h - 1]), hi = max(ohi[h], h == 0 ? phi : ohi[h - 1]);
    const bool has = lo <= hi;
    xs[h] = has ? ((lo >> 2) << 2) : 0;             // arithmetic shift: -1 -> -4
    ln[h] = has ? ((hi + 4) >> 2) - (lo >> 2) : 0;  // float4 pieces
  }
  // row skew per tile: near-rigid footprints (every row about as wide as the tile) are gathered conflict-free with a skew of
  // 8 words (four rows of an 8x4 patch land on banks 0-7, 8-15, ...); stretched / sheared ones do best with 12
  int sk4 = TL_SK4;
  if (TL_SK4_SMOOTH != TL_SK4) {
    const int widest = __reduce_max_sync(0xffffffffu, max(ln[0], max(ln[1], ln[2])));
    if (widest <= TL_SMOOTH_LEN4) sk4 = TL_SK4_SMOOTH;
  }
#pragma unroll
  for (int h = 0; h < 3; ++h) ph[h] = ((xs[h] >> 2) + sk4 * (3 * lane + h)) & (TL_AL4 - 1);
  // end phase of the last non-empty row before this lane's rows (empty rows take no space and no padding):
  // v = 8 | end phase of the lane's last non-empty row, 0 when all three are empty; inclusive "last valid" scan
  int v_e = 0;
#pragma unroll
  for (int h = 0; h < 3; ++h)
    if (ln[h] > 0) v_e = 8 | ((ph[h] + ln[h]) & (TL_AL4 - 1));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v_e, o);
    if (lane >= o && v_e == 0) v_e = u;
  }
  int e_prev = __shfl_up_sync(0xffffffffu, v_e, 1);
  e_prev = lane == 0 ? 0 : (e_prev & 7);  // nothing before: the slot starts aligned
  int pad[3];
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    pad[h] = ln[h] > 0 ? ((ph[h] - e_prev) & (TL_AL4 - 1)) : 0;
    if (ln[h] > 0) e_prev = (ph[h] + ln[h]) & (TL_AL4 - 1);
    if (TL_FILL_GAPS) {  // no gap in front of the row: the row is extended to the left by the same number of pieces instead,
      xs[h] -= 4 * pad[h];  // so the staged pieces are contiguous in shared memory and a warp's 16-byte cp.async writes
      ln[h] += pad[h];      // never hit the same banks twice (the copies of the extra pieces ride on idle lanes)
      pad[h] = 0;
    }
  }
  const int mine = ((pad[0] + pad[1] + pad[2] + ln[0] + ln[1] + ln[2]) << 16) | (ln[0] + ln[1] + ln[2]);
  int v = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  int run = v - mine;  // exclusive: alloc << 16 | pieces
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    const int r = 3 * lane + h;
    tb.rowoff4[r] = run & 0xffff;
    tb.rowx[r] = xs[h];
    tb.rowbase[r] = TL_ZPAD + 4 * (slot_start4 + (run >> 16) + pad[h]);
    run += ((pad[h] + ln[h]) << 16) | ln[h];
  }
  tb.rowoff4[TL_ROWS_P + lane] = 0x7fffffff;
  if (lane == 31) {
    tb.rowoff4[128] = 0x7fffffff;
    tb.total4 = run & 0xffff;
    tb.alloc4 = ((run >> 16) + TL_AL4 - 1) & ~(TL_AL4 - 1);
  }
}

// Everything channel independent.  `tb` (NDIRS tables) lives in the dynamic shared memory that the stages
// overwrite later: the caller must __syncthreads() between this call and the first copy.
// The slow pixels' taps are written to slowtap from the fast taps (enough for the forward).
// Host-checked: the in-plane offsets of flow / gate / blend fit in 32 bits.
// NTHR threads per CTA = NTHR/128 rows of four 8x4 warp patches per pixel slot: tile height 4 * (NTHR/128) * PPT
template <int NDIRS, bool ALIGN, bool BORDER, int PPT, int SLOTS, int NTHR = TL_THREADS>
__device__ __forceinline__ void tile_prologue(const Params& P, TileTab* tb, StageSlow& slow, Tap (*slowtap)[NDIRS],
                                              int budget_floats, int stages_needed, TileCtx<NDIRS, PPT, SLOTS>& cx) {
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cx.j = blockIdx.x * TL_TW + (warp & 3) * 8 + (lane & 7);  // a warp covers an 8x4 patch (fewer bank conflicts than 32x1)
  if (G.T == 1) {
    cx.n = blockIdx.z;
    cx.t = 0;
  } else {
    cx.n = blockIdx.z / G.T;
    cx.t = blockIdx.z - cx.n * G.T;
  }
  constexpr int WY = NTHR / 128;
  const int i0 = blockIdx.y * (4 * WY * PPT);
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    cx.irow[q] = i0 + ((warp >> 2) + WY * q) * 4 + (lane >> 3);
    cx.inimg[q] = cx.j < G.W && cx.irow[q] < G.H;
  }
  // all global loads of the prologue first (flow x/y, gate, blend weight of every (pixel, direction)): ONE exposed
  // memory latency instead of one per (q, d) (the votes / atomics below keep the compiler from hoisting them itself).
  // Out-of-image pixels of ragged tiles read the clamped in-image address and are masked afterwards.
  float lfx[PPT][NDIRS], lfy[PPT][NDIRS], lgt[PPT][NDIRS], lbl[PPT][NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    const DirP& D = P.dir[d];
    const DirAt at = dir_at(D, cx.n, cx.t);
    const int fsh = (int)D.flow_sh, gsh = (int)D.gate_sh, bsh = (int)D.blend_sh;
    const int jc = min(cx.j, G.W - 1);
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int ic = min(cx.irow[q], G.H - 1);
      const int of = ic * fsh + jc;
      lfx[q][d] = __ldg(at.flow + of);
      lfy[q][d] = __ldg(at.flow + D.flow_sc + of);
      lgt[q][d] = at.gate ? __ldg(at.gate + (ic * gsh + jc)) : 1.0f;
      lbl[q][d] = at.blend ? __ldg(at.blend + (ic * bsh + jc)) : 1.0f;
    }
  }
  int x0[PPT][NDIRS], y0[PPT][NDIRS];
  unsigned vld[PPT][NDIRS];
  for (int k = threadIdx.x; k < NDIRS * TL_ROWS_P; k += NTHR) {
    TileTab& T = tb[k >= TL_ROWS_P ? 1 : 0];
    const int r = k >= TL_ROWS_P ? k - TL_ROWS_P : k;
    T.xlo[r] = 0x7fffffff;
    T.xhi[r] = -0x7fffffff;
  }
  if (threadIdx.x < NDIRS) tb[threadIdx.x].akey32 = 0xffffffffu;
  if (threadIdx.x == 0) slow.n = 0;
  __syncthreads();
  const float bx = base_coord(cx.j, G.W, G.stepx);
  float by[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) by[q] = base_coord(cx.irow[q], G.H, G.stepy);
  const float fW = (float)G.W, fW1 = (float)(G.W - 1), fH = (float)G.H, fH1 = (float)(G.H - 1);
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    const DirP& D = P.dir[d];
    const bool gated = D.gate != nullptr;
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      TilePix& px = cx.px[q][d];
      px.tx = px.ty = px.ux = px.uy = px.bl = 0.f;
      px.clip = 0u;
      x0[q][d] = y0[q][d] = 0;
      vld[q][d] = 0u;
      if (cx.inimg[q]) {
        float fx = lfx[q][d], fy = lfy[q][d];
        if (gated) {
          fx = __fmul_rn(fx, lgt[q][d]);
          fy = __fmul_rn(fy, lgt[q][d]);
        }
        px.bl = lbl[q][d];
        // bx -/+ f as one fma: sign * f is exact, so this is the same single rounding as __fsub_rn / __fadd_rn
        const float gx = __fmaf_rn(D.sign, fx, bx), gy = __fmaf_rn(D.sign, fy, by[q]);
        const float ix = source_index_fast<ALIGN, BORDER>(gx, fW, fW1);
        const float iy = source_index_fast<ALIGN, BORDER>(gy, fH, fH1);
        if (BORDER) {  // clipped <=> the unclipped coordinate was <= 0 or >= size-1 (NaN: not clipped, as in source_index)
          const float cx_ = ALIGN ? __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), fW1) : __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.0f), fW, -1.0f), 0.5f);
          const float cy_ = ALIGN ? __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), fH1) : __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.0f), fH, -1.0f), 0.5f);
          px.clip = (unsigned)(cx_ <= 0.f || cx_ >= fW1) | ((unsigned)(cy_ <= 0.f || cy_ >= fH1) << 1);
        }
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int X = (int)fx0, Y = (int)fy0;
        px.tx = __fsub_rn(ix, fx0);
        px.ty = __fsub_rn(iy, fy0);
        px.ux = __fsub_rn(__fadd_rn(fx0, 1.0f), ix);
        px.uy = __fsub_rn(__fadd_rn(fy0, 1.0f), iy);
        const bool xin0 = (unsigned)X < (unsigned)G.W, xin1 = (unsigned)(X + 1) < (unsigned)G.W;
        const bool yin0 = (unsigned)Y < (unsigned)G.H, yin1 = (unsigned)(Y + 1) < (unsigned)G.H;
        vld[q][d] = (unsigned)(xin0 && yin0) | ((unsigned)(xin1 && yin0) << 1) | ((unsigned)(xin0 && yin1) << 2) |
                    ((unsigned)(xin1 && yin1) << 3);
        x0[q][d] = X;
        y0[q][d] = Y;
      }
      // anchor vote: the displacement (x0 - j, y0 - i) of the tile's least displaced pixel, as one 32-bit key
      // |dx|+|dy| (10 bits) | dx (11 bits) | dy (11 bits); beyond +-1023 the anchor is clamped (the tile then goes slow / generic)
      const int dx = min(max(x0[q][d] - cx.j, -1024), 1023), dy = min(max(y0[q][d] - cx.irow[q], -1024), 1023);
      const unsigned mag = (unsigned)min(abs(dx) + abs(dy), 1023);
      const unsigned key = vld[q][d] ? ((mag << 22) | (((unsigned)dx & 0x7ffu) << 11) | ((unsigned)dy & 0x7ffu)) : 0xffffffffu;
      const unsigned best = __reduce_min_sync(0xffffffffu, key);
      if (lane == 0 && best != 0xffffffffu) atomicMin(&tb[d].akey32, best);
    }
  }
  __syncthreads();
  int adx[NDIRS], ady[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    const unsigned k = tb[d].akey32;  // ~0 when no pixel of the tile has a tap: adx = ady = -1, nobody is tested
    adx[d] = ((int)(k << 10)) >> 21;
    ady[d] = ((int)(k << 21)) >> 21;
  }
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    bool far = false;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d)
      far |= vld[q][d] != 0u && (abs(x0[q][d] - cx.j - adx[d]) > TL_R || abs(y0[q][d] - cx.irow[q] - ady[d]) > TL_R);
    cx.act[q] = cx.inimg[q] && !far;
    if (far) {
      const int slot = atomicAdd(&slow.n, 1);
      if (slot < TL_MAXSLOW) {
        slow.pix[slot] = (unsigned short)(((cx.irow[q] - i0) << 5) | (cx.j - blockIdx.x * TL_TW));
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) {
          Tap& k = slowtap[slot][d];
          k.x0 = x0[q][d];
          k.y0 = y0[q][d];
          k.valid = vld[q][d];
          k.tx = cx.px[q][d].tx;
          k.ty = cx.px[q][d].ty;
          k.ux = cx.px[q][d].ux;
          k.uy = cx.px[q][d].uy;
          k.blend = cx.px[q][d].bl;
        }
      }
    }
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      if (!cx.act[q]) vld[q][d] = 0u;
      if (vld[q][d]) {
        TileTab& T = tb[d];
        const int r = y0[q][d] - (i0 + ady[d] - TL_R);  // 0 .. TL_ROWS - 3
        atomicMin(&T.xlo[r], x0[q][d]);
        atomicMax(&T.xhi[r], x0[q][d] + 1);
      }
    }
  }
  __syncthreads();
  cx.ok = slow.n <= TL_MAXSLOW;
  if (!cx.ok) return;  // wild flow (too many pixels far from the anchor): the caller goes generic; skip the tables
  if (warp < NDIRS) tile_tab_scan(tb[warp], 0);
  __syncthreads();
  const int tot0 = tb[0].total4;
  cx.total = tot0;
  int alloc = tb[0].alloc4;
  if (NDIRS > 1) {
    cx.total += tb[NDIRS - 1].total4;
    alloc += tb[NDIRS - 1].alloc4;
  }
  cx.stage_f = TL_ZPAD + 4 * alloc;
  if (cx.total > SLOTS * NTHR || stages_needed * cx.stage_f > budget_floats) cx.ok = 0;
  if (!cx.ok) return;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    const TileTab& T = tb[d];
    const int yb = i0 + ady[d] - TL_R;
    const int doff = d == 0 ? 0 : 4 * tb[0].alloc4;  // the second direction's slot follows the first one's
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      cx.px[q][d].o0 = cx.px[q][d].o1 = 0;
      cx.px[q][d].vld = vld[q][d];
      cx.px[q][d].fx1 = vld[q][d] ? (float)(x0[q][d] + 1) : 0.0f;
      cx.px[q][d].fy1 = vld[q][d] ? (float)(y0[q][d] + 1) : 0.0f;
      if (vld[q][d]) {
        const int r = y0[q][d] - yb;
        cx.px[q][d].o0 = doff + T.rowbase[r] + (x0[q][d] - T.rowx[r]);
        cx.px[q][d].o1 = doff + T.rowbase[r + 1] + (x0[q][d] - T.rowx[r + 1]);
      }
    }
  }
  // which pieces this thread copies: piece k = tid + s*256 of the two directions' piece lists back to back
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int k = threadIdx.x + s * NTHR;
    cx.info[s] = 0x8000u;
    cx.pdst[s] = 0;
    if (k < cx.total) {
      const int d = (NDIRS > 1 && k >= tot0) ? 1 : 0;
      const TileTab& T = tb[d];
      const int kk = k - (d ? tot0 : 0);
      int r = 0;  // last row with rowoff4[r] <= kk
#pragma unroll
      for (int step = 64; step > 0; step >>= 1)
        if (T.rowoff4[r + step] <= kk) r += step;
      const int e = kk - T.rowoff4[r];
      const int y = i0 + ady[d] - TL_R + r, col = T.rowx[r] + 4 * e;
      const bool in = y >= 0 && y < G.H && col >= 0 && col < G.W;  // W % 4 == 0: a piece is all in or all out
      cx.info[s] = ((unsigned)d << 31) | (in ? (((unsigned)y << 16) | (unsigned)col) : 0x8000u);
      cx.pdst[s] = (d ? 4 * tb[0].alloc4 : 0) + T.rowbase[r] + 4 * e;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// small PTX helpers: 32-bit shared addresses, predicated copies / stores (no branches in the channel loop)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tl_lds(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float tl_lds4(unsigned a) {  // [a + 4]
  float v;
  asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned long long tl_lds64(unsigned a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
// 16-byte async copy, issued only when `on`; `bytes` (0 or 16) come from memory, the rest is zero-filled
__device__ __forceinline__ void cp_async16_if(unsigned dst, const void* src, int bytes, bool on) {
  asm volatile(
      "{\n .reg .pred pp;\n setp.ne.b32 pp, %3, 0;\n @pp cp.async.cg.shared.global [%0], [%1], 16, %2;\n}\n" ::"r"(dst),
      "l"(src), "r"(bytes), "r"((int)on)
      : "memory");
}
__device__ __forceinline__ void red_shared_add(unsigned a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_shared_add_4(unsigned a, int v) { asm volatile("red.shared.add.s32 [%0+4], %1;" ::"r"(a), "r"(v) : "memory"); }

// Per-thread copy plan of one channel: element offset inside a plane of each piece (or a flag)
constexpr int PIECE_ZERO = -1, PIECE_NONE = -2;

// ---------------------------------------------------------------------------------------------
// Kernel 1: fused forward warp (+gate) (+blend), NDIRS directions, all channel groups.
// All groups share the row strides (host-checked), so a piece's offset inside a plane is channel independent.
// ---------------------------------------------------------------------------------------------
struct TileChanF {
  const float* src[2];
  float* out;
  float* z[2];  // grad_src plane of this channel per direction to zero-fill (NULL: none), see ZeroP
};

// element offset inside a grad_src plane of the 4 floats thread `tid` clears for direction *zd, or -1
template <int NDIRS>
__device__ __forceinline__ int zero_plan(const Geo& G, const ZeroP& Z, int* zd) {
  const int tid = threadIdx.x;
  *zd = tid >> 7;
  if (!Z.on || *zd >= NDIRS) return -1;
  const int i = blockIdx.y * TL_TH + ((tid >> 3) & 15), j = blockIdx.x * TL_TW + (tid & 7) * 4;
  return (i < G.H && j < G.W) ? i * Z.sh[*zd] + j : -1;
}
#ifndef TL_ZFILL_BULK
#define TL_ZFILL_BULK 0  // A/B: zero-fill by 1-D bulk copies (UBLKCP: shared zeros -> global rows) instead of LSU stores: 0.301 vs 0.305 ms, not worth it
#endif
// bytes (multiple of 16) of zeros from the CTA's shared zero line to global memory through the bulk-copy engine
__device__ __forceinline__ void bulk_zero_row(float* gdst, unsigned zsrc_s, int bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(zsrc_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_zero4_if(float* p, bool on) {
  asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %1, 0;\n @pp st.global.v4.f32 [%0], {%2,%2,%2,%2};\n}\n" ::"l"(p), "r"((int)on), "f"(0.f)
               : "memory");
}

template <int NDIRS, bool ALIGN, bool BORDER>
__global__ void __launch_bounds__(TL_THREADS, TL_FWD_MINCTA) fwd_tile_kernel(const __grid_constant__ Params P, int smem_floats, int Ctot,
                                                                 const __grid_constant__ ZeroP Z) {
  extern __shared__ float4 tl_smem4[];
  float* const smem = reinterpret_cast<float*>(tl_smem4);
  __shared__ StageSlow slow;
  __shared__ Tap slowtap[TL_MAXSLOW][NDIRS];
  __shared__ __align__(16) TileChanF tab[TL_MAXCH];
  __shared__ __align__(128) float zline[TL_TW];  // 128 bytes of zeros: source of the bulk zero-fill
  if (TL_ZFILL_BULK && Z.on && threadIdx.x < TL_TW) {
    zline[threadIdx.x] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // visible to the bulk-copy engine (barriers follow in the prologue)
  }
  int n, t, j, stage_f;
  bool act[TL_PPT];
  float w[TL_PPT][NDIRS][4], bl[TL_PPT][NDIRS];
  unsigned a0[TL_PPT][NDIRS], a1[TL_PPT][NDIRS];  // shared byte addresses (stage 0) of the nw and sw taps
  int poff[TL_SLOTS];                             // piece tid + s*256: element offset in its plane / PIECE_*
  unsigned tsel[TL_SLOTS];                        // byte offset of the piece's direction inside a table entry
  unsigned pd[TL_SLOTS];                          // shared byte address (stage 0) the piece is copied to
  int ooff[TL_PPT];
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
  {
    TileCtx<NDIRS, TL_PPT, TL_SLOTS> cx;
    tile_prologue<NDIRS, ALIGN, BORDER, TL_PPT, TL_SLOTS>(P, reinterpret_cast<TileTab*>(smem), slow, slowtap, smem_floats, TL_FD, cx);
    n = cx.n, t = cx.t, j = cx.j;
    if (!cx.ok) {  // wild flow: the tile's source footprint does not fit -> gather from global memory
#pragma unroll
      for (int q = 0; q < TL_PPT; ++q)
        if (cx.inimg[q]) fwd_generic_pixel<NDIRS>(P, n, t, cx.irow[q], j);
      int zd;
      const int zo = zero_plan<NDIRS>(P.geo, Z, &zd);
      if (zo >= 0)
        for (int g = 0; g < P.geo.n_groups; ++g)
          for (int c = 0; c < P.grp[g].C; ++c) {
            float* zp = zero_plane(Z, P.geo, g, zd, n, t, c);
            if (zp) st_zero4_if(zp + zo, true);
          }
      return;
    }
    stage_f = cx.stage_f;
#pragma unroll
    for (int q = 0; q < TL_PPT; ++q) {
      act[q] = cx.act[q];
      ooff[q] = cx.irow[q] * P.grp[0].out_sh + j;
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const TilePix& px = cx.px[q][d];
        w[q][d][0] = __fmul_rn(px.ux, px.uy);
        w[q][d][1] = __fmul_rn(px.tx, px.uy);
        w[q][d][2] = __fmul_rn(px.ux, px.ty);
        w[q][d][3] = __fmul_rn(px.tx, px.ty);
        bl[q][d] = px.bl;
        a0[q][d] = smem_s + 4u * (unsigned)px.o0;
        a1[q][d] = smem_s + 4u * (unsigned)px.o1;
      }
    }
#pragma unroll
    for (int s = 0; s < TL_SLOTS; ++s) {
      const unsigned info = cx.info[s];
      const int d = piece_dir(info);
      tsel[s] = 8u * (unsigned)d;
      pd[s] = smem_s + 4u * (unsigned)cx.pdst[s];
      poff[s] = threadIdx.x + s * TL_THREADS >= cx.total ? PIECE_NONE
                : piece_zero(info)                       ? PIECE_ZERO
                                                         : piece_y(info) * P.grp[0].src_sh[d] + piece_col(info);
    }
  }
  if (threadIdx.x < Ctot) {
    int g, c;
    chan_lookup(P, threadIdx.x, g, c);
    const GroupP& R = P.grp[g];
    TileChanF e;
#pragma unroll
    for (int d = 0; d < 2; ++d) e.src[d] = R.src[d] + n * R.src_sn[d] + t * R.src_st[d] + (long long)c * R.src_sc[d];
    e.out = R.out + n * R.out_sn + t * R.out_st + (long long)c * R.out_sc;
#pragma unroll
    for (int d = 0; d < 2; ++d) e.z[d] = (Z.on && d < NDIRS) ? zero_plane(Z, P.geo, g, d, n, t, c) : nullptr;
    tab[threadIdx.x] = e;
  }
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) has_bl[d] = P.dir[d].blend != nullptr;
  int zdir;
  const int zoff = zero_plan<NDIRS>(P.geo, Z, &zdir);
  const unsigned zsel = 8u + 8u * (unsigned)zdir;  // &tab[cf].z[zdir] relative to &tab[cf].out
  __syncthreads();  // the tables in dynamic shared memory are dead from here on; tab / slowtap are visible
  if (threadIdx.x < TL_FD * TL_ZPAD) smem[(threadIdx.x / TL_ZPAD) * stage_f + (threadIdx.x % TL_ZPAD)] = 0.f;
  const unsigned stage_b = 4u * (unsigned)stage_f;
  const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);

  auto issue = [&](int cf, unsigned soff) {  // copies of channel cf into the stage at byte offset soff
    const bool live = cf < Ctot;
    const unsigned te = tab_s + (unsigned)sizeof(TileChanF) * (unsigned)(live ? cf : 0);
#pragma unroll
    for (int s = 0; s < TL_SLOTS; ++s) {
      const float* base = reinterpret_cast<const float*>(tl_lds64(te + tsel[s]));
      cp_async16_if(pd[s] + soff, base + max(poff[s], 0), poff[s] >= 0 ? 16 : 0, live && poff[s] != PIECE_NONE);
    }
    cp_async_commit();
  };
  {
    unsigned so = 0;
#pragma unroll
    for (int p = 0; p < TL_FD - 1; ++p, so += stage_b) issue(p, so);
  }
  // zero-fill of the grad_src blocks (side job, Z.on): fire-and-forget stores while the first copies are in flight
  if (TL_ZFILL_BULK) {
    if (Z.on) {  // one bulk copy per (plane, row of the tile): 16 rows x 2 directions x Ctot planes spread over the threads
      const unsigned zs = (unsigned)__cvta_generic_to_shared(zline);
      const int j0 = blockIdx.x * TL_TW, i0z = blockIdx.y * TL_TH;
      const int bytes = 4 * min(TL_TW, P.geo.W - j0);
      for (int it = threadIdx.x; it < Ctot * NDIRS * TL_TH; it += TL_THREADS) {
        const int row = it % TL_TH, pl = it / TL_TH, d = pl % NDIRS, cf = pl / NDIRS;
        float* const zp = tab[cf].z[d];
        if (zp != nullptr && i0z + row < P.geo.H) bulk_zero_row(zp + (i0z + row) * Z.sh[d] + j0, zs, bytes);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  } else if (zoff >= 0) {
    unsigned tz = tab_s + 16u + zsel;
#pragma unroll 4
    for (int cf = 0; cf < Ctot; ++cf, tz += (unsigned)sizeof(TileChanF)) {
      float* const zp = reinterpret_cast<float*>(tl_lds64(tz));
      st_zero4_if(zp + zoff, zp != nullptr);
    }
  }
  // slow pixels, while the first copies are in flight: one (pixel, channel) item per thread
  for (int it = threadIdx.x; it < slow.n * Ctot; it += TL_THREADS) {
    const int sidx = it / Ctot, pix = slow.pix[sidx];
    fwd_slow_item<NDIRS>(P, n, t, blockIdx.y * TL_TH + (pix >> 5), blockIdx.x * TL_TW + (pix & 31), slowtap[sidx], it - sidx * Ctot);
  }
  unsigned soff = 0, poffs = (TL_FD - 1) * stage_b;  // stage of channel cf; stage the copies of channel cf+D-1 go to
  const unsigned ring_b = TL_FD * stage_b;
  unsigned te = tab_s + 16u;  // &tab[cf].out
#pragma unroll 1
  for (int cf = 0; cf < Ctot; ++cf) {
    cp_async_wait<TL_FD - 2>();
    __syncthreads();  // channel cf has landed for every thread; everyone is done with channel cf-1
    issue(cf + TL_FD - 1, poffs);
    float* const op = reinterpret_cast<float*>(tl_lds64(te));
#pragma unroll
    for (int q = 0; q < TL_PPT; ++q) {
      float r = 0.f;
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const unsigned s0 = a0[q][d] + soff, s1 = a1[q][d] + soff;
        float a = __fmul_rn(tl_lds(s0), w[q][d][0]);
        a = __fmaf_rn(tl_lds4(s0), w[q][d][1], a);
        a = __fmaf_rn(tl_lds(s1), w[q][d][2], a);
        a = __fmaf_rn(tl_lds4(s1), w[q][d][3], a);
        if (has_bl[d]) a = __fmul_rn(a, bl[q][d]);
        r = (d == 0) ? a : __fadd_rn(r, a);
      }
      st_cs_if(op + ooff[q], r, act[q]);
    }
    soff += stage_b;
    if (soff == ring_b) soff = 0;
    poffs += stage_b;
    if (poffs == ring_b) poffs = 0;
    te += (unsigned)sizeof(TileChanF);
  }
  // the bulk copies read this CTA's shared memory: they must have done so before the CTA exits
  if (TL_ZFILL_BULK && Z.on) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Kernels 2 + 3 fused (non-deterministic fast backward).  Row strides shared by all groups (host-checked).
// ---------------------------------------------------------------------------------------------
struct TileChanB {
  const float* src[2];
  float* gs[2];  // may be NULL
  const float* go;
  int g, c;  // group / channel (the non-finite path needs them)
  unsigned long long tex[2];  // texture path: object holding this channel's plane, per direction
  float row[2];               //               first texture row of the plane
  int pad_[2];
};

// inf / NaN in grad_out of this channel: exact float atomics straight to global memory (rare path, kept out of line)
__device__ __noinline__ void bwd_nonfinite_px(const Params& P, const GradP& Q, int n, int t, int i, int j, int g, int c, int d, float gw) {
  Tap k;
  compute_tap(P.geo, P.dir[d], n, t, i, j, k);
  scatter_atomic_px(Q, g, d, n, t, c, k, gw);
}

// SCATTER = false: no grad_src is wanted (the sources are data): kernel 2 only on the staged tiles - no accumulators, no scale
// vote, no flush; fewer registers and 4 instead of 6 shared-memory units per CTA, hence 3 CTAs per SM.
// TEX = true: the source taps come from the texture units (fwb_tex.cuh: one TLD4 per (pixel, direction, channel), fetched one
// channel ahead) instead of a staged copy of the footprint; shared memory then only holds the two scatter accumulators and the
// shared-memory pipe only carries the scatter.  The two pipes run side by side (tools/mb_tex.cu: TLD4 alone 18.5, 4 ATOMS alone
// 19.0, both together 20.5 cycles per warp, direction and channel).
template <int NDIRS, bool ALIGN, bool BORDER, int PPT, int SLOTS, int NTHR = TL_THREADS, bool SCATTER = true, bool TEX = false>
__global__ void __launch_bounds__(NTHR, NTHR == 512 ? 2 : ((PPT == 1 || !SCATTER) ? 3 : (TEX ? TL_BWDX_MINCTA : TL_BWD_MINCTA))) bwd_tile_kernel(const __grid_constant__ Params P, const __grid_constant__ GradP Q,
                                                                 int smem_floats, const __grid_constant__ TexP X) {
  extern __shared__ float4 tl_smem4[];
  float* const smem = reinterpret_cast<float*>(tl_smem4);
  __shared__ StageSlow slow;
  __shared__ Tap slowtap[TL_MAXSLOW][NDIRS];
  __shared__ __align__(16) TileChanB tab[TL_MAXCH];
  __shared__ unsigned amax_s[3];
  __shared__ int nchan_s, g0_s;
  __shared__ float slowacc[TL_MAXSLOW][NDIRS][3];  // gix, giy, grad_blend of the slow pixels, summed over the channels
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int n, t, j, total, stage_f;
  int irow[PPT];
  bool act[PPT];
  float tx[PPT][NDIRS], ty[PPT][NDIRS], ux[PPT][NDIRS], uy[PPT][NDIRS], bl[PPT][NDIRS];
  unsigned a0[PPT][NDIRS], a1[PPT][NDIRS];  // byte offsets inside a stage of the nw and sw taps
  unsigned info[SLOTS];
  unsigned pd[SLOTS];  // byte offset inside a stage of piece tid + s*256
  unsigned clipbits = 0u;  // 2 bits per (q, d): the coordinate gradient is zero (border clipping)
  float fx1[PPT][NDIRS], fy1[PPT][NDIRS];  // texture path: quad coordinates
  unsigned vbits = 0u;                      //               4 validity bits per (q, d)
  // grad_out of the first channel: issued before the prologue so that its latency hides behind it (the scale vote of
  // channel 0 is the first thing the pipeline needs)
  float ego[PPT];
  {
    int g0e = 0;
    while (g0e < G.n_groups - 1 && !Q.grad_out[g0e]) ++g0e;
    const int nt = blockIdx.z, ne = G.T == 1 ? nt : nt / G.T, te = nt - ne * G.T;
    const int je = blockIdx.x * TL_TW + (warp & 3) * 8 + (lane & 7);
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int ie = blockIdx.y * (4 * (NTHR / 128) * PPT) + ((warp >> 2) + (NTHR / 128) * q) * 4 + (lane >> 3);
      ego[q] = (Q.grad_out[g0e] && je < G.W && ie < G.H)
                   ? __ldcs(Q.grad_out[g0e] + ne * Q.go_sn[g0e] + te * Q.go_st[g0e] + (long long)ie * Q.go_sh[g0e] + je)
                   : 0.f;
    }
  }
  {
    TileCtx<NDIRS, PPT, SLOTS> cx;
    tile_prologue<NDIRS, ALIGN, BORDER, PPT, SLOTS, NTHR>(P, reinterpret_cast<TileTab*>(smem), slow, slowtap, smem_floats,
                                                          (TEX ? 0 : TL_BD) + (SCATTER ? 2 : 0), cx);
    n = cx.n, t = cx.t, j = cx.j;
    if (!cx.ok) {
#pragma unroll
      for (int q = 0; q < PPT; ++q)
        if (cx.inimg[q]) bwd_fused_generic_pixel<NDIRS>(P, Q, n, t, cx.irow[q], j);
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
        tx[q][d] = px.tx, ty[q][d] = px.ty, ux[q][d] = px.ux, uy[q][d] = px.uy, bl[q][d] = px.bl;
        clipbits |= px.clip << (2 * (q * NDIRS + d));
        a0[q][d] = 4u * (unsigned)px.o0;
        a1[q][d] = 4u * (unsigned)px.o1;
        fx1[q][d] = px.fx1, fy1[q][d] = px.fy1;
        vbits |= px.vld << (4 * (q * NDIRS + d));
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
        TileChanB e;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          e.src[d] = R.src[d] + n * R.src_sn[d] + t * R.src_st[d] + (long long)c * R.src_sc[d];
          float* gs = (d < NDIRS) ? Q.grad_src[g][d] : nullptr;
          e.gs[d] = gs ? gs + n * Q.gs_sn[g][d] + t * Q.gs_st[g][d] + (long long)c * Q.gs_sc[g][d] : nullptr;
          e.tex[d] = 0ull;
          e.row[d] = 0.0f;
          if (TEX && d < NDIRS) {
            const TexSrc& S = X.s[g][d];
            const int blk = n / S.nb;
            e.tex[d] = S.tex[blk];
            e.row[d] = (float)((n - blk * S.nb) * S.rows_n + t * S.rows_t + c * G.H);
          }
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
  if (threadIdx.x < 3) amax_s[threadIdx.x] = 0u;
  for (int k = threadIdx.x; k < TL_MAXSLOW * NDIRS * 3; k += NTHR) (&slowacc[0][0][0])[k] = 0.f;
  bool has_bl[NDIRS];
  float blmax[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) blmax[q] = 0.f;
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    has_bl[d] = P.dir[d].blend != nullptr;
#pragma unroll
    for (int q = 0; q < PPT; ++q) blmax[q] = fmaxf(blmax[q], has_bl[d] ? fabsf(bl[q][d]) : 1.0f);
  }
  float gix[PPT][NDIRS], giy[PPT][NDIRS], gbl[PPT][NDIRS];
#pragma unroll
  for (int q = 0; q < PPT; ++q)
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) gix[q][d] = giy[q][d] = gbl[q][d] = 0.f;
  __syncthreads();  // the tables in dynamic shared memory are dead from here on; tab / slowtap / amax_s are visible
  const int Cn = nchan_s, g0 = g0_s;
  // per-thread copy / flush plan (row strides are those of group g0: all groups agree)
  int poff[SLOTS], goff[SLOTS];  // element offset of piece tid + s*256 in its src plane / grad_src plane, or PIECE_*
  unsigned tsel[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int d = piece_dir(info[s]);
    tsel[s] = 8u * (unsigned)d;
    const bool none = threadIdx.x + s * NTHR >= total, zero = piece_zero(info[s]);
    poff[s] = none ? PIECE_NONE : zero ? PIECE_ZERO : piece_y(info[s]) * P.grp[g0].src_sh[d] + piece_col(info[s]);
    // no group has a grad_src for this direction <=> group g0 has none (host-checked): nothing to flush
    goff[s] = (!SCATTER || none || zero || Q.grad_src[g0][d] == nullptr) ? PIECE_NONE : piece_y(info[s]) * Q.gs_sh[g0][d] + piece_col(info[s]);
  }
  int gooff[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) gooff[q] = irow[q] * Q.go_sh[g0] + j;
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned stage_b = 4u * (unsigned)stage_f;
  const unsigned acc_s = smem_s + (TEX ? 0u : TL_BD * stage_b);  // two accumulators, stage layout
#pragma unroll
  for (int s = 0; s < SLOTS; ++s)
    if (goff[s] >= 0) {  // the pieces this thread flushes start from zero (the rest of the accumulators is never read)
      asm volatile("st.shared.v4.s32 [%0], {%1,%1,%1,%1};" ::"r"(acc_s + pd[s]), "r"(0) : "memory");
      asm volatile("st.shared.v4.s32 [%0], {%1,%1,%1,%1};" ::"r"(acc_s + stage_b + pd[s]), "r"(0) : "memory");
    }
  if (!TEX && threadIdx.x < TL_BD * TL_ZPAD) smem[(threadIdx.x / TL_ZPAD) * stage_f + (threadIdx.x % TL_ZPAD)] = 0.f;
  const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);

  float4 qc[PPT][NDIRS], qn[PPT][NDIRS];  // texture path: the quads of channel cf / cf + 1
  auto fetch = [&](int cf, float4 (*q)[NDIRS]) {
    const TileChanB& tc = tab[cf < Cn ? cf : 0];
#pragma unroll
    for (int qq = 0; qq < PPT; ++qq)
#pragma unroll
      for (int d = 0; d < NDIRS; ++d)
        q[qq][d] = tex2Dgather<float4>((cudaTextureObject_t)tc.tex[d], fx1[qq][d], tc.row[d] + fy1[qq][d], 0);
  };
  auto issue = [&](int cf, unsigned soff) {
    if (TEX) return;
    const bool live = cf < Cn;
    const unsigned te = tab_s + (unsigned)sizeof(TileChanB) * (unsigned)(live ? cf : 0);
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const float* base = reinterpret_cast<const float*>(tl_lds64(te + tsel[s]));
      cp_async16_if(smem_s + pd[s] + soff, base + max(poff[s], 0), poff[s] >= 0 ? 16 : 0, live && poff[s] != PIECE_NONE);
    }
    cp_async_commit();
  };
  auto load_go = [&](int cf, float* go) {
    const bool live = cf < Cn;
    const float* gp = tab[live ? cf : 0].go;
#pragma unroll
    for (int q = 0; q < PPT; ++q) go[q] = (live && act[q]) ? __ldcs(gp + gooff[q]) : 0.f;
  };
  auto vote_amax = [&](const float* go, int slot) {
    float m = 0.f;
#pragma unroll
    for (int q = 0; q < PPT; ++q)  // NaN must win the max: compare the bit patterns (non-negative floats order like uints)
      m = __uint_as_float(max(__float_as_uint(m), __float_as_uint(fabsf(go[q]) * blmax[q])));
    const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if (lane == 0 && mb != 0u) atomicMax(&amax_s[slot], mb);
  };
  // flush this thread's pieces of the accumulator at byte address `ab` (channel cf) into grad_src
  auto flush = [&](int cf, unsigned ab, float Sinv) {
    const unsigned te = tab_s + (unsigned)sizeof(TileChanB) * (unsigned)cf + 16u;  // &tab[cf].gs[0]
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if (goff[s] < 0) continue;
      const unsigned a = ab + pd[s];
      int4 u;
      asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(a));
      if ((u.x | u.y | u.z | u.w) == 0) continue;
      float* gs = reinterpret_cast<float*>(tl_lds64(te + tsel[s]));
      if (gs) red_add_v4(gs + goff[s], make_float4((float)u.x * Sinv, (float)u.y * Sinv, (float)u.z * Sinv, (float)u.w * Sinv));
      asm volatile("st.shared.v4.s32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0) : "memory");
    }
  };

  {
    unsigned so = 0;
#pragma unroll
    for (int p = 0; p < TL_BD - 1; ++p, so += stage_b) issue(p, so);
  }
  if (TEX && Cn > 0) fetch(0, qc);
  float go[PPT], gn[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) go[q] = (Cn > 0 && act[q]) ? ego[q] : 0.f;  // channel 0 (loaded before the prologue)
  if (SCATTER) vote_amax(go, 0);
  // slow pixels, while the first copies are in flight: one (pixel, channel) item per thread, all loads of an item
  // independent (one exposed memory latency, the eight warps evenly loaded); the coordinate-gradient partial sums over
  // the channels meet in shared memory and are stored after the channel loop
  {
    const int nslow = slow.n;
    const int sh0 = P.grp[g0].src_sh[0], sh1 = P.grp[g0].src_sh[NDIRS - 1];
    for (int it = threadIdx.x; it < nslow * Cn; it += NTHR) {
      const int cf = it / nslow, sidx = it - cf * nslow, pix = slow.pix[sidx];  // lanes = different pixels: no same-address
      const int si = blockIdx.y * (4 * (NTHR / 128) * PPT) + (pix >> 5), sj = blockIdx.x * TL_TW + (pix & 31);
      const TileChanB& tc = tab[cf];
      const float gout = __ldg(tc.go + si * Q.go_sh[g0] + sj);
      float va[NDIRS][4];
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const Tap& k = slowtap[sidx][d];
        const int sh = d == 0 ? sh0 : sh1;
        const float* sp = tc.src[d] + k.y0 * sh + k.x0;
        va[d][0] = ldg_if(sp, k.valid & 1u), va[d][1] = ldg_if(sp + 1, k.valid & 2u);
        va[d][2] = ldg_if(sp + sh, k.valid & 4u), va[d][3] = ldg_if(sp + sh + 1, k.valid & 8u);
      }
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        const Tap& k = slowtap[sidx][d];
        const float a = va[d][0], b = va[d][1], cc = va[d][2], dd = va[d][3];
        float gw = gout;
        if (has_bl[d]) {
          const float top = fmaf(b, k.tx, a * k.ux), bot = fmaf(dd, k.tx, cc * k.ux);
          atomicAdd(&slowacc[sidx][d][2], gout * fmaf(bot, k.ty, top * k.uy));
          gw = gout * k.blend;
        }
        atomicAdd(&slowacc[sidx][d][0], gw * fmaf(k.ty, dd - cc, k.uy * (b - a)));
        atomicAdd(&slowacc[sidx][d][1], gw * fmaf(k.tx, dd - b, k.ux * (cc - a)));
        scatter_atomic_px(Q, tc.g, d, n, t, tc.c, k, gw);
      }
    }
  }
  constexpr float MAGIC = 12582912.0f;  // 1.5 * 2^23: fma(x, y, MAGIC) holds round-to-nearest(x*y) in its low mantissa bits
  constexpr int MAGIC_BITS = 0x4B400000;
  unsigned soff = 0, poffs = (TL_BD - 1) * stage_b;
  const unsigned ring_b = TL_BD * stage_b;
  unsigned aoff = 0;  // byte offset of the accumulator of channel cf (0 / stage_b)
  float Sinv_prev = 0.f;
  bool flush_prev = false;
#pragma unroll 1
  for (int cf = 0; cf < Cn; ++cf) {
    if (!TEX) cp_async_wait<TL_BD - 2>();
    __syncthreads();  // channel cf has landed; scatter of cf-1 complete; flush of cf-2 complete
    issue(cf + TL_BD - 1, poffs);
    if (TEX && cf + 1 < Cn) fetch(cf + 1, qn);  // the quads of the next channel travel while this one is scattered
    load_go(cf + 1, gn);
    if (SCATTER && threadIdx.x == 0) amax_s[(cf + 2) % 3] = 0u;
    if (SCATTER && flush_prev) flush(cf - 1, acc_s + (aoff ^ stage_b), Sinv_prev);
    // ---- scale of this channel
    const unsigned ab = SCATTER ? amax_s[cf % 3] : 0u;
    const bool finite = ab < 0x7f800000u;
    const int sexp = min(252, max(2, 274 - (int)(ab >> 23)));  // biased exponent of 2^(20 - exponent(amax))
    const float S = __uint_as_float((unsigned)sexp << 23);
    const float Sinv = __uint_as_float((unsigned)(254 - sexp) << 23);
    const TileChanB& tc = tab[cf];
    const unsigned sb = smem_s + soff, ac = acc_s + aoff;
    bool want[NDIRS];
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) want[d] = SCATTER && tc.gs[d] != nullptr;
    const float Sf = (finite && ab != 0u) ? S : 0.f;  // 0: nothing goes to the accumulator
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        float a, b, c_, dd;
        if (TEX) {  // taps outside the image (and every tap of a slow / out-of-tile pixel) count as zero
          const unsigned v = vbits >> (4 * (q * NDIRS + d));
          a = (v & 1u) ? qc[q][d].w : 0.f, b = (v & 2u) ? qc[q][d].z : 0.f;
          c_ = (v & 4u) ? qc[q][d].x : 0.f, dd = (v & 8u) ? qc[q][d].y : 0.f;
        } else {
          a = tl_lds(sb + a0[q][d]), b = tl_lds4(sb + a0[q][d]);
          c_ = tl_lds(sb + a1[q][d]), dd = tl_lds4(sb + a1[q][d]);
        }
        float gw = go[q];
        if (has_bl[d]) {
          const float top = fmaf(b, tx[q][d], a * ux[q][d]), bot = fmaf(dd, tx[q][d], c_ * ux[q][d]);
          gbl[q][d] = fmaf(go[q], fmaf(bot, ty[q][d], top * uy[q][d]), gbl[q][d]);
          gw *= bl[q][d];
        }
        gix[q][d] = fmaf(gw, fmaf(ty[q][d], dd - c_, uy[q][d] * (b - a)), gix[q][d]);
        giy[q][d] = fmaf(gw, fmaf(tx[q][d], dd - b, ux[q][d] * (c_ - a)), giy[q][d]);
        if (want[d]) {
          const float sgw = (finite ? gw : 0.f) * Sf;
          const float gl = sgw * ux[q][d], gr = sgw * tx[q][d];
          red_shared_add(ac + a0[q][d], __float_as_int(fmaf(gl, uy[q][d], MAGIC)) - MAGIC_BITS);
          red_shared_add_4(ac + a0[q][d], __float_as_int(fmaf(gr, uy[q][d], MAGIC)) - MAGIC_BITS);
          red_shared_add(ac + a1[q][d], __float_as_int(fmaf(gl, ty[q][d], MAGIC)) - MAGIC_BITS);
          red_shared_add_4(ac + a1[q][d], __float_as_int(fmaf(gr, ty[q][d], MAGIC)) - MAGIC_BITS);
        }
      }
    }
    if (!finite) {  // inf / NaN in this channel's grad_out (rare): float atomics straight to global memory
      for (int q = 0; q < PPT; ++q)
        for (int d = 0; d < NDIRS; ++d)
          if (tc.gs[d] != nullptr && act[q]) bwd_nonfinite_px(P, Q, n, t, irow[q], j, tc.g, tc.c, d, has_bl[d] ? go[q] * bl[q][d] : go[q]);
    }
    if (SCATTER) vote_amax(gn, (cf + 1) % 3);
#pragma unroll
    for (int q = 0; q < PPT; ++q) go[q] = gn[q];
    if (TEX) {
#pragma unroll
      for (int q = 0; q < PPT; ++q)
#pragma unroll
        for (int d = 0; d < NDIRS; ++d) qc[q][d] = qn[q][d];
    }
    Sinv_prev = Sinv;
    flush_prev = finite && ab != 0u;
    soff += stage_b;
    if (soff == ring_b) soff = 0;
    poffs += stage_b;
    if (poffs == ring_b) poffs = 0;
    aoff ^= stage_b;
  }
  __syncthreads();
  if (flush_prev) flush(Cn - 1, acc_s + (aoff ^ stage_b), Sinv_prev);
  if (threadIdx.x < slow.n * NDIRS) {  // coordinate gradients of the slow pixels
    const int sidx = threadIdx.x / NDIRS, d = threadIdx.x - sidx * NDIRS, pix = slow.pix[sidx];
    const int si = blockIdx.y * (4 * (NTHR / 128) * PPT) + (pix >> 5), sj = blockIdx.x * TL_TW + (pix & 31);
    Tap k;  // the full tap (gradient multipliers, raw flow, gate): only needed here
    compute_tap(G, P.dir[d], n, t, si, sj, k);
    bwdflow_store(P, Q, d, n, t, si, sj, k, slowacc[sidx][d][0], slowacc[sidx][d][1], slowacc[sidx][d][2]);
  }

  // epilogue: coordinate gradient -> grad_flow / grad_gate / grad_blend.  The multipliers are d(ix)/d(gx) = W/2 or (W-1)/2,
  // zero where border padding clipped the coordinate; the raw flow and the gate are reloaded only when there is a gate.
  const float mxc = ALIGN ? __fmul_rn((float)(G.W - 1), 0.5f) : __fmul_rn((float)G.W, 0.5f);
  const float myc = ALIGN ? __fmul_rn((float)(G.H - 1), 0.5f) : __fmul_rn((float)G.H, 0.5f);
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    if (!act[q]) continue;
#pragma unroll
    for (int d = 0; d < NDIRS; ++d) {
      const unsigned cb = clipbits >> (2 * (q * NDIRS + d));
      Tap k;
      k.mx = (cb & 1u) ? 0.f : mxc;
      k.my = (cb & 2u) ? 0.f : myc;
      k.fx = k.fy = 0.f;
      k.gate = 1.f;
      const DirP& D = P.dir[d];
      if (D.gate != nullptr) {
        const DirAt at = dir_at(D, n, t);
        const int of = irow[q] * (int)D.flow_sh + j;
        k.fx = __ldg(at.flow + of);
        k.fy = __ldg(at.flow + D.flow_sc + of);
        k.gate = __ldg(at.gate + (irow[q] * (int)D.gate_sh + j));
      }
      bwdflow_store(P, Q, d, n, t, irow[q], j, k, gix[q][d], giy[q][d], gbl[q][d]);
    }
  }
}

}  // namespace fwb

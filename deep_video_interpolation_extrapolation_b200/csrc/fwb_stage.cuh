// fwb_stage.cuh — the ROW-SEGMENT TABLE: which cells of a plane a tile of work touches, and who loads them.
//
// The per-pixel gather of the warp (4 taps x C channels x 2 directions = 184 loads per pixel in the headline
// config) is bound by L1 wavefronts when it goes to global memory: with a rough flow the 32 taps of one warp
// instruction land on 7-11 different 128 B lines (ncu: l1tex 79 % busy, 0.30 ms for 0.10 ms of HBM traffic).
// The staged kernels (fwb_pair.cuh for kernels 1 and 2, fwb_csr.cuh for kernel 3) therefore bring the cells a
// CTA needs into shared memory first.  What they need is described by a row-segment table: for every plane row
// the [xlo, xhi] column range, 16 B aligned — the deformed image of the tile, not its bounding box (a bounding
// box over-reads 3.5x with the headline flows, the row segments 1.3-1.4x).  The segments are laid out back to
// back in a compact slot; which 16-byte pieces a thread loads is channel independent, so it is computed once
// and kept in registers.  Pixels whose taps are far from where the rest of the tile samples (border-clamped
// or wild ones) are classified SLOW against the tile's anchor displacement and served from global memory.
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

}  // namespace fwb

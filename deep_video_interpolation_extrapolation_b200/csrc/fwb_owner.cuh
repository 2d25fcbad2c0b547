// fwb_owner.cuh — kernel 3 (gradient w.r.t. the sources) as an OWNER GATHER, plus the segment tables it reads.
//
// ATen's grid_sampler_2d_backward scatters w*gOut with one global atomicAdd per tap per channel
// (184 atomics per pixel for the headline config; measured 0.35 T adds/s on B200 = 1.1 ms).  Here each CTA
// OWNS a 32x16 tile of grad_src and pulls in every contribution that lands in it:
//   * kernel 2 leaves, per 8x4 output micro-tile (one warp), the bounding box of its in-image taps
//     (8 B record) and appends pixels that stray from their micro-tile ("outliers", e.g. the border-clipped
//     ones) to a per-image list, plus the min/max micro-tile displacement per image;
//   * an owner CTA scans the records whose displacement window can reach its tile, and for the hits
//     recomputes the taps and accumulates w*gOut into shared memory;
//   * accumulation is 32-bit FIXED POINT with native shared-memory integer atomics (ATOMS.ADD; fp32
//     shared atomicAdd is a CAS loop on sm_100a): integer addition is associative, so the result is
//     bit-exact run to run no matter how warps interleave — the deterministic mode costs nothing.
//     The scale is a power of two chosen per (tile, channel) from max|gOut| over the contributing pixels
//     and the largest fan-in of the tile, so the sum cannot overflow and the quantum is <= 2^-29 of it.
//   * the tile is written once with plain coalesced stores: no memset, no global atomics, algorithmic
//     DRAM traffic (8C B/pixel written once).
#pragma once
#include "fwb_coords.cuh"

namespace fwb {

constexpr int MT_W = 8, MT_H = 4;       // micro-tile = the 32 output pixels of one warp
constexpr int OUTLIER_R = 12;           // a pixel farther than this (px) from its micro-tile's anchor displacement is an outlier
#ifndef FWB_OT_W
#define FWB_OT_W 32
#endif
#ifndef FWB_OT_H
#define FWB_OT_H 32
#endif
constexpr int OT_W = FWB_OT_W, OT_H = FWB_OT_H;  // owner tile (source pixels)
constexpr int OT_PITCH = OT_W + 8;      // == 8 (mod 32): the 4 rows of an undistorted 8x4 micro-tile hit 32 distinct banks
constexpr int OT_WORDS = OT_PITCH * OT_H;
constexpr int OWN_THREADS = 512;
constexpr int HITCAP = 1024;            // hit-list chunk
constexpr int OWN_CMAX = 40;            // channels per launch (shared-memory accumulators: OT_WORDS*4 B per channel)
constexpr int HEADROOM0 = 6;            // fixed-point integer bits reserved for the fan-in (2^6 contributions per pixel)

struct WsHeader {  // one per (direction, n*T+t)
  int dxmin, dxmax, dymin, dymax;  // range of the micro-tile anchor displacements (pixels)
  int n_outliers;
  int pad[3];
};

struct WsView {
  WsHeader* hdr;        // [D][NT]
  short4* tab;          // [D][NT][mth][mtw]   {xmin, xmax, ymin, ymax} of the in-image taps of inlier pixels
  int2* outl;           // [D][NT][cap]        {i*W+j, (y0<<16)|(x0&0xffff)}
  unsigned* gmax;       // [D][NT][ctot]       max |gOut*blend| as float bits, per image and channel (kernel 2)
  int mtw, mth;
  long long cap;        // H*W
  int NT, ctot;
};

struct WsLayout {
  size_t hdr_off, tab_off, outl_off, gmax_off, total;
  int mtw, mth;
};

static inline WsLayout ws_layout(int D, long long NT, int H, int W, int ctot) {
  WsLayout L;
  L.mtw = (W + MT_W - 1) / MT_W;
  L.mth = (H + MT_H - 1) / MT_H;
  size_t o = 0;
  L.hdr_off = o;
  o += sizeof(WsHeader) * (size_t)D * NT;
  o = (o + 255) & ~(size_t)255;
  L.gmax_off = o;
  o += sizeof(unsigned) * (size_t)D * NT * ctot;
  o = (o + 255) & ~(size_t)255;
  L.tab_off = o;
  o += sizeof(short4) * (size_t)D * NT * L.mtw * L.mth;
  o = (o + 255) & ~(size_t)255;
  L.outl_off = o;
  o += sizeof(int2) * (size_t)D * NT * H * W;
  L.total = (o + 255) & ~(size_t)255;
  return L;
}

static inline WsView ws_view(void* base, const WsLayout& L, int NT, int H, int W, int ctot) {
  WsView v;
  char* b = (char*)base;
  v.hdr = (WsHeader*)(b + L.hdr_off);
  v.tab = (short4*)(b + L.tab_off);
  v.outl = (int2*)(b + L.outl_off);
  v.gmax = (unsigned*)(b + L.gmax_off);
  v.mtw = L.mtw;
  v.mth = L.mth;
  v.cap = (long long)H * W;
  v.NT = NT;
  v.ctot = ctot;
  return v;
}

// header + gmax are contiguous at the start of the workspace: one init kernel
__global__ void ws_init_kernel(WsHeader* hdr, int n, unsigned* gmax, int ng) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    hdr[k].dxmin = hdr[k].dymin = 0x7fffffff;
    hdr[k].dxmax = hdr[k].dymax = -0x7fffffff;
    hdr[k].n_outliers = 0;
  }
  if (k < ng) gmax[k] = 0u;
}

// Whole-warp classification of one micro-tile.  `has` = this lane's pixel is inside the image and has at
// least one in-image tap.  The ANCHOR is the displacement of the lane with the smallest |dx|+|dy| (ties: lowest
// lane) — robust against the few pixels that border clipping or a wild flow throws far away.  Returns whether
// the lane is an INLIER (within OUTLIER_R of the anchor).  Same code in the emit kernel and in kernel 3.
__device__ __forceinline__ bool mt_classify(bool has, int dx, int dy, int& adx, int& ady, unsigned& has_mask) {
  has_mask = __ballot_sync(0xffffffffu, has);
  if (has_mask == 0u) {
    adx = ady = 0;
    return false;
  }
  const unsigned mag = (unsigned)min(abs(dx) + abs(dy), 0x3ffffff);
  const unsigned key = has ? ((mag << 5) | (threadIdx.x & 31u)) : 0xffffffffu;
  const int src = (int)(__reduce_min_sync(0xffffffffu, key) & 31u);
  adx = __shfl_sync(0xffffffffu, dx, src);
  ady = __shfl_sync(0xffffffffu, dy, src);
  return has && abs(dx - adx) <= OUTLIER_R && abs(dy - ady) <= OUTLIER_R;
}

// Emit side: write the micro-tile record, append outliers, fold the anchor displacement into the CTA range.
// Must be called by all 32 lanes of a warp whose lanes are the 8x4 pixels of micro-tile (mx, my).
__device__ __forceinline__ void mt_emit(const Geo& G, const WsView& ws, int d, int nt, int mx, int my, int i, int j,
                                        bool active, int x0, int y0, unsigned valid, int* s_range /*[4] smem*/) {
  const bool has = active && valid != 0u;
  int adx, ady;
  unsigned hm;
  const bool inl = mt_classify(has, x0 - j, y0 - i, adx, ady, hm);
  int xlo = 32767, xhi = -1, ylo = 32767, yhi = -1;
  if (inl) {
    xlo = max(x0, 0);
    xhi = min(x0 + 1, G.W - 1);
    ylo = max(y0, 0);
    yhi = min(y0 + 1, G.H - 1);
  }
  xlo = __reduce_min_sync(0xffffffffu, xlo);
  ylo = __reduce_min_sync(0xffffffffu, ylo);
  xhi = __reduce_max_sync(0xffffffffu, xhi);
  yhi = __reduce_max_sync(0xffffffffu, yhi);
  const unsigned om = __ballot_sync(0xffffffffu, has && !inl);
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) {
    if (mx < ws.mtw && my < ws.mth)
      ws.tab[(((size_t)d * ws.NT + nt) * ws.mth + my) * ws.mtw + mx] =
          make_short4((short)xlo, (short)xhi, (short)ylo, (short)yhi);
    if (hm) {
      atomicMin(&s_range[0], adx);
      atomicMax(&s_range[1], adx);
      atomicMin(&s_range[2], ady);
      atomicMax(&s_range[3], ady);
    }
    if (om) base = atomicAdd(&ws.hdr[(size_t)d * ws.NT + nt].n_outliers, __popc(om));
  }
  if (om) {
    base = __shfl_sync(0xffffffffu, base, 0);
    if (has && !inl) {
      const int slot = base + __popc(om & ((1u << lane) - 1u));
      ws.outl[((size_t)d * ws.NT + nt) * ws.cap + slot] = make_int2(i * G.W + j, (y0 << 16) | (x0 & 0xffff));
    }
  }
}

// Segment tables for kernel 3: one pass over flows/masks only (24 B/pixel), 8x4 micro-tile per warp.
template <int NDIRS>
__global__ void __launch_bounds__(256) emit_kernel(const __grid_constant__ Params P, const WsView ws) {
  const Geo& G = P.geo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mx = blockIdx.x * 4 + (warp & 3), my = blockIdx.y * 2 + (warp >> 2);
  const int j = mx * MT_W + (lane & 7), i = my * MT_H + (lane >> 3);
  const bool active = j < G.W && i < G.H;
  const int n = blockIdx.z / G.T, t = blockIdx.z - n * G.T;
  __shared__ int s_range[2][4];
  if (threadIdx.x < 8) s_range[threadIdx.x >> 2][threadIdx.x & 3] = (threadIdx.x & 1) ? -0x7fffffff : 0x7fffffff;
  __syncthreads();
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    k.valid = 0u;
    k.x0 = k.y0 = 0;
    if (active) compute_tap(G, P.dir[d], n, t, i, j, k);
    mt_emit(G, ws, d, blockIdx.z, mx, my, i, j, active, k.x0, k.y0, k.valid, s_range[d]);
  }
  __syncthreads();
  if (threadIdx.x < NDIRS && s_range[threadIdx.x][0] <= s_range[threadIdx.x][1]) {
    WsHeader* hd = ws.hdr + (size_t)threadIdx.x * ws.NT + blockIdx.z;
    atomicMin(&hd->dxmin, s_range[threadIdx.x][0]);
    atomicMax(&hd->dxmax, s_range[threadIdx.x][1]);
    atomicMin(&hd->dymin, s_range[threadIdx.x][2]);
    atomicMax(&hd->dymax, s_range[threadIdx.x][3]);
  }
}

struct OwnSeg {  // a run of channels of one group handled by this launch
  int g, c0, c1, cbase;  // cbase = index of channel c0 in the problem-wide channel numbering (gmax table)
};
struct OwnArgs {
  int d;         // direction
  int tshared;   // grad_src of these groups has T-stride 0: one tile accumulates all T frames
  int nseg, ctot;
  OwnSeg seg[FWB_MAX_GROUPS];
};

// One candidate pixel (lane) against the owner tile: fan-in count + fixed-point accumulation of w*gOut.
__device__ __forceinline__ void own_pixel(const Params& P, const GradP& Q, const OwnArgs& A, int n, int t, int i, int j,
                                          bool mine, const Tap& k, int sx0, int sy0, int* s_acc, int* s_cnt,
                                          const float* s_scale) {
  const int d = A.d;
  const int px = k.x0 - sx0, py = k.y0 - sy0;
  const bool cx0 = (unsigned)px < (unsigned)OT_W, cx1 = (unsigned)(px + 1) < (unsigned)OT_W;
  const bool cy0 = (unsigned)py < (unsigned)OT_H, cy1 = (unsigned)(py + 1) < (unsigned)OT_H;
  const bool in0 = mine && (k.valid & 1u) && cx0 && cy0, in1 = mine && (k.valid & 2u) && cx1 && cy0;
  const bool in2 = mine && (k.valid & 4u) && cx0 && cy1, in3 = mine && (k.valid & 8u) && cx1 && cy1;
  const bool any = in0 || in1 || in2 || in3;
  if (!__any_sync(0xffffffffu, any)) return;
  const int pos = py * OT_PITCH + px;
  const bool has_bl = P.dir[d].blend != nullptr;
  if (in0) atomicAdd(&s_cnt[pos], 1);
  if (in1) atomicAdd(&s_cnt[pos + 1], 1);
  if (in2) atomicAdd(&s_cnt[pos + OT_PITCH], 1);
  if (in3) atomicAdd(&s_cnt[pos + OT_PITCH + 1], 1);
  const float w0 = __fmul_rn(k.ux, k.uy), w1 = __fmul_rn(k.tx, k.uy), w2 = __fmul_rn(k.ux, k.ty), w3 = __fmul_rn(k.tx, k.ty);
  int cc = 0;
  for (int s = 0; s < A.nseg; ++s) {
    const int g = A.seg[s].g;
    const float* go = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)i * Q.go_sh[g] + j;
#pragma unroll 4
    for (int c = A.seg[s].c0; c < A.seg[s].c1; ++c, ++cc) {
      float gw = 0.0f;
      if (any) {
        gw = __ldg(go + (long long)c * Q.go_sc[g]);
        if (has_bl) gw = __fmul_rn(gw, k.blend);
      }
      const float gs = gw * s_scale[cc];  // exact: the scale is a power of two
      int* a = s_acc + cc * OT_WORDS + pos;
      if (in0) atomicAdd(a, __float2int_rn(w0 * gs));
      if (in1) atomicAdd(a + 1, __float2int_rn(w1 * gs));
      if (in2) atomicAdd(a + OT_PITCH, __float2int_rn(w2 * gs));
      if (in3) atomicAdd(a + OT_PITCH + 1, __float2int_rn(w3 * gs));
    }
  }
}

__global__ void __launch_bounds__(OWN_THREADS) bwd_src_owner_kernel(const __grid_constant__ Params P,
                                                                    const __grid_constant__ GradP Q, const WsView ws,
                                                                    const __grid_constant__ OwnArgs A) {
  extern __shared__ int smem[];
  const Geo& G = P.geo;
  int* s_acc = smem;                                  // [ctot][OT_WORDS]
  int* s_cnt = s_acc + A.ctot * OT_WORDS;             // [OT_WORDS]
  float* s_scale = (float*)(s_cnt + OT_WORDS);        // [ctot]
  float* s_inv = s_scale + A.ctot;                    // [ctot]
  int* s_hits = (int*)(s_inv + A.ctot);               // [HITCAP]
  int* s_misc = s_hits + HITCAP;                      // [0] nhit, [1] max fan-in

  const int d = A.d;
  const int sx0 = blockIdx.x * OT_W, sy0 = blockIdx.y * OT_H;
  const int n = A.tshared ? blockIdx.z : blockIdx.z / G.T;
  const int t_lo = A.tshared ? 0 : blockIdx.z - n * G.T, t_hi = A.tshared ? G.T : t_lo + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = OWN_THREADS / 32;

  // The accumulators are 32-bit fixed point with `headroom` integer bits for the fan-in.  If a tile turns out
  // to receive more than 2^headroom contributions in one pixel (possible overflow), redo it with more headroom.
  for (int headroom = HEADROOM0;; headroom += 8) {
    for (int k = tid; k < A.ctot * OT_WORDS + OT_WORDS; k += OWN_THREADS) smem[k] = 0;
    if (tid == 0) s_misc[1] = 0;
    // one power-of-two scale per channel from max|gOut*blend| over the frames this tile accumulates
    if (tid < A.ctot) {
      int cc = tid, s = 0;
      while (cc >= A.seg[s].c1 - A.seg[s].c0) {
        cc -= A.seg[s].c1 - A.seg[s].c0;
        ++s;
      }
      const int cg = A.seg[s].cbase + cc;
      unsigned gb = 0u;
      for (int t = t_lo; t < t_hi; ++t) gb = max(gb, ws.gmax[((size_t)d * ws.NT + n * G.T + t) * ws.ctot + cg]);
      float sc = 0.0f, inv = 0.0f;
      if (gb != 0u && gb < 0x7f800000u) {
        const int E = (int)(gb >> 23) - 127;  // |g| < 2^(E+1)
        const int e = max(-120, min(120, 30 - (E + 1) - headroom));
        sc = __uint_as_float((unsigned)(e + 127) << 23);
        inv = __uint_as_float((unsigned)(-e + 127) << 23);
      }
      s_scale[tid] = sc;
      s_inv[tid] = inv;
    }
    __syncthreads();

    for (int t = t_lo; t < t_hi; ++t) {
      const int nt = n * G.T + t;
      const WsHeader hd = ws.hdr[(size_t)d * ws.NT + nt];
      // ---- micro-tile records whose displacement window can reach this tile
      if (hd.dxmin <= hd.dxmax) {
        // candidate (i,j) with tap column x0 or x0+1 in [sx0, sx0+OT_W): j = x0 - dx, dx in [dxmin-R, dxmax+R]
        const int jlo = sx0 - 1 - (hd.dxmax + OUTLIER_R), jhi = sx0 + OT_W - 1 - (hd.dxmin - OUTLIER_R);
        const int ilo = sy0 - 1 - (hd.dymax + OUTLIER_R), ihi = sy0 + OT_H - 1 - (hd.dymin - OUTLIER_R);
        const int mx0 = max(jlo, 0) / MT_W, mx1 = min(jhi, G.W - 1) / MT_W;
        const int my0 = max(ilo, 0) / MT_H, my1 = min(ihi, G.H - 1) / MT_H;
        const int ww = mx1 - mx0 + 1, wh = my1 - my0 + 1;
        const int total = (jhi < 0 || ihi < 0 || ww <= 0 || wh <= 0) ? 0 : ww * wh;
        const short4* tab = ws.tab + ((size_t)d * ws.NT + nt) * ws.mth * ws.mtw;
        for (int base = 0; base < total; base += HITCAP) {
          if (tid == 0) s_misc[0] = 0;
          __syncthreads();
          for (int q = base + tid; q < min(base + HITCAP, total); q += OWN_THREADS) {
            const int my = my0 + q / ww, mx = mx0 + q % ww;
            const short4 r = tab[(size_t)my * ws.mtw + mx];
            if (r.x <= r.y && r.x < sx0 + OT_W && r.y >= sx0 && r.z < sy0 + OT_H && r.w >= sy0)
              s_hits[atomicAdd(&s_misc[0], 1)] = (my << 16) | mx;
          }
          __syncthreads();
          const int nhit = s_misc[0];
          for (int h = warp; h < nhit; h += NW) {
            const int my = s_hits[h] >> 16, mx = s_hits[h] & 0xffff;
            const int j = mx * MT_W + (lane & 7), i = my * MT_H + (lane >> 3);
            const bool active = j < G.W && i < G.H;
            Tap k;
            k.valid = 0u;
            k.x0 = k.y0 = 0;
            if (active) compute_tap(G, P.dir[d], n, t, i, j, k);
            int adx, ady;
            unsigned hm;
            const bool inl = mt_classify(active && k.valid != 0u, k.x0 - j, k.y0 - i, adx, ady, hm);
            own_pixel(P, Q, A, n, t, i, j, inl, k, sx0, sy0, s_acc, s_cnt, s_scale);
          }
          __syncthreads();
        }
      }
      // ---- outlier pixels of this image/direction
      const int nout = hd.n_outliers;
      const int2* ol = ws.outl + ((size_t)d * ws.NT + nt) * ws.cap;
      for (int base = 0; base < nout; base += HITCAP) {
        if (tid == 0) s_misc[0] = 0;
        __syncthreads();
        for (int q = base + tid; q < min(base + HITCAP, nout); q += OWN_THREADS) {
          const int2 e = ol[q];
          const int y0 = e.y >> 16, x0 = (int)(short)(e.y & 0xffff);
          if (x0 < sx0 + OT_W && x0 + 1 >= sx0 && y0 < sy0 + OT_H && y0 + 1 >= sy0) s_hits[atomicAdd(&s_misc[0], 1)] = e.x;
        }
        __syncthreads();
        const int nhit = s_misc[0];
        for (int h0 = warp * 32; h0 < nhit; h0 += NW * 32) {
          const int h = h0 + lane;
          const bool active = h < nhit;
          int i = 0, j = 0;
          Tap k;
          k.valid = 0u;
          k.x0 = k.y0 = 0;
          if (active) {
            const int pix = s_hits[h];
            i = pix / G.W;
            j = pix - i * G.W;
            compute_tap(G, P.dir[d], n, t, i, j, k);
          }
          own_pixel(P, Q, A, n, t, i, j, active, k, sx0, sy0, s_acc, s_cnt, s_scale);
        }
        __syncthreads();
      }
    }
    // largest fan-in of the tile: did the fixed-point headroom hold?
    int m = 0;
    for (int k = tid; k < OT_WORDS; k += OWN_THREADS) m = max(m, s_cnt[k]);
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) atomicMax(&s_misc[1], m);
    __syncthreads();
    const int M = s_misc[1];
    __syncthreads();
    if (M <= (1 << headroom) || headroom >= 30) break;
  }

  // ---- write the tile once: plain coalesced stores
  int cc = 0;
  for (int s = 0; s < A.nseg; ++s) {
    const int g = A.seg[s].g;
    float* gsb = Q.grad_src[g][d] + n * Q.gs_sn[g][d] + (A.tshared ? 0 : t_lo * Q.gs_st[g][d]);
    for (int c = A.seg[s].c0; c < A.seg[s].c1; ++c, ++cc) {
      const float inv = s_inv[cc];
      for (int k = tid; k < OT_W * OT_H; k += OWN_THREADS) {
        const int yy = k / OT_W, xx = k % OT_W;
        const int y = sy0 + yy, x = sx0 + xx;
        if (y < G.H && x < G.W)
          __stcs(gsb + (long long)c * Q.gs_sc[g][d] + (long long)y * Q.gs_sh[g][d] + x,
                 (float)s_acc[cc * OT_WORDS + yy * OT_PITCH + xx] * inv);
      }
    }
  }
}

static inline size_t own_smem_bytes(int ctot) {
  return sizeof(int) * ((size_t)ctot * OT_WORDS + OT_WORDS + 2 * (size_t)ctot + HITCAP + 8);
}

}  // namespace fwb

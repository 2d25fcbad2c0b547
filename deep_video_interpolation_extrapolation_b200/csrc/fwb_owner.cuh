// fwb_owner.cuh — kernel 3 (gradient w.r.t. the sources) as an OWNER GATHER, plus the segment tables it reads.
//
// ATen's grid_sampler_2d_backward scatters w*gOut with one global atomicAdd per tap per channel (184 atomics
// per pixel in the headline config; measured ~0.35 T adds/s on B200 = 1.1 ms, and L2 atomics top out near
// 0.6 T float-adds/s even as red.v4).  Here each CTA OWNS a 32x16 tile of grad_src and pulls in every
// contribution that lands in it — no memset, no global atomics, the tile is written once (8C B/pixel):
//   * emit_kernel leaves, per 8x4 output micro-tile (one warp), the bounding box of its in-image taps (8 B
//     record), appends the few pixels that stray far from their micro-tile ("outliers", e.g. border-clipped
//     ones) to a per-image list, and records the min/max micro-tile displacement per image;
//   * an owner CTA scans the records whose displacement window can reach its tile, sorts the hits, and
//     compacts — in a fixed order — the candidate pixels that really have a tap inside the tile into 16-byte
//     tap records in shared memory (phase 1);
//   * phase 2: warp w owns the float4 accumulators of channels 4w..4w+3 for the whole tile and walks the
//     records: 4 coalesced loads of gOut, then per tap one LDS.128 / 4 FFMA / STS.128 — a quarter of the
//     shared-memory instructions of a per-channel scatter, and no atomics at all (fp32 shared atomicAdd is a
//     CAS loop on sm_100a).  Lanes of one batch that hit the same pixel are serialised by a precomputed rank.
//   * every order (hits, records, ranks, frames) is fixed, so the fp32 result is bit-exact run to run: the
//     deterministic mode is simply the default path.
#pragma once
#include "fwb_coords.cuh"

namespace fwb {

constexpr int MT_W = 8, MT_H = 4;       // micro-tile = the 32 output pixels of one warp
constexpr int OUTLIER_R = 12;           // a pixel farther than this (px) from its micro-tile's anchor displacement is an outlier
constexpr int OT_W = 32, OT_H = 16;     // owner tile (source pixels)
constexpr int OT_PIX = OT_W * OT_H;
constexpr int OWN_MAXGRP = 8;           // float4 channel groups (= phase-2 warps) per launch: 32 channels
constexpr int OWN_RC = 768;             // tap records per chunk
constexpr int OWN_HITCAP = 512;         // hit-list chunk

struct WsHeader {  // one per (direction, n*T+t)
  int dxmin, dxmax, dymin, dymax;  // range of the micro-tile anchor displacements (pixels)
  int n_outliers;
  int pad[3];
};

struct WsView {
  WsHeader* hdr;  // [D][NT]
  short4* tab;    // [D][NT][mth][mtw]   {xmin, xmax, ymin, ymax} of the in-image taps of inlier pixels
  int2* outl;     // [D][NT][cap]        {i*W+j, (y0<<16)|(x0&0xffff)}
  int mtw, mth;
  long long cap;  // H*W
  int NT;
};

struct WsLayout {
  size_t hdr_off, tab_off, outl_off, total;
  int mtw, mth;
};

static inline WsLayout ws_layout(int D, long long NT, int H, int W) {
  WsLayout L;
  L.mtw = (W + MT_W - 1) / MT_W;
  L.mth = (H + MT_H - 1) / MT_H;
  size_t o = 0;
  L.hdr_off = o;
  o += sizeof(WsHeader) * (size_t)D * NT;
  o = (o + 255) & ~(size_t)255;
  L.tab_off = o;
  o += sizeof(short4) * (size_t)D * NT * L.mtw * L.mth;
  o = (o + 255) & ~(size_t)255;
  L.outl_off = o;
  o += sizeof(int2) * (size_t)D * NT * H * W;
  L.total = (o + 255) & ~(size_t)255;
  return L;
}

static inline WsView ws_view(void* base, const WsLayout& L, int NT, int H, int W) {
  WsView v;
  char* b = (char*)base;
  v.hdr = (WsHeader*)(b + L.hdr_off);
  v.tab = (short4*)(b + L.tab_off);
  v.outl = (int2*)(b + L.outl_off);
  v.mtw = L.mtw;
  v.mth = L.mth;
  v.cap = (long long)H * W;
  v.NT = NT;
  return v;
}

__global__ void ws_init_kernel(WsHeader* hdr, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    hdr[k].dxmin = hdr[k].dymin = 0x7fffffff;
    hdr[k].dxmax = hdr[k].dymax = -0x7fffffff;
    hdr[k].n_outliers = 0;
  }
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
  int g, c0, c1;
};
struct OwnArgs {
  int d;        // direction
  int tshared;  // grad_src of these groups has T-stride 0: one tile accumulates all T frames
  int nseg, nchan, ngrp;
  OwnSeg seg[FWB_MAX_GROUPS];
};

struct OwnSmem {
  float4* acc;     // [ngrp][OT_PIX]
  uint4* rec;      // [OWN_RC]  {tx bits, ty bits, packed, (i<<16)|j};  packed = (px+1) | (py+1)<<6 | inbits<<12 | rank<<16
  float* blend;    // [OWN_RC]
  int* hits;       // [OWN_HITCAP] unsorted
  int* hits2;      // [OWN_HITCAP] sorted
  int* bmax;       // [OWN_RC/32] largest rank of each batch of 32 records
  int* wcnt;       // [32] per-warp record counts of a phase-1 round
  int* misc;       // [4]
};

// sort the `n` distinct values of s.hits into s.hits2 (rank by counting; n <= OWN_HITCAP)
__device__ __forceinline__ void own_sort_hits(const OwnSmem& s, int n) {
  for (int a = threadIdx.x; a < n; a += blockDim.x) {
    const int v = s.hits[a];
    int r = 0;
    for (int b = 0; b < n; ++b) r += (s.hits[b] < v);
    s.hits2[r] = v;
  }
}

// phase 1 (one round): every warp brings one candidate lane-set; lanes with a tap inside the tile are appended
// to the record chunk in (warp, lane) order.  All threads of the CTA must call this.  Returns the new count.
__device__ __forceinline__ int own_append(const OwnSmem& s, int nrec, bool take, const Tap& k, int px, int py, unsigned bits,
                                          int i, int j) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (lane == 0) s.wcnt[warp] = __popc(m);
  __syncthreads();
  int base = nrec, total = 0;
  for (int w = 0; w < nw; ++w) {
    const int c = s.wcnt[w];
    if (w < warp) base += c;
    total += c;
  }
  if (take) {
    const int idx = base + __popc(m & ((1u << lane) - 1u));
    s.rec[idx] = make_uint4(__float_as_uint(k.tx), __float_as_uint(k.ty),
                            (unsigned)(px + 1) | ((unsigned)(py + 1) << 6) | (bits << 12), ((unsigned)i << 16) | (unsigned)j);
    s.blend[idx] = k.blend;
  }
  __syncthreads();
  return nrec + total;
}

// which taps of pixel tap `k` fall inside the owner tile at (sx0, sy0)
__device__ __forceinline__ unsigned own_bits(const Tap& k, bool mine, int sx0, int sy0, int& px, int& py) {
  px = k.x0 - sx0;
  py = k.y0 - sy0;
  if (!mine) return 0u;
  const bool cx0 = (unsigned)px < (unsigned)OT_W, cx1 = (unsigned)(px + 1) < (unsigned)OT_W;
  const bool cy0 = (unsigned)py < (unsigned)OT_H, cy1 = (unsigned)(py + 1) < (unsigned)OT_H;
  unsigned b = 0u;
  if ((k.valid & 1u) && cx0 && cy0) b |= 1u;
  if ((k.valid & 2u) && cx1 && cy0) b |= 2u;
  if ((k.valid & 4u) && cx0 && cy1) b |= 4u;
  if ((k.valid & 8u) && cx1 && cy1) b |= 8u;
  return b;
}

struct OwnChan {  // the (up to) 4 channels a phase-2 warp owns
  const float* gp[4];  // grad_out plane base for (n, c); + t*gst + i*gsh + j
  long long gst[4];
  int gsh[4];
  float* op[4];        // grad_src plane base for (n, [t], c)
  int osh[4];
  bool ok[4];
};

// phases 1.5 + 2 over the current record chunk (all threads call; frame t)
__device__ __forceinline__ void own_flush(const OwnSmem& s, int nrec, int t, int ngrp, const OwnChan& ch, bool has_bl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nb = (nrec + 31) >> 5;
  // ---- phase 1.5: lanes of one batch that target the same pixel get ranks 0,1,2..; they are applied in turns
  for (int b = warp; b < nb; b += nw) {
    const int idx = b * 32 + lane;
    const bool valid = idx < nrec;
    const unsigned pk = valid ? s.rec[idx].z : 0u;
    const unsigned key = valid ? (pk & 0xfffu) : (0x10000u | (unsigned)lane);
    const unsigned m = __match_any_sync(0xffffffffu, key);
    const int rank = __popc(m & ((1u << lane) - 1u));
    if (valid) s.rec[idx].z = pk | ((unsigned)rank << 16);
    const int mr = __reduce_max_sync(0xffffffffu, rank);
    if (lane == 0) s.bmax[b] = mr;
  }
  __syncthreads();
  // ---- phase 2: warp w accumulates channels 4w..4w+3 of every record
  if (warp < ngrp) {
    float4* acc = s.acc + warp * OT_PIX;
    uint4 r_n = make_uint4(0, 0, 0, 0);
    float g_n[4] = {0.f, 0.f, 0.f, 0.f};
    auto load = [&](int b) {
      const int idx = b * 32 + lane;
      const bool valid = idx < nrec;
      r_n = valid ? s.rec[idx] : make_uint4(0, 0, 0, 0);
      const int i = (int)(r_n.w >> 16), j = (int)(r_n.w & 0xffffu);
      const float bl = (valid && has_bl) ? s.blend[idx] : 1.0f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float g = 0.0f;
        if (valid && ch.ok[q]) g = __ldg(ch.gp[q] + t * ch.gst[q] + (long long)i * ch.gsh[q] + j);
        g_n[q] = has_bl ? __fmul_rn(g, bl) : g;
      }
    };
    if (nb > 0) load(0);
    for (int b = 0; b < nb; ++b) {
      const uint4 r = r_n;
      const float g0 = g_n[0], g1 = g_n[1], g2 = g_n[2], g3 = g_n[3];
      if (b + 1 < nb) load(b + 1);  // software prefetch of the next batch's gOut
      const float tx = __uint_as_float(r.x), ty = __uint_as_float(r.y);
      const float ux = 1.0f - tx, uy = 1.0f - ty;
      const float w0 = ux * uy, w1 = tx * uy, w2 = ux * ty, w3 = tx * ty;
      const int px = (int)(r.z & 63u) - 1, py = (int)((r.z >> 6) & 63u) - 1;
      const unsigned bits = (r.z >> 12) & 15u;
      const int rank = (int)(r.z >> 16);
      const int pos = py * OT_W + px;
      const int mr = s.bmax[b];
      for (int rr = 0; rr <= mr; ++rr) {
        const bool on = rank == rr;
#define FWB_RMW(BIT, OFF, WGT)                                   \
  if (on && (bits & BIT)) {                                      \
    float4 a = acc[pos + (OFF)];                                 \
    a.x = fmaf(WGT, g0, a.x);                                    \
    a.y = fmaf(WGT, g1, a.y);                                    \
    a.z = fmaf(WGT, g2, a.z);                                    \
    a.w = fmaf(WGT, g3, a.w);                                    \
    acc[pos + (OFF)] = a;                                        \
  }                                                              \
  __syncwarp();
        FWB_RMW(1u, 0, w0)
        FWB_RMW(2u, 1, w1)
        FWB_RMW(4u, OT_W, w2)
        FWB_RMW(8u, OT_W + 1, w3)
#undef FWB_RMW
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) bwd_src_owner_kernel(const __grid_constant__ Params P, const __grid_constant__ GradP Q,
                                                            const WsView ws, const __grid_constant__ OwnArgs A) {
  extern __shared__ float4 smem4[];
  const Geo& G = P.geo;
  OwnSmem s;
  s.acc = smem4;
  s.rec = (uint4*)(s.acc + A.ngrp * OT_PIX);
  s.blend = (float*)(s.rec + OWN_RC);
  s.hits = (int*)(s.blend + OWN_RC);
  s.hits2 = s.hits + OWN_HITCAP;
  s.bmax = s.hits2 + OWN_HITCAP;
  s.wcnt = s.bmax + (OWN_RC / 32);
  s.misc = s.wcnt + 32;

  const int d = A.d;
  const int sx0 = blockIdx.x * OT_W, sy0 = blockIdx.y * OT_H;
  const int n = A.tshared ? blockIdx.z : blockIdx.z / G.T;
  const int t_lo = A.tshared ? 0 : blockIdx.z - n * G.T, t_hi = A.tshared ? G.T : t_lo + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nw = nthr >> 5;
  const bool has_bl = P.dir[d].blend != nullptr;

  // channels of this warp (phase 2 / final store)
  OwnChan ch;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int cc = warp * 4 + q;
    ch.ok[q] = warp < A.ngrp && cc < A.nchan;
    ch.gp[q] = nullptr;
    ch.op[q] = nullptr;
    ch.gst[q] = 0;
    ch.gsh[q] = ch.osh[q] = 0;
    if (ch.ok[q]) {
      int sg = 0;
      while (cc >= A.seg[sg].c1 - A.seg[sg].c0) {
        cc -= A.seg[sg].c1 - A.seg[sg].c0;
        ++sg;
      }
      const int g = A.seg[sg].g, c = A.seg[sg].c0 + cc;
      ch.gp[q] = Q.grad_out[g] + n * Q.go_sn[g] + (long long)c * Q.go_sc[g];
      ch.gst[q] = Q.go_st[g];
      ch.gsh[q] = Q.go_sh[g];
      ch.op[q] = Q.grad_src[g][d] + n * Q.gs_sn[g][d] + (A.tshared ? 0 : t_lo * Q.gs_st[g][d]) + (long long)c * Q.gs_sc[g][d];
      ch.osh[q] = Q.gs_sh[g][d];
    }
  }

  for (int k = tid; k < A.ngrp * OT_PIX; k += nthr) s.acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  for (int t = t_lo; t < t_hi; ++t) {
    const int nt = n * G.T + t;
    const WsHeader hd = ws.hdr[(size_t)d * ws.NT + nt];
    int nrec = 0;
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
      for (int base = 0; base < total; base += OWN_HITCAP) {
        if (tid == 0) s.misc[0] = 0;
        __syncthreads();
        for (int q = base + tid; q < min(base + OWN_HITCAP, total); q += nthr) {
          const int my = my0 + q / ww, mx = mx0 + q % ww;
          const short4 r = tab[(size_t)my * ws.mtw + mx];
          if (r.x <= r.y && r.x < sx0 + OT_W && r.y >= sx0 && r.z < sy0 + OT_H && r.w >= sy0)
            s.hits[atomicAdd(&s.misc[0], 1)] = (my << 16) | mx;
        }
        __syncthreads();
        const int nhit = s.misc[0];
        own_sort_hits(s, nhit);
        __syncthreads();
        for (int h0 = 0; h0 < nhit; h0 += nw) {
          if (nrec + nw * 32 > OWN_RC) {
            own_flush(s, nrec, t, A.ngrp, ch, has_bl);
            nrec = 0;
          }
          const int h = h0 + warp;
          Tap k;
          k.valid = 0u;
          k.x0 = k.y0 = 0;
          k.tx = k.ty = 0.f;
          k.blend = 1.f;
          int i = 0, j = 0;
          bool active = false;
          if (h < nhit) {
            const int my = s.hits2[h] >> 16, mx = s.hits2[h] & 0xffff;
            j = mx * MT_W + (lane & 7);
            i = my * MT_H + (lane >> 3);
            active = j < G.W && i < G.H;
            if (active) compute_tap(G, P.dir[d], n, t, i, j, k);
          }
          int adx, ady, px, py;
          unsigned hm;
          const bool inl = mt_classify(active && k.valid != 0u, k.x0 - j, k.y0 - i, adx, ady, hm);
          const unsigned bits = own_bits(k, inl, sx0, sy0, px, py);
          nrec = own_append(s, nrec, bits != 0u, k, px, py, bits, i, j);
        }
      }
    }
    // ---- outlier pixels of this image/direction, in ascending pixel order (id ranges that fit the hit list)
    const int nout = hd.n_outliers;
    const int2* ol = ws.outl + ((size_t)d * ws.NT + nt) * ws.cap;
    const int HW = G.H * G.W;
    for (int lo = 0; lo < HW && nout > 0;) {
      int hi = HW, cnt;
      for (;;) {
        if (tid == 0) s.misc[0] = 0;
        __syncthreads();
        for (int q = tid; q < nout; q += nthr) {
          const int2 e = ol[q];
          const int y0 = e.y >> 16, x0 = (int)(short)(e.y & 0xffff);
          if (e.x >= lo && e.x < hi && x0 < sx0 + OT_W && x0 + 1 >= sx0 && y0 < sy0 + OT_H && y0 + 1 >= sy0) {
            const int slot = atomicAdd(&s.misc[0], 1);
            if (slot < OWN_HITCAP) s.hits[slot] = e.x;
          }
        }
        __syncthreads();
        cnt = s.misc[0];
        __syncthreads();
        if (cnt <= OWN_HITCAP) break;
        hi = lo + max((hi - lo) >> 1, 1);  // a range of <= OWN_HITCAP ids always fits (ids are distinct)
      }
      own_sort_hits(s, cnt);
      __syncthreads();
      for (int h0 = 0; h0 < cnt; h0 += nw * 32) {
        if (nrec + nw * 32 > OWN_RC) {
          own_flush(s, nrec, t, A.ngrp, ch, has_bl);
          nrec = 0;
        }
        const int h = h0 + warp * 32 + lane;
        Tap k;
        k.valid = 0u;
        k.x0 = k.y0 = 0;
        k.tx = k.ty = 0.f;
        k.blend = 1.f;
        int i = 0, j = 0;
        const bool active = h < cnt;
        if (active) {
          const int pix = s.hits2[h];
          i = pix / G.W;
          j = pix - i * G.W;
          compute_tap(G, P.dir[d], n, t, i, j, k);
        }
        int px, py;
        const unsigned bits = own_bits(k, active, sx0, sy0, px, py);
        nrec = own_append(s, nrec, bits != 0u, k, px, py, bits, i, j);
      }
      lo = hi;
    }
    own_flush(s, nrec, t, A.ngrp, ch, has_bl);
  }

  // ---- write the tile once: plain coalesced stores, warp w its 4 channel planes
  if (warp < A.ngrp) {
    const float4* acc = s.acc + warp * OT_PIX;
    for (int pos = lane; pos < OT_PIX; pos += 32) {
      const int y = sy0 + pos / OT_W, x = sx0 + pos % OT_W;
      if (y < G.H && x < G.W) {
        const float4 a = acc[pos];
        if (ch.ok[0]) __stcs(ch.op[0] + (long long)y * ch.osh[0] + x, a.x);
        if (ch.ok[1]) __stcs(ch.op[1] + (long long)y * ch.osh[1] + x, a.y);
        if (ch.ok[2]) __stcs(ch.op[2] + (long long)y * ch.osh[2] + x, a.z);
        if (ch.ok[3]) __stcs(ch.op[3] + (long long)y * ch.osh[3] + x, a.w);
      }
    }
  }
}

static inline size_t own_smem_bytes(int ngrp) {
  return sizeof(float4) * (size_t)ngrp * OT_PIX + sizeof(uint4) * OWN_RC + sizeof(float) * OWN_RC +
         sizeof(int) * (2 * OWN_HITCAP + OWN_RC / 32 + 32 + 8);
}

}  // namespace fwb

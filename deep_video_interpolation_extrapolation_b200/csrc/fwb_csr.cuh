// fwb_csr.cuh — kernel 3 (gradient w.r.t. the sources) as an OWNER GATHER over per-pixel contributor lists.
//
// ATen's grid_sampler_2d_backward scatters w*gOut with one global atomicAdd per tap per channel (184 float
// atomics per pixel in the headline config; B200's L2 retires ~0.35 T of them per second = 1.1 ms, and the
// order of the adds changes from run to run).  A shared-memory scatter is bound by read-modify-write traffic
// and by same-address conflicts.  Here the scatter is turned into a GATHER:
//   * a CTA OWNS a 32x32 tile of grad_src of one (n, t, direction) and finds every output pixel with a tap in
//     it (micro-tile tap boxes + outlier list left by emit_kernel, see fwb_owner.cuh): "records";
//   * the records are inverted — channel independent, once per tile — into a CSR: for every owned source pixel
//     the list of (weight, contributor) pairs, each list sorted by record number (a fixed order);
//   * the contributors' grad_out values are staged per channel plane into shared memory with 16-byte cp.async
//     (row segments of the contributing output rows; far-away contributors get one cell each), double
//     buffered like kernels 1 and 2;
//   * every thread then sums its 4 source pixels' lists from shared memory into registers and writes
//     grad_src once, coalesced: no atomics, no read-modify-write, no memset.
// Every order (records, lists, flushes, frames) is fixed, so the fp32 result is bit-exact run to run:
// deterministic mode is the default path, at no cost.
#pragma once
#include "fwb_coords.cuh"
#include "fwb_owner.cuh"
#include "fwb_stage.cuh"

namespace fwb {

constexpr int CS_TW = 32, CS_TH = 32, CS_PIX = CS_TW * CS_TH;
constexpr int CS_THREADS = 256, CS_PPT = 4;  // thread owns 4 pixels of the tile (picked by list length, see B5)
constexpr int CS_BW = CS_TW + 1;             // base cells: (x0, y0) of a record lies in [-1, 31]^2 relative to the tile
constexpr int CS_NB = CS_BW * (CS_TH + 1);   // 1089
constexpr int CS_NBP = 1280;                 // padded: 5 per thread in the scan
constexpr int CS_RC = 2048;                  // records per flush
constexpr int CS_HITCAP = 512;               // micro-tile / outlier hits per scan chunk
constexpr int CS_MAXOUT = 256;               // far contributors per flush (one staged cell each)
constexpr int CS_OUTROUND = 64;              // outlier candidates examined per round
constexpr int CS_CCMAX = 4;                  // grad_out planes per chunk
constexpr int CS_LONG = 17;                  // a list of this many records or more is walked by a whole warp

struct CsrArgs {
  int d;         // direction
  int tshared;   // grad_src of these groups has T-stride 0: the tile accumulates all T frames
  int gout_vec;  // every grad_out pointer is 16-byte aligned with strides % 4 == 0 (cp.async staging allowed)
  unsigned group_mask;  // groups served by this launch
  int stage_floats;     // floats of shared memory available for staging
  int long_len;         // a list of this many records or more (<= 31) is walked by a whole warp
};

struct CsrSmem {
  int* boff;            // [CS_NBP + 1] start of every base cell's record list (sorted records)
  int* bcur;            // [CS_NBP]     count, then fill cursor
  unsigned short* tmp;  // [CS_RC]      records grouped by base cell, unordered inside a cell
  float4* sw;           // [CS_RC]      sorted records: tap weights (nw, ne, sw, se) * blend
  unsigned* so;         // [CS_RC]      sorted records: byte offset inside a staged plane, or (i << 16) | j
  int* hits;            // [CS_HITCAP]
  int* hits2;           // [CS_HITCAP]
  int* wcnt;            // [32] scratch
  int* misc;            // [8]
  int* bkt;             // [32] list-length histogram / bucket cursors
  unsigned short* perm; // [CS_PIX] owned pixels ordered by list length (a warp gets 32 lists of equal length)
  // records (aliased with the staging area: dead once the sorted copies exist)
  unsigned* r_ij;  // (i << 16) | j
  unsigned* r_pk;  // (px+1) | (py+1) << 6 | bits << 12
  float* r_tx;
  float* r_ty;
  float* r_bl;
  float* stage;
};

static inline size_t csr_fixed_bytes() {
  return sizeof(int) * (CS_NBP + 4 + CS_NBP) + sizeof(unsigned short) * CS_RC + 16 * (size_t)CS_RC + 4 * (size_t)CS_RC +
         sizeof(int) * (2 * CS_HITCAP + 32 + 8 + 32) + sizeof(unsigned short) * CS_PIX;
}
static inline size_t csr_record_bytes() { return (size_t)CS_RC * 20; }

__device__ __forceinline__ CsrSmem csr_carve(float4* base) {
  CsrSmem s;
  char* p = reinterpret_cast<char*>(base);
  s.sw = reinterpret_cast<float4*>(p);
  p += 16 * (size_t)CS_RC;
  s.boff = reinterpret_cast<int*>(p);
  p += sizeof(int) * (CS_NBP + 4);
  s.bcur = reinterpret_cast<int*>(p);
  p += sizeof(int) * CS_NBP;
  s.so = reinterpret_cast<unsigned*>(p);
  p += 4 * (size_t)CS_RC;
  s.tmp = reinterpret_cast<unsigned short*>(p);
  p += sizeof(unsigned short) * CS_RC;
  s.hits = reinterpret_cast<int*>(p);
  p += sizeof(int) * CS_HITCAP;
  s.hits2 = reinterpret_cast<int*>(p);
  p += sizeof(int) * CS_HITCAP;
  s.wcnt = reinterpret_cast<int*>(p);
  p += sizeof(int) * 32;
  s.misc = reinterpret_cast<int*>(p);
  p += sizeof(int) * 8;
  s.bkt = reinterpret_cast<int*>(p);
  p += sizeof(int) * 32;
  s.perm = reinterpret_cast<unsigned short*>(p);
  p += sizeof(unsigned short) * CS_PIX;
  s.stage = reinterpret_cast<float*>(p);  // 16-byte aligned: every size above is a multiple of 16
  s.r_ij = reinterpret_cast<unsigned*>(p);
  s.r_pk = s.r_ij + CS_RC;
  s.r_tx = reinterpret_cast<float*>(s.r_pk + CS_RC);
  s.r_ty = s.r_tx + CS_RC;
  s.r_bl = s.r_ty + CS_RC;
  return s;
}

__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
  return v;
}

// sort the `n` distinct values of s.hits into s.hits2 (rank by counting; n <= CS_HITCAP)
__device__ __forceinline__ void csr_sort_hits(const CsrSmem& s, int n) {
  for (int a = threadIdx.x; a < n; a += blockDim.x) {
    const int v = s.hits[a];
    int r = 0;
    for (int b = 0; b < n; ++b) r += (s.hits[b] < v);
    s.hits2[r] = v;
  }
}

// which taps of `k` fall inside the owner tile at (sx0, sy0)
__device__ __forceinline__ unsigned csr_bits(const Tap& k, bool mine, int sx0, int sy0, int& px, int& py) {
  px = k.x0 - sx0;
  py = k.y0 - sy0;
  if (!mine) return 0u;
  const bool cx0 = (unsigned)px < (unsigned)CS_TW, cx1 = (unsigned)(px + 1) < (unsigned)CS_TW;
  const bool cy0 = (unsigned)py < (unsigned)CS_TH, cy1 = (unsigned)(py + 1) < (unsigned)CS_TH;
  unsigned b = 0u;
  if ((k.valid & 1u) && cx0 && cy0) b |= 1u;
  if ((k.valid & 2u) && cx1 && cy0) b |= 2u;
  if ((k.valid & 4u) && cx0 && cy1) b |= 4u;
  if ((k.valid & 8u) && cx1 && cy1) b |= 8u;
  return b;
}

// append the taking lanes' records in (warp, lane) order.  All threads call.  nrec/nent are CTA-uniform.
__device__ __forceinline__ void csr_append(const CsrSmem& s, int& nrec, int& nent, bool take, const Tap& k, int px, int py,
                                           unsigned bits, int i, int j) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, take);
  int ne = take ? __popc(bits) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ne += __shfl_xor_sync(0xffffffffu, ne, o);
  if (lane == 0) s.wcnt[warp] = __popc(m) | (ne << 8);
  __syncthreads();
  int base = nrec, total = 0, tent = 0;
  for (int w = 0; w < nw; ++w) {
    const int c = s.wcnt[w] & 0xff;
    if (w < warp) base += c;
    total += c;
    tent += s.wcnt[w] >> 8;
  }
  if (take) {
    const int idx = base + __popc(m & ((1u << lane) - 1u));
    s.r_ij[idx] = ((unsigned)i << 16) | (unsigned)j;
    s.r_pk[idx] = (unsigned)(px + 1) | ((unsigned)(py + 1) << 6) | (bits << 12);
    s.r_tx[idx] = k.tx;
    s.r_ty[idx] = k.ty;
    s.r_bl[idx] = k.blend;
  }
  __syncthreads();
  nrec += total;
  nent += tent;
}

// Invert records [0, nrec) (the first `nin` come from micro-tiles, the rest are far contributors) into per-cell
// lists, stage grad_out and accumulate the tile.  All threads call.  `accumulate`: add to grad_src, not store.
//
// A record with base cell b = (x0, y0) feeds the owned pixels b (nw tap), b+(1,0) (ne), b+(0,1) (sw), b+(1,1) (se).
// So it is enough to group the RECORDS by base cell (one entry per record, not four): pixel c then walks the
// two contiguous ranges  [base c-(1,0), base c]  (taking ne, then nw)  and  [base c-(1,1), base c-(0,1)]  (se, sw).
// Inside a base cell the records are ordered by record number, which is a fixed order -> bit-exact run to run.
__device__ __noinline__ void csr_flush(const Params& P, const GradP& Q, const CsrArgs& A, int n, int t, int nrec, int nin,
                                       bool accumulate) {
  extern __shared__ float4 cs_smem4[];  // carved here (not passed in) so that every access compiles to LDS/STS
  __shared__ StageTab gtab;
  const CsrSmem s = csr_carve(cs_smem4);
  const Geo& G = P.geo;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = A.d;
  const bool tsh = A.tshared != 0;
  // ---- B1: records per base cell; C1: rows/segments of the contributing output pixels
  for (int k = tid; k < CS_NBP; k += CS_THREADS) s.bcur[k] = 0;
  if (tid < 32) s.bkt[tid] = 0;
  stage_tab_init(&gtab, 1);
  __syncthreads();
  int ilo = 0x7fffffff, ihi = -0x7fffffff;
  for (int r = tid; r < nrec; r += CS_THREADS) {
    const unsigned pk = s.r_pk[r];
    atomicAdd(&s.bcur[(int)((pk >> 6) & 63u) * CS_BW + (int)(pk & 63u)], 1);
    if (r < nin) {
      const int i = (int)(s.r_ij[r] >> 16), j = (int)(s.r_ij[r] & 0xffffu);
      ilo = min(ilo, i);
      ihi = max(ihi, i);
      atomicMin(&gtab.xlo[i & (ST_ROWS - 1)], j);
      atomicMax(&gtab.xhi[i & (ST_ROWS - 1)], j);
    }
  }
  ilo = __reduce_min_sync(0xffffffffu, ilo);
  ihi = __reduce_max_sync(0xffffffffu, ihi);
  if (lane == 0 && ilo <= ihi) {
    atomicMin(&gtab.ymin, ilo);
    atomicMax(&gtab.ymax, ihi);
  }
  __syncthreads();
  // ---- B2: exclusive scan of the counts -> boff; reset the cursors
  {
    const int b5 = tid * 5;
    int v[5], sum = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      v[k] = s.bcur[b5 + k];
      sum += v[k];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s.wcnt[warp] = inc;
    __syncthreads();
    int ex = inc - sum;
    for (int w = 0; w < warp; ++w) ex += s.wcnt[w];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      s.boff[b5 + k] = ex;
      ex += v[k];
      s.bcur[b5 + k] = 0;
    }
    if (tid == CS_THREADS - 1) s.boff[CS_NBP] = ex;
  }
  if (warp == 0) stage_tab_scan(gtab);  // C2: prefix of the row segments (j ranges, 16 B aligned)
  __syncthreads();
  // ---- plan: can grad_out be staged?
  const int nfar = nrec - nin;
  const int seg_floats = 4 * gtab.total4;
  const int plane = ST_ZPAD + seg_floats + ((nfar + 3) & ~3);
  int cc = 0;
  const bool staged = A.gout_vec && gtab.ok && (cc = min(CS_CCMAX, A.stage_floats / (2 * plane))) >= 1;
  // ---- B3: group the records by base cell (order inside a cell arbitrary here)
  for (int r = tid; r < nrec; r += CS_THREADS) {
    const unsigned pk = s.r_pk[r];
    const int bc = (int)((pk >> 6) & 63u) * CS_BW + (int)(pk & 63u);
    s.tmp[s.boff[bc] + atomicAdd(&s.bcur[bc], 1)] = (unsigned short)r;
  }
  // list length of every owned pixel -> histogram (B5)
#pragma unroll
  for (int q = 0; q < CS_PPT; ++q) {
    const int pos = tid + CS_THREADS * q;
    const int bA = ((pos >> 5) + 1) * CS_BW + (pos & 31) + 1, bB = bA - CS_BW;
    const int len = (s.boff[bA + 1] - s.boff[bA - 1]) + (s.boff[bB + 1] - s.boff[bB - 1]);
    atomicAdd(&s.bkt[min(len, A.long_len)], 1);
  }
  __syncthreads();
  // ---- B4: rank every record inside its base cell (by record number) and emit the sorted copy: one thread
  // per record, so a long list costs its members a longer scan but never serialises on one thread
  for (int k = tid; k < nrec; k += CS_THREADS) {
    const int r = s.tmp[k];
    const unsigned pk = s.r_pk[r];
    const int bc = (int)((pk >> 6) & 63u) * CS_BW + (int)(pk & 63u);
    const int b = s.boff[bc], e = s.boff[bc + 1];
    int rank = 0;
    for (int m = b; m < e; ++m) rank += (s.tmp[m] < r);
    const float tx = s.r_tx[r], ty = s.r_ty[r], bl = s.r_bl[r];
    const float ux = 1.0f - tx, uy = 1.0f - ty;
    s.sw[b + rank] = make_float4(ux * uy * bl, tx * uy * bl, ux * ty * bl, tx * ty * bl);
    const unsigned ij = s.r_ij[r];
    unsigned o = ij;
    if (staged) {
      if (r < nin) {
        const int rr = (int)(ij >> 16) - gtab.ymin;
        o = 4u * (unsigned)(ST_ZPAD + 4 * gtab.rowoff4[rr] + ((int)(ij & 0xffffu) - gtab.rowx[rr]));
      } else {
        o = 4u * (unsigned)(ST_ZPAD + seg_floats + (r - nin));
      }
    }
    s.so[b + rank] = o;
  }
  // ---- B5: order the owned pixels by list length, so that the 32 lists a warp walks in lock step are
  // equally long (with a rough flow the lengths vary 4x inside a row: 26 % lane efficiency otherwise)
  if (warp == 0) {
    const int v = s.bkt[lane];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    s.bkt[lane] = inc - v;
  }
  // loads of this thread per plane: up to ST_SLOTS 16-byte pieces of the row segments + 1 far cell
  StageLoad ld;
  int far_ij = -1;
  if (staged) {
    stage_assign(gtab, G.H, G.W, ld);
    if (tid < nfar) far_ij = (int)s.r_ij[nin + tid];
  }
  __syncthreads();  // records are dead from here on (the staging area overwrites them)
  const int long_start = s.bkt[A.long_len];  // perm[long_start ..] are the pixels with long lists
  __syncthreads();
#pragma unroll
  for (int q = 0; q < CS_PPT; ++q) {
    const int pos = tid + CS_THREADS * q;
    const int bA = ((pos >> 5) + 1) * CS_BW + (pos & 31) + 1, bB = bA - CS_BW;
    const int len = (s.boff[bA + 1] - s.boff[bA - 1]) + (s.boff[bB + 1] - s.boff[bB - 1]);
    s.perm[atomicAdd(&s.bkt[min(len, A.long_len)], 1)] = (unsigned short)pos;
  }
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(s.stage);
  if (staged)
    for (int k = tid; k < 2 * cc * ST_ZPAD; k += CS_THREADS) s.stage[(k / ST_ZPAD) * plane + (k % ST_ZPAD)] = 0.f;
  __syncthreads();
  // this thread's 4 pixels: the two record ranges and where the tap changes inside each
  // range 1 = [r1b, r1b + len1): ne taps up to r1m, then nw; range 2 starts at r2b: se taps up to r2m, then sw.
  // Both are walked by ONE loop of `tot` iterations (equal across a warp thanks to B5).
  int opos[CS_PPT], r1b[CS_PPT], r1m[CS_PPT], len1[CS_PPT], d2[CS_PPT], r2m[CS_PPT], tot[CS_PPT];
#pragma unroll
  for (int q = 0; q < CS_PPT; ++q) {
    // ranks are dealt to the warps in alternating direction so that every warp gets a similar total
    const int rk = CS_THREADS * q + ((q & 1) ? CS_THREADS - 1 - tid : tid);
    opos[q] = s.perm[rk];
    const int bA = ((opos[q] >> 5) + 1) * CS_BW + (opos[q] & 31) + 1, bB = bA - CS_BW;
    r1b[q] = s.boff[bA - 1];  // base (px-1, py): ne tap
    r1m[q] = s.boff[bA];      // base (px, py):   nw tap
    len1[q] = s.boff[bA + 1] - r1b[q];
    const int r2b = s.boff[bB - 1];  // base (px-1, py-1): se tap
    r2m[q] = s.boff[bB];             // base (px, py-1):   sw tap
    tot[q] = len1[q] + s.boff[bB + 1] - r2b;
    d2[q] = r2b - len1[q];
    if (rk >= long_start) opos[q] = -1;  // served by the warp-cooperative pass below
  }

  for (int g = 0; g < G.n_groups; ++g) {
    if (!((A.group_mask >> g) & 1u)) continue;
    const GroupP& R = P.grp[g];
    const float* gp = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g];
    const int gsh = Q.go_sh[g], gsc = Q.go_sc[g];
    float* opb = Q.grad_src[g][d] + n * Q.gs_sn[g][d] + (tsh ? 0 : t * Q.gs_st[g][d]);
    const int osh = Q.gs_sh[g][d], osc = Q.gs_sc[g][d];
    int lgoff[ST_SLOTS];
#pragma unroll
    for (int sl = 0; sl < ST_SLOTS; ++sl) lgoff[sl] = staged ? (ld.ycol[sl] >> 16) * gsh + (ld.ycol[sl] & 0xffff) : 0;
    const int fgoff = far_ij >= 0 ? (far_ij >> 16) * gsh + (far_ij & 0xffff) : 0;

    auto issue = [&](int c0, int stage) {
      const int nch = min(cc, R.C - c0);
      unsigned pb = sbase + 4u * (unsigned)(stage * cc * plane);
      for (int u = 0; u < nch; ++u) {
        const float* pp = gp + (long long)(c0 + u) * gsc;
#pragma unroll
        for (int sl = 0; sl < ST_SLOTS; ++sl)
          if (ld.bytes[sl] >= 0)
            cp_async16(pb + 4u * (unsigned)(ST_ZPAD + 4 * (tid + sl * CS_THREADS)), pp + lgoff[sl], ld.bytes[sl]);
        if (far_ij >= 0) cp_async4(pb + 4u * (unsigned)(ST_ZPAD + seg_floats + tid), pp + fgoff);
        pb += 4u * (unsigned)plane;
      }
      cp_async_commit();
    };

    const int step = staged ? cc : CS_CCMAX;
    int stage = 0;
    if (staged) issue(0, 0);
    for (int c0 = 0; c0 < R.C; c0 += step) {
      const int nch = min(step, R.C - c0);
      if (staged) {
        if (c0 + step < R.C) {
          issue(c0 + step, stage ^ 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncthreads();
      }
      // staged: shared-space byte addresses of the chunk's planes; planes past nch alias plane 0 (finite data,
      // result unused)
      const unsigned sp0 = sbase + 4u * (unsigned)(stage * cc * plane);
      const unsigned sp1 = sp0 + 4u * (unsigned)(plane * (nch > 1));
      const unsigned sp2 = sp0 + 8u * (unsigned)(plane * (nch > 2));
      const unsigned sp3 = sp0 + 12u * (unsigned)(plane * (nch > 3));
      const float* gp0 = gp + (long long)c0 * gsc;
#pragma unroll
      for (int q = 0; q < CS_PPT; ++q) {
        if (opos[q] < 0) continue;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int it = 0; it < tot[q]; ++it) {
          const bool first = it < len1[q];
          const int k = first ? r1b[q] + it : it + d2[q];
          const float4 w4 = s.sw[k];
          const unsigned so = s.so[k];
          const bool lo = k < (first ? r1m[q] : r2m[q]);
          const float wa = lo ? w4.y : w4.x, wb = lo ? w4.w : w4.z;
          const float w = first ? wa : wb;
          if (staged) {
            acc0 = fmaf(w, lds_f32(sp0 + so), acc0);
            acc1 = fmaf(w, lds_f32(sp1 + so), acc1);
            acc2 = fmaf(w, lds_f32(sp2 + so), acc2);
            acc3 = fmaf(w, lds_f32(sp3 + so), acc3);
          } else {
            const float* cell = gp0 + (int)(so >> 16) * gsh + (int)(so & 0xffffu);
            acc0 = fmaf(w, __ldg(cell), acc0);
            if (nch > 1) acc1 = fmaf(w, __ldg(cell + gsc), acc1);
            if (nch > 2) acc2 = fmaf(w, __ldg(cell + 2ll * gsc), acc2);
            if (nch > 3) acc3 = fmaf(w, __ldg(cell + 3ll * gsc), acc3);
          }
        }
        const int oy = blockIdx.y * CS_TH + (opos[q] >> 5), ox = blockIdx.x * CS_TW + (opos[q] & 31);
        if (ox < G.W && oy < G.H) {
          float* o = opb + (long long)c0 * osc + oy * osh + ox;
          const float acc[CS_CCMAX] = {acc0, acc1, acc2, acc3};
#pragma unroll
          for (int u = 0; u < CS_CCMAX; ++u)
            if (u < nch) {
              float v = acc[u];
              if (accumulate) v += o[(long long)u * osc];
              o[(long long)u * osc] = v;
            }
        }
      }
      // long lists: one warp per pixel, lane l takes records l, l+32, ...; fixed butterfly order
      for (int lp = long_start + warp; lp < CS_PIX; lp += CS_THREADS / 32) {
        const int pos = s.perm[lp];
        const int bA = ((pos >> 5) + 1) * CS_BW + (pos & 31) + 1, bB = bA - CS_BW;
        const int a1b = s.boff[bA - 1], a1m = s.boff[bA], l1 = s.boff[bA + 1] - a1b;
        const int a2b = s.boff[bB - 1], a2m = s.boff[bB], tl = l1 + s.boff[bB + 1] - a2b;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int it = lane; it < tl; it += 32) {
          const bool first = it < l1;
          const int k = first ? a1b + it : a2b + it - l1;
          const float4 w4 = s.sw[k];
          const unsigned so = s.so[k];
          const bool lo = k < (first ? a1m : a2m);
          const float wa = lo ? w4.y : w4.x, wb = lo ? w4.w : w4.z;
          const float w = first ? wa : wb;
          if (staged) {
            acc0 = fmaf(w, lds_f32(sp0 + so), acc0);
            acc1 = fmaf(w, lds_f32(sp1 + so), acc1);
            acc2 = fmaf(w, lds_f32(sp2 + so), acc2);
            acc3 = fmaf(w, lds_f32(sp3 + so), acc3);
          } else {
            const float* cell = gp0 + (int)(so >> 16) * gsh + (int)(so & 0xffffu);
            acc0 = fmaf(w, __ldg(cell), acc0);
            if (nch > 1) acc1 = fmaf(w, __ldg(cell + gsc), acc1);
            if (nch > 2) acc2 = fmaf(w, __ldg(cell + 2ll * gsc), acc2);
            if (nch > 3) acc3 = fmaf(w, __ldg(cell + 3ll * gsc), acc3);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
          acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
          acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
          acc3 += __shfl_xor_sync(0xffffffffu, acc3, o);
        }
        const int oy = blockIdx.y * CS_TH + (pos >> 5), ox = blockIdx.x * CS_TW + (pos & 31);
        if (lane < nch && ox < G.W && oy < G.H) {
          float* o = opb + (long long)(c0 + lane) * osc + oy * osh + ox;
          float v = lane == 0 ? acc0 : lane == 1 ? acc1 : lane == 2 ? acc2 : acc3;
          if (accumulate) v += *o;
          *o = v;
        }
      }
      if (staged) {
        __syncthreads();
        stage ^= 1;
      }
    }
  }
  __syncthreads();  // the staging area becomes the record area again
}

__global__ void __launch_bounds__(CS_THREADS, 2) bwd_src_csr_kernel(const __grid_constant__ Params P,
                                                                  const __grid_constant__ GradP Q, const WsView ws,
                                                                  const __grid_constant__ CsrArgs A) {
  extern __shared__ float4 cs_smem4[];
  const Geo& G = P.geo;
  const CsrSmem s = csr_carve(cs_smem4);
  const int d = A.d;
  const int sx0 = blockIdx.x * CS_TW, sy0 = blockIdx.y * CS_TH;
  const int n = A.tshared ? blockIdx.z : blockIdx.z / G.T;
  const int t_lo = A.tshared ? 0 : blockIdx.z - n * G.T, t_hi = A.tshared ? G.T : t_lo + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nw = nthr >> 5;

  bool written = false;  // has this tile been stored yet (then later flushes accumulate)

  for (int t = t_lo; t < t_hi; ++t) {
    const int nt = n * G.T + t;
    const WsHeader hd = ws.hdr[(size_t)d * ws.NT + nt];
    int nrec = 0, nent = 0;
    // ---- micro-tile records whose displacement window can reach this tile
    if (hd.dxmin <= hd.dxmax) {
      // candidate (i,j) with tap column x0 or x0+1 in [sx0, sx0+TW): j = x0 - dx, dx in [dxmin-R, dxmax+R]
      const int jlo = sx0 - 1 - (hd.dxmax + OUTLIER_R), jhi = sx0 + CS_TW - 1 - (hd.dxmin - OUTLIER_R);
      const int ilo = sy0 - 1 - (hd.dymax + OUTLIER_R), ihi = sy0 + CS_TH - 1 - (hd.dymin - OUTLIER_R);
      const int mx0 = max(jlo, 0) / MT_W, mx1 = min(jhi, G.W - 1) / MT_W;
      const int my0 = max(ilo, 0) / MT_H, my1 = min(ihi, G.H - 1) / MT_H;
      const int ww = mx1 - mx0 + 1, wh = my1 - my0 + 1;
      const int total = (jhi < 0 || ihi < 0 || ww <= 0 || wh <= 0) ? 0 : ww * wh;
      const short4* tab = ws.tab + ((size_t)d * ws.NT + nt) * ws.mth * ws.mtw;
      for (int base = 0; base < total; base += CS_HITCAP) {
        if (tid == 0) s.misc[0] = 0;
        __syncthreads();
        for (int q = base + tid; q < min(base + CS_HITCAP, total); q += nthr) {
          const int my = my0 + q / ww, mx = mx0 + q % ww;
          const short4 r = tab[(size_t)my * ws.mtw + mx];
          if (r.x <= r.y && r.x < sx0 + CS_TW && r.y >= sx0 && r.z < sy0 + CS_TH && r.w >= sy0)
            s.hits[atomicAdd(&s.misc[0], 1)] = (my << 16) | mx;
        }
        __syncthreads();
        const int nhit = s.misc[0];
        csr_sort_hits(s, nhit);
        __syncthreads();
        for (int h0 = 0; h0 < nhit; h0 += nw) {
          if (nrec + nw * 32 > CS_RC) {
            csr_flush(P, Q, A, n, t, nrec, nrec, written);
            written = true;
            nrec = nent = 0;
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
          const unsigned bits = csr_bits(k, inl, sx0, sy0, px, py);
          csr_append(s, nrec, nent, bits != 0u, k, px, py, bits, i, j);
        }
      }
    }
    const int nin = nrec;
    // ---- outlier pixels of this image/direction, in ascending pixel order (id ranges that fit the hit list)
    const int nout = hd.n_outliers;
    const int2* ol = ws.outl + ((size_t)d * ws.NT + nt) * ws.cap;
    const int HW = G.H * G.W;
    int nin_cur = nin;  // records [0, nin_cur) are micro-tile records of the current batch
    for (int lo = 0; lo < HW && nout > 0;) {
      int hi = HW, cnt;
      for (;;) {
        if (tid == 0) s.misc[0] = 0;
        __syncthreads();
        for (int q = tid; q < nout; q += nthr) {
          const int2 e = ol[q];
          const int y0 = e.y >> 16, x0 = (int)(short)(e.y & 0xffff);
          if (e.x >= lo && e.x < hi && x0 < sx0 + CS_TW && x0 + 1 >= sx0 && y0 < sy0 + CS_TH && y0 + 1 >= sy0) {
            const int slot = atomicAdd(&s.misc[0], 1);
            if (slot < CS_HITCAP) s.hits[slot] = e.x;
          }
        }
        __syncthreads();
        cnt = s.misc[0];
        __syncthreads();
        if (cnt <= CS_HITCAP) break;
        hi = lo + max((hi - lo) >> 1, 1);  // a range of <= CS_HITCAP ids always fits (ids are distinct)
      }
      csr_sort_hits(s, cnt);
      __syncthreads();
      for (int h0 = 0; h0 < cnt; h0 += CS_OUTROUND) {
        if (nrec + CS_OUTROUND > CS_RC || (nrec - nin_cur) + CS_OUTROUND > CS_MAXOUT) {
          csr_flush(P, Q, A, n, t, nrec, nin_cur, written);
          written = true;
          nrec = nent = 0;
          nin_cur = 0;
        }
        const int h = h0 + tid;
        Tap k;
        k.valid = 0u;
        k.x0 = k.y0 = 0;
        k.tx = k.ty = 0.f;
        k.blend = 1.f;
        int i = 0, j = 0;
        const bool active = tid < CS_OUTROUND && h < cnt;
        if (active) {
          const int pix = s.hits2[h];
          i = pix / G.W;
          j = pix - i * G.W;
          compute_tap(G, P.dir[d], n, t, i, j, k);
        }
        int px, py;
        const unsigned bits = csr_bits(k, active, sx0, sy0, px, py);
        csr_append(s, nrec, nent, bits != 0u, k, px, py, bits, i, j);
      }
      lo = hi;
    }
    // the frame's last batch; also the one that writes zeros when the tile has no contributor at all
    if (nrec > 0 || !written) {
      csr_flush(P, Q, A, n, t, nrec, nin_cur, written);
      written = true;
    }
  }
}

static inline size_t csr_smem_bytes(size_t stage_bytes) {
  const size_t a = csr_record_bytes();
  return csr_fixed_bytes() + (stage_bytes > a ? stage_bytes : a);
}

}  // namespace fwb

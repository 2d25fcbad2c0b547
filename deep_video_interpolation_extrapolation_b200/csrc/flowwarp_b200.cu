// flowwarp_b200.cu — sm_100a kernels + C-ABI of the flow-warp / blend hot path.
//
// Reference semantics: utils/net_utils.py:89-129 and nets/OpticalUnet.py:7-15,123-146 of
// lzhangbj/deep_video_interpolation_extrapolation (see include/flowwarp_b200.h for the map).
// HBM-bound gather/scatter: no tensor cores.  One thread per output pixel so that a warp's 32
// taps of one plane are (nearly) one 128 B line; all channels of all groups are looped inside the
// thread so the coordinate math, weights and validity are computed once per pixel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/flowwarp_b200.h"
#include "fwb_coords.cuh"
#include "fwb_generic.cuh"
#include "fwb_owner.cuh"
#include "fwb_tile.cuh"
#include "fwb_cl.cuh"
#include "fwb_bwdx.cuh"
#include "fwb_blend.cuh"
#include "fwb_label.cuh"
#include "fwb_loss.cuh"

namespace fwb {

// Thread -> pixel mapping.  A warp covers an 8x4 "micro-tile" (not 32x1): with rough flows the 32 taps of a
// 32x1 row land on ~16 different source rows = ~16 L1 wavefronts per load; an 8x4 patch halves that.  A CTA
// is 8 warps = 4x2 micro-tiles = 32x8 pixels.  The micro-tile is also the unit of the segment tables that
// kernel 3 (owner gather) consumes.
constexpr int BX = 32;  // CTA width in pixels
constexpr int BY = 8;   // CTA height in pixels
constexpr int NTHREADS = 256;

__device__ __forceinline__ void thread_pixel(int& i, int& j) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  j = blockIdx.x * BX + (warp & 3) * MT_W + (lane & 7);
  i = blockIdx.y * BY + (warp >> 2) * MT_H + (lane >> 3);
}

// ---------------------------------------------------------------------------------------------
// Kernel 1 (generic): one thread per pixel, gather from global memory (fwb_generic.cuh).
// ---------------------------------------------------------------------------------------------
template <int NDIRS>
__global__ void __launch_bounds__(NTHREADS) fwd_kernel(const __grid_constant__ Params P) {
  const Geo& G = P.geo;
  int i, j;
  thread_pixel(i, j);
  if (j >= G.W || i >= G.H) return;
  const int n = blockIdx.z / G.T, t = blockIdx.z - n * G.T;
  fwd_generic_pixel<NDIRS>(P, n, t, i, j);
}

// Debug / parity kernel: integer indices, validity bits, float coordinates of one direction.
__global__ void indices_kernel(const __grid_constant__ Params P, int d, int* __restrict__ x0,
                               int* __restrict__ y0, uint8_t* __restrict__ valid, float* __restrict__ ix,
                               float* __restrict__ iy) {
  const Geo& G = P.geo;
  int i, j;
  thread_pixel(i, j);
  if (j >= G.W || i >= G.H) return;
  const int n = blockIdx.z / G.T, t = blockIdx.z - n * G.T;
  Tap k;
  compute_tap(G, P.dir[d], n, t, i, j, k);
  const long long o = ((long long)blockIdx.z * G.H + i) * G.W + j;
  if (x0) x0[o] = k.x0;
  if (y0) y0[o] = k.y0;
  if (valid) valid[o] = (uint8_t)k.valid;
  if (ix) ix[o] = k.ix;
  if (iy) iy[o] = k.iy;
}

// ---------------------------------------------------------------------------------------------
// Kernel 2 (generic): gradient w.r.t. flow / gate / blend weight — a pure gather, one thread per pixel.
// ---------------------------------------------------------------------------------------------
template <int NDIRS>
__global__ void __launch_bounds__(NTHREADS) bwd_flow_kernel(const __grid_constant__ Params P,
                                                            const __grid_constant__ GradP Q) {
  const Geo& G = P.geo;
  int i, j;
  thread_pixel(i, j);
  if (j >= G.W || i >= G.H) return;
  const int n = blockIdx.z / G.T, t = blockIdx.z - n * G.T;
  bwdflow_generic_pixel<NDIRS>(P, Q, n, t, i, j);
}

// ---------------------------------------------------------------------------------------------
// Kernel 3a: gradient w.r.t. the sources by global atomics (the "spill" path; non-deterministic).
// grad_src must be zero on entry (zero_rows_kernel below).
// ---------------------------------------------------------------------------------------------
__global__ void zero_rows_kernel(float* __restrict__ base, long long sn, long long st, int sc, int sh, int N,
                                 int T, int C, int H, int W) {
  // blockIdx.x enumerates (n,t,c), blockIdx.y strides over the rows of that plane
  const int ntc = blockIdx.x;
  const int c = ntc % C, nt = ntc / C;
  const int n = nt / T, t = nt - n * T;
  float* plane = base + n * sn + t * st + (long long)c * sc;
  for (int i = blockIdx.y; i < H; i += gridDim.y) {
    float* row = plane + (long long)i * sh;
    for (int j = threadIdx.x; j < W; j += blockDim.x) row[j] = 0.0f;
  }
}

template <int NDIRS>
__global__ void __launch_bounds__(NTHREADS) bwd_src_atomic_kernel(const __grid_constant__ Params P,
                                                                const __grid_constant__ GradP Q) {
  const Geo& G = P.geo;
  int i, j;
  thread_pixel(i, j);
  if (j >= G.W || i >= G.H) return;
  const int n = blockIdx.z / G.T, t = blockIdx.z - n * G.T;
  float w[NDIRS][4], bl[NDIRS];
  int x0[NDIRS], y0[NDIRS];
  unsigned v[NDIRS];
  bool has_bl[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) {
    Tap k;
    compute_tap(G, P.dir[d], n, t, i, j, k);
    w[d][0] = __fmul_rn(k.ux, k.uy);
    w[d][1] = __fmul_rn(k.tx, k.uy);
    w[d][2] = __fmul_rn(k.ux, k.ty);
    w[d][3] = __fmul_rn(k.tx, k.ty);
    x0[d] = k.x0;
    y0[d] = k.y0;
    v[d] = k.valid;
    bl[d] = k.blend;
    has_bl[d] = P.dir[d].blend != nullptr;
  }
  for (int g = 0; g < G.n_groups; ++g) {
    const GroupP& R = P.grp[g];
    if (!Q.grad_out[g]) continue;
    const float* go = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)i * Q.go_sh[g] + j;
    for (int c = 0; c < R.C; ++c) {
      const float gout = __ldcs(go + (long long)c * Q.go_sc[g]);
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) {
        float* gs = Q.grad_src[g][d];
        if (!gs) continue;
        const int sh = Q.gs_sh[g][d];
        gs += n * Q.gs_sn[g][d] + t * Q.gs_st[g][d] + (long long)c * Q.gs_sc[g][d] + (long long)y0[d] * sh + x0[d];
        const float gw = has_bl[d] ? gout * bl[d] : gout;
        if (v[d] & 1u) atomicAdd(gs, w[d][0] * gw);
        if (v[d] & 2u) atomicAdd(gs + 1, w[d][1] * gw);
        if (v[d] & 4u) atomicAdd(gs + sh, w[d][2] * gw);
        if (v[d] & 8u) atomicAdd(gs + sh + 1, w[d][3] * gw);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool stride_ok(long long C, long long sc, long long H, long long sh);
static int validate(const fwb_problem* p) {
  if (!p) return FWB_E_NULL;
  if (p->N < 0 || p->T < 1 || p->H < 1 || p->W < 1) return FWB_E_SHAPE;
  if ((long long)p->N * p->T > 65535) return FWB_E_SHAPE;
  if (p->H > 32767 || p->W > 32767) return FWB_E_RANGE;
  if (p->n_dirs < 1 || p->n_dirs > 2) return FWB_E_DIRS;
  if (p->n_groups < 1 || p->n_groups > FWB_MAX_GROUPS) return FWB_E_GROUPS;
  if (p->padding_mode != FWB_PAD_ZEROS && p->padding_mode != FWB_PAD_BORDER) return FWB_E_MODE;
  if (p->align_corners != 0 && p->align_corners != 1) return FWB_E_MODE;
  for (int d = 0; d < p->n_dirs; ++d) {
    if (!p->dir[d].flow) return FWB_E_NULL;
    if (p->dir[d].sign != 1.0f && p->dir[d].sign != -1.0f) return FWB_E_MODE;
    if (((uintptr_t)p->dir[d].flow | (uintptr_t)p->dir[d].gate | (uintptr_t)p->dir[d].blend) & 3u)
      return FWB_E_ALIGN;
  }
  for (int g = 0; g < p->n_groups; ++g) {
    const fwb_group* R = &p->grp[g];
    if (R->C < 1) return FWB_E_SHAPE;
    for (int d = 0; d < p->n_dirs; ++d) {
      if (!R->src[d]) return FWB_E_NULL;
      if ((uintptr_t)R->src[d] & 3u) return FWB_E_ALIGN;
      // in-image offsets are 32-bit in the kernels
      if (!stride_ok(R->C, R->src_sc[d], p->H, R->src_sh[d])) return FWB_E_SHAPE;
    }
    if (R->out && !stride_ok(R->C, R->out_sc, p->H, R->out_sh)) return FWB_E_SHAPE;
  }
  return 0;
}

static void to_params(const fwb_problem* p, Params& P) {
  Geo& G = P.geo;
  G.N = p->N;
  G.T = p->T;
  G.H = p->H;
  G.W = p->W;
  G.n_dirs = p->n_dirs;
  G.n_groups = p->n_groups;
  G.pad_border = p->padding_mode == FWB_PAD_BORDER;
  G.align = p->align_corners;
  G.stepx = p->W > 1 ? 2.0f / (float)(p->W - 1) : 0.0f;
  G.stepy = p->H > 1 ? 2.0f / (float)(p->H - 1) : 0.0f;
  for (int d = 0; d < 2; ++d) {
    const fwb_dir& s = p->dir[d < p->n_dirs ? d : 0];
    DirP& D = P.dir[d];
    D.flow = s.flow;
    D.flow_sn = s.flow_sn;
    D.flow_sc = s.flow_sc;
    D.flow_st = s.flow_st;
    D.flow_sh = s.flow_sh;
    D.gate = s.gate;
    D.gate_sn = s.gate_sn;
    D.gate_st = s.gate_st;
    D.gate_sh = s.gate_sh;
    D.blend = s.blend;
    D.blend_sn = s.blend_sn;
    D.blend_st = s.blend_st;
    D.blend_sh = s.blend_sh;
    D.sign = s.sign;
  }
  for (int g = 0; g < FWB_MAX_GROUPS; ++g) {
    const fwb_group& s = p->grp[g < p->n_groups ? g : 0];
    GroupP& R = P.grp[g];
    R.C = s.C;
    for (int d = 0; d < 2; ++d) {
      const int e = d < p->n_dirs ? d : 0;
      R.src[d] = s.src[e];
      R.src_sn[d] = s.src_sn[e];
      R.src_st[d] = s.src_st[e];
      R.src_sc[d] = (int)s.src_sc[e];
      R.src_sh[d] = (int)s.src_sh[e];
    }
    R.out = s.out;
    R.out_sn = s.out_sn;
    R.out_st = s.out_st;
    R.out_sc = (int)s.out_sc;
    R.out_sh = (int)s.out_sh;
  }
}

// in-plane offsets are 32-bit in the kernels: a channel / row stride must be non-negative and the farthest in-plane offset
// (C channels, H + 2 rows) must fit
static bool stride_ok(long long C, long long sc, long long H, long long sh) {
  return sc >= 0 && sh >= 0 && sc <= 2147483647LL && sh <= 2147483647LL && C * sc + (H + 2) * sh <= 2147483647LL;
}

static int to_grads(const fwb_problem* p, const fwb_grads* q, GradP& Q) {
  if (!q) return FWB_E_NULL;
  for (int g = 0; g < p->n_groups; ++g) {
    if (q->grad_out[g] && !stride_ok(p->grp[g].C, q->go_sc[g], p->H, q->go_sh[g])) return FWB_E_SHAPE;
    for (int d = 0; d < p->n_dirs; ++d)
      if (q->grad_src[g][d] && !stride_ok(p->grp[g].C, q->gs_sc[g][d], p->H, q->gs_sh[g][d])) return FWB_E_SHAPE;
  }
  for (int g = 0; g < FWB_MAX_GROUPS; ++g) {
    const bool on = g < p->n_groups;
    Q.grad_out[g] = on ? q->grad_out[g] : nullptr;
    Q.go_sn[g] = q->go_sn[g];
    Q.go_st[g] = q->go_st[g];
    Q.go_sc[g] = (int)q->go_sc[g];
    Q.go_sh[g] = (int)q->go_sh[g];
    if ((uintptr_t)Q.grad_out[g] & 3u) return FWB_E_ALIGN;
    for (int d = 0; d < 2; ++d) {
      Q.grad_src[g][d] = (on && d < p->n_dirs) ? q->grad_src[g][d] : nullptr;
      Q.gs_sn[g][d] = q->gs_sn[g][d];
      Q.gs_st[g][d] = q->gs_st[g][d];
      Q.gs_sc[g][d] = (int)q->gs_sc[g][d];
      Q.gs_sh[g][d] = (int)q->gs_sh[g][d];
      if ((uintptr_t)Q.grad_src[g][d] & 3u) return FWB_E_ALIGN;
    }
  }
  for (int d = 0; d < 2; ++d) {
    const bool on = d < p->n_dirs;
    Q.grad_flow[d] = on ? q->grad_flow[d] : nullptr;
    Q.gf_sn[d] = q->gf_sn[d];
    Q.gf_sc[d] = q->gf_sc[d];
    Q.gf_st[d] = q->gf_st[d];
    Q.gf_sh[d] = q->gf_sh[d];
    Q.grad_gate[d] = on ? q->grad_gate[d] : nullptr;
    Q.gg_sn[d] = q->gg_sn[d];
    Q.gg_st[d] = q->gg_st[d];
    Q.gg_sh[d] = q->gg_sh[d];
    Q.grad_blend[d] = on ? q->grad_blend[d] : nullptr;
    Q.gb_sn[d] = q->gb_sn[d];
    Q.gb_st[d] = q->gb_st[d];
    Q.gb_sh[d] = q->gb_sh[d];
  }
  return 0;
}

static dim3 pixel_grid(const fwb_problem* p) {
  return dim3((p->W + BX - 1) / BX, (p->H + BY - 1) / BY, p->N * p->T);
}

// Kernel selection.  Defaults are the fastest measured variants:
//   forward            texture-gather kernel (fwb_tex.cuh) on dense sources  ("notex": shared-memory tile kernel, fwb_tile.cuh;
//                                                                             "notile" / "generic": one-thread-per-pixel LDG gather)
//   backward, fast     kernels 2+3 fused: texture gather + integer-atomic shared-memory scatter on dense sources (fwb_tile.cuh,
//                      TEX = true); kernel 2 alone on the texture units when no source gradient is wanted (fwb_tex.cuh)
//                      ("sorted": kernel 2 on textures + kernel 3 as a sorted shared-memory gather, fwb_bwdx.cuh, two launches;
//                       "notex": staged tiles + integer-atomic scatter; "nofuse": split kernels, atomics-free)
//   backward, determ.  generic kernel 2 + owner-gather kernel 3
// FWB_KERNELS=<comma separated words> and the FWB_TILE_* sizes switch variants for A/B measurements and for the tests that
// keep every variant parity-checked.  The environment is read ONCE per process (thread-safe); fwb_reload_env() re-reads it.
enum : unsigned { KN_NOFUSE = 8u, KN_GENERIC = 16u, KN_NOTILE = 32u, KN_NOZFUSE = 128u, KN_NOTEX = 256u, KN_SORTED = 512u, KN_NOCL = 1024u };
struct EnvCfg {
  unsigned knobs;
  int tile_fwd_kb, tile_bwd_kb, tile_bwdf_kb, tile_bwdx_kb, tile_bwd_ppt, bwdx_ppt, cl_kb, cl_minc;
};
static EnvCfg g_env;
static std::atomic<int> g_env_ready{0};
static std::mutex g_env_mu;
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
static void env_load_locked() {
  EnvCfg e;
  e.knobs = 0u;
  const char* v = getenv("FWB_KERNELS");
  if (v && *v) {
    if (strstr(v, "nofuse")) e.knobs |= KN_NOFUSE;
    if (strstr(v, "generic")) e.knobs |= KN_GENERIC;
    if (strstr(v, "notile")) e.knobs |= KN_NOTILE;
    if (strstr(v, "nozfuse")) e.knobs |= KN_NOZFUSE;
    if (strstr(v, "notex")) e.knobs |= KN_NOTEX;
    if (strstr(v, "sorted")) e.knobs |= KN_SORTED;
    if (strstr(v, "nocl")) e.knobs |= KN_NOCL;
  }
  e.tile_bwd_ppt = env_int("FWB_TILE_BWD_PPT", 2);
  e.tile_fwd_kb = env_int("FWB_TILE_FWD_KB", 52);
  e.tile_bwd_kb = env_int("FWB_TILE_BWD_KB", e.tile_bwd_ppt == 1 ? 48 : 80);
  e.tile_bwdf_kb = env_int("FWB_TILE_BWDF_KB", 52);
  e.tile_bwdx_kb = env_int("FWB_TILE_BWDX_KB", 46);
  e.bwdx_ppt = env_int("FWB_BWDX_PPT", 2);
  e.cl_kb = env_int("FWB_CL_KB", 95);
  e.cl_minc = env_int("FWB_CL_MINC", 12);
  g_env = e;
  g_env_ready.store(1, std::memory_order_release);
}
static const EnvCfg& env() {
  if (!g_env_ready.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(g_env_mu);
    if (!g_env_ready.load(std::memory_order_relaxed)) env_load_locked();
  }
  return g_env;
}
static unsigned knobs() { return env().knobs; }
static bool force_generic() { return (knobs() & KN_GENERIC) != 0u; }

// the staged kernels move 16-byte pieces of the source planes with cp.async: every source pointer must be
// 16-byte aligned and every stride a multiple of 4 elements
static bool stage_ok(const fwb_problem* p) {
  if (force_generic() || (p->W & 3)) return false;  // a 16-byte piece is all inside or all outside the image
  for (int g = 0; g < p->n_groups; ++g)
    for (int d = 0; d < p->n_dirs; ++d) {
      const fwb_group* R = &p->grp[g];
      if (((uintptr_t)R->src[d] & 15u) || (R->src_sn[d] & 3) || (R->src_st[d] & 3) || (R->src_sc[d] & 3) ||
          (R->src_sh[d] & 3))
        return false;
    }
  return true;
}

static int total_channels(const fwb_problem* p) {
  int c = 0;
  for (int g = 0; g < p->n_groups; ++g) c += p->grp[g].C;
  return c;
}

// the tile kernels keep one in-plane offset per piece / pixel: every group must use the same row strides
static bool tile_dirs_ok(const fwb_problem* p) {  // in-plane offsets of flow / gate / blend are 32-bit in the tile kernels
  for (int d = 0; d < p->n_dirs; ++d) {
    const fwb_dir& D = p->dir[d];
    const long long m = 2147483647LL, h = p->H + 1;
    if (D.flow_sh < 0 || D.flow_sh * h >= m) return false;
    if (D.gate && (D.gate_sh < 0 || D.gate_sh * h >= m)) return false;
    if (D.blend && (D.blend_sh < 0 || D.blend_sh * h >= m)) return false;
  }
  return true;
}
static bool tile_fwd_ok(const fwb_problem* p) {
  if (total_channels(p) > TL_MAXCH || !tile_dirs_ok(p)) return false;
  for (int g = 1; g < p->n_groups; ++g) {
    if (p->grp[g].out_sh != p->grp[0].out_sh) return false;
    for (int d = 0; d < p->n_dirs; ++d)
      if (p->grp[g].src_sh[d] != p->grp[0].src_sh[d]) return false;
  }
  return (long long)(p->H + 1) * p->grp[0].out_sh < 2147483647LL;
}
static bool tile_bwd_ok(const fwb_problem* p, const fwb_grads* q) {
  int g0 = -1, C = 0;
  for (int g = 0; g < p->n_groups; ++g) {
    if (!q->grad_out[g]) continue;
    if (g0 < 0) g0 = g;
    C += p->grp[g].C;
    if (q->go_sh[g] != q->go_sh[g0]) return false;
    for (int d = 0; d < p->n_dirs; ++d) {
      if (p->grp[g].src_sh[d] != p->grp[g0].src_sh[d]) return false;
      if (q->grad_src[g][d] && q->gs_sh[g][d] != q->gs_sh[g0][d]) return false;
      if (q->grad_src[g][d] && !q->grad_src[g0][d]) return false;  // gs_sh[g0][d] must be meaningful
    }
  }
  if (g0 < 0 || C > TL_MAXCH || !tile_dirs_ok(p)) return false;
  for (int d = 0; d < p->n_dirs; ++d)
    if ((long long)(p->H + 1) * q->gs_sh[g0][d] >= 2147483647LL) return false;
  return (long long)(p->H + 1) * q->go_sh[g0] < 2147483647LL;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-(kernel, device) attribute: set it only when a launch needs more than
// what this process already asked for on the current device (one cudaFuncSetAttribute per kernel and device, not per launch)
struct SmemAttr {
  const void* fn;
  int dev, bytes;
};
static std::mutex g_attr_mu;
static std::vector<SmemAttr> g_attr;
static int set_smem_ptr(const void* fn, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> lk(g_attr_mu);
  for (SmemAttr& a : g_attr)
    if (a.fn == fn && a.dev == dev) {
      if (a.bytes >= bytes) return 0;
      e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e == cudaSuccess) a.bytes = bytes;
      return (int)e;
    }
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) g_attr.push_back({fn, dev, bytes});
  return (int)e;
}
template <typename K>
static int set_smem(K kernel, int bytes) {
  return set_smem_ptr(reinterpret_cast<const void*>(kernel), bytes);
}

// ---------------------------------------------------------------------------------------------
// Texture objects over dense source tensors (fwb_tex.cuh).  A texture object is a descriptor (address, extent, pitch): it
// stays valid for as long as the memory does, so objects are created once per (pointer, extent, pitch, device) and kept in a
// process-wide, mutex-guarded cache that is never evicted while kernels may be in flight: when it is full the caller simply
// takes the shared-memory tile path.  fwb_release_cache() destroys the objects (caller guarantees the device is idle).
// ---------------------------------------------------------------------------------------------
struct TexEntry {
  const void* ptr;
  int width, rows, dev;
  size_t pitch;
  cudaTextureObject_t obj;
};
struct TexLimits {
  int dev, maxw, maxh, align, palign;
};
static std::mutex g_tex_mu;
static std::vector<TexEntry> g_tex;
static std::vector<TexLimits> g_texlim;
constexpr size_t TEX_CACHE_MAX = 4096;

static bool tex_limits(int dev, TexLimits& L) {  // g_tex_mu held
  for (const TexLimits& l : g_texlim)
    if (l.dev == dev) {
      L = l;
      return true;
    }
  L.dev = dev;
  if (cudaDeviceGetAttribute(&L.maxw, cudaDevAttrMaxTexture2DLinearWidth, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&L.maxh, cudaDevAttrMaxTexture2DLinearHeight, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&L.align, cudaDevAttrTextureAlignment, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&L.palign, cudaDevAttrTexturePitchAlignment, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  g_texlim.push_back(L);
  return true;
}

static bool tex_get(const float* ptr, int width, int rows, size_t pitch, int dev, unsigned long long* out) {  // g_tex_mu held
  for (const TexEntry& e : g_tex)
    if (e.ptr == ptr && e.width == width && e.rows == rows && e.pitch == pitch && e.dev == dev) {
      *out = (unsigned long long)e.obj;
      return true;
    }
  if (g_tex.size() >= TEX_CACHE_MAX) return false;
  cudaResourceDesc rd;
  memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypePitch2D;
  rd.res.pitch2D.devPtr = const_cast<float*>(ptr);
  rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  rd.res.pitch2D.width = (size_t)width;
  rd.res.pitch2D.height = (size_t)rows;
  rd.res.pitch2D.pitchInBytes = pitch;
  cudaTextureDesc td;
  memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  // creating a descriptor enqueues nothing; allow it while the calling thread is capturing a CUDA graph
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  (void)cudaThreadExchangeStreamCaptureMode(&mode);
  cudaTextureObject_t obj = 0;
  const cudaError_t e = cudaCreateTextureObject(&obj, &rd, &td, nullptr);
  (void)cudaThreadExchangeStreamCaptureMode(&mode);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  g_tex.push_back({ptr, width, rows, dev, pitch, obj});
  *out = (unsigned long long)obj;
  return true;
}

// Can every source of the problem be read through textures?  Fills X (texture objects, block geometry) if so.
static bool tex_prepare(const fwb_problem* p, TexP& X) {
  if (knobs() & (KN_NOTEX | KN_GENERIC)) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  std::lock_guard<std::mutex> lk(g_tex_mu);
  TexLimits L;
  if (!tex_limits(dev, L) || p->W > L.maxw) return false;
  memset(&X, 0, sizeof(X));
  for (int g = 0; g < p->n_groups; ++g)
    for (int d = 0; d < p->n_dirs; ++d) {
      const fwb_group& R = p->grp[g];
      const long long sh = R.src_sh[d], sc = R.src_sc[d], st = R.src_st[d], sn = R.src_sn[d], C = R.C;
      const int Tn = (p->T == 1 || st == 0) ? 1 : p->T;
      if (sh < p->W || (sh * 4) % L.palign || sc != (long long)p->H * sh) return false;
      if (Tn > 1 && st != C * sc) return false;
      if (p->N > 1 && sn != Tn * C * sc) return false;
      const long long rows_n = (long long)Tn * C * p->H;
      if (rows_n > L.maxh || rows_n > (1 << 23)) return false;
      if ((uintptr_t)R.src[d] % (uintptr_t)L.align) return false;
      long long nb = L.maxh / rows_n;
      if (nb > p->N) nb = p->N;
      if (nb > 1024) nb = 1024;
      while (nb > 1 && ((nb * sn * 4) % L.align)) --nb;  // every block of clips must start on a texture-aligned address
      if (nb < p->N && ((nb * sn * 4) % L.align)) return false;
      const long long nblk = (p->N + nb - 1) / nb;
      if (nblk > TX_MAXBLK) return false;
      TexSrc& S = X.s[g][d];
      S.nb = (int)nb;
      S.rows_n = (int)rows_n;
      S.rows_t = Tn > 1 ? (int)(C * p->H) : 0;
      for (long long b = 0; b < nblk; ++b) {
        const long long n0 = b * nb, cnt = (p->N - n0 < nb) ? p->N - n0 : nb;
        if (!tex_get(R.src[d] + n0 * sn, p->W, (int)(cnt * rows_n), (size_t)sh * 4, dev, &S.tex[b])) return false;
      }
    }
  return true;
}

}  // namespace fwb

using namespace fwb;

// zero every grad_src plane present (memset when contiguous, a row kernel otherwise)
static int32_t zero_grad_src(const fwb_problem* p, const GradP& Q, cudaStream_t s) {
  for (int gi = 0; gi < p->n_groups; ++gi)
    for (int d = 0; d < p->n_dirs; ++d) {
      float* gs = Q.grad_src[gi][d];
      if (!gs) continue;
      const int Tn = Q.gs_st[gi][d] == 0 ? 1 : p->T;
      const long long C = p->grp[gi].C, HW = (long long)p->H * p->W;
      if (Q.gs_sh[gi][d] == p->W && Q.gs_sc[gi][d] == HW && (Tn == 1 || Q.gs_st[gi][d] == C * HW) &&
          (p->N == 1 || Q.gs_sn[gi][d] == Tn * C * HW)) {  // contiguous: one memset at full write bandwidth
        const int32_t rc = (int32_t)cudaMemsetAsync(gs, 0, (size_t)(p->N * Tn * C * HW) * sizeof(float), s);
        if (rc) return rc;
        continue;
      }
      const dim3 zg((unsigned)(p->N * Tn * p->grp[gi].C), (unsigned)(p->H < 64 ? p->H : 64));
      zero_rows_kernel<<<zg, 256, 0, s>>>(gs, Q.gs_sn[gi][d], Q.gs_st[gi][d], Q.gs_sc[gi][d], Q.gs_sh[gi][d], p->N, Tn,
                                          p->grp[gi].C, p->H, p->W);
    }
  return (int32_t)cudaGetLastError();
}

static int32_t run_backward_src(const fwb_problem* p, const fwb_grads* g, void* workspace, size_t workspace_bytes,
                                void* stream) {
  int rc = validate(p);
  if (rc) return rc;
  Params P;
  GradP Q;
  to_params(p, P);
  rc = to_grads(p, g, Q);
  if (rc) return rc;
  if (p->N == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int NT = p->N * p->T;
  if (!(p->flags & FWB_FLAG_ATOMIC_SRC)) {
    // owner gather (deterministic): needs the tables the backward_flow call left in the workspace
    const WsLayout L = ws_layout(p->n_dirs, NT, p->H, p->W);
    if (!workspace || workspace_bytes < L.total || ((uintptr_t)workspace & 15u)) return FWB_E_WORKSPACE;
    const WsView ws = ws_view(workspace, L, NT, p->H, p->W);
    if ((rc = set_smem(bwd_src_owner_kernel, (int)own_smem_bytes(OWN_MAXGRP)))) return rc;
    for (int d = 0; d < p->n_dirs; ++d)
      for (int shared = 0; shared < 2; ++shared) {
        // channel runs of the groups that want grad_src for this direction with this T-sharing,
        // <= 4*OWN_MAXGRP channels per launch
        OwnArgs A = {};
        A.d = d;
        A.tshared = shared;
        auto flush = [&]() -> int32_t {
          if (A.nseg == 0) return 0;
          A.ngrp = (A.nchan + 3) / 4;
          const int nwarps = A.ngrp < 4 ? 4 : A.ngrp;
          const dim3 grid((p->W + OT_W - 1) / OT_W, (p->H + OT_H - 1) / OT_H, shared ? p->N : NT);
          bwd_src_owner_kernel<<<grid, nwarps * 32, own_smem_bytes(A.ngrp), s>>>(P, Q, ws, A);
          A.nseg = 0;
          A.nchan = 0;
          return (int32_t)cudaGetLastError();
        };
        for (int gi = 0; gi < p->n_groups; ++gi) {
          if (!Q.grad_src[gi][d] || !Q.grad_out[gi]) continue;
          const int is_shared = (Q.gs_st[gi][d] == 0 && p->T > 1) ? 1 : 0;
          if (is_shared != shared) continue;
          int c0 = 0;
          while (c0 < p->grp[gi].C) {
            const int take = min(p->grp[gi].C - c0, 4 * OWN_MAXGRP - A.nchan);
            A.seg[A.nseg].g = gi;
            A.seg[A.nseg].c0 = c0;
            A.seg[A.nseg].c1 = c0 + take;
            A.nseg++;
            A.nchan += take;
            c0 += take;
            if (A.nchan == 4 * OWN_MAXGRP || A.nseg == FWB_MAX_GROUPS)
              if ((rc = flush())) return rc;
          }
        }
        if ((rc = flush())) return rc;
      }
    // a group whose grad_out is NULL contributes nothing: its grad_src is zero
    for (int gi = 0; gi < p->n_groups; ++gi)
      for (int d = 0; d < p->n_dirs; ++d)
        if (Q.grad_src[gi][d] && !Q.grad_out[gi]) {
          const int Tn = Q.gs_st[gi][d] == 0 ? 1 : p->T;
          const dim3 zg((unsigned)(p->N * Tn * p->grp[gi].C), (unsigned)(p->H < 64 ? p->H : 64));
          zero_rows_kernel<<<zg, 256, 0, s>>>(Q.grad_src[gi][d], Q.gs_sn[gi][d], Q.gs_st[gi][d], Q.gs_sc[gi][d],
                                              Q.gs_sh[gi][d], p->N, Tn, p->grp[gi].C, p->H, p->W);
        }
    return (int32_t)cudaGetLastError();
  }
  // ---- global-atomic scatter (ATen-style; non-deterministic; kept for A/B measurements)
  for (int gi = 0; gi < p->n_groups; ++gi)
    for (int d = 0; d < p->n_dirs; ++d) {
      float* gs = Q.grad_src[gi][d];
      if (!gs) continue;
      const int Tn = Q.gs_st[gi][d] == 0 ? 1 : p->T;
      const dim3 zg((unsigned)(p->N * Tn * p->grp[gi].C), (unsigned)(p->H < 64 ? p->H : 64));
      zero_rows_kernel<<<zg, 256, 0, s>>>(gs, Q.gs_sn[gi][d], Q.gs_st[gi][d], Q.gs_sc[gi][d], Q.gs_sh[gi][d],
                                          p->N, Tn, p->grp[gi].C, p->H, p->W);
    }
  const dim3 grid = pixel_grid(p), block(NTHREADS);
  if (p->n_dirs == 2)
    bwd_src_atomic_kernel<2><<<grid, block, 0, s>>>(P, Q);
  else
    bwd_src_atomic_kernel<1><<<grid, block, 0, s>>>(P, Q);
  return (int32_t)cudaGetLastError();
}

extern "C" {

int32_t fwb_version(void) { return FWB_VERSION; }

void fwb_reload_env(void) {
  std::lock_guard<std::mutex> lk(g_env_mu);
  env_load_locked();
}

void fwb_release_cache(void) {
  std::lock_guard<std::mutex> lk(g_tex_mu);
  for (const TexEntry& e : g_tex) (void)cudaDestroyTextureObject(e.obj);
  g_tex.clear();
}

const char* fwb_strerror(int32_t code) {
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case 0: return "ok";
    case FWB_E_NULL: return "a required pointer is NULL";
    case FWB_E_SHAPE: return "N/T/H/W/C out of range (empty spatial dims are an error)";
    case FWB_E_DIRS: return "n_dirs must be 1 or 2";
    case FWB_E_GROUPS: return "n_groups must be in 1..FWB_MAX_GROUPS";
    case FWB_E_MODE: return "unknown padding_mode / align_corners / sign";
    case FWB_E_ALIGN: return "pointer not 4-byte aligned";
    case FWB_E_WORKSPACE: return "workspace missing or too small";
    case FWB_E_RANGE: return "H or W above 32767";
    default: return "unknown error";
  }
}

}  // extern "C"

// the forward, optionally with the zero-fill of zq->grad_src as a side job (zq == NULL: plain forward)
static int32_t run_forward(const fwb_problem* p, const fwb_grads* zq, void* stream) {
  int rc = validate(p);
  if (rc) return rc;
  for (int g = 0; g < p->n_groups; ++g) {
    if (!p->grp[g].out) return FWB_E_NULL;
    if ((uintptr_t)p->grp[g].out & 3u) return FWB_E_ALIGN;
  }
  Params P;
  to_params(p, P);
  GradP Q;
  if (zq && (rc = to_grads(p, zq, Q))) return rc;
  if (p->N == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  TexP X;
  const bool tex = tex_prepare(p, X);
  const bool tile = !tex && !(knobs() & KN_NOTILE) && stage_ok(p) && tile_fwd_ok(p);
  ZeroP Z;
  memset(&Z, 0, sizeof(Z));
  if (zq) {
    // in-kernel zero-fill needs 16-byte aligned rows and one row stride per direction; anything else: memsets first
    bool fuse = (tex || tile) && !(p->W & 3) && !(knobs() & KN_NOZFUSE), any = false;
    int sh[2] = {0, 0};
    for (int g = 0; g < p->n_groups; ++g)
      for (int d = 0; d < p->n_dirs; ++d) {
        float* gs = Q.grad_src[g][d];
        if (!gs) continue;
        any = true;
        if (((uintptr_t)gs & 15u) || (Q.gs_sn[g][d] & 3) || (Q.gs_st[g][d] & 3) || (Q.gs_sc[g][d] & 3) || (Q.gs_sh[g][d] & 3)) fuse = false;
        if (sh[d] == 0) sh[d] = Q.gs_sh[g][d];
        if (sh[d] != Q.gs_sh[g][d] || (long long)(p->H + 1) * Q.gs_sh[g][d] >= 2147483647LL) fuse = false;
        Z.gs[g][d] = gs;
        Z.sn[g][d] = Q.gs_sn[g][d];
        Z.st[g][d] = Q.gs_st[g][d];
        Z.sc[g][d] = Q.gs_sc[g][d];
      }
    Z.sh[0] = sh[0];
    Z.sh[1] = sh[1];
    Z.on = (fuse && any) ? 1 : 0;
    if (any && !Z.on && (rc = zero_grad_src(p, Q, s))) return rc;
  }
  if (tex) {
    const dim3 grid((p->W + TXF_TW - 1) / TXF_TW, (p->H + TXF_TH - 1) / TXF_TH, p->N * p->T);
    if (p->n_dirs == 2)
      fwd_tex_kernel<2><<<grid, TXF_THREADS, 0, s>>>(P, X, Z, total_channels(p));
    else
      fwd_tex_kernel<1><<<grid, TXF_THREADS, 0, s>>>(P, X, Z, total_channels(p));
    return (int32_t)cudaGetLastError();
  }
  if (tile) {
    const int sb = env().tile_fwd_kb * 1024;
    const int Ctot = total_channels(p);
    const dim3 grid((p->W + TL_TW - 1) / TL_TW, (p->H + TL_TH - 1) / TL_TH, p->N * p->T);
#define FWB_LAUNCH_TFWD(D, A, B)                                                     \
  do {                                                                               \
    if ((rc = set_smem(fwd_tile_kernel<D, A, B>, sb))) return rc;                    \
    fwd_tile_kernel<D, A, B><<<grid, TL_THREADS, sb, s>>>(P, sb / 4, Ctot, Z);       \
  } while (0)
    const int key = (p->n_dirs == 2 ? 4 : 0) | (p->align_corners ? 2 : 0) | (p->padding_mode == FWB_PAD_BORDER ? 1 : 0);
    switch (key) {
      case 0: FWB_LAUNCH_TFWD(1, false, false); break;
      case 1: FWB_LAUNCH_TFWD(1, false, true); break;
      case 2: FWB_LAUNCH_TFWD(1, true, false); break;
      case 3: FWB_LAUNCH_TFWD(1, true, true); break;
      case 4: FWB_LAUNCH_TFWD(2, false, false); break;
      case 5: FWB_LAUNCH_TFWD(2, false, true); break;
      case 6: FWB_LAUNCH_TFWD(2, true, false); break;
      default: FWB_LAUNCH_TFWD(2, true, true); break;
    }
#undef FWB_LAUNCH_TFWD
    return (int32_t)cudaGetLastError();
  }
  const dim3 grid = pixel_grid(p), block(NTHREADS);
  if (p->n_dirs == 2)
    fwd_kernel<2><<<grid, block, 0, s>>>(P);
  else
    fwd_kernel<1><<<grid, block, 0, s>>>(P);
  return (int32_t)cudaGetLastError();
}

extern "C" {

int32_t fwb_warp_blend_forward(const fwb_problem* p, void* stream) { return run_forward(p, nullptr, stream); }

int32_t fwb_warp_blend_forward_zero(const fwb_problem* p, const fwb_grads* g, void* stream) {
  if (!g) return FWB_E_NULL;
  return run_forward(p, g, stream);
}

static int label_params(const fwb_label_problem* p, bool bwd, LabelP& L) {
  if (!p) return FWB_E_NULL;
  if (p->N < 0 || p->T < 1 || p->H < 1 || p->W < 1 || p->K < 1 || p->K > 256) return FWB_E_SHAPE;
  if ((long long)p->N * p->T > 65535) return FWB_E_SHAPE;
  if (p->H > 32767 || p->W > 32767) return FWB_E_RANGE;
  if (p->n_dirs < 1 || p->n_dirs > 2) return FWB_E_DIRS;
  if (p->padding_mode != FWB_PAD_ZEROS && p->padding_mode != FWB_PAD_BORDER) return FWB_E_MODE;
  if (p->align_corners != 0 && p->align_corners != 1) return FWB_E_MODE;
  fwb_problem q;
  memset(&q, 0, sizeof(q));
  q.N = p->N, q.T = p->T, q.H = p->H, q.W = p->W, q.n_dirs = p->n_dirs, q.n_groups = 1;
  q.padding_mode = p->padding_mode, q.align_corners = p->align_corners;
  for (int d = 0; d < p->n_dirs; ++d) {
    if (!p->dir[d].flow || !p->labels[d]) return FWB_E_NULL;
    if (p->dir[d].sign != 1.0f && p->dir[d].sign != -1.0f) return FWB_E_MODE;
    if (((uintptr_t)p->dir[d].flow | (uintptr_t)p->dir[d].gate | (uintptr_t)p->dir[d].blend) & 3u) return FWB_E_ALIGN;
    if (p->lab_sh[d] < 0 || (long long)(p->H + 2) * p->lab_sh[d] > 2147483647LL) return FWB_E_SHAPE;
    q.dir[d] = p->dir[d];
  }
  Params P;
  to_params(&q, P);
  L.geo = P.geo;
  L.K = p->K;
  for (int d = 0; d < 2; ++d) {
    const int e = d < p->n_dirs ? d : 0;
    L.dir[d] = P.dir[d];
    L.lab[d] = p->labels[e];
    L.lab_sn[d] = p->lab_sn[e], L.lab_st[d] = p->lab_st[e], L.lab_sh[d] = (int)p->lab_sh[e];
    const bool on = d < p->n_dirs;
    L.grad_flow[d] = on ? p->grad_flow[d] : nullptr;
    L.gf_sn[d] = p->gf_sn[d], L.gf_sc[d] = p->gf_sc[d], L.gf_st[d] = p->gf_st[d], L.gf_sh[d] = p->gf_sh[d];
    L.grad_gate[d] = on ? p->grad_gate[d] : nullptr;
    L.gg_sn[d] = p->gg_sn[d], L.gg_st[d] = p->gg_st[d], L.gg_sh[d] = p->gg_sh[d];
    L.grad_blend[d] = on ? p->grad_blend[d] : nullptr;
    L.gb_sn[d] = p->gb_sn[d], L.gb_st[d] = p->gb_st[d], L.gb_sh[d] = p->gb_sh[d];
    if (((uintptr_t)L.grad_flow[d] | (uintptr_t)L.grad_gate[d] | (uintptr_t)L.grad_blend[d]) & 3u) return FWB_E_ALIGN;
  }
  L.out = p->out, L.out_sn = p->out_sn, L.out_st = p->out_st, L.out_sc = (int)p->out_sc, L.out_sh = (int)p->out_sh;
  L.go = p->grad_out, L.go_sn = p->go_sn, L.go_st = p->go_st, L.go_sc = (int)p->go_sc, L.go_sh = (int)p->go_sh;
  L.accumulate = p->accumulate;
  if (!bwd) {
    if (!p->out) return FWB_E_NULL;
    if ((uintptr_t)p->out & 3u) return FWB_E_ALIGN;
    if (p->out_sc < 0 || p->out_sh < 0 || (long long)p->K * p->out_sc + (long long)(p->H + 2) * p->out_sh > 2147483647LL) return FWB_E_SHAPE;
  } else {
    if (!p->grad_out) return FWB_E_NULL;
    if ((uintptr_t)p->grad_out & 3u) return FWB_E_ALIGN;
    if (p->go_sc < 0 || p->go_sh < 0 || (long long)p->K * p->go_sc + (long long)(p->H + 2) * p->go_sh > 2147483647LL) return FWB_E_SHAPE;
  }
  return 0;
}

int32_t fwb_label_warp_blend_forward(const fwb_label_problem* p, void* stream) {
  LabelP L;
  int rc = label_params(p, false, L);
  if (rc) return rc;
  if (p->N == 0) return 0;
  const dim3 grid((p->W + 31) / 32, (p->H + 7) / 8, p->N * p->T);
  if (p->n_dirs == 2)
    label_fwd_kernel<2><<<grid, LB_THREADS, 0, (cudaStream_t)stream>>>(L);
  else
    label_fwd_kernel<1><<<grid, LB_THREADS, 0, (cudaStream_t)stream>>>(L);
  return (int32_t)cudaGetLastError();
}

int32_t fwb_label_warp_blend_backward(const fwb_label_problem* p, void* stream) {
  LabelP L;
  int rc = label_params(p, true, L);
  if (rc) return rc;
  if (p->N == 0) return 0;
  const dim3 grid((p->W + 31) / 32, (p->H + 7) / 8, p->N * p->T);
  if (p->n_dirs == 2)
    label_bwd_kernel<2><<<grid, LB_THREADS, 0, (cudaStream_t)stream>>>(L);
  else
    label_bwd_kernel<1><<<grid, LB_THREADS, 0, (cudaStream_t)stream>>>(L);
  return (int32_t)cudaGetLastError();
}

int32_t fwb_mask_blend_forward(const fwb_blend* b, void* stream) {
  int rc = blend_validate(b, false);
  if (rc) return rc;
  if (b->N == 0) return 0;
  BlendP B;
  blend_params(b, B);
  cudaStream_t s = (cudaStream_t)stream;
  if (blend_vec_ok(b, false))
    mask_blend_fwd_kernel<4><<<dim3((b->W / 4 + 127) / 128, b->H, b->N * b->T), 128, 0, s>>>(B);
  else
    mask_blend_fwd_kernel<1><<<dim3((b->W + 127) / 128, b->H, b->N * b->T), 128, 0, s>>>(B);
  return (int32_t)cudaGetLastError();
}

int32_t fwb_mask_blend_backward(const fwb_blend* b, void* stream) {
  int rc = blend_validate(b, true);
  if (rc) return rc;
  if (b->N == 0 || (!b->grad_input && !b->grad_mask && !b->grad_noise)) return 0;
  BlendP B;
  blend_params(b, B);
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = blend_vec_ok(b, true);
  if (B.gi || B.gm || (B.gn && B.T == 1)) {
    if (vec)
      mask_blend_bwd_kernel<4><<<dim3((b->W / 4 + 127) / 128, b->H, b->N * b->T), 128, 0, s>>>(B);
    else
      mask_blend_bwd_kernel<1><<<dim3((b->W + 127) / 128, b->H, b->N * b->T), 128, 0, s>>>(B);
  }
  if (B.gn && B.T > 1 && B.Cn > 0) {
    if (vec)
      mask_blend_gnoise_kernel<4><<<dim3((b->W / 4 + 127) / 128, b->H, b->N * B.Cn), 128, 0, s>>>(B);
    else
      mask_blend_gnoise_kernel<1><<<dim3((b->W + 127) / 128, b->H, b->N * B.Cn), 128, 0, s>>>(B);
  }
  return (int32_t)cudaGetLastError();
}

int32_t fwb_sample_indices(const fwb_problem* p, int32_t d, int32_t* x0, int32_t* y0, uint8_t* valid,
                           float* ix, float* iy, void* stream) {
  int rc = validate(p);
  if (rc) return rc;
  if (d < 0 || d >= p->n_dirs) return FWB_E_DIRS;
  if (p->N == 0) return 0;
  Params P;
  to_params(p, P);
  indices_kernel<<<pixel_grid(p), dim3(NTHREADS), 0, (cudaStream_t)stream>>>(P, d, x0, y0, valid, ix, iy);
  return (int32_t)cudaGetLastError();
}

size_t fwb_workspace_bytes(const fwb_problem* p) {
  if (validate(p)) return 0;
  return ws_layout(p->n_dirs, (long long)p->N * p->T, p->H, p->W).total;
}

int32_t fwb_warp_blend_backward_flow(const fwb_problem* p, const fwb_grads* g, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  int rc = validate(p);
  if (rc) return rc;
  Params P;
  GradP Q;
  to_params(p, P);
  rc = to_grads(p, g, Q);
  if (rc) return rc;
  if (p->N == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int NT = p->N * p->T;
  const bool fused_req = (p->flags & FWB_FLAG_FUSED_BWD) && !(p->flags & (FWB_FLAG_DETERMINISTIC | FWB_FLAG_ATOMIC_SRC));
  if (fused_req) {
    // any_src: a group wants grad_src AND has a grad_out (it scatters); any_gs: some grad_src plane exists at all (a plane
    // whose group has no grad_out still has to come back as zeros)
    bool any_src = false, any_gs = false, ok = stage_ok(p) && !(knobs() & KN_NOFUSE);
    for (int gi = 0; gi < p->n_groups; ++gi)
      for (int d = 0; d < p->n_dirs; ++d) {
        if (!Q.grad_src[gi][d]) continue;
        any_gs = true;
        if (!Q.grad_out[gi]) continue;
        any_src = true;
        if (((uintptr_t)Q.grad_src[gi][d] & 15u) || (Q.gs_sn[gi][d] & 3) || (Q.gs_st[gi][d] & 3) || (Q.gs_sc[gi][d] & 3) ||
            (Q.gs_sh[gi][d] & 3))
          ok = false;  // red.global.add.v4.f32 needs 16-byte aligned rows
      }
    // a backward without any grad_src (the sources are data: the usual training case) takes the tile kernel too: it then
    // neither scatters nor flushes, i.e. it is kernel 2 on the staged tiles
    const bool tile_ok = !(knobs() & KN_NOTILE) && tile_bwd_ok(p, g);
    if (ok && tile_ok) {
      // grad_src is accumulated with reductions: zero every plane first (also those of groups without grad_out) unless the
      // caller did (FWB_FLAG_GRAD_SRC_ZEROED, e.g. fwb_warp_blend_forward_zero)
      if (any_gs && !(p->flags & FWB_FLAG_GRAD_SRC_ZEROED) && (rc = zero_grad_src(p, Q, s))) return rc;
      if (tile_ok) {
        const int ppt = env().tile_bwd_ppt;  // pixels per thread: 1 = 32x8 tiles, 3 CTAs/SM; 2 = 32x16 tiles, 2 CTAs/SM
        TexP X;
        const bool tex = ppt != 1 && tex_prepare(p, X);
        if (!tex) memset(&X, 0, sizeof(X));
        if (tex && (!any_src || (knobs() & KN_SORTED))) {
          // no source gradient wanted (the sources are data: the usual training case): kernel 2 alone, on the texture units.
          // "sorted" (A/B): kernel 2 on the texture units, then kernel 3 as a sorted shared-memory gather (fwb_bwdx.cuh)
          int want = 0;
          for (int d = 0; d < p->n_dirs; ++d) want |= (Q.grad_flow[d] || Q.grad_gate[d] || Q.grad_blend[d]);
          if (want) {
            const dim3 g2((p->W + TXF_TW - 1) / TXF_TW, (p->H + TXF_TH - 1) / TXF_TH, p->N * p->T);
            if (p->n_dirs == 2)
              bwd_flow_tex_kernel<2><<<g2, TXF_THREADS, 0, s>>>(P, Q, X);
            else
              bwd_flow_tex_kernel<1><<<g2, TXF_THREADS, 0, s>>>(P, Q, X);
            if ((rc = (int32_t)cudaGetLastError())) return rc;
          }
          if (any_src) {
            const int sbx = env().tile_bwdx_kb * 1024;
            const int xp = env().bwdx_ppt == 1 ? 1 : 2;
            const dim3 g3((p->W + TL_TW - 1) / TL_TW, (p->H + 8 * xp - 1) / (8 * xp), p->N * p->T);
#define FWB_LAUNCH_SRT(D, A, B)                                                            \
  do {                                                                                     \
    if (xp == 1) {                                                                         \
      if ((rc = set_smem(bwd_src_sorted_kernel<D, A, B, 1>, sbx))) return rc;              \
      bwd_src_sorted_kernel<D, A, B, 1><<<g3, BX_THREADS, sbx, s>>>(P, Q, sbx / 4);        \
    } else {                                                                               \
      if ((rc = set_smem(bwd_src_sorted_kernel<D, A, B, 2>, sbx))) return rc;              \
      bwd_src_sorted_kernel<D, A, B, 2><<<g3, BX_THREADS, sbx, s>>>(P, Q, sbx / 4);        \
    }                                                                                      \
  } while (0)
            const int key3 = (p->n_dirs == 2 ? 4 : 0) | (p->align_corners ? 2 : 0) | (p->padding_mode == FWB_PAD_BORDER ? 1 : 0);
            const int mc = env_int("FWB_BX_MINCTA", 2);  // A/B (temporary): occupancy variants of the headline instantiation
            if (key3 == 5 && mc != 2) {
              if (xp == 2 && mc == 3) {
                if ((rc = set_smem(bwd_src_sorted_kernel<2, false, true, 2, 3>, sbx))) return rc;
                bwd_src_sorted_kernel<2, false, true, 2, 3><<<g3, BX_THREADS, sbx, s>>>(P, Q, sbx / 4);
              } else if (xp == 1 && mc == 3) {
                if ((rc = set_smem(bwd_src_sorted_kernel<2, false, true, 1, 3>, sbx))) return rc;
                bwd_src_sorted_kernel<2, false, true, 1, 3><<<g3, BX_THREADS, sbx, s>>>(P, Q, sbx / 4);
              } else if (xp == 1 && mc == 4) {
                if ((rc = set_smem(bwd_src_sorted_kernel<2, false, true, 1, 4>, sbx))) return rc;
                bwd_src_sorted_kernel<2, false, true, 1, 4><<<g3, BX_THREADS, sbx, s>>>(P, Q, sbx / 4);
              } else {
                if ((rc = set_smem(bwd_src_sorted_kernel<2, false, true, 2, 4>, sbx))) return rc;
                bwd_src_sorted_kernel<2, false, true, 2, 4><<<g3, BX_THREADS, sbx, s>>>(P, Q, sbx / 4);
              }
              return (int32_t)cudaGetLastError();
            }
            switch (key3) {
              case 0: FWB_LAUNCH_SRT(1, false, false); break;
              case 1: FWB_LAUNCH_SRT(1, false, true); break;
              case 2: FWB_LAUNCH_SRT(1, true, false); break;
              case 3: FWB_LAUNCH_SRT(1, true, true); break;
              case 4: FWB_LAUNCH_SRT(2, false, false); break;
              case 5: FWB_LAUNCH_SRT(2, false, true); break;
              case 6: FWB_LAUNCH_SRT(2, true, false); break;
              default: FWB_LAUNCH_SRT(2, true, true); break;
            }
#undef FWB_LAUNCH_SRT
          }
          return (int32_t)cudaGetLastError();
        }
        int Cgo = 0;
        for (int gi = 0; gi < p->n_groups; ++gi)
          if (Q.grad_out[gi]) Cgo += p->grp[gi].C;
        if (tex && any_src && !(knobs() & KN_NOCL) && Cgo >= env().cl_minc && Cgo <= CL_MAXC) {
          // channel-per-lane scatter + kernel 2 on the texture units, warp-specialised (fwb_cl.cuh): the default for wide
          // channel sets (lanes = channels: a narrow set leaves most lanes idle, the pixel-per-lane kernel below serves it)
          const int gos_b = 4 * ((Cgo * CL_GS + 3) & ~3);
          const int dyn = env().cl_kb * 1024;
          const int acc_words = (dyn - gos_b - 16 - 4 * CL_SLOTS * (CL_NPIX / 2)) / 4;  // slack after the last plane
          bool cl_fits = acc_words >= (Cgo + 1) * (CL_ZPAD + 64);
          for (int gi = 0; gi < p->n_groups; ++gi)
            if (Q.grad_out[gi] && (long long)Q.go_sc[gi] * p->grp[gi].C >= 2147483647LL) cl_fits = false;
          for (int gi = 0; gi < p->n_groups; ++gi)
            for (int d = 0; d < p->n_dirs; ++d)
              if (Q.grad_src[gi][d] && ((long long)Q.gs_sc[gi][d] * (p->grp[gi].C + 1) + (long long)(p->H + 1) * Q.gs_sh[gi][d]) >= 2147483647LL) cl_fits = false;
          if (cl_fits) {
            const dim3 gc((p->W + CL_TW - 1) / CL_TW, (p->H + CL_TH - 1) / CL_TH, p->N * p->T);
            int go16 = !(p->W & 3);  // 16-byte staging of the grad_out tile: every plane row 16-byte aligned
            for (int gi = 0; gi < p->n_groups; ++gi)
              if (Q.grad_out[gi] && (((uintptr_t)Q.grad_out[gi] & 15u) || (Q.go_sn[gi] & 3) || (Q.go_st[gi] & 3) || (Q.go_sc[gi] & 3) || (Q.go_sh[gi] & 3)))
                go16 = 0;
#define FWB_LAUNCH_CL(D, A, B)                                                \
  do {                                                                        \
    if ((rc = set_smem(bwd_cl_kernel<D, A, B>, dyn))) return rc;              \
    bwd_cl_kernel<D, A, B><<<gc, CL_THREADS, dyn, s>>>(P, Q, X, acc_words, go16); \
  } while (0)
            const int keyc = (p->n_dirs == 2 ? 4 : 0) | (p->align_corners ? 2 : 0) | (p->padding_mode == FWB_PAD_BORDER ? 1 : 0);
            switch (keyc) {
              case 0: FWB_LAUNCH_CL(1, false, false); break;
              case 1: FWB_LAUNCH_CL(1, false, true); break;
              case 2: FWB_LAUNCH_CL(1, true, false); break;
              case 3: FWB_LAUNCH_CL(1, true, true); break;
              case 4: FWB_LAUNCH_CL(2, false, false); break;
              case 5: FWB_LAUNCH_CL(2, false, true); break;
              case 6: FWB_LAUNCH_CL(2, true, false); break;
              default: FWB_LAUNCH_CL(2, true, true); break;
            }
#undef FWB_LAUNCH_CL
            return (int32_t)cudaGetLastError();
          }
        }
        const int sb = env().tile_bwd_kb * 1024;
        const dim3 grid((p->W + TL_TW - 1) / TL_TW, (p->H + 8 * ppt - 1) / (8 * ppt), p->N * p->T);
#define FWB_LAUNCH_TBWD(D, A, B)                                                                  \
  do {                                                                                            \
    if (ppt == 1) {                                                                               \
      if ((rc = set_smem(bwd_tile_kernel<D, A, B, 1, 2>, sb))) return rc;                         \
      bwd_tile_kernel<D, A, B, 1, 2><<<grid, TL_THREADS, sb, s>>>(P, Q, sb / 4, X);               \
    } else if (!any_src) { /* flow-only backward: no accumulators, 3 CTAs/SM */                   \
      const int sbn = env().tile_bwdf_kb * 1024;                                                  \
      if ((rc = set_smem(bwd_tile_kernel<D, A, B, 2, 3, TL_THREADS, false>, sbn))) return rc;     \
      bwd_tile_kernel<D, A, B, 2, 3, TL_THREADS, false><<<grid, TL_THREADS, sbn, s>>>(P, Q, sbn / 4, X); \
    } else if (tex) { /* texture gather + integer-atomic scatter: shared memory holds the two accumulators only */ \
      const int sbx = env().tile_bwdx_kb * 1024;                                                  \
      if ((rc = set_smem(bwd_tile_kernel<D, A, B, 2, 3, TL_THREADS, true, true>, sbx))) return rc; \
      bwd_tile_kernel<D, A, B, 2, 3, TL_THREADS, true, true><<<grid, TL_THREADS, sbx, s>>>(P, Q, sbx / 4, X); \
    } else {                                                                                      \
      if ((rc = set_smem(bwd_tile_kernel<D, A, B, 2, 3>, sb))) return rc;                         \
      bwd_tile_kernel<D, A, B, 2, 3><<<grid, TL_THREADS, sb, s>>>(P, Q, sb / 4, X);               \
    }                                                                                             \
  } while (0)
        const int key = (p->n_dirs == 2 ? 4 : 0) | (p->align_corners ? 2 : 0) | (p->padding_mode == FWB_PAD_BORDER ? 1 : 0);
        switch (key) {
          case 0: FWB_LAUNCH_TBWD(1, false, false); break;
          case 1: FWB_LAUNCH_TBWD(1, false, true); break;
          case 2: FWB_LAUNCH_TBWD(1, true, false); break;
          case 3: FWB_LAUNCH_TBWD(1, true, true); break;
          case 4: FWB_LAUNCH_TBWD(2, false, false); break;
          case 5: FWB_LAUNCH_TBWD(2, false, true); break;
          case 6: FWB_LAUNCH_TBWD(2, true, false); break;
          default: FWB_LAUNCH_TBWD(2, true, true); break;
        }
#undef FWB_LAUNCH_TBWD
        return (int32_t)cudaGetLastError();
      }
    }
  }
  // (not fused, or the tile kernel cannot take this problem: split path — kernel 2, then kernel 3 below)
  // the segment tables kernel 3 needs are produced whenever a workspace is supplied
  if (workspace && !(p->flags & FWB_FLAG_ATOMIC_SRC)) {
    const WsLayout L = ws_layout(p->n_dirs, NT, p->H, p->W);
    if (workspace_bytes < L.total || ((uintptr_t)workspace & 15u)) return FWB_E_WORKSPACE;
    const WsView ws = ws_view(workspace, L, NT, p->H, p->W);
    const int nh = p->n_dirs * NT;
    ws_init_kernel<<<(nh + 255) / 256, 256, 0, s>>>(ws.hdr, nh);
    const dim3 eg((p->W + 31) / 32, (p->H + 7) / 8, NT);
    if (p->n_dirs == 2)
      emit_kernel<2><<<eg, 256, 0, s>>>(P, ws);
    else
      emit_kernel<1><<<eg, 256, 0, s>>>(P, ws);
  }
  int want = 0;
  for (int d = 0; d < p->n_dirs; ++d) want |= (Q.grad_flow[d] || Q.grad_gate[d] || Q.grad_blend[d]);
  if (want) {
    TexP X2;
    if (tex_prepare(p, X2)) {  // kernel 2 on the texture units (dense sources), also in the deterministic split path
      const dim3 g2((p->W + TXF_TW - 1) / TXF_TW, (p->H + TXF_TH - 1) / TXF_TH, p->N * p->T);
      if (p->n_dirs == 2)
        bwd_flow_tex_kernel<2><<<g2, TXF_THREADS, 0, s>>>(P, Q, X2);
      else
        bwd_flow_tex_kernel<1><<<g2, TXF_THREADS, 0, s>>>(P, Q, X2);
    } else {
      const dim3 grid = pixel_grid(p), block(NTHREADS);
      if (p->n_dirs == 2)
        bwd_flow_kernel<2><<<grid, block, 0, s>>>(P, Q);
      else
        bwd_flow_kernel<1><<<grid, block, 0, s>>>(P, Q);
    }
  }
  if ((rc = (int32_t)cudaGetLastError())) return rc;
  if (fused_req) {  // the fused kernel could not take this problem: finish grad_src here, as the flag promises
    bool any_src = false;
    for (int gi = 0; gi < p->n_groups; ++gi)
      for (int d = 0; d < p->n_dirs; ++d) any_src |= Q.grad_src[gi][d] != nullptr;
    if (any_src) return run_backward_src(p, g, workspace, workspace_bytes, stream);
  }
  return 0;
}

int32_t fwb_warp_blend_backward_src(const fwb_problem* p, const fwb_grads* g, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  // with FWB_FLAG_FUSED_BWD the backward_flow call has already produced grad_src
  if (p && (p->flags & FWB_FLAG_FUSED_BWD) && !(p->flags & (FWB_FLAG_DETERMINISTIC | FWB_FLAG_ATOMIC_SRC))) {
    const int rc = validate(p);
    return rc;
  }
  return run_backward_src(p, g, workspace, workspace_bytes, stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// flow-regularisation losses (fwb_loss.cuh)
// ---------------------------------------------------------------------------------------------
static LossView to_view(const fwb_view* v) {
  LossView o = {nullptr, 0, 0, 0, 0};
  if (v && v->ptr) o = {v->ptr, v->sn, v->st, v->sc, v->sh};
  return o;
}
static int32_t loss_dims_ok(int32_t N, int32_t T, int32_t C, int32_t H, int32_t W) {
  if (N < 0 || T < 1 || C < 1 || H < 1 || W < 1) return FWB_E_SHAPE;
  if ((long long)N * T > 65535 || H > 65535) return FWB_E_SHAPE;
  return 0;
}

extern "C" {

size_t fwb_loss_partials_bytes(int32_t N, int32_t T, int32_t H) {
  if (N < 1 || T < 1 || H < 1) return sizeof(float2);
  return (size_t)N * T * H * sizeof(float2);
}

int32_t fwb_flowgrad_loss_forward(const fwb_view* flow, const fwb_view* image, int32_t N, int32_t T, int32_t C, int32_t H,
                                  int32_t W, void* partials, float* loss, void* stream) {
  int32_t rc = loss_dims_ok(N, T, C, H, W);
  if (rc) return rc;
  if (!flow || !flow->ptr || !image || !image->ptr || !partials || !loss) return FWB_E_NULL;
  cudaStream_t s = (cudaStream_t)stream;
  if (N == 0) return (int32_t)cudaMemsetAsync(loss, 0, sizeof(float), s);
  flowgrad_fwd_kernel<<<dim3(H, N * T), LS_THREADS, 0, s>>>(to_view(flow), to_view(image), T, C, H, W, (float2*)partials);
  // per frame: mean over N*2*H*(W-1) and N*2*(H-1)*W elements; the frames are summed and divided by T
  const double cx = (double)N * 2.0 * H * (W - 1), cy = (double)N * 2.0 * (H - 1) * W;
  loss_final_kernel<<<1, LS_THREADS, 0, s>>>((const float2*)partials, (long long)N * T * H, cx > 0 ? 1.0 / (cx * T) : 0.0,
                                             cy > 0 ? 1.0 / (cy * T) : 0.0, loss);
  return (int32_t)cudaGetLastError();
}

int32_t fwb_flowgrad_loss_backward(const fwb_view* flow, const fwb_view* image, int32_t N, int32_t T, int32_t C, int32_t H,
                                   int32_t W, const float* grad_loss, const fwb_view* grad_flow, void* stream) {
  int32_t rc = loss_dims_ok(N, T, C, H, W);
  if (rc) return rc;
  if (!flow || !flow->ptr || !image || !image->ptr || !grad_loss || !grad_flow || !grad_flow->ptr) return FWB_E_NULL;
  if (N == 0) return 0;
  const double cx = (double)N * 2.0 * H * (W - 1), cy = (double)N * 2.0 * (H - 1) * W;
  flowgrad_bwd_kernel<<<dim3((W + LS_THREADS - 1) / LS_THREADS, H, N * T), LS_THREADS, 0, (cudaStream_t)stream>>>(
      to_view(flow), to_view(image), T, C, H, W, grad_loss, cx > 0 ? (float)(1.0 / (cx * T)) : 0.f, cy > 0 ? (float)(1.0 / (cy * T)) : 0.f,
      grad_flow->ptr, grad_flow->sn, grad_flow->st, grad_flow->sc, grad_flow->sh);
  return (int32_t)cudaGetLastError();
}

int32_t fwb_masked_abs_forward(const fwb_view* a, const fwb_view* b, const fwb_view* mask, int32_t N, int32_t T, int32_t C, int32_t H,
                              int32_t W, void* partials, float* loss, void* stream) {
  int32_t rc = loss_dims_ok(N, T, C, H, W);
  if (rc) return rc;
  if (!a || !a->ptr || !b || !b->ptr || !partials || !loss) return FWB_E_NULL;
  cudaStream_t s = (cudaStream_t)stream;
  if (N == 0) return (int32_t)cudaMemsetAsync(loss, 0, sizeof(float), s);
  masked_l1_fwd_kernel<<<dim3(H, N * T), LS_THREADS, 0, s>>>(to_view(a), to_view(b), to_view(mask), T, C, H, W, (float2*)partials);
  loss_final_kernel<<<1, LS_THREADS, 0, s>>>((const float2*)partials, (long long)N * T * H, 1.0 / ((double)N * C * H * W), 0.0, loss);
  return (int32_t)cudaGetLastError();
}

int32_t fwb_masked_abs_backward(const fwb_view* a, const fwb_view* b, const fwb_view* mask, int32_t N, int32_t T, int32_t C, int32_t H,
                               int32_t W, const float* grad_loss, const fwb_view* grad_a, const fwb_view* grad_b,
                               const fwb_view* grad_mask, void* stream) {
  int32_t rc = loss_dims_ok(N, T, C, H, W);
  if (rc) return rc;
  if (!a || !a->ptr || !b || !b->ptr || !grad_loss) return FWB_E_NULL;
  if (N == 0) return 0;
  masked_l1_bwd_kernel<<<dim3((W + LS_THREADS - 1) / LS_THREADS, H, N * T), LS_THREADS, 0, (cudaStream_t)stream>>>(
      to_view(a), to_view(b), to_view(mask), T, C, H, W, grad_loss, (float)(1.0 / ((double)N * C * H * W)), to_view(grad_a), to_view(grad_b),
      to_view(grad_mask));
  return (int32_t)cudaGetLastError();
}

}  // extern "C"

#ifdef CL_PROF
// debug build only: phase timeline of bwd_cl_kernel (64 clock sums), cleared after reading
extern "C" int fwb_debug_cl_prof(unsigned long long* out64) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyFromSymbol(out64, fwb::cl_prof_acc, sizeof(unsigned long long) * 64);
  if (e != cudaSuccess) return (int)e;
  static unsigned long long z[64];
  return (int)cudaMemcpyToSymbol(fwb::cl_prof_acc, z, sizeof(z));
}
#endif

// fwb_coords.cuh — flow -> sample coordinate arithmetic shared by every kernel of the library.
//
// Bit-exact contract (see include/flowwarp_b200.h and DESIGN.md "Coordinate arithmetic"):
//   base grid   utils/net_utils.py:99-103 (torch.linspace on the CPU), nets/OpticalUnet.py:7-15
//   grid        utils/net_utils.py:111,118,126; nets/OpticalUnet.py:127-130
//   unnormalise torch:include/ATen/native/cuda/GridSampler.cuh:21-31 (FMA-contracted by nvcc in ATen's
//               binary; measured with tools/probe_coords.py)
//   border clip torch:include/ATen/native/cuda/GridSampler.cuh:53-83, downgrade :138-147
// Every rounding is spelled with an explicit intrinsic so that nvcc cannot re-associate or contract.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flowwarp_b200.h"

namespace fwb {

struct DirP {
  const float* flow;
  long long flow_sn, flow_sc, flow_st, flow_sh;
  const float* gate;
  long long gate_sn, gate_st, gate_sh;
  const float* blend;
  long long blend_sn, blend_st, blend_sh;
  float sign;
};

struct GroupP {
  int C;
  const float* src[2];
  long long src_sn[2], src_st[2];
  int src_sc[2], src_sh[2];
  float* out;
  long long out_sn, out_st;
  int out_sc, out_sh;
};

struct Geo {
  int N, T, H, W;
  int n_dirs, n_groups;
  int pad_border, align;
  float stepx, stepy;  // fl32(2/(W-1)), fl32(2/(H-1)) computed on the host like torch.linspace does
};

// torch.linspace(-1, 1, n)[i] on the CPU (the reference always builds the base grid on the host)
__device__ __forceinline__ float base_coord(int i, int n, float step) {
  if (n <= 1) return -1.0f;
  return (i < (n >> 1)) ? __fmaf_rn(step, (float)i, -1.0f) : __fmaf_rn(-step, (float)(n - 1 - i), 1.0f);
}

// grid value -> source pixel coordinate; mult = d(coordinate)/d(grid value)
__device__ __forceinline__ float source_index(float g, int size, bool align, bool border, float& mult) {
  float c;
  const float fs = (float)size, fs1 = (float)(size - 1);
  if (align) {
    c = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), fs1);
    mult = __fmul_rn(fs1, 0.5f);
  } else {
    c = __fmul_rn(__fmaf_rn(__fadd_rn(g, 1.0f), fs, -1.0f), 0.5f);
    mult = __fmul_rn(fs, 0.5f);
  }
  if (border) {
    if (c <= 0.0f) {
      c = 0.0f;
      mult = 0.0f;
    } else if (c >= fs1) {
      c = fs1;
      mult = 0.0f;
    } else if (c != c) {
      c = 0.0f;
    }
  }
  if (!(c <= 2147483648.0f) || !(c >= -2147483648.0f) || isinf(c)) c = -100.0f;
  return c;
}

struct Tap {
  float ix, iy, mx, my;
  int x0, y0;
  unsigned valid;  // bit0 nw, bit1 ne, bit2 sw, bit3 se
  float tx, ty;    // ix - x0, iy - y0
  float ux, uy;    // (x0+1) - ix, (y0+1) - iy
  float fx, fy, gate, blend;
};

// The direction's tensors at one (n, t): hoists the 64-bit batch/frame offsets out of the per-pixel code.
struct DirAt {
  const float* flow;
  const float* gate;   // nullptr: no gate
  const float* blend;  // nullptr: no blend weight
};
__device__ __forceinline__ DirAt dir_at(const DirP& D, int n, int t) {
  DirAt a;
  a.flow = D.flow + n * D.flow_sn + t * D.flow_st;
  a.gate = D.gate ? D.gate + n * D.gate_sn + t * D.gate_st : nullptr;
  a.blend = D.blend ? D.blend + n * D.blend_sn + t * D.blend_st : nullptr;
  return a;
}

// Everything channel-independent about one (pixel, direction).
__device__ __forceinline__ void compute_tap_at(const Geo& G, const DirP& D, const DirAt& A, int i, int j, Tap& k) {
  const float* fp = A.flow + (long long)i * D.flow_sh + j;
  float fx = __ldg(fp), fy = __ldg(fp + D.flow_sc);
  k.fx = fx;
  k.fy = fy;
  k.gate = 1.0f;
  if (A.gate) {
    k.gate = __ldg(A.gate + (long long)i * D.gate_sh + j);
    fx = __fmul_rn(fx, k.gate);
    fy = __fmul_rn(fy, k.gate);
  }
  k.blend = A.blend ? __ldg(A.blend + (long long)i * D.blend_sh + j) : 1.0f;
  const float bx = base_coord(j, G.W, G.stepx), by = base_coord(i, G.H, G.stepy);
  const float gx = D.sign < 0.0f ? __fsub_rn(bx, fx) : __fadd_rn(bx, fx);
  const float gy = D.sign < 0.0f ? __fsub_rn(by, fy) : __fadd_rn(by, fy);
  k.ix = source_index(gx, G.W, G.align, G.pad_border, k.mx);
  k.iy = source_index(gy, G.H, G.align, G.pad_border, k.my);
  const float fx0 = floorf(k.ix), fy0 = floorf(k.iy);
  k.x0 = (int)fx0;
  k.y0 = (int)fy0;
  k.tx = __fsub_rn(k.ix, fx0);
  k.ty = __fsub_rn(k.iy, fy0);
  k.ux = __fsub_rn(__fadd_rn(fx0, 1.0f), k.ix);
  k.uy = __fsub_rn(__fadd_rn(fy0, 1.0f), k.iy);
  const bool xin0 = (unsigned)k.x0 < (unsigned)G.W, xin1 = (unsigned)(k.x0 + 1) < (unsigned)G.W;
  const bool yin0 = (unsigned)k.y0 < (unsigned)G.H, yin1 = (unsigned)(k.y0 + 1) < (unsigned)G.H;
  k.valid = (unsigned)(xin0 && yin0) | ((unsigned)(xin1 && yin0) << 1) | ((unsigned)(xin0 && yin1) << 2) |
            ((unsigned)(xin1 && yin1) << 3);
}

__device__ __forceinline__ void compute_tap(const Geo& G, const DirP& D, int n, int t, int i, int j, Tap& k) {
  compute_tap_at(G, D, dir_at(D, n, t), i, j, k);
}

// ---------------------------------------------------------------------------------------------
// The same arithmetic with the padding / align mode as template parameters and only the outputs the staged
// kernels need (no gradient multipliers, no raw flow): roughly a third of the instructions.  Bit-identical to
// compute_tap for x0, y0, valid, tx, ty, ux, uy, blend (every rounding is the same operation in the same order).
// ---------------------------------------------------------------------------------------------
struct FastTap {
  int x0, y0;
  unsigned valid;
  float tx, ty, ux, uy, blend;
};

template <bool ALIGN, bool BORDER>
__device__ __forceinline__ float source_index_fast(float g, float fs, float fs1) {
  float c = ALIGN ? __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), fs1) : __fmul_rn(__fmaf_rn(__fadd_rn(g, 1.0f), fs, -1.0f), 0.5f);
  if (BORDER) {
    c = fminf(fmaxf(c, 0.0f), fs1);  // NaN -> 0 (fmaxf), <= 0 -> 0, >= size-1 -> size-1: source_index's clip
  } else {
    c = (c <= 2147483648.0f && c >= -2147483648.0f) ? c : -100.0f;  // non-finite / beyond int: far outside
  }
  return c;
}

// bx = base_coord(j, W, stepx) is the same for every pixel of a thread (lane = column): passed in.
template <bool ALIGN, bool BORDER>
__device__ __forceinline__ void compute_tap_fast(const Geo& G, const DirP& D, const DirAt& A, float bx, int i, int j, FastTap& k) {
  const float* fp = A.flow + (long long)i * D.flow_sh + j;
  float fx = __ldg(fp), fy = __ldg(fp + D.flow_sc);
  if (A.gate) {
    const float gate = __ldg(A.gate + (long long)i * D.gate_sh + j);
    fx = __fmul_rn(fx, gate);
    fy = __fmul_rn(fy, gate);
  }
  k.blend = A.blend ? __ldg(A.blend + (long long)i * D.blend_sh + j) : 1.0f;
  const float by = base_coord(i, G.H, G.stepy);
  // bx -/+ f as one fma: sign * f is exact, so this is the same single rounding as __fsub_rn / __fadd_rn
  const float gx = __fmaf_rn(D.sign, fx, bx), gy = __fmaf_rn(D.sign, fy, by);
  const float ix = source_index_fast<ALIGN, BORDER>(gx, (float)G.W, (float)(G.W - 1));
  const float iy = source_index_fast<ALIGN, BORDER>(gy, (float)G.H, (float)(G.H - 1));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  k.x0 = (int)fx0;
  k.y0 = (int)fy0;
  k.tx = __fsub_rn(ix, fx0);
  k.ty = __fsub_rn(iy, fy0);
  k.ux = __fsub_rn(__fadd_rn(fx0, 1.0f), ix);
  k.uy = __fsub_rn(__fadd_rn(fy0, 1.0f), iy);
  const bool xin0 = (unsigned)k.x0 < (unsigned)G.W, xin1 = (unsigned)(k.x0 + 1) < (unsigned)G.W;
  const bool yin0 = (unsigned)k.y0 < (unsigned)G.H, yin1 = (unsigned)(k.y0 + 1) < (unsigned)G.H;
  k.valid = (unsigned)(xin0 && yin0) | ((unsigned)(xin1 && yin0) << 1) | ((unsigned)(xin0 && yin1) << 2) |
            ((unsigned)(xin1 && yin1) << 3);
}

struct GradP {
  const float* grad_out[FWB_MAX_GROUPS];
  long long go_sn[FWB_MAX_GROUPS], go_st[FWB_MAX_GROUPS];
  int go_sc[FWB_MAX_GROUPS], go_sh[FWB_MAX_GROUPS];
  float* grad_src[FWB_MAX_GROUPS][2];
  long long gs_sn[FWB_MAX_GROUPS][2], gs_st[FWB_MAX_GROUPS][2];
  int gs_sc[FWB_MAX_GROUPS][2], gs_sh[FWB_MAX_GROUPS][2];
  float* grad_flow[2];
  long long gf_sn[2], gf_sc[2], gf_st[2], gf_sh[2];
  float* grad_gate[2];
  long long gg_sn[2], gg_st[2], gg_sh[2];
  float* grad_blend[2];
  long long gb_sn[2], gb_st[2], gb_sh[2];
};

struct Params {
  Geo geo;
  DirP dir[2];
  GroupP grp[FWB_MAX_GROUPS];
};

}  // namespace fwb

// fwb_util.cuh — small PTX helpers and the global-memory scatter shared by the tile kernels (fwb_tile.cuh).
#pragma once
#include "fwb_coords.cuh"
#include "fwb_generic.cuh"

namespace fwb {

constexpr int ST_MAXSLOW = 128;  // capacity of the slow-pixel list of a tile

struct StageSlow {  // pixels of the tile that are served from global memory (border-clamped / wild ones)
  int n;
  unsigned short pix[ST_MAXSLOW];  // (row in tile << 5) | column in tile
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// predicated streaming store (no branch around it)
__device__ __forceinline__ void st_cs_if(float* p, float v, bool pred) {
  asm volatile("{\n .reg .pred pp;\n setp.ne.b32 pp, %2, 0;\n @pp st.global.cs.f32 [%0], %1;\n}\n" ::"l"(p), "f"(v), "r"((int)pred)
               : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// global-atomic scatter of one pixel's 4 taps of one direction for channel c: grad_src[tap] += w_tap * gw
// (ATen grid_sampler_2d_backward's atomicAdd scatter, reached from utils/net_utils.py:113)
__device__ __forceinline__ void scatter_atomic_px(const GradP& Q, int g, int d, int n, int t, int c, const Tap& k, float gw) {
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
      red_add_v2(p0, w[2 * r] * gw, w[2 * r + 1] * gw);
    } else {
#pragma unroll
      for (int q = 2 * r; q < 2 * r + 2; ++q)
        if (k.valid & (1u << q)) atomicAdd(gs + off[q], w[q] * gw);
    }
  }
}

// one pixel, all channels of all groups: global-atomic scatter into grad_src (tiles that do not fit in shared memory)
template <int NDIRS>
__device__ __forceinline__ void bwd_src_generic_pixel(const Params& P, const GradP& Q, int n, int t, int i, int j) {
  Tap k[NDIRS];
#pragma unroll
  for (int d = 0; d < NDIRS; ++d) compute_tap(P.geo, P.dir[d], n, t, i, j, k[d]);
  for (int g = 0; g < P.geo.n_groups; ++g) {
    if (!Q.grad_out[g]) continue;
    const float* go = Q.grad_out[g] + n * Q.go_sn[g] + t * Q.go_st[g] + (long long)i * Q.go_sh[g] + j;
    for (int c = 0; c < P.grp[g].C; ++c) {
      const float gout = __ldg(go + (long long)c * Q.go_sc[g]);
#pragma unroll
      for (int d = 0; d < NDIRS; ++d) scatter_atomic_px(Q, g, d, n, t, c, k[d], P.dir[d].blend ? gout * k[d].blend : gout);
    }
  }
}

// ... preceded by kernel 2's generic body (the fused kernels)
template <int NDIRS>
__device__ __forceinline__ void bwd_fused_generic_pixel(const Params& P, const GradP& Q, int n, int t, int i, int j) {
  bwdflow_generic_pixel<NDIRS>(P, Q, n, t, i, j);
  bwd_src_generic_pixel<NDIRS>(P, Q, n, t, i, j);
}

}  // namespace fwb

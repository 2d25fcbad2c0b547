// fwb_loss.cuh — the flow-regularisation losses that consume the warp.
//
// Their Python source is deleted from the reference; the formulas below are read from the bytecode that survives in
// __pycache__/losses.cpython-36.pyc (TrainingLoss._flowgradloss, line 413; TrainingLoss._flowconsist, line 481) and are built
// on gradientx / gradienty (utils/net_utils.py:243-248) and on the FlowWrapper held as self.flowwarp (runners/VAEer.py:53):
//
//   _flowgradloss(flow, image):  flow *= 128; image *= 256
//       weightx = exp(-mean_c |gradientx(image)|), weighty likewise (keepdim over the channel axis)
//       return mean |gradientx(flow) * weightx| + mean |gradienty(flow) * weighty|          (edge-aware smoothness)
//   _flowconsist(flow, flowback, mask_fw, mask_bw):
//       prev = mean( mask_bw * |flowwarp(flow, -flowback) - flowback| ),  next = mean( mask_fw * |flowwarp(flowback, flow) - flow| )
//       return prev + next                                                                  (forward / backward consistency)
//
// The warps of _flowconsist are the library's own kernels (a 2-channel source); what is added here are the two fused
// reductions: one streaming pass each, partial sums per CTA, a fixed-order final sum in double (deterministic, no atomics).
// Tensors are [N,T,C,H,W] views with element strides (sn, st, sc, sh), W-stride 1: a frame slice flow[:, :, i] of the
// reference's [N,2,T,H,W] layout needs no copy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fwb {

constexpr int LS_THREADS = 256;

struct LossView {
  const float* p;
  long long sn, st, sc, sh;
};

__device__ __forceinline__ float ls_block_sum(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < LS_THREADS / 32 ? sm[lane] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

// weight of the horizontal / vertical edge at (y, x): exp(-mean_c |256 img[c,y,x] - 256 img[c,y,x+1]|) (resp. y+1)
__device__ __forceinline__ float ls_weight(const float* im, long long sc, int C, long long step) {
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += fabsf(256.0f * im[c * sc] - 256.0f * im[c * sc + step]);
  return expf(-(s / (float)C));
}

// ---- _flowgradloss, forward: one CTA per (n*T + t, row); part[cta] = (sum of |dx * wx|, sum of |dy * wy|) of that row
__global__ void __launch_bounds__(LS_THREADS) flowgrad_fwd_kernel(LossView F, LossView I, int T, int C, int H, int W, float2* part) {
  __shared__ float sm[LS_THREADS / 32];
  const int y = blockIdx.x, nt = blockIdx.y, n = nt / T, t = nt - n * T;
  const float* f = F.p + n * F.sn + t * F.st + (long long)y * F.sh;
  const float* im = I.p + n * I.sn + t * I.st + (long long)y * I.sh;
  float sx = 0.f, sy = 0.f;
  for (int x = threadIdx.x; x < W; x += LS_THREADS) {
    if (x + 1 < W) {
      const float w = ls_weight(im + x, I.sc, C, 1);
#pragma unroll
      for (int k = 0; k < 2; ++k) sx += fabsf((128.0f * f[k * F.sc + x] - 128.0f * f[k * F.sc + x + 1]) * w);
    }
    if (y + 1 < H) {
      const float w = ls_weight(im + x, I.sc, C, I.sh);
#pragma unroll
      for (int k = 0; k < 2; ++k) sy += fabsf((128.0f * f[k * F.sc + x] - 128.0f * f[k * F.sc + F.sh + x]) * w);
    }
  }
  sx = ls_block_sum(sx, sm);
  sy = ls_block_sum(sy, sm);
  if (threadIdx.x == 0) part[(long long)nt * H + y] = make_float2(sx, sy);
}

// final sum of the per-CTA partials in a fixed order (double), loss = sum.x * inv_x + sum.y * inv_y (float, as torch adds two means)
__global__ void __launch_bounds__(LS_THREADS) loss_final_kernel(const float2* part, long long nparts, double inv_x, double inv_y, float* loss) {
  __shared__ double smx[LS_THREADS], smy[LS_THREADS];
  double ax = 0.0, ay = 0.0;
  for (long long q = threadIdx.x; q < nparts; q += LS_THREADS) {
    ax += (double)part[q].x;
    ay += (double)part[q].y;
  }
  smx[threadIdx.x] = ax;
  smy[threadIdx.x] = ay;
  __syncthreads();
  for (int o = LS_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      smx[threadIdx.x] += smx[threadIdx.x + o];
      smy[threadIdx.x] += smy[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(smx[0] * inv_x) + (float)(smy[0] * inv_y);
}

__device__ __forceinline__ float ls_sign(float v) { return v > 0.f ? 1.0f : (v < 0.f ? -1.0f : 0.0f); }

// ---- _flowgradloss, backward w.r.t. flow (the image is data): one thread per pixel, both flow channels
//   d loss / d flow[k,y,x] = 128 gx [sgn(u_x(y,x)) wx(y,x) - sgn(u_x(y,x-1)) wx(y,x-1)] + 128 gy [sgn(u_y(y,x)) wy(y,x) - sgn(u_y(y-1,x)) wy(y-1,x)]
// with u = (flow difference) * weight and gx = grad_loss / count_x, gy = grad_loss / count_y.
__global__ void __launch_bounds__(LS_THREADS) flowgrad_bwd_kernel(LossView F, LossView I, int T, int C, int H, int W, const float* grad_loss,
                                                                  float inv_x, float inv_y, float* gf, long long g_sn, long long g_st,
                                                                  long long g_sc, long long g_sh) {
  const int x = blockIdx.x * LS_THREADS + threadIdx.x, y = blockIdx.y, nt = blockIdx.z, n = nt / T, t = nt - n * T;
  if (x >= W) return;
  const float* f = F.p + n * F.sn + t * F.st + (long long)y * F.sh + x;
  const float* im = I.p + n * I.sn + t * I.st + (long long)y * I.sh + x;
  const float g = __ldg(grad_loss), gx = 128.0f * g * inv_x, gy = 128.0f * g * inv_y;
  float acc[2] = {0.f, 0.f};
  if (x + 1 < W) {
    const float w = ls_weight(im, I.sc, C, 1);
#pragma unroll
    for (int k = 0; k < 2; ++k) acc[k] += gx * w * ls_sign((128.0f * f[k * F.sc] - 128.0f * f[k * F.sc + 1]) * w);
  }
  if (x > 0) {
    const float w = ls_weight(im - 1, I.sc, C, 1);
#pragma unroll
    for (int k = 0; k < 2; ++k) acc[k] -= gx * w * ls_sign((128.0f * f[k * F.sc - 1] - 128.0f * f[k * F.sc]) * w);
  }
  if (y + 1 < H) {
    const float w = ls_weight(im, I.sc, C, I.sh);
#pragma unroll
    for (int k = 0; k < 2; ++k) acc[k] += gy * w * ls_sign((128.0f * f[k * F.sc] - 128.0f * f[k * F.sc + F.sh]) * w);
  }
  if (y > 0) {
    const float w = ls_weight(im - I.sh, I.sc, C, I.sh);
#pragma unroll
    for (int k = 0; k < 2; ++k) acc[k] -= gy * w * ls_sign((128.0f * f[k * F.sc - F.sh] - 128.0f * f[k * F.sc]) * w);
  }
  float* o = gf + n * g_sn + t * g_st + (long long)y * g_sh + x;
  o[0] = acc[0];
  o[g_sc] = acc[1];
}

// ---- mean(mask * |a - b|) (the two terms of _flowconsist), forward: one CTA per (n*T + t, row); mask [N,T,1,H,W] or NULL
__global__ void __launch_bounds__(LS_THREADS) masked_l1_fwd_kernel(LossView A, LossView B, LossView M, int T, int C, int H, int W, float2* part) {
  __shared__ float sm[LS_THREADS / 32];
  const int y = blockIdx.x, nt = blockIdx.y, n = nt / T, t = nt - n * T;
  const float* a = A.p + n * A.sn + t * A.st + (long long)y * A.sh;
  const float* b = B.p + n * B.sn + t * B.st + (long long)y * B.sh;
  const float* m = M.p ? M.p + n * M.sn + t * M.st + (long long)y * M.sh : nullptr;
  float s = 0.f;
  for (int x = threadIdx.x; x < W; x += LS_THREADS) {
    const float mk = m ? m[x] : 1.0f;
    for (int c = 0; c < C; ++c) s += m ? mk * fabsf(a[c * A.sc + x] - b[c * B.sc + x]) : fabsf(a[c * A.sc + x] - b[c * B.sc + x]);
  }
  s = ls_block_sum(s, sm);
  if (threadIdx.x == 0) part[(long long)nt * H + y] = make_float2(s, 0.f);
}

// backward: grad_a = g * mask * sgn(a - b) / count, grad_b = -grad_a, grad_mask = g * sum_c |a - b| / count (each optional)
__global__ void __launch_bounds__(LS_THREADS) masked_l1_bwd_kernel(LossView A, LossView B, LossView M, int T, int C, int H, int W,
                                                                   const float* grad_loss, float inv, LossView GA, LossView GB, LossView GM) {
  const int x = blockIdx.x * LS_THREADS + threadIdx.x, y = blockIdx.y, nt = blockIdx.z, n = nt / T, t = nt - n * T;
  if (x >= W) return;
  const float g = __ldg(grad_loss) * inv;
  const float* a = A.p + n * A.sn + t * A.st + (long long)y * A.sh + x;
  const float* b = B.p + n * B.sn + t * B.st + (long long)y * B.sh + x;
  const float mk = M.p ? M.p[n * M.sn + t * M.st + (long long)y * M.sh + x] : 1.0f;
  float* ga = GA.p ? const_cast<float*>(GA.p) + n * GA.sn + t * GA.st + (long long)y * GA.sh + x : nullptr;
  float* gb = GB.p ? const_cast<float*>(GB.p) + n * GB.sn + t * GB.st + (long long)y * GB.sh + x : nullptr;
  float sum = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = a[c * A.sc] - b[c * B.sc];
    sum += fabsf(d);
    const float v = g * mk * ls_sign(d);
    if (ga) ga[c * GA.sc] = v;
    if (gb) gb[c * GB.sc] = -v;
  }
  if (GM.p) const_cast<float*>(GM.p)[n * GM.sn + t * GM.st + (long long)y * GM.sh + x] = g * sum;
}

}  // namespace fwb

/*
 * flowwarp_oracle.c — CPU restatement of the reference's flow-warp + blend path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports, links or executes this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and
 * there only as the checker / the CPU baseline.  The product path is the CUDA library.
 *
 * PARITY PIN (pinned): the reference ships no tests, golden vectors or fixtures of its own for this
 * path, so this restatement is pinned to OUTPUTS OF THE UNMODIFIED REFERENCE run
 * in the build container: tests/golden/make_golden.py (and make_golden_refine.py) import
 * /root/reference/utils/net_utils.py as it is, run
 * FlowWrapper / warp / warp_back (+ autograd) on seeded inputs and stores the results in
 * tests/golden/ (npz files); tests/test_oracle_golden.py checks this file against them (exact sample
 * coordinates via the parity-image probe, outputs and gradients within the stated tolerances).
 *
 * What is restated (file:line under /root/reference, `torch:` = installed PyTorch 2.11 headers):
 *   base grid      utils/net_utils.py:99-103  (torch.linspace(-1,1,n) on the CPU; -1 when n == 1)
 *                  nets/OpticalUnet.py:7-15   (same values through a K=1 matmul)
 *   gated flow     utils/net_utils.py:118,126 (flow * mask, sign flip for warp_back)
 *   grid           utils/net_utils.py:111     (base_grid - flow);  nets/OpticalUnet.py:127-130 (+back)
 *   sampler        utils/net_utils.py:113, nets/OpticalUnet.py:132-139 -> aten::grid_sampler_2d:
 *                  torch:include/ATen/native/cuda/GridSampler.cuh:21-31 (unnormalize),
 *                  :53-57 (clip), :138-147 (safe_downgrade), :223-226 (within_bounds_2d),
 *                  torch:_decomp/decompositions.py:4515-4537 (4-tap sum, nw ne sw se order)
 *   blend          nets/OpticalUnet.py:141-146 (mask * warped, per direction; summed here)
 *   backward       torch:include/ATen/native/GridSampler.h:43-83 (multipliers, clip gradient),
 *                  ATen grid_sampler_2d_backward (gix/giy accumulation, scatter of w*gOut)
 *
 * Arithmetic notes (measured): ATen evaluates `(g+1)*size-1` with one rounding
 * (FMA contraction) on both its CPU-vectorised and CUDA builds; torch.linspace's CPU kernel is
 * `fma(step,i,-1)` for i < n/2 and `fma(-step,n-1-i,+1)` above.  Compile with -ffp-contract=off:
 * every fused operation below is an explicit fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/flowwarp_b200.h"

#include <pthread.h>

/* plain pthreads fan-out (the image's gcc ships without libgomp) */
static int g_threads = 1;
void fwo_set_num_threads(int32_t k) { g_threads = k < 1 ? 1 : (k > 256 ? 256 : k); }
int32_t fwo_num_threads(void) { return g_threads; }

typedef struct {
  void (*fn)(const void*, int, int);
  const void* arg;
  int begin, end;
} job_t;

static void* job_main(void* v) {
  job_t* j = (job_t*)v;
  j->fn(j->arg, j->begin, j->end);
  return NULL;
}

/* run fn(arg, b, e) over [0,total) split into g_threads contiguous chunks */
static void parallel_for(void (*fn)(const void*, int, int), const void* arg, int total) {
  int k = g_threads < total ? g_threads : total;
  if (k <= 1) {
    fn(arg, 0, total);
    return;
  }
  pthread_t th[256];
  job_t jobs[256];
  for (int i = 0; i < k; ++i) {
    jobs[i].fn = fn;
    jobs[i].arg = arg;
    jobs[i].begin = (int)((long long)total * i / k);
    jobs[i].end = (int)((long long)total * (i + 1) / k);
    pthread_create(&th[i], NULL, job_main, &jobs[i]);
  }
  for (int i = 0; i < k; ++i) pthread_join(th[i], NULL);
}

/* utils/net_utils.py:100,102 — torch.linspace(-1, 1, n)[i]; `torch.Tensor([-1])` when n == 1 */
float fwo_base_coord(int i, int n) {
  if (n <= 1) return -1.0f;
  const float step = 2.0f / (float)(n - 1);
  if (i < n / 2) return fmaf(step, (float)i, -1.0f);
  return fmaf(-step, (float)(n - 1 - i), 1.0f);
}

/* torch:include/ATen/native/cuda/GridSampler.cuh:21-31 */
static float unnormalize(float g, int size, int align_corners) {
  if (align_corners) return ((g + 1.0f) / 2.0f) * (float)(size - 1);
  return fmaf(g + 1.0f, (float)size, -1.0f) / 2.0f;
}

/* torch:include/ATen/native/cuda/GridSampler.cuh:53-57,138-147,150-170; the gradient factor of the
 * clip follows :62-83 (borders count as out of range). */
static float source_index(float g, int size, int pad, int align_corners, float* dmult) {
  float c = unnormalize(g, size, align_corners);
  float m = align_corners ? (float)(size - 1) / 2.0f : (float)size / 2.0f;
  if (pad == FWB_PAD_BORDER) {
    const float hi = (float)(size - 1);
    if (c <= 0.0f) {
      c = 0.0f;
      m = 0.0f;
    } else if (c >= hi) {
      c = hi;
      m = 0.0f;
    } else if (c != c) { /* NaN: forward clip gives max(NaN,0)=0; keep multiplier (ATen: NaN fails both tests) */
      c = 0.0f;
    }
  }
  if (!(c <= 2147483646.0f) || !(c >= -2147483648.0f) || !isfinite((double)c)) c = -100.0f;
  if (dmult) *dmult = m;
  return c;
}

typedef struct {
  float ix, iy, mx, my; /* source coordinates and d(ix)/d(gx), d(iy)/d(gy) */
  int x0, y0;
  unsigned valid;        /* bit0 nw, bit1 ne, bit2 sw, bit3 se */
  float wnw, wne, wsw, wse;
  float fx, fy;          /* raw flow */
  float gate, blend;     /* 1 when absent */
} tap_t;

static void compute_tap(const fwb_problem* p, int d, int n, int t, int i, int j, tap_t* k) {
  const fwb_dir* D = &p->dir[d];
  const int64_t fo = n * D->flow_sn + t * D->flow_st + i * D->flow_sh + j;
  float fx = D->flow[fo], fy = D->flow[fo + D->flow_sc];
  k->fx = fx;
  k->fy = fy;
  k->gate = 1.0f;
  if (D->gate) {
    k->gate = D->gate[n * D->gate_sn + t * D->gate_st + i * D->gate_sh + j];
    fx = fx * k->gate; /* utils/net_utils.py:118 */
    fy = fy * k->gate;
  }
  k->blend = D->blend ? D->blend[n * D->blend_sn + t * D->blend_st + i * D->blend_sh + j] : 1.0f;
  const float bx = fwo_base_coord(j, p->W), by = fwo_base_coord(i, p->H);
  const float gx = D->sign < 0 ? bx - fx : bx + fx; /* utils/net_utils.py:111 / OpticalUnet.py:129 */
  const float gy = D->sign < 0 ? by - fy : by + fy;
  k->ix = source_index(gx, p->W, p->padding_mode, p->align_corners, &k->mx);
  k->iy = source_index(gy, p->H, p->padding_mode, p->align_corners, &k->my);
  const float fx0 = floorf(k->ix), fy0 = floorf(k->iy);
  k->x0 = (int)fx0;
  k->y0 = (int)fy0;
  const float x1 = fx0 + 1.0f, y1 = fy0 + 1.0f;
  k->wnw = (x1 - k->ix) * (y1 - k->iy);
  k->wne = (k->ix - fx0) * (y1 - k->iy);
  k->wsw = (x1 - k->ix) * (k->iy - fy0);
  k->wse = (k->ix - fx0) * (k->iy - fy0);
  const int xin0 = k->x0 >= 0 && k->x0 < p->W, xin1 = k->x0 + 1 >= 0 && k->x0 + 1 < p->W;
  const int yin0 = k->y0 >= 0 && k->y0 < p->H, yin1 = k->y0 + 1 >= 0 && k->y0 + 1 < p->H;
  k->valid = (unsigned)((xin0 && yin0) | ((xin1 && yin0) << 1) | ((xin0 && yin1) << 2) |
                        ((xin1 && yin1) << 3));
}

static int check(const fwb_problem* p) {
  if (!p) return FWB_E_NULL;
  if (p->N < 0 || p->T < 1 || p->H < 1 || p->W < 1) return FWB_E_SHAPE;
  if (p->n_dirs < 1 || p->n_dirs > 2) return FWB_E_DIRS;
  if (p->n_groups < 1 || p->n_groups > FWB_MAX_GROUPS) return FWB_E_GROUPS;
  if (p->padding_mode != FWB_PAD_ZEROS && p->padding_mode != FWB_PAD_BORDER) return FWB_E_MODE;
  return 0;
}

/* Debug outputs: floor indices, validity bits and the float coordinates of direction d. */
int32_t fwo_sample_indices(const fwb_problem* p, int32_t d, int32_t* x0, int32_t* y0, uint8_t* valid,
                           float* ix, float* iy) {
  int rc = check(p);
  if (rc) return rc;
  for (int n = 0; n < p->N; ++n)
    for (int t = 0; t < p->T; ++t)
      for (int i = 0; i < p->H; ++i)
        for (int j = 0; j < p->W; ++j) {
          tap_t k;
          compute_tap(p, d, n, t, i, j, &k);
          const int64_t o = (((int64_t)n * p->T + t) * p->H + i) * p->W + j;
          if (x0) x0[o] = k.x0;
          if (y0) y0[o] = k.y0;
          if (valid) valid[o] = (uint8_t)k.valid;
          if (ix) ix[o] = k.ix;
          if (iy) iy[o] = k.iy;
        }
  return 0;
}

/* Forward: utils/net_utils.py:93-121 / nets/OpticalUnet.py:123-146, all groups and directions. */
static void forward_rows(const void* arg, int begin, int end) {
  const fwb_problem* p = (const fwb_problem*)arg;
  for (int r = begin; r < end; ++r) {
    {
      const int nt = r / p->H, i = r % p->H;
      const int n = nt / p->T, t = nt % p->T;
      for (int j = 0; j < p->W; ++j) {
        tap_t k[2];
        for (int d = 0; d < p->n_dirs; ++d) compute_tap(p, d, n, t, i, j, &k[d]);
        for (int g = 0; g < p->n_groups; ++g) {
          const fwb_group* G = &p->grp[g];
          for (int c = 0; c < G->C; ++c) {
            float o = 0.0f;
            for (int d = 0; d < p->n_dirs; ++d) {
              const float* s = G->src[d] + n * G->src_sn[d] + t * G->src_st[d] + c * G->src_sc[d];
              const int64_t r0 = (int64_t)k[d].y0 * G->src_sh[d], r1 = r0 + G->src_sh[d];
              const int x0 = k[d].x0;
              float a = 0.0f; /* nw, ne, sw, se — torch:_decomp/decompositions.py:4515-4537 */
              if (k[d].valid & 1u) a = fmaf(s[r0 + x0], k[d].wnw, a);
              if (k[d].valid & 2u) a = fmaf(s[r0 + x0 + 1], k[d].wne, a);
              if (k[d].valid & 4u) a = fmaf(s[r1 + x0], k[d].wsw, a);
              if (k[d].valid & 8u) a = fmaf(s[r1 + x0 + 1], k[d].wse, a);
              if (p->dir[d].blend) a = a * k[d].blend; /* nets/OpticalUnet.py:145-146 */
              o = (d == 0) ? a : o + a;
            }
            G->out[n * G->out_sn + t * G->out_st + c * G->out_sc + i * G->out_sh + j] = o;
          }
        }
      }
    }
  }
}

int32_t fwo_warp_blend_forward(const fwb_problem* p) {
  int rc = check(p);
  if (rc) return rc;
  parallel_for(forward_rows, p, p->N * p->T * p->H);
  return 0;
}

typedef struct {
  const fwb_problem* p;
  const fwb_grads* q;
} bwd_arg_t;
static void backward_batch(const void* arg, int begin, int end);

/* Backward: every gradient the autograd graph of the reference path produces.
 * grad_src buffers are ZEROED here first (they are accumulated into). */
int32_t fwo_warp_blend_backward(const fwb_problem* p, const fwb_grads* q) {
  int rc = check(p);
  if (rc) return rc;
  if (!q) return FWB_E_NULL;
  /* zero grad_src */
  for (int g = 0; g < p->n_groups; ++g)
    for (int d = 0; d < p->n_dirs; ++d) {
      float* gs = q->grad_src[g][d];
      if (!gs) continue;
      const int Tn = q->gs_st[g][d] == 0 ? 1 : p->T;
      for (int n = 0; n < p->N; ++n)
        for (int t = 0; t < Tn; ++t)
          for (int c = 0; c < p->grp[g].C; ++c)
            for (int i = 0; i < p->H; ++i)
              memset(gs + n * q->gs_sn[g][d] + t * q->gs_st[g][d] + c * q->gs_sc[g][d] +
                         i * q->gs_sh[g][d],
                     0, sizeof(float) * (size_t)p->W);
    }
  bwd_arg_t a = {p, q};
  parallel_for(backward_batch, &a, p->N);
  return 0;
}

static void backward_batch(const void* arg, int begin, int end) {
  const fwb_problem* p = ((const bwd_arg_t*)arg)->p;
  const fwb_grads* q = ((const bwd_arg_t*)arg)->q;
  for (int n = begin; n < end; ++n)
    for (int t = 0; t < p->T; ++t)
      for (int i = 0; i < p->H; ++i)
        for (int j = 0; j < p->W; ++j)
          for (int d = 0; d < p->n_dirs; ++d) {
            tap_t k;
            compute_tap(p, d, n, t, i, j, &k);
            const float y1 = (float)k.y0 + 1.0f, x1 = (float)k.x0 + 1.0f;
            const float fy0 = (float)k.y0, fx0 = (float)k.x0;
            float gix = 0.0f, giy = 0.0f, gbl = 0.0f;
            for (int g = 0; g < p->n_groups; ++g) {
              const fwb_group* G = &p->grp[g];
              const float* go = q->grad_out[g];
              if (!go) continue;
              float* gs = q->grad_src[g][d];
              for (int c = 0; c < G->C; ++c) {
                const float gout =
                    go[n * q->go_sn[g] + t * q->go_st[g] + c * q->go_sc[g] + i * q->go_sh[g] + j];
                /* d(out)/d(warp_d) = blend_d  (nets/OpticalUnet.py:145-146) */
                const float gw = p->dir[d].blend ? gout * k.blend : gout;
                const float* s = G->src[d] + n * G->src_sn[d] + t * G->src_st[d] + c * G->src_sc[d];
                const int64_t r0 = (int64_t)k.y0 * G->src_sh[d], r1 = r0 + G->src_sh[d];
                float vnw = 0, vne = 0, vsw = 0, vse = 0;
                if (k.valid & 1u) vnw = s[r0 + k.x0];
                if (k.valid & 2u) vne = s[r0 + k.x0 + 1];
                if (k.valid & 4u) vsw = s[r1 + k.x0];
                if (k.valid & 8u) vse = s[r1 + k.x0 + 1];
                if (p->dir[d].blend) {
                  float a = 0.0f;
                  a = fmaf(vnw, k.wnw, a);
                  a = fmaf(vne, k.wne, a);
                  a = fmaf(vsw, k.wsw, a);
                  a = fmaf(vse, k.wse, a);
                  gbl = fmaf(gout, a, gbl);
                }
                /* ATen grid_sampler_2d_backward: OOB tap values are 0 */
                gix -= vnw * (y1 - k.iy) * gw;
                giy -= vnw * (x1 - k.ix) * gw;
                gix += vne * (y1 - k.iy) * gw;
                giy -= vne * (k.ix - fx0) * gw;
                gix -= vsw * (k.iy - fy0) * gw;
                giy += vsw * (x1 - k.ix) * gw;
                gix += vse * (k.iy - fy0) * gw;
                giy += vse * (k.ix - fx0) * gw;
                if (gs) {
                  float* o = gs + n * q->gs_sn[g][d] + t * q->gs_st[g][d] + c * q->gs_sc[g][d];
                  const int64_t q0 = (int64_t)k.y0 * q->gs_sh[g][d], q1 = q0 + q->gs_sh[g][d];
                  if (k.valid & 1u) o[q0 + k.x0] += k.wnw * gw;
                  if (k.valid & 2u) o[q0 + k.x0 + 1] += k.wne * gw;
                  if (k.valid & 4u) o[q1 + k.x0] += k.wsw * gw;
                  if (k.valid & 8u) o[q1 + k.x0 + 1] += k.wse * gw;
                }
              }
            }
            /* grad_grid = (mx*gix, my*giy) — torch:include/ATen/native/GridSampler.h:43-54;
             * grid = base -/+ flow_eff  =>  grad_flow_eff = sign * grad_grid */
            float gfx = k.mx * gix, gfy = k.my * giy;
            if (p->dir[d].sign < 0) {
              gfx = -gfx;
              gfy = -gfy;
            }
            if (q->grad_gate[d] && p->dir[d].gate) /* d(flow*gate)/d(gate), utils/net_utils.py:118 */
              q->grad_gate[d][n * q->gg_sn[d] + t * q->gg_st[d] + i * q->gg_sh[d] + j] =
                  gfx * k.fx + gfy * k.fy;
            if (q->grad_flow[d]) {
              float* o = q->grad_flow[d] + n * q->gf_sn[d] + t * q->gf_st[d] + i * q->gf_sh[d] + j;
              o[0] = p->dir[d].gate ? gfx * k.gate : gfx;
              o[q->gf_sc[d]] = p->dir[d].gate ? gfy * k.gate : gfy;
            }
            if (q->grad_blend[d] && p->dir[d].blend)
              q->grad_blend[d][n * q->gb_sn[d] + t * q->gb_st[d] + i * q->gb_sh[d] + j] = gbl;
          }
}

/* ---------------------------------------------------------------------------------------------
 * Flat entry point for the full-size parity tests: contiguous fp32 arrays, plain pointers and sizes.
 * The problem description is packed HERE, in C, from the shapes alone — it does not go through the
 * product's Python struct packer (_problem.py), so a packing bug on the product side cannot cancel
 * out in a CUDA-vs-oracle comparison.  The bidirectional warp + mask-weighted blend of
 * nets/OpticalUnet.py:123-146 (direction 0: grid = base - flow0, direction 1: grid = base + flow1),
 * T = 1, n_groups channel groups that share flows and blend masks.
 *   src0[g], src1[g]   [N, C[g], H, W]      flow0, flow1   [N, 2, H, W]     blend0, blend1  [N, H, W] or NULL
 *   out[g]             [N, C[g], H, W]      (forward; skipped when out == NULL)
 *   gout[g]            [N, C[g], H, W]      (backward; skipped when gout == NULL)
 *   gsrc0[g], gsrc1[g] [N, C[g], H, W]      gflow0, gflow1 [N, 2, H, W]     gblend0, gblend1 [N, H, W]
 * ------------------------------------------------------------------------------------------- */
int32_t fwo_bidir_contig(int32_t N, int32_t H, int32_t W, int32_t n_groups, const int32_t* C,
                         const float* const* src0, const float* const* src1, const float* flow0,
                         const float* flow1, const float* blend0, const float* blend1, int32_t pad,
                         int32_t align_corners, float* const* out, const float* const* gout,
                         float* const* gsrc0, float* const* gsrc1, float* gflow0, float* gflow1,
                         float* gblend0, float* gblend1) {
  if (n_groups < 1 || n_groups > FWB_MAX_GROUPS) return FWB_E_GROUPS;
  fwb_problem p;
  fwb_grads q;
  memset(&p, 0, sizeof p);
  memset(&q, 0, sizeof q);
  const int64_t HW = (int64_t)H * W;
  p.N = N, p.T = 1, p.H = H, p.W = W;
  p.n_dirs = 2, p.n_groups = n_groups;
  p.padding_mode = pad, p.align_corners = align_corners, p.flags = 0;
  const float* flows[2] = {flow0, flow1};
  const float* blends[2] = {blend0, blend1};
  float* gflows[2] = {gflow0, gflow1};
  float* gblends[2] = {gblend0, gblend1};
  for (int d = 0; d < 2; ++d) {
    p.dir[d].flow = flows[d];
    p.dir[d].flow_sn = 2 * HW, p.dir[d].flow_sc = HW, p.dir[d].flow_st = 0, p.dir[d].flow_sh = W;
    p.dir[d].blend = blends[d];
    p.dir[d].blend_sn = HW, p.dir[d].blend_st = 0, p.dir[d].blend_sh = W;
    p.dir[d].sign = d == 0 ? -1.0f : 1.0f;
    q.grad_flow[d] = gflows[d];
    q.gf_sn[d] = 2 * HW, q.gf_sc[d] = HW, q.gf_st[d] = 0, q.gf_sh[d] = W;
    q.grad_blend[d] = gblends[d];
    q.gb_sn[d] = HW, q.gb_st[d] = 0, q.gb_sh[d] = W;
  }
  for (int g = 0; g < n_groups; ++g) {
    fwb_group* G = &p.grp[g];
    G->C = C[g];
    const float* s[2] = {src0[g], src1[g]};
    for (int d = 0; d < 2; ++d) {
      G->src[d] = s[d];
      G->src_sn[d] = C[g] * HW, G->src_st[d] = 0, G->src_sc[d] = HW, G->src_sh[d] = W;
    }
    if (out) {
      G->out = out[g];
      G->out_sn = C[g] * HW, G->out_st = 0, G->out_sc = HW, G->out_sh = W;
    }
    if (gout) {
      q.grad_out[g] = gout[g];
      q.go_sn[g] = C[g] * HW, q.go_st[g] = 0, q.go_sc[g] = HW, q.go_sh[g] = W;
      float* gs[2] = {gsrc0 ? gsrc0[g] : NULL, gsrc1 ? gsrc1[g] : NULL};
      for (int d = 0; d < 2; ++d) {
        q.grad_src[g][d] = gs[d];
        q.gs_sn[g][d] = C[g] * HW, q.gs_st[g][d] = C[g] * HW /* T = 1: never stepped */,
        q.gs_sc[g][d] = HW, q.gs_sh[g][d] = W;
      }
    }
  }
  int rc = 0;
  if (out) rc = fwo_warp_blend_forward(&p);
  if (!rc && gout) rc = fwo_warp_blend_backward(&p, &q);
  return rc;
}

// variant_kernels.cu — the kernels behind the non-default constructor flags and optimizers of the reference:
//   * is_deconv=False (models/unet.py:189-191): nn.UpsamplingBilinear2d(scale_factor=2) [bilinear, align_corners=True]
//     followed by a 1x1 conv.  Both are linear and the bilinear weights sum to one, so the 1x1 conv (bias included) runs
//     FIRST, on the low-resolution tensor (4x fewer pixels, on the tensor cores through unpp_conv_tc with taps = 1), and
//     these kernels do the x2 upsample of its output and the exact adjoint for the backward pass;
//   * the flat-buffer forms of every optimizer trainer/trainer.py:344-376 can select.
// All are coalesced 16-byte-per-thread CUDA-core kernels bounded by HBM bandwidth.
#include "common.h"
#include "../../include/unpp.h"
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

namespace {

inline int grid_for(long total, int block, int per_sm = 8) {
  long g = (total + block - 1) / block;
  const long cap = long(unpp::num_sms()) * per_sm;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void fma8(float (&acc)[8], const uint4& u, float w) {
  acc[0] = fmaf(w, bf_lo(u.x), acc[0]), acc[1] = fmaf(w, bf_hi(u.x), acc[1]);
  acc[2] = fmaf(w, bf_lo(u.y), acc[2]), acc[3] = fmaf(w, bf_hi(u.y), acc[3]);
  acc[4] = fmaf(w, bf_lo(u.z), acc[4]), acc[5] = fmaf(w, bf_hi(u.z), acc[5]);
  acc[6] = fmaf(w, bf_lo(u.w), acc[6]), acc[7] = fmaf(w, bf_hi(u.w), acc[7]);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}

// Source coordinate of output index o for align_corners=True, the way ATen computes it (UpSample.h
// area_pixel_compute_source_index): src = o * (in-1)/(out-1) in float, i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0.
struct Tap {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Tap tap_of(int o, int in, float scale) {
  const float src = scale * float(o);
  Tap t;
  t.i0 = int(src);
  t.i1 = t.i0 + (t.i0 < in - 1 ? 1 : 0);
  t.l1 = src - float(t.i0);
  t.l0 = 1.f - t.l1;
  return t;
}

// y[N,2H,2W,C] = bilinear_x2(x[N,H,W,C]); one thread = 8 channels of one output pixel.
__global__ void bilinear_up2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int C8, float sy, float sx) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const long total = long(N) * 2 * H * 2 * W * C8;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    long t = i;
    const int c8 = int(t % C8);
    t /= C8;
    const int ox = int(t % (2 * W));
    t /= 2 * W;
    const int oy = int(t % (2 * H)), n = int(t / (2 * H));
    const Tap ty = tap_of(oy, H, sy), tx = tap_of(ox, W, sx);
    const uint4* base = x + long(n) * H * W * C8 + c8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    fma8(acc, __ldg(base + (long(ty.i0) * W + tx.i0) * C8), ty.l0 * tx.l0);
    fma8(acc, __ldg(base + (long(ty.i0) * W + tx.i1) * C8), ty.l0 * tx.l1);
    fma8(acc, __ldg(base + (long(ty.i1) * W + tx.i0) * C8), ty.l1 * tx.l0);
    fma8(acc, __ldg(base + (long(ty.i1) * W + tx.i1) * C8), ty.l1 * tx.l1);
    y[i] = pack8(acc);
  }
}

// Weight with which output index o reads input index i (0 for most o): the adjoint gathers with the SAME float weights
// the forward kernel scatters with, so <up(x), g> == <x, up^T(g)> holds to rounding.
__device__ __forceinline__ float weight_of(int o, int i, int in, float scale) {
  const Tap t = tap_of(o, in, scale);
  return (t.i0 == i ? t.l0 : 0.f) + (t.i1 == i ? t.l1 : 0.f);
}

// dx[N,H,W,C] = bilinear_x2^T(dy[N,2H,2W,C]); one thread = 8 channels of one input pixel.  With scale = (H-1)/(2H-1) < 1/2
// only output rows 2i-2 .. 2i+3 can touch input row i (same for columns).
__global__ void bilinear_up2x_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int N, int H, int W, int C8, float sy, float sx) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const long total = long(N) * H * W * C8;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    long t = i;
    const int c8 = int(t % C8);
    t /= C8;
    const int ix = int(t % W);
    t /= W;
    const int iy = int(t % H), n = int(t / H);
    float wx[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int ox = 2 * ix - 2 + j;
      wx[j] = (ox >= 0 && ox < 2 * W) ? weight_of(ox, ix, W, sx) : 0.f;
    }
    const uint4* base = dy + long(n) * 2 * H * 2 * W * C8 + c8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 6; ++k) {
      const int oy = 2 * iy - 2 + k;
      if (oy < 0 || oy >= 2 * H) continue;
      const float wy = weight_of(oy, iy, H, sy);
      if (wy == 0.f) continue;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (wx[j] == 0.f) continue;
        fma8(acc, __ldg(base + (long(oy) * 2 * W + (2 * ix - 2 + j)) * C8), wy * wx[j]);
      }
    }
    dx[i] = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------
// Flat-buffer optimizers.  sc[0..3] are the per-step scalars (host-computed or written by optim_prep_kernel):
//   ADAMW:    sc0 = lr*sqrt(1-b2^t)/(1-b1^t)
//   ADAM:     sc0 = lr/(1-b1^t), sc1 = sqrt(1-b2^t)
//   ADABOUND: sc0 = lr*sqrt(1-b2^t)/(1-b1^t), sc1 = lower bound, sc2 = upper bound
//   SGD/SGDW: sc0 = lr, sc3 = 1 on the first step (momentum buffer := gradient)
struct OptimScalars {
  float v[4];
};

__device__ __host__ inline void optim_scalars(const UnppOptimArgs& a, double lr, unsigned long long t, float* sc) {
  const double bc1 = 1.0 - pow(double(a.beta1), double(t)), bc2 = 1.0 - pow(double(a.beta2), double(t));
  sc[0] = sc[1] = sc[2] = 0.f;
  sc[3] = t == 1ull ? 1.f : 0.f;
  switch (a.kind) {
    case UNPP_OPT_ADAMW: sc[0] = float(lr * sqrt(bc2) / bc1); break;
    case UNPP_OPT_ADAM: sc[0] = float(lr / bc1), sc[1] = float(sqrt(bc2)); break;
    case UNPP_OPT_ADABOUND: {
      const double flr = double(a.final_lr) * lr / double(a.base_lr);  // adabound.py:119
      sc[0] = float(lr * sqrt(bc2) / bc1);
      sc[1] = float(flr * (1.0 - 1.0 / (double(a.gamma) * double(t) + 1.0)));
      sc[2] = float(flr * (1.0 + 1.0 / (double(a.gamma) * double(t))));
      break;
    }
    default: sc[0] = float(lr); break;
  }
}

__global__ void optim_prep_kernel(UnppOptimArgs a, unsigned long long* counter, const float* lr_dev, float* sc) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const unsigned long long t = *counter + 1ull;
  *counter = t;
  optim_scalars(a, double(lr_dev ? *lr_dev : a.lr), t, sc);
}

// One element of any optimizer (see UnppOptimArgs): shared by the flat-buffer kernel and the multi-tensor kernel.
__device__ __forceinline__ void optim_update(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s1, float* __restrict__ s2, long i,
                                             const UnppOptimArgs& a, const float (&sc)[4]) {
  const float wd = a.weight_decay, b1 = a.beta1, b2 = a.beta2, eps = a.eps;
  const float po = p[i];
  float gr = g[i] * a.grad_scale;
  float pn = po;
  switch (a.kind) {
    case UNPP_OPT_ADAMW: {  // tools/optimizers/adamw.py:38-100: decay = wd * p_old, NOT scaled by lr
      const float mm = s1[i] * b1 + (1.f - b1) * gr, vv = s2[i] * b2 + (1.f - b2) * gr * gr;
      s1[i] = mm, s2[i] = vv;
      pn = po - sc[0] * (mm / (sqrtf(vv) + eps));
      if (wd != 0.f) pn -= po * wd;
      break;
    }
    case UNPP_OPT_ADAM: {  // torch.optim.Adam (trainer.py:351-355): L2 decay folded into the gradient
      if (wd != 0.f) gr = fmaf(wd, po, gr);
      const float mm = s1[i] * b1 + (1.f - b1) * gr, vv = s2[i] * b2 + (1.f - b2) * gr * gr;
      s1[i] = mm, s2[i] = vv;
      pn = po - sc[0] * (mm / (sqrtf(vv) / sc[1] + eps));
      break;
    }
    case UNPP_OPT_ADABOUND: {  // tools/optimizers/adabound.py:57-122 (amsbound = False)
      if (wd != 0.f) gr = fmaf(wd, po, gr);
      const float mm = s1[i] * b1 + (1.f - b1) * gr, vv = s2[i] * b2 + (1.f - b2) * gr * gr;
      s1[i] = mm, s2[i] = vv;
      const float rate = fminf(fmaxf(sc[0] / (sqrtf(vv) + eps), sc[1]), sc[2]);
      pn = po - rate * mm;
      break;
    }
    case UNPP_OPT_SGD: {  // torch.optim.SGD (trainer.py:345-349): beta1 = momentum, dampening 0, no Nesterov
      if (wd != 0.f) gr = fmaf(wd, po, gr);
      if (b1 != 0.f) {
        gr = sc[3] != 0.f ? gr : s1[i] * b1 + gr;
        s1[i] = gr;
      }
      pn = po - sc[0] * gr;
      break;
    }
    default: {  // UNPP_OPT_SGDW — tools/optimizers/sgdw.py:77-110 AS SHIPPED: the momentum buffer is maintained
      // (beta1 = momentum, beta2 = dampening) but the step direction is never applied; p only decays by wd * p.
      if (b1 != 0.f) s1[i] = sc[3] != 0.f ? gr : s1[i] * b1 + (1.f - b2) * gr;
      if (wd != 0.f) pn = po - wd * po;
      break;
    }
  }
  p[i] = pn;
}

__global__ void optim_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s1, float* __restrict__ s2, long n,
                                  UnppOptimArgs a, OptimScalars host_sc, const float* __restrict__ sc_dev) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  float sc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) sc[k] = sc_dev ? sc_dev[k] : host_sc.v[k];
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x) optim_update(p, g, s1, s2, i, a, sc);
}

// The same update for MANY tensors in one launch (the drop-in optimizers on a model whose parameters are separate allocations: 74 tensors,
// the reference issues ~10 launches per tensor): block b owns 1024 consecutive elements of the tensor whose block range contains b.
__global__ void __launch_bounds__(256) optim_step_multi_kernel(const UnppOptimTensor* __restrict__ table, int ntensors, UnppOptimArgs a, OptimScalars host_sc) {
  unpp::pdl_wait();
  unpp::pdl_trigger();
  int t = 0;
  while (t < ntensors - 1 && int(blockIdx.x) >= __ldg(&table[t].block_end)) ++t;
  const UnppOptimTensor e = table[t];
  const int first = t ? __ldg(&table[t - 1].block_end) : 0;
  const long base = long(int(blockIdx.x) - first) * 1024;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long i = base + k * 256 + threadIdx.x;
    if (i < e.n) optim_update(e.p, e.g, e.state1, e.state2, i, a, host_sc.v);
  }
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

static int check_up(const void* a, const void* b, int N, int H, int W, int C, const char* who) {
  if (!a || !b || N < 1 || H < 1 || W < 1 || C < 8 || (C & 7)) return unpp::fail(UNPP_ERR_BAD_ARG, "%s: null pointer, empty grid or C not a multiple of 8", who);
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) return unpp::fail(UNPP_ERR_BAD_ARG, "%s: pointers must be 16-byte aligned", who);
  return UNPP_OK;
}
static inline float ac_scale(int in) { return in > 1 ? float(in - 1) / float(2 * in - 1) : 0.f; }  // ATen area_pixel_compute_scale, align_corners

extern "C" int unpp_bilinear_up2x(const void* x, void* y, int N, int H, int W, int C, unpp_stream_t stream) {
  if (int rc = check_up(x, y, N, H, W, C, "bilinear_up2x")) return rc;
  unpp::launch(bilinear_up2x_kernel, grid_for(long(N) * 4 * H * W * (C / 8), 256), 256, 0, STREAM(stream), reinterpret_cast<const uint4*>(x),
               reinterpret_cast<uint4*>(y), N, H, W, C / 8, ac_scale(H), ac_scale(W));
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("bilinear_up2x: launch");
  return UNPP_OK;
}

extern "C" int unpp_bilinear_up2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, unpp_stream_t stream) {
  if (int rc = check_up(dy, dx, N, H, W, C, "bilinear_up2x_bwd")) return rc;
  unpp::launch(bilinear_up2x_bwd_kernel, grid_for(long(N) * H * W * (C / 8), 256), 256, 0, STREAM(stream), reinterpret_cast<const uint4*>(dy),
               reinterpret_cast<uint4*>(dx), N, H, W, C / 8, ac_scale(H), ac_scale(W));
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("bilinear_up2x_bwd: launch");
  return UNPP_OK;
}

extern "C" int unpp_optim_step(float* p, const float* g, float* state1, float* state2, long n, const UnppOptimArgs* a, int step,
                               uint64_t* step_counter, const float* lr_dev, float* scalars_scratch, unpp_stream_t stream) {
  if (!p || !g || n < 1 || !a) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step: bad argument");
  if (a->kind < UNPP_OPT_ADAMW || a->kind > UNPP_OPT_SGDW) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step: unknown optimizer kind %d", a->kind);
  const bool two = a->kind == UNPP_OPT_ADAMW || a->kind == UNPP_OPT_ADAM || a->kind == UNPP_OPT_ADABOUND;
  if (two && (!state1 || !state2)) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step: Adam-family optimizers need both moment buffers");
  if (!two && a->beta1 != 0.f && !state1) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step: momentum needs a buffer");
  if (a->kind == UNPP_OPT_ADABOUND && !(a->base_lr > 0.f)) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step: AdaBound needs base_lr > 0");
  OptimScalars hs = {};
  if (step_counter) {  // step count (and optionally lr) live on the device: CUDA-graph friendly
    if (!scalars_scratch) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step: a device step counter needs scalars_scratch[4]");
    unpp::launch(optim_prep_kernel, 1, 1, 0, STREAM(stream), *a, reinterpret_cast<unsigned long long*>(step_counter), lr_dev, scalars_scratch);
  } else {
    if (step < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step: step is 1-based");
    if (lr_dev) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step: lr_dev needs the device step counter");
    optim_scalars(*a, double(a->lr), (unsigned long long)step, hs.v);
    scalars_scratch = nullptr;
  }
  unpp::launch(optim_step_kernel, grid_for(n, 256), 256, 0, STREAM(stream), p, g, state1, state2, n, *a, hs, static_cast<const float*>(scalars_scratch));
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("optim_step: launch");
  return UNPP_OK;
}

extern "C" int unpp_optim_step_multi(const UnppOptimTensor* table_dev, int ntensors, int total_blocks, const UnppOptimArgs* a, int step, unpp_stream_t stream) {
  if (!table_dev || ntensors < 1 || total_blocks < 1 || !a || step < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step_multi: bad argument");
  if (a->kind < UNPP_OPT_ADAMW || a->kind > UNPP_OPT_SGDW) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step_multi: unknown optimizer kind %d", a->kind);
  if (a->kind == UNPP_OPT_ADABOUND && !(a->base_lr > 0.f)) return unpp::fail(UNPP_ERR_BAD_ARG, "optim_step_multi: AdaBound needs base_lr > 0");
  OptimScalars hs = {};
  optim_scalars(*a, double(a->lr), (unsigned long long)step, hs.v);
  unpp::launch(optim_step_multi_kernel, total_blocks, 256, 0, STREAM(stream), table_dev, ntensors, *a, hs);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("optim_step_multi: launch");
  return UNPP_OK;
}

extern "C" int unpp_sizeof_optim_args(void) { return int(sizeof(UnppOptimArgs)); }
extern "C" int unpp_sizeof_optim_tensor(void) { return int(sizeof(UnppOptimTensor)); }

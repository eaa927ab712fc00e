// train_kernels.cu — the memory-bound kernels of the training step that sit between the
// tensor-core convolutions: BatchNorm batch statistics finalisation and apply (+ReLU, +2x2 pool),
// BatchNorm / max-pool backward, the fused head backward (sigmoid' + 1x1 head dgrad/wgrad + dropout
// mask, optionally with the MSE loss and its gradient computed in place), deterministic partial-sum
// reductions, the flat-buffer AdamW of the reference, and the dropout keep-mask generator.
// All are coalesced, vectorised (16 B per thread) CUDA-core kernels bounded by HBM bandwidth.
#include "common.h"
#include "../../include/unpp.h"
#include <cuda_bf16.h>
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
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf_lo(u.x), f[1] = bf_hi(u.x), f[2] = bf_lo(u.y), f[3] = bf_hi(u.y);
  f[4] = bf_lo(u.z), f[5] = bf_hi(u.z), f[6] = bf_lo(u.w), f[7] = bf_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack2(f[0], f[1]), u.y = pack2(f[2], f[3]), u.z = pack2(f[4], f[5]), u.w = pack2(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------------
// Deterministic partial-sum reductions.  A block of 256 threads owns 32 consecutive outputs; its 8
// warps sum disjoint, interleaved slices of the partials (p = w, w+8, ...), the 8 slice sums are then
// added in a fixed order — 8x more loads in flight than one thread per output, same result every run.
template <typename F>
__device__ __forceinline__ float sliced_sum(int nparts, F load) {
  __shared__ float red[8][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  // eight loads in flight per thread, added in the original order (bit-identical to the plain loop, ~6x less latency)
  float s = 0.f;
  int p = w;
  for (; p + 56 < nparts; p += 64) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = load(p + 8 * k);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  for (; p < nparts; p += 8) s += load(p);
  red[w][l] = s;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][l];
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// out[i] (+)= scale * sum_p partial[p * stride + i]
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int nparts, long stride, int n, float scale,
                                                              float* __restrict__ out, int accumulate) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  for (int base = blockIdx.x * 32; base < n; base += gridDim.x * 32) {
    const int i = base + (threadIdx.x & 31);
    const float t = sliced_sum(nparts, [&](int p) { return i < n ? __ldg(partial + p * stride + i) : 0.f; });
    if (threadIdx.x < 32 && i < n) out[i] = accumulate ? out[i] + t * scale : t * scale;
  }
}

// wgrad partial [nparts][taps][cin_total][cout] -> dst[co*s_co + ci*s_ci + tap*s_tap] for ci < ci_count
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, int taps, int cin_total, int cout,
                                                           float* __restrict__ dst, int ci_begin, int ci_count, long s_co, long s_ci, long s_tap,
                                                           float scale) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const int n = taps * ci_count * cout;
  const long stride = long(taps) * cin_total * cout;
  for (int base = blockIdx.x * 32; base < n; base += gridDim.x * 32) {
    const int i = base + (threadIdx.x & 31);
    const int co = i % cout, ci = (i / cout) % ci_count, tap = i / (cout * ci_count);
    const float* q = partial + (long(tap) * cin_total + ci_begin + ci) * cout + co;
    const float t = sliced_sum(nparts, [&](int p) { return i < n ? __ldg(q + p * stride) : 0.f; });
    if (threadIdx.x < 32 && i < n) dst[co * s_co + ci * s_ci + tap * s_tap] = t * scale;
  }
}

// All deferred reductions of a training step in one launch: block b owns 128 consecutive outputs (32 lanes x float4
// along co; jobs whose partials are not 16-byte aligned fall back to 32 scalar outputs per block pass) of the job
// whose block range contains b (the table is small: a linear scan of the running totals).  The eight warps sum
// interleaved slices of the partials, the slice sums are added in a fixed order: deterministic.
__global__ void __launch_bounds__(256) reduce_batched_kernel(const UnppReduceJob* __restrict__ table, int njobs) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  __shared__ float4 red4[8][32];
  int j = 0;
  while (j < njobs - 1 && int(blockIdx.x) >= __ldg(&table[j].block_end)) ++j;
  const UnppReduceJob job = table[j];
  const int first = j ? __ldg(&table[j - 1].block_end) : 0;
  const int n = job.taps * job.ci_count * job.cout;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const bool vec = !(job.cout & 3) && !(job.stride & 3) && !(reinterpret_cast<uintptr_t>(job.partial) & 15);
  for (int sub = 0; sub < (vec ? 1 : 4); ++sub) {
    const int i = (int(blockIdx.x) - first) * 128 + (vec ? l * 4 : sub * 32 + l);
    const int co = i % job.cout, ci = (i / job.cout) % job.ci_count, tap = i / (job.cout * job.ci_count);
    const float* q = job.partial + (long(tap) * job.cin_total + job.ci_begin + ci) * job.cout + co;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      if (vec) {
        int p = w;
        for (; p + 56 < job.nparts; p += 64) {  // eight 16-byte loads in flight, added in the original order
          float4 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(q + (p + 8 * k) * job.stride));
#pragma unroll
          for (int k = 0; k < 8; ++k) s.x += v[k].x, s.y += v[k].y, s.z += v[k].z, s.w += v[k].w;
        }
        for (; p < job.nparts; p += 8) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(q + p * job.stride));
          s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
        }
      } else {
        for (int p = w; p < job.nparts; p += 8) s.x += __ldg(q + p * job.stride);
      }
    }
    red4[w][l] = s;
    __syncthreads();
    if (w == 0 && i < n) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) t.x += red4[k][l].x, t.y += red4[k][l].y, t.z += red4[k][l].z, t.w += red4[k][l].w;
      float* d = job.dst + co * job.s_co + ci * job.s_ci + tap * job.s_tap;
      d[0] = t.x * job.scale;
      if (vec) d[job.s_co] = t.y * job.scale, d[2 * job.s_co] = t.z * job.scale, d[3 * job.s_co] = t.w * job.scale;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// BatchNorm2d training statistics (reference models/unet.py:133, nn.BatchNorm2d eps 1e-5 momentum 0.1):
// reduce the per-CTA (sum, sum of squares) partials of the conv epilogue, produce mean / inverse std
// and the fused affine (scale, shift), and update the running statistics (unbiased variance).
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partial, int nparts, int C, float count,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                                                          float* __restrict__ mean, float* __restrict__ istd, float* __restrict__ scale,
                                                          float* __restrict__ shift) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  // block = 32 channels x 8 slices of the partials; sums in double (2*C*nparts values: free, and it removes the
  // E[z^2]-m^2 cancellation); fixed summation order
  __shared__ double r1[8][32], r2[8][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, c = blockIdx.x * 32 + l;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    int p = w;
    for (; p + 56 < nparts; p += 64) {  // sixteen loads in flight, added in the original order
      float a[8], b[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = __ldg(partial + (size_t(p + 8 * k) * 2 + 0) * C + c), b[k] = __ldg(partial + (size_t(p + 8 * k) * 2 + 1) * C + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) s1 += double(a[k]), s2 += double(b[k]);
    }
    for (; p < nparts; p += 8) {
      s1 += double(__ldg(partial + (size_t(p) * 2 + 0) * C + c));
      s2 += double(__ldg(partial + (size_t(p) * 2 + 1) * C + c));
    }
  }
  r1[w][l] = s1, r2[w][l] = s2;
  __syncthreads();
  if (w != 0 || c >= C) return;
  s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s1 += r1[k][l], s2 += r2[k][l];
  const double m = s1 / count;
  double var = s2 / count - m * m;
  if (var < 0.0) var = 0.0;
  const float is = float(1.0 / sqrt(var + double(eps)));
  mean[c] = float(m);
  istd[c] = is;
  const float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - float(m) * sc;
  if (running_mean) {
    const double unbiased = count > 1.f ? var * double(count) / double(count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * float(m);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * float(unbiased);
  }
}

// y = relu(z * scale[c] + shift[c]) (bf16 NHWC, 8 channels per thread); optional 2x2 max-pooled copy.
__global__ void bn_relu_kernel(const uint4* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift, uint4* __restrict__ y,
                               long total, int C8) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const long stride = long(gridDim.x) * blockDim.x, i0 = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (stride % C8 == 0) {  // the thread keeps its eight channels for the whole loop: their coefficients live in registers
    const int c0 = int(i0 % C8) * 8;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sc[k] = __ldg(scale + c0 + k), sh[k] = __ldg(shift + c0 + k);
    for (long i = i0; i < total; i += stride) {
      float f[8];
      unpack8(__ldg(z + i), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
      y[i] = pack8(f);
    }
    return;
  }
  for (long i = i0; i < total; i += stride) {
    const int c0 = int(i % C8) * 8;
    float f[8];
    unpack8(__ldg(z + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], __ldg(scale + c0 + k), __ldg(shift + c0 + k)), 0.f);
    y[i] = pack8(f);
  }
}
__global__ void bn_relu_pool_kernel(const uint4* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                                    uint4* __restrict__ y, uint4* __restrict__ pooled, int N, int H, int W, int C8) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const int Ho = H / 2, Wo = W / 2;
  const long total = long(N) * Ho * Wo * C8;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    long t = i;
    const int c8 = int(t % C8);
    t /= C8;
    const int xo = int(t % Wo);
    t /= Wo;
    const int yo = int(t % Ho);
    const long n = t / Ho;
    float sc[8], sh[8], mx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sc[k] = __ldg(scale + c8 * 8 + k), sh[k] = __ldg(shift + c8 * 8 + k), mx[k] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long idx = ((n * H + 2 * yo + (q >> 1)) * W + 2 * xo + (q & 1)) * C8 + c8;
      float f[8];
      unpack8(__ldg(z + idx), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
      const uint4 packed = pack8(f);
      y[idx] = packed;
      float r[8];
      unpack8(packed, r);  // pool the bf16-rounded values, exactly what a pool over the stored y would see
#pragma unroll
      for (int k = 0; k < 8; ++k) mx[k] = fmaxf(mx[k], r[k]);
    }
    pooled[i] = pack8(mx);
  }
}

// ------------------------------------------------------------------------------------------
// MaxPool2d(2) backward: the gradient of each pooled element goes to the FIRST maximum of its 2x2
// window in row-major order (what ATen's max_pool2d_with_indices records); the other three get 0.
__global__ void maxpool_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dp, uint4* __restrict__ dx, int N, int H, int W, int C8) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const int Ho = H / 2, Wo = W / 2;
  const long total = long(N) * Ho * Wo * C8;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    long t = i;
    const int c8 = int(t % C8);
    t /= C8;
    const int xo = int(t % Wo);
    t /= Wo;
    const int yo = int(t % Ho);
    const long n = t / Ho;
    float v[4][8], g[8];
    long idx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      idx[q] = ((n * H + 2 * yo + (q >> 1)) * W + 2 * xo + (q & 1)) * C8 + c8;
      unpack8(__ldg(x + idx[q]), v[q]);
    }
    unpack8(__ldg(dp + i), g);
    float o[4][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int best = 0;
      float bv = v[0][k];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (v[q][k] > bv) bv = v[q][k], best = q;
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][k] = (q == best) ? g[k] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) dx[idx[q]] = pack8(o[q]);
  }
}

// BatchNorm backward, apply pass: dz = gamma*istd * (dyh - s1/M - xhat * s2/M), xhat = (z-mean)*istd,
// with s1 = sum dyh (= dbeta), s2 = sum dyh*xhat (= dgamma) reduced beforehand (conv_tc epilogue stats).
__global__ void bn_bwd_apply_kernel(const uint4* __restrict__ dyh, const uint4* __restrict__ z, const float* __restrict__ mean,
                                    const float* __restrict__ istd, const float* __restrict__ gamma, const float* __restrict__ sums, float inv_count,
                                    uint4* __restrict__ dz, long total, int C8) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const int C = C8 * 8;
  const long stride = long(gridDim.x) * blockDim.x, i0 = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (stride % C8 == 0) {
    // The thread keeps its eight channels for the whole loop: dz = a*dyh + b*z + c with per-channel coefficients in registers
    //   a = gamma*istd, b = -gamma*istd^2*s2/M, c = gamma*istd*(mean*istd*s2/M - s1/M).
    const int c0 = int(i0 % C8) * 8;
    float ca[8], cb[8], cc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      const float is = __ldg(istd + c), gm = __ldg(gamma + c) * is, s1 = __ldg(sums + c) * inv_count, s2 = __ldg(sums + C + c) * inv_count;
      ca[k] = gm, cb[k] = -gm * is * s2, cc[k] = gm * (__ldg(mean + c) * is * s2 - s1);
    }
    for (long i = i0; i < total; i += stride) {
      float g[8], zz[8], o[8];
      unpack8(__ldg(dyh + i), g);
      unpack8(__ldg(z + i), zz);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(cb[k], zz[k], fmaf(ca[k], g[k], cc[k]));
      dz[i] = pack8(o);
    }
    return;
  }
  for (long i = i0; i < total; i += stride) {
    const int c0 = int(i % C8) * 8;
    float g[8], zz[8], o[8];
    unpack8(__ldg(dyh + i), g);
    unpack8(__ldg(z + i), zz);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      const float is = __ldg(istd + c);
      const float xh = (zz[k] - __ldg(mean + c)) * is;
      o[k] = __ldg(gamma + c) * is * (g[k] - __ldg(sums + c) * inv_count - xh * __ldg(sums + C + c) * inv_count);
    }
    dz[i] = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------
// Head backward.  Forward was heat = sigmoid(Wh . (x * mask * scale) + bh) with x = X_0k (bf16 NHWC,
// 16 channels), reference models/unet.py:283-286.  Given either the upstream gradient dheat (fp32
// NCHW) or — fused MSE mode — the target T (then dheat = coef * (p - T), loss += (p - T)^2), computes
//   dlogit = dheat * p * (1 - p)
//   dx[c]  = [x[c] > 0] * mask[c]*scale * sum_cls dlogit[cls] * Wh[cls][c]   (bf16 NHWC; x = relu(.) so the ReLU mask of the
//            producing conv is applied here) ;  dxsum[c] += dx[c]  (bias gradient of that conv when the head is its only consumer)
//   dWh[cls][c] += dlogit[cls] * x[c]*mask[c]*scale ;  dbh[cls] += dlogit[cls] ; loss partial
// One thread per pixel; per-thread register accumulators, one shuffle+smem reduction per CTA at
// the end, written to partial[blockIdx.x][NCLS*16 + NCLS + 1 + 16] = dW, db, loss, dxsum.
// FOCAL is a compile-time switch: the powf/logf path costs registers and instructions the MSE / upstream path must not pay.
template <int NCLS, bool FOCAL>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ heat, const float* __restrict__ dheat, const float* __restrict__ target,
                                                       float gamma, float coef, const uint2* __restrict__ x, const uint16_t* __restrict__ mask, float drop_scale,
                                                       const float* __restrict__ head_w, uint2* __restrict__ dx, float* __restrict__ partial,
                                                       int N, long HW) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  // Four threads per pixel, each owning four of the 16 channels (8 B of x / dx, 4 B of the keep-mask): 25 accumulators
  // per thread instead of 85, so three blocks per SM are resident and enough loads are in flight to stream at HBM speed.
  // The four lanes of a pixel read the same heat / target words (one broadcast transaction).
  constexpr int NACC = NCLS * 16 + NCLS + 1 + 16, LOSS = NCLS * 16 + NCLS;
  const int sub = threadIdx.x & 3;
  const float lead = sub == 0 ? 1.f : 0.f;  // per-pixel sums (bias gradient, loss) are counted by one lane of the four
  float accw[NCLS][4], accb[NCLS], accx[4], accl = 0.f, w[NCLS][4];
#pragma unroll
  for (int c = 0; c < NCLS; ++c) {
    accb[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) accw[c][k] = 0.f, w[c][k] = __ldg(head_w + c * 16 + 4 * sub + k);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) accx[k] = 0.f;
  // One step of a lane quad = four consecutive pixels (HW % 4 == 0, so they share an image): the heat / target words of
  // a class are ONE float4, and all 16 loads of the step are issued before any is used (the kernel is latency-bound).
  const long quads = long(N) * HW / 4, step = long(gridDim.x) * 64L, HW4 = HW / 4;
  long i4 = blockIdx.x * 64L + (threadIdx.x >> 2);
  long n = i4 / HW4, p4 = i4 % HW4;  // (image, pixel quad) advanced incrementally: no 64-bit division in the loop
  const long step_n = step / HW4, step_p = step % HW4;
  for (; i4 < quads; i4 += step, n += step_n, p4 += step_p) {
    if (p4 >= HW4) p4 -= HW4, ++n;
    float4 hh[NCLS], tt[NCLS];
#pragma unroll
    for (int c = 0; c < NCLS; ++c) {
      const long o = (n * NCLS + c) * HW + 4 * p4;
      hh[c] = __ldg(reinterpret_cast<const float4*>(heat + o));
      tt[c] = __ldg(reinterpret_cast<const float4*>((target ? target : dheat) + o));
    }
    uint2 xq[4];
    uint32_t mq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      xq[u] = __ldg(x + 4 * (4 * i4 + u) + sub);
      mq[u] = mask ? uint32_t(__ldg(mask + 4 * i4 + u)) >> (4 * sub) : 0xFu;  // the keep bits of this lane's four channels
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint2 xr = xq[u];
      const uint32_t mw = mq[u];
      float dl[NCLS];
#pragma unroll
      for (int c = 0; c < NCLS; ++c) {
        const float pr = (&hh[c].x)[u], tv = (&tt[c].x)[u];
        float dh;
        if (target) {
          const float d = pr - tv;
          if constexpr (!FOCAL) {  // MSE: loss += d^2, d loss / d p = coef * d
            accl = fmaf(lead * d, d, accl);
            dh = coef * d;
          } else {  // FocalLoss_BCE_2d (tools/losses/focal_loss.py:264-301): a = |p-t|, e = 1-a+1e-20, loss += -a^gamma * log(e)
            const float a = fabsf(d), e = 1.f - a + 1e-20f;
            // gamma = 3 (the trainer's setting, trainer.py:426) needs no pow; log through the fast intrinsic (rel. error ~1e-6)
            const float le = __logf(e), pg1 = gamma == 3.f ? a * a : (a > 0.f ? __powf(a, gamma - 1.f) : 0.f);
            accl += lead * (-(pg1 * a) * le);
            dh = coef * copysignf(-gamma * pg1 * le + (pg1 * a) / e, d);
          }
        } else {
          dh = tv;  // upstream gradient
        }
        dl[c] = dh * pr * (1.f - pr);
        accb[c] = fmaf(lead, dl[c], accb[c]);
      }
      const float xv[4] = {bf_lo(xr.x), bf_hi(xr.x), bf_lo(xr.y), bf_hi(xr.y)};
      float g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float keep = mask ? (((mw >> k) & 1u) ? drop_scale : 0.f) : 1.f;
        const float xd = xv[k] * keep;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < NCLS; ++c) {
          accw[c][k] = fmaf(dl[c], xd, accw[c][k]);
          s = fmaf(dl[c], w[c][k], s);
        }
        g[k] = xv[k] > 0.f ? s * keep : 0.f;
        g[k] = __bfloat162float(__float2bfloat16_rn(g[k]));  // sum exactly what is stored
        accx[k] += g[k];
      }
      dx[4 * (4 * i4 + u) + sub] = make_uint2(pack2(g[0], g[1]), pack2(g[2], g[3]));
    }
  }
  // fixed-order reduction: lanes of equal channel group (xor 4, 8, 16), then the eight warps
  __shared__ float red[8][NACC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto lanes = [](float v) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
#pragma unroll
  for (int c = 0; c < NCLS; ++c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v = lanes(accw[c][k]);
      if (lane < 4) red[warp][c * 16 + 4 * sub + k] = v;
    }
    const float b = lanes(accb[c]);
    if (lane == 0) red[warp][NCLS * 16 + c] = b;
  }
  {
    const float l = lanes(accl);
    if (lane == 0) red[warp][LOSS] = l;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = lanes(accx[k]);
    if (lane < 4) red[warp][LOSS + 1 + 4 * sub + k] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NACC; i += blockDim.x) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][i];
    partial[size_t(blockIdx.x) * NACC + i] = s;
  }
}

// ------------------------------------------------------------------------------------------
// The reference's AdamW (tools/optimizers/adamw.py:38-100), one pass over the flat parameter buffer:
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p = p - step_size * m / (sqrt(v)+eps) - wd * p_old
// (decay is NOT scaled by lr, and uses the pre-update p).  step_size is computed on the host in
// double like the reference does in Python floats.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n, float b1,
                             float b2, float eps, float step_size, const float* __restrict__ step_size_dev, float wd, float grad_scale) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  if (step_size_dev) step_size = *step_size_dev;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x) {
    const float gr = g[i] * grad_scale;
    const float mm = m[i] * b1 + (1.f - b1) * gr;
    const float vv = v[i] * b2 + (1.f - b2) * gr * gr;
    m[i] = mm, v[i] = vv;
    const float denom = sqrtf(vv) + eps;
    const float po = p[i];
    float pn = po - step_size * (mm / denom);
    if (wd != 0.f) pn -= po * wd;
    p[i] = pn;
  }
}

// Device-side step bookkeeping so that a whole training step can live in one CUDA graph: bumps the
// step counter and derives AdamW's bias-corrected step size from it (adamw.py:86-88).
__global__ void adamw_prep_kernel(unsigned long long* counter, float lr, float b1, float b2, float* step_size) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const unsigned long long t = *counter + 1ull;
  *counter = t;
  const double bc1 = 1.0 - pow(double(b1), double(t)), bc2 = 1.0 - pow(double(b2), double(t));
  *step_size = float(double(lr) * sqrt(bc2) / bc1);
}

// Dropout keep-mask (nn.Dropout(p), models/unet.py:254): u8 1 = keep, counter-based hash RNG
// (not torch's Philox stream: parity runs pass explicit masks; this serves throughput runs).
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
  return x;
}
__global__ void dropout_mask_kernel(uint4* __restrict__ out, long n8, uint32_t seed_lo, uint32_t seed_hi, uint32_t thresh16,
                                    const unsigned long long* __restrict__ counter) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  if (counter) {  // per-step stream: fold the device step counter into the seed
    const unsigned long long c = (*counter + 1ull) * 0x9E3779B97F4A7C15ull;
    seed_lo ^= uint32_t(c), seed_hi ^= uint32_t(c >> 32);
  }
  // one thread = eight pixels = eight 16-bit keep words (one 16-byte store); a hash yields two 16-bit uniforms
  for (long t = blockIdx.x * long(blockDim.x) + threadIdx.x; t < n8; t += long(gridDim.x) * blockDim.x) {
    uint32_t wds[4] = {0, 0, 0, 0};
#pragma unroll
    for (int px = 0; px < 8; ++px) {
      const long i = t * 8 + px;
      uint32_t bits = 0;
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        const uint32_t r = mix32(mix32(uint32_t(i) * 8u + h + seed_lo) ^ (uint32_t(i >> 29) + seed_hi));
        bits |= ((r & 0xFFFF) >= thresh16 ? 1u : 0u) << (2 * h);
        bits |= ((r >> 16) >= thresh16 ? 1u : 0u) << (2 * h + 1);
      }
      wds[px >> 1] |= bits << (16 * (px & 1));
    }
    out[t] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
  }
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int unpp_reduce_partials(const float* partial, int nparts, long stride, int n, float scale, float* out, int accumulate,
                                    unpp_stream_t stream) {
  if (!partial || !out || nparts < 1 || n < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "reduce_partials: bad argument");
  unpp::launch(reduce_partials_kernel, grid_for((long(n) + 31) / 32 * 256, 256), 256, 0, STREAM(stream), partial, nparts, stride, n, scale, out, accumulate);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("reduce_partials: launch");
  return UNPP_OK;
}

extern "C" int unpp_reduce_batched(const UnppReduceJob* table, int njobs, int total_blocks, unpp_stream_t stream) {
  if (!table || njobs < 1 || total_blocks < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "reduce_batched: bad argument");
  unpp::launch(reduce_batched_kernel, total_blocks, 256, 0, STREAM(stream), table, njobs);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("reduce_batched: launch");
  return UNPP_OK;
}
extern "C" int unpp_sizeof_reduce_job(void) { return int(sizeof(UnppReduceJob)); }

extern "C" int unpp_wgrad_reduce(const float* partial, int nparts, int taps, int cin_total, int cout, float* dst, int ci_begin, int ci_count,
                                 long s_co, long s_ci, long s_tap, float scale, unpp_stream_t stream) {
  if (!partial || !dst || nparts < 1 || taps < 1 || ci_count < 1 || ci_begin < 0 || ci_begin + ci_count > cin_total || cout < 1)
    return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad_reduce: bad argument");
  const int n = taps * ci_count * cout;
  unpp::launch(wgrad_reduce_kernel, grid_for((long(n) + 31) / 32 * 256, 256), 256, 0, STREAM(stream), partial, nparts, taps, cin_total, cout, dst, ci_begin, ci_count, s_co, s_ci,
                                                                   s_tap, scale);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("wgrad_reduce: launch");
  return UNPP_OK;
}

extern "C" int unpp_bn_finalize(const float* partial, int nparts, int C, float count, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, float momentum, float eps, float* mean, float* istd, float* scale, float* shift,
                                unpp_stream_t stream) {
  if (!partial || !gamma || !beta || !mean || !istd || !scale || !shift || nparts < 1 || C < 1 || !(count >= 1.f))
    return unpp::fail(UNPP_ERR_BAD_ARG, "bn_finalize: bad argument");
  unpp::launch(bn_finalize_kernel, (C + 31) / 32, 256, 0, STREAM(stream), partial, nparts, C, count, gamma, beta, running_mean, running_var, momentum, eps,
                                                               mean, istd, scale, shift);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("bn_finalize: launch");
  return UNPP_OK;
}

extern "C" int unpp_bn_relu(const void* z, const float* scale, const float* shift, void* y, void* pooled, int N, int H, int W, int C,
                            unpp_stream_t stream) {
  if (!z || !scale || !shift || !y || N < 1 || H < 1 || W < 1 || C % 8) return unpp::fail(UNPP_ERR_BAD_ARG, "bn_relu: bad argument");
  if (pooled) {
    if ((H & 1) || (W & 1)) return unpp::fail(UNPP_ERR_BAD_ARG, "bn_relu: pooling needs even H and W");
    const long total = long(N) * (H / 2) * (W / 2) * (C / 8);
    unpp::launch(bn_relu_pool_kernel, grid_for(total, 256), 256, 0, STREAM(stream), reinterpret_cast<const uint4*>(z), scale, shift,
                                                                          reinterpret_cast<uint4*>(y), reinterpret_cast<uint4*>(pooled), N, H, W,
                                                                          C / 8);
  } else {
    const long total = long(N) * H * W * (C / 8);
    unpp::launch(bn_relu_kernel, grid_for(total, 256), 256, 0, STREAM(stream), reinterpret_cast<const uint4*>(z), scale, shift, reinterpret_cast<uint4*>(y),
                                                                     total, C / 8);
  }
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("bn_relu: launch");
  return UNPP_OK;
}

extern "C" int unpp_maxpool2x2_bwd(const void* x, const void* dpooled, void* dx, int N, int H, int W, int C, unpp_stream_t stream) {
  if (!x || !dpooled || !dx || N < 1 || H < 2 || W < 2 || (H & 1) || (W & 1) || C % 8)
    return unpp::fail(UNPP_ERR_BAD_ARG, "maxpool2x2_bwd: need even H, W and C %% 8 == 0");
  const long total = long(N) * (H / 2) * (W / 2) * (C / 8);
  unpp::launch(maxpool_bwd_kernel, grid_for(total, 256), 256, 0, STREAM(stream), reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dpooled),
                                                                       reinterpret_cast<uint4*>(dx), N, H, W, C / 8);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("maxpool2x2_bwd: launch");
  return UNPP_OK;
}

extern "C" int unpp_bn_bwd_apply(const void* dyh, const void* z, const float* mean, const float* istd, const float* gamma, const float* sums,
                                 float count, void* dz, int N, int H, int W, int C, unpp_stream_t stream) {
  if (!dyh || !z || !mean || !istd || !gamma || !sums || !dz || N < 1 || H < 1 || W < 1 || C % 8 || !(count >= 1.f))
    return unpp::fail(UNPP_ERR_BAD_ARG, "bn_bwd_apply: bad argument");
  const long total = long(N) * H * W * (C / 8);
  unpp::launch(bn_bwd_apply_kernel, grid_for(total, 256), 256, 0, STREAM(stream), reinterpret_cast<const uint4*>(dyh), reinterpret_cast<const uint4*>(z), mean,
                                                                        istd, gamma, sums, 1.f / count, reinterpret_cast<uint4*>(dz), total, C / 8);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("bn_bwd_apply: launch");
  return UNPP_OK;
}

extern "C" int unpp_head_bwd_grid(int N, int H, int W) { return grid_for(long(N) * H * W, 256, 2); }

extern "C" int unpp_head_bwd(const float* heat, const float* dheat, const float* target, int loss_kind, float gamma, float coef, const void* x, const uint16_t* drop_mask,
                             float drop_scale, const float* head_w, int classes, void* dx, float* partial, int N, int H, int W,
                             unpp_stream_t stream) {
  if (!heat || (!dheat && !target) || !x || !head_w || !dx || !partial || N < 1 || H < 1 || W < 1)
    return unpp::fail(UNPP_ERR_BAD_ARG, "head_bwd: bad argument");
  if (loss_kind != 0 && loss_kind != 1) return unpp::fail(UNPP_ERR_BAD_ARG, "head_bwd: loss_kind must be 0 (MSE) or 1 (focal BCE)");
  if ((long(H) * W) % 4 || ((reinterpret_cast<uintptr_t>(heat) | reinterpret_cast<uintptr_t>(target ? target : dheat)) & 15))
    return unpp::fail(UNPP_ERR_BAD_ARG, "head_bwd: H*W must be a multiple of 4 and heat / target / dheat 16-byte aligned");
  const int grid = unpp_head_bwd_grid(N, H, W);
  const long HW = long(H) * W;
#define LAUNCH(NC)                                                                                                                        \
  if (loss_kind == 1)                                                                                                                     \
    unpp::launch(head_bwd_kernel<NC, true>, grid, 256, 0, STREAM(stream), heat, dheat, target, gamma, coef, reinterpret_cast<const uint2*>(x),           \
                                                        drop_mask, drop_scale, head_w,                  \
                                                        reinterpret_cast<uint2*>(dx), partial, N, HW);                                     \
  else                                                                                                                                    \
    unpp::launch(head_bwd_kernel<NC, false>, grid, 256, 0, STREAM(stream), heat, dheat, target, gamma, coef, reinterpret_cast<const uint2*>(x),          \
                                                        drop_mask, drop_scale, head_w,                  \
                                                        reinterpret_cast<uint2*>(dx), partial, N, HW)
  switch (classes) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    case 5: LAUNCH(5); break;
    case 6: LAUNCH(6); break;
    case 7: LAUNCH(7); break;
    case 8: LAUNCH(8); break;
    default: return unpp::fail(UNPP_ERR_UNSUPPORTED, "head_bwd: 1..8 classes supported");
  }
#undef LAUNCH
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("head_bwd: launch");
  return UNPP_OK;
}

extern "C" int unpp_adamw(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                          int step, float grad_scale, unpp_stream_t stream) {
  if (!p || !g || !m || !v || n < 1 || step < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "adamw: bad argument");
  const double bc1 = 1.0 - pow(double(beta1), step), bc2 = 1.0 - pow(double(beta2), step);
  const float step_size = float(double(lr) * sqrt(bc2) / bc1);
  unpp::launch(adamw_kernel, grid_for(n, 256), 256, 0, STREAM(stream), p, g, m, v, n, beta1, beta2, eps, step_size, nullptr, weight_decay, grad_scale);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("adamw: launch");
  return UNPP_OK;
}

extern "C" int unpp_adamw_dev(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, uint64_t* step_counter, float* step_size_scratch, float grad_scale, unpp_stream_t stream) {
  if (!p || !g || !m || !v || n < 1 || !step_counter || !step_size_scratch) return unpp::fail(UNPP_ERR_BAD_ARG, "adamw_dev: bad argument");
  unpp::launch(adamw_prep_kernel, 1, 1, 0, STREAM(stream), reinterpret_cast<unsigned long long*>(step_counter), lr, beta1, beta2, step_size_scratch);
  unpp::launch(adamw_kernel, grid_for(n, 256), 256, 0, STREAM(stream), p, g, m, v, n, beta1, beta2, eps, 0.f, step_size_scratch, weight_decay, grad_scale);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("adamw_dev: launch");
  return UNPP_OK;
}

extern "C" int unpp_dropout_mask(uint16_t* mask, long npix, float p_drop, uint64_t seed, const uint64_t* step_counter, unpp_stream_t stream) {
  if (!mask || npix < 8 || (npix & 7) || (reinterpret_cast<uintptr_t>(mask) & 15) || !(p_drop >= 0.f) || !(p_drop < 1.f))
    return unpp::fail(UNPP_ERR_BAD_ARG, "dropout_mask: npix must be a positive multiple of 8, mask 16-byte aligned, 0 <= p < 1");
  const uint32_t thresh = uint32_t(double(p_drop) * 65536.0 + 0.5);
  unpp::launch(dropout_mask_kernel, grid_for(npix / 8, 256), 256, 0, STREAM(stream), reinterpret_cast<uint4*>(mask), npix / 8, uint32_t(seed), uint32_t(seed >> 32),
               thresh, reinterpret_cast<const unsigned long long*>(step_counter));
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("dropout_mask: launch");
  return UNPP_OK;
}

// aux_kernels.cu — the memory-bound helpers around the tensor-core convolution:
//   weight packing (fp32 OIHW state_dict -> bf16 UMMA B-operand layout),
//   NCHW fp32 -> NHWC bf16 input conversion, 2x2 max-pool, and heat-map peak extraction.
// All are plain coalesced/vectorised CUDA-core kernels (HBM-bound byte movers).
#include <cooperative_groups.h>
#include "common.h"
#include "../../include/unpp.h"
#include "b2_blocks.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace {

// ------------------------------------------------------------------------------------------
// weight packing: one thread per packed 8-element K group (16 B store)
__constant__ b2::UnitTable<b2::kMainUnits> c_main_units = b2::main_units();
__constant__ b2::UnitTable<b2::kLowUnits> c_low_units = b2::low_units();

// kinds 4 / 5 / 6: only the non-zero (position, pixel) blocks of the 2x2-blocked layouts are stored (b2_blocks.h).
// Per K = 16 slab: `units` 16-column blocks of 512 B; inside the run of an MMA: [2 k8][N = 16 * nblk columns][8 channels].
__device__ __forceinline__ void pack_entry_b2(const UnppPackArgs& a) {
  const bool low = a.kind == 6;
  const int units = low ? b2::kLowUnits : b2::kMainUnits;
  const int k8_count = a.k_count / 8;
  const long total = long(units) * 16 * k8_count;
  for (long idx = blockIdx.x * long(blockDim.x) + threadIdx.x; idx < total; idx += long(gridDim.x) * blockDim.x) {
    const int c = int(idx & 15), u = int((idx >> 4) % units), k8 = int((idx >> 4) / units);
    const b2::Blk blk = low ? c_low_units.of[u] : c_main_units.of[u];
    const int q = blk.b0 + (u - blk.cum);  // pixel of the 2x2 block
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k = k8 * 8 + kk;
      float w;
      if (low) {  // composed transposed-conv weights [4*16][Cin][3][3], output channel = q * 16 + co (engine.compose_deconv_conv)
        w = a.src[(size_t(q * 16 + c) * a.src_I + (a.k_begin + k)) * 9 + blk.pos];
      } else {
        const int r = (blk.pos >> 2) - (q >> 1), s = (blk.pos & 3) - (q & 1);  // in [0, 2] by construction
        if (a.kind == 4) {
          w = (a.k_begin + k) < a.src_I ? a.src[(size_t(c) * a.src_I + (a.k_begin + k)) * 9 + r * 3 + s] : 0.f;
          if (a.scale) w *= a.scale[c];
        } else {
          w = a.src[(size_t(a.k_begin + k) * a.src_I + (a.n_begin + c)) * 9 + (2 - r) * 3 + (2 - s)];
        }
      }
      v[kk] = __float2bfloat16_rn(w);
    }
    const int kd = a.k_dst8 + k8, slab = kd >> 1, half = kd & 1, N = 16 * blk.nblk;
    const size_t dst_bytes = size_t(slab) * units * b2::kUnitBytes + size_t(blk.cum) * b2::kUnitBytes + size_t(half) * N * 16 + size_t((u - blk.cum) * 16 + c) * 16;
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.dst) + dst_bytes) = *reinterpret_cast<const uint4*>(v);
  }
}

// kind 7: first layer of the 2x2-blocked path with 4-channel (8-byte) pixels.  One K = 16 step is half a window row: per
// window row dy the K axis is 4 chunks of 8 = (pixel pair, 4 channels): chunk c holds window columns dx = 2c - 1, 2c
// (dx = -1 and dx >= 4 lie outside the 4x4 window: zero), chunk 3 is all zero.  Layout [dy][chunk][64 columns][8] bf16.
__device__ __forceinline__ void pack_entry_c4(const UnppPackArgs& a) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 4 * 4 * 64; idx += gridDim.x * blockDim.x) {
    const int n = idx & 63, chunk = (idx >> 6) & 3, dy = idx >> 8;
    const int q = n >> 4, co = n & 15, r = dy - (q >> 1);
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ci = e & 3, dx = 2 * chunk + (e >> 2) - 1, sft = dx - (q & 1);
      float w = 0.f;
      if (chunk < 3 && ci < a.src_I && r >= 0 && r <= 2 && sft >= 0 && sft <= 2) {
        w = a.src[(size_t(co) * a.src_I + ci) * 9 + r * 3 + sft];
        if (a.scale) w *= a.scale[co];
      }
      v[e] = __float2bfloat16_rn(w);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dst) + size_t(idx) * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

__device__ __forceinline__ void pack_entry(const UnppPackArgs& a) {
  if (a.kind == 7) {
    pack_entry_c4(a);
    return;
  }
  if (a.kind >= 4) {
    pack_entry_b2(a);
    return;
  }
  const int nt_count = a.n_total / a.n_tile;
  const int k8_count = a.k_count / 8;
  const long total = long(nt_count) * a.taps * k8_count * a.n_tile;
  for (long idx = blockIdx.x * long(blockDim.x) + threadIdx.x; idx < total; idx += long(gridDim.x) * blockDim.x) {
    long t = idx;
    const int nl = int(t % a.n_tile);
    t /= a.n_tile;
    const int k8 = int(t % k8_count);
    t /= k8_count;
    const int tap = int(t % a.taps);
    const int nt = int(t / a.taps);
    const int n = nt * a.n_tile + nl;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k = k8 * 8 + kk;
      float w;
      if (a.kind == 0) {  // B[n=co][tap][k=ci] = W[co][k_begin+ci][tap]; input channels beyond src_I are zero padding
        w = (a.k_begin + k) < a.src_I ? a.src[(size_t(n) * a.src_I + (a.k_begin + k)) * a.taps + tap] : 0.f;
        if (a.scale) w *= a.scale[n];
      } else if (a.kind == 1) {  // B[n=ci][tap][k=co] = W[co][n_begin+ci][taps-1-tap]
        w = a.src[(size_t(a.k_begin + k) * a.src_I + (a.n_begin + n)) * a.taps + (a.taps - 1 - tap)];
      } else if (a.kind == 2) {  // B[n=pq*Cout+co][0][k=ci] = Wd[ci][co][pq];  src [Cin][Cout][2][2]
        const int cout = a.src_I, pq = n / cout, co = n % cout;
        w = a.src[(size_t(a.k_begin + k) * cout + co) * 4 + pq];
      } else if (a.kind == 3) {  // B[n=ci][tap=pq][k=co] = Wd[ci][co][pq]
        w = a.src[(size_t(a.n_begin + n) * a.src_I + (a.k_begin + k)) * 4 + tap];
      }
      v[kk] = __float2bfloat16_rn(w);
    }
    const size_t dst = (((size_t(nt) * a.taps + tap) * a.k8_total + a.k_dst8 + k8) * a.n_tile + nl) * 8;
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dst) + dst) = *reinterpret_cast<const uint4*>(v);
  }
}

__global__ void pack_weights_kernel(UnppPackArgs a) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger(); pack_entry(a); }
// one launch for a whole table of pack jobs (device-resident): blockIdx.y selects the job
__global__ void pack_weights_batched_kernel(const UnppPackArgs* __restrict__ table) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const UnppPackArgs a = table[blockIdx.y];
  pack_entry(a);
}

// ------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC bf16 with channel padding.  A thread converts four consecutive pixels (one float4 per plane, 128
// contiguous output bytes) when H*W is a multiple of 4; the generic path does one pixel per thread.
template <int CPAD>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C, long HW) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const long stride = long(gridDim.x) * blockDim.x, i0 = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (!(HW & 3) && !(reinterpret_cast<uintptr_t>(x) & 15)) {
    const long HW4 = HW >> 2, quads = long(N) * HW4;
    long n = i0 / HW4, p4 = i0 % HW4;  // advanced incrementally: no 64-bit division in the loop
    const long step_n = stride / HW4, step_p = stride % HW4;
    for (long i = i0; i < quads; i += stride, n += step_n, p4 += step_p) {
      if (p4 >= HW4) p4 -= HW4, ++n;
      float4 f[CPAD];
#pragma unroll
      for (int c = 0; c < CPAD; ++c) f[c] = c < C ? __ldg(reinterpret_cast<const float4*>(x + (n * C + c) * HW) + p4) : make_float4(0.f, 0.f, 0.f, 0.f);
      uint4* o = reinterpret_cast<uint4*>(out + i * 4 * CPAD);
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        __align__(16) __nv_bfloat16 v[CPAD];
#pragma unroll
        for (int c = 0; c < CPAD; ++c) v[c] = __float2bfloat16_rn((&f[c].x)[px]);
#pragma unroll
        for (int k = 0; k < CPAD / 8; ++k) o[px * (CPAD / 8) + k] = reinterpret_cast<const uint4*>(v)[k];
      }
    }
    return;
  }
  const long total = long(N) * HW;
  for (long i = i0; i < total; i += stride) {
    const long n = i / HW, p = i % HW;
    __align__(16) __nv_bfloat16 v[CPAD];
#pragma unroll
    for (int c = 0; c < CPAD; ++c) v[c] = __float2bfloat16_rn(c < C ? __ldg(x + (n * C + c) * HW + p) : 0.f);
    uint4* o = reinterpret_cast<uint4*>(out + i * CPAD);
#pragma unroll
    for (int k = 0; k < CPAD / 8; ++k) o[k] = reinterpret_cast<const uint4*>(v)[k];
  }
}

// NCHW fp32 (C <= 4) -> NHWC bf16 with 4 channels per pixel (8 B): the input of the first-layer mode of conv_tc.  A thread
// converts four consecutive pixels (one float4 per plane in, 32 contiguous bytes out) when H*W is a multiple of 4.
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C, long HW) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const long stride = long(gridDim.x) * blockDim.x, i0 = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (!(HW & 3) && !(reinterpret_cast<uintptr_t>(x) & 15)) {
    const long HW4 = HW >> 2, quads = long(N) * HW4;
    for (long i = i0; i < quads; i += stride) {
      const long n = i / HW4, p4 = i % HW4;
      float4 f[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) f[c] = c < C ? __ldg(reinterpret_cast<const float4*>(x + (n * C + c) * HW) + p4) : make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t w[8];
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        __nv_bfloat162 lo = __floats2bfloat162_rn((&f[0].x)[px], (&f[1].x)[px]), hi = __floats2bfloat162_rn((&f[2].x)[px], (&f[3].x)[px]);
        w[2 * px] = *reinterpret_cast<uint32_t*>(&lo), w[2 * px + 1] = *reinterpret_cast<uint32_t*>(&hi);
      }
      uint4* o = reinterpret_cast<uint4*>(out + i * 16);
      o[0] = make_uint4(w[0], w[1], w[2], w[3]);
      o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
    return;
  }
  const long total = long(N) * HW;
  for (long i = i0; i < total; i += stride) {
    const long n = i / HW, p = i % HW;
    float f[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) f[c] = c < C ? __ldg(x + (n * C + c) * HW + p) : 0.f;
    __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(out + i * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// 8-bit images -> NHWC bf16, torchvision ToTensor's scaling fused in: value = float(byte) / 255 (one fp32 division, like
// ``img.float().div(255)``), then the same bf16 rounding the fp32 path applies.  The reference decodes 8-bit RGB with PIL and
// converts on the CPU (datasets/datasets_base.py:71-72) before the H2D copy of fp32 tensors (trainer/trainer.py:109): moving
// the bytes instead is 4x less PCIe traffic.  HWC = 1: x is [N,H,W,C] (what PIL / OpenCV hold); 0: [N,C,H,W].
// One thread per pixel; the warp's reads are contiguous in either layout.
template <int CPAD, bool HWC>
__global__ void u8_to_nhwc_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C, long HW) {
  unpp::pdl_wait();
  unpp::pdl_trigger();
  const long stride = long(gridDim.x) * blockDim.x, total = long(N) * HW;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const long n = i / HW, p = i % HW;
    __align__(16) __nv_bfloat16 v[CPAD];
#pragma unroll
    for (int c = 0; c < CPAD; ++c) {
      float f = 0.f;
      if (c < C) f = float(HWC ? __ldg(x + i * C + c) : __ldg(x + (n * C + c) * HW + p)) / 255.f;
      v[c] = __float2bfloat16_rn(f);
    }
    if constexpr (CPAD == 4) {
      *reinterpret_cast<uint2*>(out + i * 4) = *reinterpret_cast<const uint2*>(v);
    } else {
      uint4* o = reinterpret_cast<uint4*>(out + i * CPAD);
#pragma unroll
      for (int k = 0; k < CPAD / 8; ++k) o[k] = reinterpret_cast<const uint4*>(v)[k];
    }
  }
}

// ------------------------------------------------------------------------------------------
// 2x2 max pool on NHWC bf16, 8 channels (16 B) per thread.
__device__ __forceinline__ uint4 max8(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162* pa = reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162* pb = reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}
__global__ void maxpool2x2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int N, int H, int W, int C8) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const int Ho = H / 2, Wo = W / 2;
  const long total = long(N) * Ho * Wo * C8;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    long t = i;
    const int c = int(t % C8);
    t /= C8;
    const int xo = int(t % Wo);
    t /= Wo;
    const int yo = int(t % Ho);
    const long n = t / Ho;
    const uint4* r0 = x + ((n * H + 2 * yo) * W + 2 * xo) * C8 + c;
    const uint4* r1 = r0 + long(W) * C8;
    out[i] = max8(max8(__ldg(r0), __ldg(r0 + C8)), max8(__ldg(r1), __ldg(r1 + C8)));
  }
}

// ------------------------------------------------------------------------------------------
// Per-plane arg-max with "first maximum in row-major order" tie-breaking
// (np.where(h == h.max()) -> [0] of tools/misc/heatmap.py:173-176).  One CTA per plane; each
// thread scans a strided set of float4s keeping (value, smallest index); warp-shuffle then
// cross-warp reduction with the same total order: larger value wins, equal value -> smaller
// index wins.  NaNs never win (comparison false), matching a max over finite sigmoid outputs.
__device__ __forceinline__ void take(float& bv, int& bi, float v, int i) {
  if (v > bv || (v == bv && i < bi)) bv = v, bi = i;
}
__global__ void __launch_bounds__(256) argmax_kernel(const float* __restrict__ heat, int HW, int W, int32_t* __restrict__ xy,
                                                     float* __restrict__ val) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  const float* h = heat + size_t(blockIdx.x) * HW;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0) {
    const float4* h4 = reinterpret_cast<const float4*>(h);
    for (int i = threadIdx.x; i < HW / 4; i += blockDim.x) {
      const float4 v = __ldg(h4 + i);
      take(bv, bi, v.x, 4 * i), take(bv, bi, v.y, 4 * i + 1), take(bv, bi, v.z, 4 * i + 2), take(bv, bi, v.w, 4 * i + 3);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) take(bv, bi, __ldg(h + i), i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    take(bv, bi, ov, oi);
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sv[warp] = bv, si[warp] = bi;
  __syncthreads();
  if (warp == 0) {
    bv = lane < (blockDim.x >> 5) ? sv[lane] : -INFINITY;
    bi = lane < (blockDim.x >> 5) ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      take(bv, bi, ov, oi);
    }
    if (lane == 0) {
      if (bi == 0x7fffffff) bi = 0;  // all-NaN / -inf plane: np.argmax would also report 0
      xy[2 * blockIdx.x] = bi % W;      // x first (heatmap.py:178)
      xy[2 * blockIdx.x + 1] = bi / W;  // then y
      if (val) val[blockIdx.x] = bv;
    }
  }
}

// The same arg-max for few, large planes (1024x1024 at batch 16 = 64 planes of 4 MB: one CTA per plane leaves 84 SMs idle and
// streams at 0.8 TB/s): every plane is cut into `splits` contiguous segments, one CTA each, writing its (value, index) pair;
// a second launch folds the pairs of a plane.  (value desc, index asc) is a strict total order, so the result does not depend
// on the split count or on the folding order: bit-identical to the one-CTA kernel.
__global__ void __launch_bounds__(256) argmax_split_kernel(const float* __restrict__ heat, int HW, int seg, float* __restrict__ pval, int* __restrict__ pidx) {
  unpp::pdl_wait();
  unpp::pdl_trigger();
  const int plane = blockIdx.y, s = blockIdx.x;
  const float* h = heat + size_t(plane) * HW;
  const int b = s * seg, e = min(b + seg, HW);  // seg % 4 == 0
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0) {
    const float4* h4 = reinterpret_cast<const float4*>(h);
    int i = b / 4 + threadIdx.x;
    const int e4 = e / 4;
    for (; i + 3 * 256 < e4; i += 4 * 256) {  // four 16-byte loads in flight per thread
      const float4 v0 = __ldg(h4 + i), v1 = __ldg(h4 + i + 256), v2 = __ldg(h4 + i + 512), v3 = __ldg(h4 + i + 768);
      take(bv, bi, v0.x, 4 * i), take(bv, bi, v0.y, 4 * i + 1), take(bv, bi, v0.z, 4 * i + 2), take(bv, bi, v0.w, 4 * i + 3);
      take(bv, bi, v1.x, 4 * i + 1024), take(bv, bi, v1.y, 4 * i + 1025), take(bv, bi, v1.z, 4 * i + 1026), take(bv, bi, v1.w, 4 * i + 1027);
      take(bv, bi, v2.x, 4 * i + 2048), take(bv, bi, v2.y, 4 * i + 2049), take(bv, bi, v2.z, 4 * i + 2050), take(bv, bi, v2.w, 4 * i + 2051);
      take(bv, bi, v3.x, 4 * i + 3072), take(bv, bi, v3.y, 4 * i + 3073), take(bv, bi, v3.z, 4 * i + 3074), take(bv, bi, v3.w, 4 * i + 3075);
    }
    for (; i < e4; i += 256) {
      const float4 v = __ldg(h4 + i);
      take(bv, bi, v.x, 4 * i), take(bv, bi, v.y, 4 * i + 1), take(bv, bi, v.z, 4 * i + 2), take(bv, bi, v.w, 4 * i + 3);
    }
  } else {
    for (int i = b + threadIdx.x; i < e; i += 256) take(bv, bi, __ldg(h + i), i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    take(bv, bi, ov, oi);
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sv[warp] = bv, si[warp] = bi;
  __syncthreads();
  if (warp == 0) {
    bv = lane < 8 ? sv[lane] : -INFINITY;
    bi = lane < 8 ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      take(bv, bi, ov, oi);
    }
    if (lane == 0) pval[size_t(plane) * gridDim.x + s] = bv, pidx[size_t(plane) * gridDim.x + s] = bi;
  }
}

__global__ void __launch_bounds__(128) argmax_fold_kernel(const float* __restrict__ pval, const int* __restrict__ pidx, int planes, int splits, int W,
                                                          int32_t* __restrict__ xy, float* __restrict__ val) {
  unpp::pdl_wait();
  unpp::pdl_trigger();
  const int plane = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;  // one warp per plane
  if (plane >= planes) return;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int s = lane; s < splits; s += 32) take(bv, bi, pval[size_t(plane) * splits + s], pidx[size_t(plane) * splits + s]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    take(bv, bi, ov, oi);
  }
  if (lane == 0) {
    if (bi == 0x7fffffff) bi = 0;
    xy[2 * plane] = bi % W, xy[2 * plane + 1] = bi / W;
    if (val) val[plane] = bv;
  }
}

// ------------------------------------------------------------------------------------------
// Top-`num` peaks of a heat-map plane: the multi-point form of Heatmap.extract_points_ (tools/misc/heatmap.py:148-208), which
// thresholds the plane at 0.5, splits it into regions (OpenCV watershed), takes every region's maximum (first index), sorts the
// regions by that maximum, brightest first, and keeps `num` of them — retrying once at 0.9 x threshold when nothing is found.
// Here a region's representative is a strict local maximum of the thresholded plane under the total order (value desc, index
// asc) over its 8-neighbourhood: for separated blobs (what the targets of helper.create_heatmap look like) this is the same
// point set in the same order; touching blobs that the watershed would cut and a plain maximum search would merge keep one
// peak each as long as each has its own local maximum.  One CTA per plane, `num` selection rounds over the (L2-resident) plane:
// round r takes the best peak that comes after the (r-1)-th in the order, so nothing is stored and nothing can overflow.
__device__ __forceinline__ bool peak_at(const float* __restrict__ h, int H, int W, int p, float v) {
  const int y = p / W, x = p - y * W;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      if (!dy && !dx) continue;
      const int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const int q = yy * W + xx;
      const float u = __ldg(h + q);
      if (u > v || (u == v && q < p)) return false;
    }
  }
  return true;
}
__global__ void __launch_bounds__(256) topk_peaks_kernel(const float* __restrict__ heat, int H, int W, int num, float threshold, int32_t* __restrict__ xy,
                                                         float* __restrict__ val, int32_t* __restrict__ count) {
  unpp::pdl_wait();
  unpp::pdl_trigger();
  const int HW = H * W;
  const float* h = heat + size_t(blockIdx.x) * HW;
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ float s_lastv;
  __shared__ int s_lasti;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int found = 0;
  for (int attempt = 0; attempt < 2 && found == 0; ++attempt) {
    const float thr = attempt ? threshold * 0.9f : threshold;  // heatmap.py:187-190: one retry on a lower threshold
    float lastv = INFINITY;
    int lasti = -1;
    for (int r = 0; r < num; ++r) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int p = threadIdx.x; p < HW; p += blockDim.x) {
        const float v = __ldg(h + p);
        if (!(v >= thr)) continue;                                     // heatm[heatm < threshold] = 0 (heatmap.py:158)
        if (!(v < lastv || (v == lastv && p > lasti))) continue;      // strictly after the previous pick in (value desc, index asc)
        if (!(v > bv || (v == bv && p < bi))) continue;               // not better than what this thread already holds
        if (peak_at(h, H, W, p, v)) bv = v, bi = p;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        take(bv, bi, ov, oi);
      }
      if (lane == 0) sv[warp] = bv, si[warp] = bi;
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) take(bv, bi, sv[k], si[k]);
        s_lastv = bv, s_lasti = bi;
      }
      __syncthreads();
      lastv = s_lastv, lasti = s_lasti;
      __syncthreads();
      if (lasti == 0x7fffffff) break;  // no further peak
      if (threadIdx.x == 0) {
        xy[(size_t(blockIdx.x) * num + r) * 2] = lasti % W;      // x first (heatmap.py:178)
        xy[(size_t(blockIdx.x) * num + r) * 2 + 1] = lasti / W;
        val[size_t(blockIdx.x) * num + r] = lastv;
      }
      ++found;
    }
  }
  if (threadIdx.x == 0) {
    count[blockIdx.x] = found;
    for (int r = found; r < num; ++r) xy[(size_t(blockIdx.x) * num + r) * 2] = xy[(size_t(blockIdx.x) * num + r) * 2 + 1] = -1, val[size_t(blockIdx.x) * num + r] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// Target heat-map synthesis of the reference trainer (tools/misc/helper.py:87-172, called on the CPU every
// step at trainer/trainer.py:122-123): 7 key points -> 4 planes, point groups {0}, {1,2,3}, {4}, {5..};
// per point exp(-0.5 * dist / 3) with the Euclidean DISTANCE (not squared) in float64; planes 0 and 2 are
// assigned (helper.py:106,142), planes 1 and 3 are summed (float32 += float64) and ALWAYS divided by their
// maximum (helper.py:122-123,158-159) — also when plane 3 holds a single point (6 key points).
// A thread-block CLUSTER of 8 CTAs per (n, channel) plane (1024 CTAs at batch 32: the fp64 sqrt / exp chains need ~50 warps per SM
// in flight; one CTA per plane ran at 8): every CTA owns one eighth of the pixels, pass 1 writes the sums and reduces the CTA's
// maximum, the eight maxima are exchanged through distributed shared memory, pass 2 normalises the CTA's own pixels.
constexpr int kHmCluster = 8;
__global__ void __cluster_dims__(kHmCluster, 1, 1) __launch_bounds__(256) create_heatmap_kernel(const float* __restrict__ kp, int npts, int H, int W, float* __restrict__ out) {
  unpp::pdl_wait();  // see common.h: everything below may read what earlier kernels of the stream wrote
  unpp::pdl_trigger();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int plane_idx = blockIdx.y, rank = blockIdx.x;  // gridDim.x == cluster size: blockIdx.x is the rank inside the cluster
  const int n = plane_idx >> 2, ch = plane_idx & 3;
  const int p0 = ch == 0 ? 0 : ch == 1 ? 1 : ch == 2 ? 4 : 5, p1 = ch == 0 ? 1 : ch == 1 ? 4 : ch == 2 ? 5 : npts;
  float* plane = out + size_t(plane_idx) * H * W;
  double cx[4], cy[4];
  const int np = p1 - p0;
  for (int i = 0; i < 4; ++i) {
    cx[i] = i < np ? double(kp[(size_t(n) * npts + p0 + i) * 2]) : 0.0;
    cy[i] = i < np ? double(kp[(size_t(n) * npts + p0 + i) * 2 + 1]) : 0.0;
  }
  const int HW = H * W, seg = (HW + kHmCluster - 1) / kHmCluster, b = rank * seg, e = min(b + seg, HW);
  const bool summed = ch & 1;  // planes 1 and 3: "+=" then "/ max"
  float mx = 0.f;
  for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
    const double x = double(i % W), y = double(i / W);
    float acc = 0.f;
    for (int k = 0; k < np && k < 4; ++k) {
      const double d = sqrt((x - cx[k]) * (x - cx[k]) + (y - cy[k]) * (y - cy[k]));
      const double g = exp(-0.5 * d / 3.0);
      acc = summed ? float(double(acc) + g) : float(g);
    }
    plane[i] = acc;
    mx = fmaxf(mx, acc);
  }
  __shared__ float smax[8];
  __shared__ float cta_max;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    mx = smax[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) mx = fmaxf(mx, smax[k]);
    cta_max = mx;
  }
  cluster.sync();  // every CTA's maximum is published
  mx = 0.f;
  for (int r = 0; r < kHmCluster; ++r) mx = fmaxf(mx, *cluster.map_shared_rank(&cta_max, r));
  cluster.sync();  // nobody leaves (and frees its shared memory) while a peer still reads it
  if (!summed) return;
  for (int i = b + threadIdx.x; i < e; i += blockDim.x) plane[i] = plane[i] / mx;  // each thread re-reads only what it wrote
}

// ------------------------------------------------------------------------------------------
// conv3x3(ConvTranspose2d_k2s2(x)) as ONE 3x3 conv over the low-resolution x (models/unet.py:187,199-201: the k2s2 upsample
// never overlaps).  comp[(2*jy+jx)*Co + co][ci][ty][tx] = sum over the conv taps (r, s) whose upsampled position
// (jy + r - 1, jx + s - 1) lies in low-resolution cell (ty - 1, tx - 1), and over u, of Wc[co][u][r][s] * Wd[ci][u][p][q] with
// (p, q) the parity of that position; table[3*rc + cc][co] = b_conv[co] + the upsample bias through the taps that stay inside
// the image for a first / interior / last row (rc) and column (cc): zero padding is applied AFTER the upsample.
// One thread per output element, fp32 accumulation of at most 4 * Cu products in a fixed order.
__global__ void compose_deconv_conv_kernel(const float* __restrict__ wc, int Ctot, int Cu, const float* __restrict__ wd, const float* __restrict__ b_up,
                                           const float* __restrict__ b_conv, int Co, int Ci, float* __restrict__ comp, float* __restrict__ table) {
  unpp::pdl_wait();
  unpp::pdl_trigger();
  const int ncomp = 4 * Co * Ci * 9, total = ncomp + 9 * Co;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < ncomp) {
      const int tx = i % 3, ty = (i / 3) % 3, ci = (i / 9) % Ci, row = i / (9 * Ci);
      const int co = row % Co, jx = (row / Co) & 1, jy = (row / Co) >> 1;
      float acc = 0.f;
      for (int r = 0; r < 3; ++r) {
        const int uy = jy + r - 1, dyl = uy < 0 ? -1 : uy >> 1, pp = uy - 2 * dyl;
        if (dyl + 1 != ty) continue;
        for (int sft = 0; sft < 3; ++sft) {
          const int ux = jx + sft - 1, dxl = ux < 0 ? -1 : ux >> 1, qq = ux - 2 * dxl;
          if (dxl + 1 != tx) continue;
          for (int u = 0; u < Cu; ++u) acc = fmaf(__ldg(wc + ((size_t(co) * Ctot + u) * 3 + r) * 3 + sft), __ldg(wd + ((size_t(ci) * Cu + u) * 2 + pp) * 2 + qq), acc);
        }
      }
      comp[i] = acc;
    } else {
      const int j = i - ncomp, co = j % Co, cls = j / Co, rc = cls / 3, cc = cls % 3;
      float acc = __ldg(b_conv + co);
      for (int r = (rc == 0 ? 1 : 0); r < (rc == 2 ? 2 : 3); ++r)
        for (int sft = (cc == 0 ? 1 : 0); sft < (cc == 2 ? 2 : 3); ++sft)
          for (int u = 0; u < Cu; ++u) acc = fmaf(__ldg(wc + ((size_t(co) * Ctot + u) * 3 + r) * 3 + sft), __ldg(b_up + u), acc);
      table[j] = acc;
    }
  }
}

// Eval-mode BatchNorm2d folded into the preceding conv (models/unet.py:132-133): y = (conv + b - running_mean) * gamma / sqrt(running_var + eps) + beta
//   scale[c] = gamma[c] / sqrt(running_var[c] + eps)  (multiplied into the packed weights),  bias[c] = (b[c] - running_mean[c]) * scale[c] + beta[c]
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rmean, const float* __restrict__ rvar,
                               const float* __restrict__ cbias, float eps, int C, float* __restrict__ scale, float* __restrict__ bias) {
  unpp::pdl_wait();
  unpp::pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rvar[c] + eps);
  scale[c] = sc;
  bias[c] = (cbias[c] - rmean[c]) * sc + beta[c];
}

inline int grid_for(long total, int block) {
  long g = (total + block - 1) / block;
  const long cap = long(unpp::num_sms()) * 16;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int unpp_pack_weights(const UnppPackArgs* a, unpp_stream_t stream) {
  if (!a || !a->src || !a->dst) return unpp::fail(UNPP_ERR_BAD_ARG, "pack_weights: null pointer");
  if (a->kind < 0 || a->kind > 7) return unpp::fail(UNPP_ERR_BAD_ARG, "pack_weights: bad kind");
  if (a->kind == 7 && (a->taps != 4 || a->n_total != 64 || a->n_tile != 64 || a->k_count != 32 || a->k8_total != 4 || a->k_dst8 || a->src_O != 16 || a->src_I > 4))
    return unpp::fail(UNPP_ERR_BAD_ARG, "pack_weights: kind 7 (first layer) needs taps=4, n_total=n_tile=64, k_count=32 and a [16][<=4][3][3] source");
  if (a->kind >= 4 && a->kind < 7 && (a->taps != (a->kind == 6 ? 9 : 16) || a->n_total != 64 || a->n_tile != 64))
    return unpp::fail(UNPP_ERR_BAD_ARG, "pack_weights: 2x2-blocked kinds need taps=16 (kind 6: 9), n_total=n_tile=64");
  if (a->n_tile < 8 || a->n_total % a->n_tile || a->k_count % 8 || a->k_count < 8 || a->taps < 1)
    return unpp::fail(UNPP_ERR_BAD_ARG, "pack_weights: n_total %% n_tile, k_count %% 8 must be 0");
  if (a->k_dst8 + a->k_count / 8 > a->k8_total) return unpp::fail(UNPP_ERR_BAD_ARG, "pack_weights: K range exceeds k8_total");
  const long total = a->kind == 7 ? 1024 : a->kind >= 4 ? long(a->kind == 6 ? b2::kLowUnits : b2::kMainUnits) * 16 * (a->k_count / 8) : long(a->n_total) * a->taps * (a->k_count / 8);
  unpp::launch(pack_weights_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), *a);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("pack_weights: launch");
  return UNPP_OK;
}

extern "C" int unpp_pack_weights_batched(const UnppPackArgs* table_dev, int n, unpp_stream_t stream) {
  if (!table_dev || n < 1 || n > 65535) return unpp::fail(UNPP_ERR_BAD_ARG, "pack_weights_batched: bad argument");
  unpp::launch(pack_weights_batched_kernel, dim3(8, n), 256, 0, reinterpret_cast<cudaStream_t>(stream), table_dev);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("pack_weights_batched: launch");
  return UNPP_OK;
}

extern "C" int unpp_nchw_to_nhwc(const float* x, void* out, int N, int C, int H, int W, int Cpad, unpp_stream_t stream) {
  if (!x || !out || N < 1 || C < 1 || H < 1 || W < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "nchw_to_nhwc: bad argument");
  if ((Cpad != 16 && Cpad != 4) || C > Cpad) return unpp::fail(UNPP_ERR_UNSUPPORTED, "nchw_to_nhwc: Cpad must be 16 or 4 and C <= Cpad");
  const long total = (long(H) * W) % 4 ? long(N) * H * W : long(N) * H * W / 4;
  if (Cpad == 4) {
    unpp::launch(nchw_to_nhwc4_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, reinterpret_cast<__nv_bfloat16*>(out), N, C, long(H) * W);
    if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("nchw_to_nhwc: launch");
    return UNPP_OK;
  }
  unpp::launch(nchw_to_nhwc_kernel<16>, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      x, reinterpret_cast<__nv_bfloat16*>(out), N, C, long(H) * W);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("nchw_to_nhwc: launch");
  return UNPP_OK;
}

extern "C" int unpp_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var, const float* conv_bias, float eps, int C,
                            float* scale, float* bias, unpp_stream_t stream) {
  if (!gamma || !beta || !running_mean || !running_var || !conv_bias || !scale || !bias || C < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "bn_fold: bad argument");
  unpp::launch(bn_fold_kernel, (C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream), gamma, beta, running_mean, running_var, conv_bias, eps, C, scale, bias);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("bn_fold: launch");
  return UNPP_OK;
}

extern "C" int unpp_compose_deconv_conv(const float* w_conv, int Ctot, int Cu, const float* w_up, const float* b_up, const float* b_conv, int Co, int Ci,
                                        float* comp, float* table, unpp_stream_t stream) {
  if (!w_conv || !w_up || !b_up || !b_conv || !comp || !table || Co < 1 || Ci < 1 || Cu < 1 || Cu > Ctot)
    return unpp::fail(UNPP_ERR_BAD_ARG, "compose_deconv_conv: bad argument");
  const long total = 4L * Co * Ci * 9 + 9L * Co;
  unpp::launch(compose_deconv_conv_kernel, grid_for(total, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream), w_conv, Ctot, Cu, w_up, b_up, b_conv, Co, Ci, comp, table);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("compose_deconv_conv: launch");
  return UNPP_OK;
}

extern "C" int unpp_u8_to_nhwc(const uint8_t* x, void* out, int N, int C, int H, int W, int Cpad, int hwc, unpp_stream_t stream) {
  if (!x || !out || N < 1 || C < 1 || H < 1 || W < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "u8_to_nhwc: bad argument");
  if ((Cpad != 16 && Cpad != 4) || C > Cpad) return unpp::fail(UNPP_ERR_UNSUPPORTED, "u8_to_nhwc: Cpad must be 16 or 4 and C <= Cpad");
  const long total = long(N) * H * W;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (Cpad == 4 && hwc) unpp::launch(u8_to_nhwc_kernel<4, true>, grid_for(total, 256), 256, 0, st, x, o, N, C, long(H) * W);
  else if (Cpad == 4) unpp::launch(u8_to_nhwc_kernel<4, false>, grid_for(total, 256), 256, 0, st, x, o, N, C, long(H) * W);
  else if (hwc) unpp::launch(u8_to_nhwc_kernel<16, true>, grid_for(total, 256), 256, 0, st, x, o, N, C, long(H) * W);
  else unpp::launch(u8_to_nhwc_kernel<16, false>, grid_for(total, 256), 256, 0, st, x, o, N, C, long(H) * W);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("u8_to_nhwc: launch");
  return UNPP_OK;
}

extern "C" int unpp_maxpool2x2(const void* x, void* out, int N, int H, int W, int C, unpp_stream_t stream) {
  if (!x || !out || N < 1 || H < 2 || W < 2 || (H & 1) || (W & 1) || C % 8) return unpp::fail(UNPP_ERR_BAD_ARG, "maxpool2x2: need even H, W and C %% 8 == 0");
  const long total = long(N) * (H / 2) * (W / 2) * (C / 8);
  unpp::launch(maxpool2x2_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), N, H, W, C / 8);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("maxpool2x2: launch");
  return UNPP_OK;
}

extern "C" int unpp_create_heatmap(const float* keypoints, int N, int npts, int H, int W, float* out, unpp_stream_t stream) {
  if (!keypoints || !out || N < 1 || H < 1 || W < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "create_heatmap: bad argument");
  if (npts < 6 || npts > 9) return unpp::fail(UNPP_ERR_UNSUPPORTED, "create_heatmap: the reference's grouping needs 6..9 key points (7 in the trainer)");
  if (N * 4 > 65535) return unpp::fail(UNPP_ERR_UNSUPPORTED, "create_heatmap: at most 16383 images per call");
  unpp::launch(create_heatmap_kernel, dim3(kHmCluster, N * 4), 256, 0, reinterpret_cast<cudaStream_t>(stream), keypoints, npts, H, W, out);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("create_heatmap: launch");
  return UNPP_OK;
}

extern "C" int unpp_topk_peaks(const float* heat, int planes, int H, int W, int num, float threshold, int32_t* xy, float* val, int32_t* count, unpp_stream_t stream) {
  if (!heat || !xy || !val || !count || planes < 0 || H < 1 || W < 1 || num < 1 || num > 64) return unpp::fail(UNPP_ERR_BAD_ARG, "topk_peaks: bad argument (1 <= num <= 64)");
  if (long(H) * W > 0x7ffffff0L) return unpp::fail(UNPP_ERR_UNSUPPORTED, "topk_peaks: plane too large");
  if (planes == 0) return UNPP_OK;
  unpp::launch(topk_peaks_kernel, planes, 256, 0, reinterpret_cast<cudaStream_t>(stream), heat, H, W, num, threshold, xy, val, count);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("topk_peaks: launch");
  return UNPP_OK;
}

extern "C" int unpp_argmax_splits(int planes, int H, int W) {
  // enough CTAs for ~4 per SM, segments of at least 16 K values; 1 = the one-CTA-per-plane kernel is the better choice
  if (planes < 1 || H < 1 || W < 1) return 1;
  const long HW = long(H) * W;
  long want = (4L * unpp::num_sms() + planes - 1) / planes, cap = HW / 16384;
  if (want > cap) want = cap;
  if (want > 256) want = 256;
  return want < 2 ? 1 : int(want);
}

extern "C" int unpp_argmax_peaks_split(const float* heat, int planes, int H, int W, int32_t* xy, float* val, void* workspace, int splits, unpp_stream_t stream) {
  if (!heat || !xy || planes < 0 || H < 1 || W < 1 || splits < 1 || splits > 1024) return unpp::fail(UNPP_ERR_BAD_ARG, "argmax_peaks_split: bad argument");
  if (long(H) * W > 0x7ffffff0L || planes > 65535) return unpp::fail(UNPP_ERR_UNSUPPORTED, "argmax_peaks_split: plane too large / too many planes");
  if (planes == 0) return UNPP_OK;
  if (splits == 1) return unpp_argmax_peaks(heat, planes, H, W, xy, val, stream);
  if (!workspace) return unpp::fail(UNPP_ERR_BAD_ARG, "argmax_peaks_split: workspace of planes * splits * 8 bytes needed");
  const int HW = H * W;
  const int seg = ((HW + splits - 1) / splits + 3) & ~3;
  float* pval = static_cast<float*>(workspace);
  int* pidx = reinterpret_cast<int*>(pval + size_t(planes) * splits);
  unpp::launch(argmax_split_kernel, dim3(splits, planes), 256, 0, reinterpret_cast<cudaStream_t>(stream), heat, HW, seg, pval, pidx);
  unpp::launch(argmax_fold_kernel, (planes + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream), static_cast<const float*>(pval), static_cast<const int*>(pidx), planes,
               splits, W, xy, val);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("argmax_peaks_split: launch");
  return UNPP_OK;
}

extern "C" int unpp_argmax_peaks(const float* heat, int planes, int H, int W, int32_t* xy, float* val, unpp_stream_t stream) {
  if (!heat || !xy || planes < 0 || H < 1 || W < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "argmax_peaks: bad argument");
  if (long(H) * W > 0x7ffffff0L) return unpp::fail(UNPP_ERR_UNSUPPORTED, "argmax_peaks: plane too large");
  if (planes == 0) return UNPP_OK;
  unpp::launch(argmax_kernel, planes, 256, 0, reinterpret_cast<cudaStream_t>(stream), heat, H * W, W, xy, val);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("argmax_peaks: launch");
  return UNPP_OK;
}

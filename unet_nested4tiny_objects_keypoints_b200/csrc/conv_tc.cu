// conv_tc.cu — fused implicit-GEMM convolution for UNet++ on Blackwell tensor cores.
//
// One kernel serves every GEMM-shaped op of the network (3x3 conv of a virtual concat, the k2s2
// transposed conv as a pointwise GEMM with a scatter epilogue, and all dgrads with flipped packed
// weights):
//
//   * activations are NHWC bf16.  A persistent CTA walks 16-row x TW-column output tiles; for each
//     tile and each <=64-channel chunk of each source tensor ONE TMA box load brings the
//     (16+2) x (TW+2) halo tile into shared memory as pixel-major rows with the hardware swizzle
//     that matches the chunk width (32/64/128 B).  The concat is never materialised: the K loop
//     just walks the chunk table.
//   * the nine filter taps are NOT nine copies: each tap is the same smem tile seen through a
//     matrix descriptor whose start address is shifted by (r*P + s) pixels (P = TW+2) and whose
//     stride-byte-offset is one tile row, so a 128-row MMA covers a 16x8 pixel patch
//     (probe/umma_probe.cu m6/m7/m8 verify this addressing on B200).
//   * tcgen05.mma (M=128, N=n_tile, K=16, bf16 -> fp32) accumulates in TMEM; two accumulator sets
//     let the epilogue of tile i overlap the MMAs of tile i+1.
//   * epilogue warps read TMEM (tcgen05.ld 32x32b), apply bias / ReLU / ReLU-mask / addend, the
//     fused 1x1 head + sigmoid, and store NHWC bf16 (and NCHW fp32 heatmaps).
//
// Warp roles: warp 0 = TMA producer, warps 1,10.. = MMA issuers (warp 1 also allocates TMEM), warps 2..9 = epilogue
// (two per TMEM lane quadrant).  The kernel is compiled per epilogue variant (conv / deconv scatter,
// fused head, training extras) so that the inference epilogue carries no run-time feature tests.
#include <cstdlib>
#include "sm100.cuh"
#include "common.h"
#include "../../include/unpp.h"
#include "b2_blocks.h"

using namespace sm100;

namespace {

constexpr int kMaxChunks = 8;
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 8;
// Issuing threads: each owns the sub-tile accumulators j = i, i+n, i+2n (disjoint TMEM columns).  The inference
// variants (<= 108 registers) run four of them (416 threads); the training variants need up to 168 registers
// and run three (384 threads).
// (A fourth issuer for the training variants does not pay: registers are allocated for a multiple of four warps, so a 13-warp
// CTA is limited to 128 registers per thread whatever __maxnreg__ says — 152 or 144 fail to launch — and at 128 the training
// epilogue spills ~600 bytes per thread: 3.51 -> 3.95 ms per step.)
// The head conv of the training forward (HEAD && TRAIN: dropout keep bits, nothing else) is as light as the inference head: four issuers.
constexpr bool heavy_epilogue(bool head, bool train) { return train && !head; }
constexpr int mma_warps(bool heavy) { return heavy ? 3 : 4; }
constexpr int block_threads(bool heavy) { return 64 + 32 * kEpiWarps + 32 * (mma_warps(heavy) - 1); }

struct ConvTcParams {
  CUtensorMap maps[UNPP_MAX_SRC];
  int nchunk;
  int ch_map[kMaxChunks];   // which tensor map
  int ch_c0[kMaxChunks];    // first channel inside that source
  int ch_span[kMaxChunks];  // bytes per pixel row of the chunk: 32 / 64 / 128
  int ch_wk8[kMaxChunks];   // offset of the chunk in the packed-weight K axis, in 8-channel units
  int N, H, W;
  int TW, nsub, tiles_x, tiles_y, ntiles;
  int taps, ncols, k8_total;
  int stage_bytes, nstage, w_bytes, w_smem_bytes;
  int tmem_cols;
  const __nv_bfloat16* wpacked;
  // epilogue
  int mode, relu, cout;  // cout = channels of `out` (n_total for conv, n_total/4 for deconv)
  const float* bias;
  __nv_bfloat16* out;
  const float* head_w;
  const float* head_b;
  float* heat;
  float* logit;
  int head_classes;
  const uint16_t* drop_mask;
  float drop_scale;
  const __nv_bfloat16* addend;
  const __nv_bfloat16* relu_mask_src;
  float* stats_partial;
  const __nv_bfloat16* stats_aux;
  const float* aux_mean;
  const float* aux_istd;
  int nacc;      // accumulator sets in TMEM (2 or 4): the epilogue of tile i overlaps the MMAs of tiles i+1 .. i+nacc-1
  int b2, b2_P;  // 2x2 output blocking: flag, pixel PAIRS per staged tile row (TW/2 + 2)
  int c4;        // with b2: first-layer mode, the one source has 4 channels (8 B pixels, 16 B pair rows, no swizzle); b2_P = TW/2 + 3
  // fused transposed conv (with b2): one extra K chunk per tile read from the low-resolution tensor
  CUtensorMap lowmap;
  int low_on, low_P, low_k8, low_w_off, low_w_bytes;  // low_P = low-res pixels per staged tile row (TW/2 + 2); offsets in bytes
  const __nv_bfloat16* low_wpacked;
  int bias9;     // bias is a [9][16] (row class, column class) table
  __nv_bfloat16* pooled;  // fused 2x2 max pool of the output (inference epilogue), or null
  int pf;   // L2 prefetch distance of the TMA producer in (tile, chunk) steps beyond the stage ring; 0 = off
  int dbg;  // UNPP_DBG experiment bits (0 in production): 1 = epilogue skips its work, 2 = no MMAs issued, 4 = no TMA tile loads,
            // 8 = epilogue does not store its bf16 output, 16 = epilogue does not read TMEM
};

__device__ __forceinline__ uint32_t layout_type_of_span(int span) { return span == 128 ? 2u : span == 64 ? 4u : 6u; }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// Sum over the 32 lanes of v[c] for c = lane >> 1, by recursive halving (16 shuffles instead of
// 80): after the four exchange steps lane l holds channel (l >> 1) summed over 16 lanes; the last
// step folds the two lanes that share a channel.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2], a1;
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float mine = h4 ? v[8 + i] : v[i], other = h4 ? v[i] : v[8 + i];
    a8[i] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float mine = h3 ? a8[4 + i] : a8[i], other = h3 ? a8[i] : a8[4 + i];
    a4[i] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float mine = h2 ? a4[2 + i] : a4[i], other = h2 ? a4[i] : a4[2 + i];
    a2[i] = mine + __shfl_xor_sync(0xffffffffu, other, 4);
  }
  {
    const float mine = h1 ? a2[1] : a2[0], other = h1 ? a2[0] : a2[1];
    a1 = mine + __shfl_xor_sync(0xffffffffu, other, 2);
  }
  return a1 + __shfl_xor_sync(0xffffffffu, a1, 1);
}

// Epilogue of one 16-column group of one 8x16-pixel... rather 128-pixel sub-tile: v[16] are the fp32
// accumulators of this thread's pixel for GEMM columns [gcol, gcol+16).
// Register-resident copy of the epilogue parameters (read once from the parameter bank).
struct EpiArgs {
  int H, W, cout, head_classes, dbg, bias9;
  __nv_bfloat16* out;
  __nv_bfloat16* pooled;
  float* heat;
  float* logit;
  const uint16_t* drop_mask;
  float drop_scale;
  const __nv_bfloat16* addend;
  const __nv_bfloat16* relu_mask_src;
  float* stats_partial;
  const __nv_bfloat16* stats_aux;
  const float* aux_mean;
  const float* aux_istd;
};

// Training-only operands of one unit, fetched from global memory BEFORE the TMEM load is waited for.
struct TrainOperands {
  uint4 a0, a1, m0, m1, x0, x1;
};
template <bool TRAIN>
__device__ __forceinline__ void prefetch_train(const EpiArgs& p, TrainOperands& t, size_t pix, int gcol, bool valid) {
  if constexpr (TRAIN) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    t.a0 = t.a1 = t.m0 = t.m1 = t.x0 = t.x1 = z;
    if (valid) {
      const size_t off = pix * p.cout + gcol;
      if (p.addend) {
        const uint4* q = reinterpret_cast<const uint4*>(p.addend + off);
        t.a0 = __ldg(q), t.a1 = __ldg(q + 1);
      }
      if (p.relu_mask_src) {
        const uint4* q = reinterpret_cast<const uint4*>(p.relu_mask_src + off);
        t.m0 = __ldg(q), t.m1 = __ldg(q + 1);
      }
      if (p.stats_aux) {
        const uint4* q = reinterpret_cast<const uint4*>(p.stats_aux + off);
        t.x0 = __ldg(q), t.x1 = __ldg(q + 1);
      }
    }
  }
}

// Epilogue of one unit = one 16-column group of one 128-pixel sub-tile: v[16] are the fp32 accumulators
// of this thread's pixel for GEMM columns [gcol, gcol+16).  Statistics (training): sa1 += value,
// sa2 += value^2 or value*aux, per thread when reg_stats (reduced over the warp once per kernel),
// else reduced over the warp here into the warp's shared-memory slots.
template <bool DECONV, bool HEAD, bool TRAIN>
__device__ __forceinline__ void epilogue_group(const EpiArgs& p, float (&v)[16], const TrainOperands& t, int n, int y, int x, bool valid,
                                               int gcol, int c0, const float* s_bias, const float* s_head, float relu_floor, bool reg_stats,
                                               float (&sa1)[16], float (&sa2)[16], float* s_stats1, float* s_stats2, int lane) {
  {
    int boff = DECONV ? gcol % p.cout : c0;
    if (!DECONV && p.bias9) boff = (((y == 0) ? 0 : (y == p.H - 1 ? 2 : 1)) * 3 + ((x == 0) ? 0 : (x == p.W - 1 ? 2 : 1))) * 16;
    const float4* b4 = reinterpret_cast<const float4*>(s_bias + boff);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 b = b4[k];
      v[4 * k] += b.x, v[4 * k + 1] += b.y, v[4 * k + 2] += b.z, v[4 * k + 3] += b.w;
    }
  }
  if constexpr (DECONV) {
    if (valid) {
      const int pq = gcol / p.cout, co0 = gcol % p.cout;
      const size_t opix = (size_t(n) * (2 * p.H) + (2 * y + (pq >> 1))) * (2 * p.W) + (2 * x + (pq & 1));
      uint4 o0, o1;
      o0.x = pack_bf16x2(v[0], v[1]), o0.y = pack_bf16x2(v[2], v[3]), o0.z = pack_bf16x2(v[4], v[5]), o0.w = pack_bf16x2(v[6], v[7]);
      o1.x = pack_bf16x2(v[8], v[9]), o1.y = pack_bf16x2(v[10], v[11]), o1.z = pack_bf16x2(v[12], v[13]), o1.w = pack_bf16x2(v[14], v[15]);
      uint4* op = reinterpret_cast<uint4*>(p.out + opix * p.cout + co0);
      op[0] = o0;
      op[1] = o1;
    }
    return;
  }
  const size_t pix = (size_t(n) * p.H + y) * p.W + x;
  if constexpr (TRAIN) {
    if (p.addend) {
      const uint32_t aw[8] = {t.a0.x, t.a0.y, t.a0.z, t.a0.w, t.a1.x, t.a1.y, t.a1.z, t.a1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) v[2 * k] += bf16_lo(aw[k]), v[2 * k + 1] += bf16_hi(aw[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], relu_floor);
  if constexpr (TRAIN) {
    if (p.relu_mask_src) {
      const uint32_t mw[8] = {t.m0.x, t.m0.y, t.m0.z, t.m0.w, t.m1.x, t.m1.y, t.m1.z, t.m1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!(bf16_lo(mw[k]) > 0.f)) v[2 * k] = 0.f;
        if (!(bf16_hi(mw[k]) > 0.f)) v[2 * k + 1] = 0.f;
      }
    }
    if (p.stats_partial) {
      // statistics of the bf16-rounded value that is stored; invalid pixels contribute 0.  Second
      // statistic: v*v (BN batch variance) or v*aux (BN backward: sum dyh*z, turned into
      // sum dyh*xhat = istd*(sum dyh*z - mean*sum dyh) when the CTA writes its partials).
      float s1[16], s2[16];
      const uint32_t xw[8] = {t.x0.x, t.x0.y, t.x0.z, t.x0.w, t.x1.x, t.x1.y, t.x1.z, t.x1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t pk = valid ? pack_bf16x2(v[2 * k], v[2 * k + 1]) : 0u;  // exactly the two bf16 values that get stored
        const float r0 = bf16_lo(pk), r1 = bf16_hi(pk);
        s1[2 * k] = r0, s1[2 * k + 1] = r1;
        s2[2 * k] = r0 * (p.stats_aux ? bf16_lo(xw[k]) : r0), s2[2 * k + 1] = r1 * (p.stats_aux ? bf16_hi(xw[k]) : r1);
      }
      if (reg_stats) {
#pragma unroll
        for (int k = 0; k < 16; ++k) sa1[k] += s1[k], sa2[k] += s2[k];
      } else {
        const float r1 = warp_reduce16(s1, lane), r2 = warp_reduce16(s2, lane);
        if ((lane & 1) == 0) {  // lane holds channel (lane >> 1); each warp owns its own slots
          s_stats1[c0 + (lane >> 1)] += r1;
          s_stats2[c0 + (lane >> 1)] += r2;
        }
      }
    }
  }
  if (p.out && valid && !(p.dbg & 8)) {
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]), o0.y = pack_bf16x2(v[2], v[3]), o0.z = pack_bf16x2(v[4], v[5]), o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]), o1.y = pack_bf16x2(v[10], v[11]), o1.z = pack_bf16x2(v[12], v[13]), o1.w = pack_bf16x2(v[14], v[15]);
    uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.cout + gcol);
    op[0] = o0;
    op[1] = o1;
  }
  if constexpr (HEAD) {
    if (valid) {
      if constexpr (TRAIN) {
        if (p.drop_mask) {
          const uint32_t dm = __ldg(p.drop_mask + pix);  // 16 keep bits of this pixel
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = ((dm >> k) & 1u) ? v[k] * p.drop_scale : 0.f;
        }
      }
      const size_t plane = size_t(p.H) * p.W;
      size_t o = size_t(n) * p.head_classes * plane + size_t(y) * p.W + x;
      for (int cls = 0; cls < p.head_classes; ++cls, o += plane) {
        const float4* w4 = reinterpret_cast<const float4*>(s_head + cls * 16);
        float acc = s_head[8 * 16 + cls];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 w = w4[k];
          acc = fmaf(w.x, v[4 * k], acc), acc = fmaf(w.y, v[4 * k + 1], acc), acc = fmaf(w.z, v[4 * k + 2], acc), acc = fmaf(w.w, v[4 * k + 3], acc);
        }
        if constexpr (TRAIN) {
          if (p.logit) p.logit[o] = acc;
        }
        p.heat[o] = 1.f / (1.f + __expf(-acc));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Lean epilogue step: G groups of 16 accumulator columns whose outputs (and training operands) are 32*G contiguous bytes.
//   G = 2: the 2x2 output-blocked path.  A thread owns one 2x2 block (accumulator row); the two warps of a lane quadrant
//          take the upper / lower pixel row of the blocks, so a step = the two horizontally adjacent pixels (x0, x0 + 1).
//   G = 1: the classic path with <= 32 GEMM columns per CTA: a step = 16 channels of one pixel.
// `eoff` = element offset of the step's first value in the NHWC output (and in every same-shaped operand).
template <int G>
struct EpiOperands {  // training-only operands of the step, fetched BEFORE the TMEM load is waited for
  uint4 a[2 * G], m[2 * G], x[2 * G];
  uint32_t dbits;  // head dropout: 16 keep bits per pixel, the (up to two) pixels of the step in one word
};
template <int G, bool HEAD, bool TRAIN>
__device__ __forceinline__ void epi_prefetch(const EpiArgs& p, EpiOperands<G>& t, size_t eoff, bool valid) {
  if constexpr (TRAIN && HEAD) {
    // the head conv of the training forward: bias + ReLU + dropout + 1x1 head, none of the backward operands (the launch rejects them)
    t.dbits = 0xFFFFFFFFu;
    if (valid && p.drop_mask)
      t.dbits = G == 2 ? __ldg(reinterpret_cast<const uint32_t*>(p.drop_mask + (eoff >> 4))) : uint32_t(__ldg(p.drop_mask + (eoff >> 4)));
  } else if constexpr (TRAIN) {
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 2 * G; ++i) t.a[i] = t.m[i] = t.x[i] = z;
    t.dbits = 0xFFFFFFFFu;
    if (valid) {
      if (p.addend) {
        const uint4* q = reinterpret_cast<const uint4*>(p.addend + eoff);
#pragma unroll
        for (int i = 0; i < 2 * G; ++i) t.a[i] = __ldg(q + i);
      }
      if (p.relu_mask_src) {
        const uint4* q = reinterpret_cast<const uint4*>(p.relu_mask_src + eoff);
#pragma unroll
        for (int i = 0; i < 2 * G; ++i) t.m[i] = __ldg(q + i);
      }
      if (p.stats_aux) {
        const uint4* q = reinterpret_cast<const uint4*>(p.stats_aux + eoff);
#pragma unroll
        for (int i = 0; i < 2 * G; ++i) t.x[i] = __ldg(q + i);
      }
      if constexpr (HEAD) {
        if (p.drop_mask) {
          // (the head conv has 16 channels: eoff / 16 = pixel index; two adjacent pixels, x0 even: one aligned 32-bit word)
          t.dbits = G == 2 ? __ldg(reinterpret_cast<const uint32_t*>(p.drop_mask + (eoff >> 4))) : uint32_t(__ldg(p.drop_mask + (eoff >> 4)));
        }
      }
    }
  }
}

// element-wise max of two packed bf16x2 words (the 2x2 max pool works on the rounded values that are stored)
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// Ask L2 for the step's training operands (issued for the whole tile while the warp would otherwise idle on the accumulator
// barrier): the loads of epi_prefetch then find them in L2 instead of paying the DRAM latency once per step.
template <int G, bool HEAD, bool TRAIN>
__device__ __forceinline__ void epi_l2_prefetch(const EpiArgs& p, size_t eoff, bool valid) {
  if constexpr (TRAIN && !HEAD) {
    if (!valid) return;
    if (p.addend) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.addend + eoff));
    if (p.relu_mask_src) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.relu_mask_src + eoff));
    if (p.stats_aux) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.stats_aux + eoff));
  }
}

template <int G, bool HEAD, bool TRAIN>
__device__ __forceinline__ void epi_finish(const EpiArgs& p, const uint32_t (&raw)[16 * G], const EpiOperands<G>& t, const float (&bias_r)[16], const float* s_bias,
                                           const float* s_head, bool relu, size_t eoff, bool valid, int n, int yy, int x0, float (&sa1)[16],
                                           float (&sa2)[16], uint32_t (&words)[8 * G]) {
  float hv[HEAD ? G : 1][16];  // post-activation values of the pixels (fused head only)
#pragma unroll
  for (int px = 0; px < G; ++px) {
    float v[16];
    if (G == 2 && p.bias9) {  // (row class, column class) bias table of the fused transposed conv
      const int x = x0 + px;
      const int boff = (((yy == 0) ? 0 : (yy == p.H - 1 ? 2 : 1)) * 3 + ((x == 0) ? 0 : (x == p.W - 1 ? 2 : 1))) * 16;
      const float4* b4 = reinterpret_cast<const float4*>(s_bias + boff);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 b = b4[k];
        v[4 * k] = __uint_as_float(raw[16 * px + 4 * k]) + b.x, v[4 * k + 1] = __uint_as_float(raw[16 * px + 4 * k + 1]) + b.y;
        v[4 * k + 2] = __uint_as_float(raw[16 * px + 4 * k + 2]) + b.z, v[4 * k + 3] = __uint_as_float(raw[16 * px + 4 * k + 3]) + b.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(raw[16 * px + k]) + bias_r[k];
    }
    if constexpr (!TRAIN && !HEAD) {
      // plain inference epilogue: the ReLU clamp is folded into the bf16 conversion
#pragma unroll
      for (int k = 0; k < 8; ++k) words[8 * px + k] = relu ? cvt_bf16x2_relu(v[2 * k], v[2 * k + 1]) : cvt_bf16x2(v[2 * k], v[2 * k + 1]);
    } else {
      if constexpr (TRAIN && !HEAD) {
        if (p.addend) {
          const uint32_t aw[8] = {t.a[2 * px].x, t.a[2 * px].y, t.a[2 * px].z, t.a[2 * px].w, t.a[2 * px + 1].x, t.a[2 * px + 1].y, t.a[2 * px + 1].z, t.a[2 * px + 1].w};
#pragma unroll
          for (int k = 0; k < 8; ++k) v[2 * k] += bf16_lo(aw[k]), v[2 * k + 1] += bf16_hi(aw[k]);
        }
      }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
      }
      if constexpr (TRAIN && !HEAD) {
        if (p.relu_mask_src) {
          const uint32_t mw[8] = {t.m[2 * px].x, t.m[2 * px].y, t.m[2 * px].z, t.m[2 * px].w, t.m[2 * px + 1].x, t.m[2 * px + 1].y, t.m[2 * px + 1].z, t.m[2 * px + 1].w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (!(bf16_lo(mw[k]) > 0.f)) v[2 * k] = 0.f;
            if (!(bf16_hi(mw[k]) > 0.f)) v[2 * k + 1] = 0.f;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) words[8 * px + k] = cvt_bf16x2(v[2 * k], v[2 * k + 1]);
      if constexpr (TRAIN && !HEAD) {
        if (p.stats_partial && valid) {
          // statistics of the bf16-rounded values that are stored; second statistic v*v (BN batch variance) or v*aux (BN backward)
          const uint32_t xw[8] = {t.x[2 * px].x, t.x[2 * px].y, t.x[2 * px].z, t.x[2 * px].w, t.x[2 * px + 1].x, t.x[2 * px + 1].y, t.x[2 * px + 1].z, t.x[2 * px + 1].w};
          const bool aux = p.stats_aux != nullptr;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float r0 = bf16_lo(words[8 * px + k]), r1 = bf16_hi(words[8 * px + k]);
            sa1[2 * k] += r0, sa1[2 * k + 1] += r1;
            sa2[2 * k] = fmaf(r0, aux ? bf16_lo(xw[k]) : r0, sa2[2 * k]), sa2[2 * k + 1] = fmaf(r1, aux ? bf16_hi(xw[k]) : r1, sa2[2 * k + 1]);
          }
        }
      }
      if constexpr (HEAD) {
#pragma unroll
        for (int k = 0; k < 16; ++k) hv[px][k] = v[k];
        if constexpr (TRAIN) {
          if (p.drop_mask) {
#pragma unroll
            for (int k = 0; k < 16; ++k) hv[px][k] = ((t.dbits >> (16 * px + k)) & 1u) ? v[k] * p.drop_scale : 0.f;
          }
        }
      }
    }
  }
  if (!valid) return;
  if (p.out) {
    uint8_t* op = reinterpret_cast<uint8_t*>(p.out + eoff);
#pragma unroll
    for (int px = 0; px < G; ++px) st_global_v8(op + 32 * px, words + 8 * px);
  }
  if constexpr (HEAD && G == 2) {
    const size_t plane = size_t(p.H) * p.W;
    size_t o = size_t(n) * p.head_classes * plane + size_t(yy) * p.W + x0;
    for (int cls = 0; cls < p.head_classes; ++cls, o += plane) {
      const float4* w4 = reinterpret_cast<const float4*>(s_head + cls * 16);
      float a0 = s_head[8 * 16 + cls], a1 = a0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 w = w4[k];
        a0 = fmaf(w.x, hv[0][4 * k], a0), a0 = fmaf(w.y, hv[0][4 * k + 1], a0), a0 = fmaf(w.z, hv[0][4 * k + 2], a0), a0 = fmaf(w.w, hv[0][4 * k + 3], a0);
        a1 = fmaf(w.x, hv[G - 1][4 * k], a1), a1 = fmaf(w.y, hv[G - 1][4 * k + 1], a1), a1 = fmaf(w.z, hv[G - 1][4 * k + 2], a1), a1 = fmaf(w.w, hv[G - 1][4 * k + 3], a1);
      }
      if constexpr (TRAIN) {
        if (p.logit) *reinterpret_cast<float2*>(p.logit + o) = make_float2(a0, a1);
      }
      *reinterpret_cast<float2*>(p.heat + o) = make_float2(__fdividef(1.f, 1.f + __expf(-a0)), __fdividef(1.f, 1.f + __expf(-a1)));
    }
  }
}

template <bool DECONV, bool HEAD, bool TRAIN>
__global__ void __launch_bounds__(block_threads(heavy_epilogue(HEAD, TRAIN)), 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  constexpr bool kHeavy = heavy_epilogue(HEAD, TRAIN);  // backward operands / statistics in the epilogue: 168 registers, three issuers
  constexpr int kMmaWarps = mma_warps(kHeavy), kThreads = block_threads(kHeavy);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_acc_full[4], bar_acc_empty[4], bar_w;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[256];
  __shared__ __align__(16) float s_head[8 * 16 + 8];
  __shared__ float s_stats[kHeavy ? kEpiWarps : 1][2][kHeavy ? 256 : 1];  // per epilogue warp per-channel partial sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile_idx = blockIdx.y;  // which n_tile slice of the GEMM N axis this CTA owns
  // the swizzle patterns of TMA and UMMA are functions of the absolute shared address: keep every
  // stage 1024-aligned whatever the static shared footprint turns out to be
  uint8_t* const w_smem = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* const stage0 = w_smem + p.w_smem_bytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstage; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], kMmaWarps);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bar_acc_full[i], kMmaWarps);
      mbar_init(&bar_acc_empty[i], kEpiWarps);
    }
    mbar_init(&bar_w, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_slot, p.tmem_cols);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < UNPP_MAX_SRC; ++i)
      if (i < p.nchunk) tma_prefetch_desc(&p.maps[p.ch_map[i]]);
    if (p.low_on) tma_prefetch_desc(&p.lowmap);
  }
  // Everything above is on-chip set-up and overlaps the tail of the preceding kernel (programmatic dependent launch);
  // from here on global memory written by earlier kernels is read.
  unpp::pdl_wait();
  unpp::pdl_trigger();
  // stage the per-column bias (conv: this CTA's n_tile slice; deconv: all Cout) and the 1x1 head
  {
    const int nb = p.bias9 ? 9 * 16 : ((DECONV || p.b2) ? p.cout : p.ncols);
    for (int i = threadIdx.x; i < nb; i += kThreads) s_bias[i] = p.bias ? __ldg(p.bias + (DECONV ? 0 : ntile_idx * p.ncols) + i) : 0.f;
    if constexpr (HEAD) {
      for (int i = threadIdx.x; i < p.head_classes * 16; i += kThreads) s_head[i] = __ldg(p.head_w + i);
      for (int i = threadIdx.x; i < p.head_classes; i += kThreads) s_head[8 * 16 + i] = __ldg(p.head_b + i);
    }
    if constexpr (kHeavy) {
      for (int i = threadIdx.x; i < kEpiWarps * 2 * 256; i += kThreads) (&s_stats[0][0][0])[i] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int pad = p.taps == 9 ? 1 : 0;
  const int P = p.TW + 2 * pad;    // tile pitch in pixels
  const int rows = 16 + 2 * pad;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar_w, p.w_bytes + (p.low_on ? p.low_w_bytes : 0));
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpacked) + size_t(ntile_idx) * p.w_bytes;
      for (int off = 0; off < p.w_bytes; off += 16384) {
        int n = p.w_bytes - off < 16384 ? p.w_bytes - off : 16384;
        bulk_load(w_smem + off, wsrc + off, n, &bar_w);
      }
      if (p.low_on) {
        const uint8_t* lsrc = reinterpret_cast<const uint8_t*>(p.low_wpacked);
        for (int off = 0; off < p.low_w_bytes; off += 16384) {
          int n = p.low_w_bytes - off < 16384 ? p.low_w_bytes - off : 16384;
          bulk_load(w_smem + p.low_w_off + off, lsrc + off, n, &bar_w);
        }
      }
      // Box of step (tile, chunk): tensor map and coordinates, shared by the load and by the L2 prefetch that runs ahead of it.
      auto box_of = [&](int tile, int c, const CUtensorMap*& map, int& c0, int& c1, int& c2, int& c3) {
        const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y;
        c3 = tile / (p.tiles_x * p.tiles_y);
        if (c == p.nchunk) map = &p.lowmap, c0 = 0, c1 = tx * (p.TW >> 1) - 1, c2 = ty * 16 - 1;
        else if (p.b2) map = &p.maps[p.ch_map[c]], c0 = 0, c1 = tx * (p.TW >> 1) - 1, c2 = ty * 32 - 1;
        else map = &p.maps[p.ch_map[c]], c0 = p.ch_c0[c], c1 = tx * p.TW - pad, c2 = ty * 16 - pad;
      };
      const int nper = p.nchunk + p.low_on;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, n = tile / (p.tiles_x * p.tiles_y);
        for (int c = 0; c < p.nchunk + p.low_on; ++c, ++it) {
          const int s = it % p.nstage, ph = (it / p.nstage) & 1;
          if (p.pf && !(p.dbg & 4)) {
            // The stage ring bounds the bytes a CTA has in flight (two 74 KB stages next to the resident weights of the
            // many-source layers); pulling the boxes of later steps into L2 now decouples the DRAM fetch from that bound:
            // the load issued when the stage frees is an L2 hit.
            const int e = it + p.nstage - 1 + p.pf;
            const int ptile = blockIdx.x + (e / nper) * gridDim.x;
            if (ptile < p.ntiles) {
              const CUtensorMap* m;
              int c0, c1, c2, c3;
              box_of(ptile, e % nper, m, c0, c1, c2, c3);
              tma_prefetch_4d(m, c0, c1, c2, c3);
            }
          }
          mbar_wait(&bar_empty[s], ph ^ 1);
          if (p.dbg & 4) {
            mbar_arrive(&bar_full[s]);
            continue;
          }
          if (c == p.nchunk) {  // low-resolution halo tile of the fused transposed conv: 18 rows x (TW/2 + 2) pixels of 64 B
            mbar_arrive_expect_tx(&bar_full[s], 18 * p.low_P * 64);
            tma_load_4d(&p.lowmap, &bar_full[s], stage0 + size_t(s) * p.stage_bytes, 0, tx * (p.TW >> 1) - 1, ty * 16 - 1, n);
            continue;
          }
          if (p.b2) {  // pixel-pair rows (64 B), 34 image rows x (TW/2 + 2) pairs around a 32 x TW output tile
            mbar_arrive_expect_tx(&bar_full[s], 34 * p.b2_P * (p.c4 ? 16 : 64));
            tma_load_4d(&p.maps[p.ch_map[c]], &bar_full[s], stage0 + size_t(s) * p.stage_bytes, 0, tx * (p.TW >> 1) - 1, ty * 32 - 1, n);
            continue;
          }
          mbar_arrive_expect_tx(&bar_full[s], rows * P * p.ch_span[c]);
          tma_load_4d(&p.maps[p.ch_map[c]], &bar_full[s], stage0 + size_t(s) * p.stage_bytes, p.ch_c0[c], tx * p.TW - pad,
                      ty * 16 - pad, n);
        }
      }
    }
  } else if (warp == 1 || warp >= 2 + kEpiWarps) {
    // ------------------------------------------------------------------ MMA issuers (warp 1 and warps 10..)
    // One thread can only sustain ~1 MMA per 85 cycles through the uniform datapath (measured: the
    // tensor pipe takes ~40); three or four issuing threads, each with its own accumulators, remove that limit.
    const int mw = warp == 1 ? 0 : warp - (1 + kEpiWarps);
    const bool has0 = mw < p.nsub, has1 = mw + kMmaWarps < p.nsub, has2 = kMmaWarps < 4 && mw + 2 * kMmaWarps < p.nsub;  // (8 sub-tiles at most: a third slot only exists with three issuers)
    // every kernel parameter used below is copied to a register first: the loop must not touch
    // the parameter bank between MMA issues
    mbar_wait(&bar_w, 0);
    const int ncols = p.ncols, nsub = p.nsub, taps = p.taps, k8_total = p.k8_total, nchunk = p.nchunk, nstage = p.nstage;
    const int ntiles = p.ntiles, stage_bytes = p.stage_bytes, dbg = p.dbg, b2 = p.b2, b2_P = p.b2_P, nacc = p.nacc;
    const int low_on = p.low_on, low_P = p.low_P, low_k8 = p.low_k8, low_w_off = p.low_w_off, c4 = p.c4;
    const uint32_t idesc = make_idesc_bf16(128, ncols), idesc32 = make_idesc_bf16(128, 32), idesc16 = make_idesc_bf16(128, 16);
    const uint32_t w_addr = smem_u32(w_smem);
    const uint32_t stage_addr0 = smem_u32(stage0);
    int spans[kMaxChunks], wk8s[kMaxChunks];
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) spans[c] = p.ch_span[c], wk8s[c] = p.ch_wk8[c];
    int it = 0, tile_it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
      const int b = tile_it % nacc, aph = (tile_it / nacc) & 1;
      mbar_wait(&bar_acc_empty[b], aph ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + uint32_t(b * nsub * ncols);
#pragma unroll 1
      for (int c = 0; c < nchunk + low_on; ++c, ++it) {
        const int s = it % nstage, ph = (it / nstage) & 1;
        int span = spans[0], wk8 = wk8s[0];
#pragma unroll
        for (int cc = 1; cc < kMaxChunks; ++cc)
          if (cc == c) span = spans[cc], wk8 = wk8s[cc];
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        // Issue loops: fully unrolled over the window positions / taps, descriptors formed from a constant high
        // word and a 32-bit low word (start address + LBO) so that every MMA costs two independent adds, not a
        // serial chain through the uniform datapath (the single issuing thread is the limiter at small N).
        if (c == nchunk) {
          if (elect_one() && !(dbg & 2)) {
            // fused transposed conv: GEMM row = block = one LOW-resolution pixel (8 consecutive pixels = 8 rows of 64 B,
            // next row group one low-res row down); classic 3x3 taps over the low-res halo tile, two K=16 slabs per
            // tap (32 channels), the 64 columns are (pixel of the 2x2 block, co) with composed weights.
            const uint64_t ad = make_sdesc(stage_addr0 + uint32_t(s) * stage_bytes, 16, uint32_t(low_P * 64), 4);
            const uint64_t bd = make_sdesc(w_addr + uint32_t(low_w_off), 0, 128, 0);
            const uint32_t a_hi = uint32_t(ad >> 32), b_hi = uint32_t(bd >> 32), b_lo = uint32_t(bd);
            const uint32_t a_lo0 = uint32_t(ad) + uint32_t(mw * 32), a_lo1 = a_lo0 + uint32_t(kMmaWarps * 32);
            const uint32_t acc0 = acc + uint32_t(mw * 64), acc1 = acc + uint32_t((mw + kMmaWarps) * 64);
            const uint32_t row_step = uint32_t(low_P * 4);
            const int kslabs = low_k8 >> 1;
            for (int ks = 0; ks < kslabs; ++ks) {
              const uint32_t b_ks = b_lo + (uint32_t(ks * b2::kLowUnits * b2::kUnitBytes) >> 4);
#pragma unroll
              for (int i = 0; i < b2::kLowBlks; ++i) {  // one MMA per run of non-zero pixel blocks (b2_blocks.h); this chunk is never the first
                const b2::Blk blk = b2::low_blk(i);
                const uint32_t ao = uint32_t(blk.pos / 3) * row_step + uint32_t((blk.pos % 3) * 4 + ks * 2);
                const uint32_t nb = uint32_t(blk.nblk);
                const uint64_t bdesc = (uint64_t(b_hi) << 32) | (b_ks + (uint32_t(blk.cum * b2::kUnitBytes) >> 4) + (nb << 20));  // LBO = 16 * nblk * 16 B
                const uint32_t idn = nb == 4 ? idesc : nb == 2 ? idesc32 : idesc16, dcol = uint32_t(blk.b0 * 16);
                if (has0) umma_bf16(acc0 + dcol, (uint64_t(a_hi) << 32) | (a_lo0 + ao), bdesc, idn, 1u);
                if constexpr (kMmaWarps < 4) {  // (b2 tiles have at most four sub-tiles: with four issuers nobody owns a second one)
                  if (has1) umma_bf16(acc1 + dcol, (uint64_t(a_hi) << 32) | (a_lo1 + ao), bdesc, idn, 1u);
                }
              }
            }
          }
        } else if (c4) {
          if (elect_one() && !(dbg & 2)) {
            // First layer (<= 4 input channels, 8-byte pixels): 2x2 output blocks again, but one K = 16 step is a whole
            // window ROW segment — the 16-byte pair rows are unswizzled K-major core-matrix rows, GEMM row j (block j) starts
            // at pair j and its two 16-byte K chunks are pairs (j, j+1) resp. (j+2, j+3): LBO = 16 B makes consecutive GEMM
            // rows overlapping windows of the same staged image row.  Window pixels -2 .. +5 around the block (weights are
            // zero outside -1 .. +2): 2 MMAs of N = 64 per window row, 8 per 512 output pixels (20 on the 16-channel path).
            const uint64_t ad = make_sdesc(stage_addr0 + uint32_t(s) * stage_bytes, 16, uint32_t(2 * b2_P * 16), 0);
            const uint64_t bd = make_sdesc(w_addr, 1024, 128, 0);  // packed [dy][4 K chunks][64 columns][8] (pack kind 7)
            const uint32_t a_hi = uint32_t(ad >> 32), b_hi = uint32_t(bd >> 32), b_lo = uint32_t(bd);
            const uint32_t a_lo0 = uint32_t(ad) + uint32_t(mw * 8), a_lo1 = a_lo0 + uint32_t(kMmaWarps * 8);
            const uint32_t acc0 = acc + uint32_t(mw * 64), acc1 = acc + uint32_t((mw + kMmaWarps) * 64);
#pragma unroll
            for (int dy = 0; dy < 4; ++dy) {
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const uint32_t ao = uint32_t(dy * b2_P + 2 * hf);
                const uint64_t bdesc = (uint64_t(b_hi) << 32) | (b_lo + uint32_t((dy * 4096 + hf * 2048) >> 4));
                const uint32_t accum = (dy | hf) ? 1u : 0u;
                if (has0) umma_bf16(acc0, (uint64_t(a_hi) << 32) | (a_lo0 + ao), bdesc, idesc, accum);
                if constexpr (kMmaWarps < 4) {
                  if (has1) umma_bf16(acc1, (uint64_t(a_hi) << 32) | (a_lo1 + ao), bdesc, idesc, accum);
                }
              }
            }
          }
        } else if (b2) {
          if (elect_one() && !(dbg & 2)) {
            // 2x2 output blocks: GEMM row = block (8 consecutive pixel pairs = 8 rows of 64 B, next row group two
            // image rows down), K walks the 4x4 input window: pixel (dy, dx) of the window is 32 B into / past the
            // pair row, i.e. a start-address shift of (dy * pairs_per_row * 4 + (dx + 1) * 2) 16-byte units.
            const uint64_t ad = make_sdesc(stage_addr0 + uint32_t(s) * stage_bytes, 16, uint32_t(2 * b2_P * 64), 4);
            const uint64_t bd = make_sdesc(w_addr + uint32_t((wk8 >> 1) * b2::kMainUnits * b2::kUnitBytes), 0, 128, 0);
            const uint32_t a_hi = uint32_t(ad >> 32), b_hi = uint32_t(bd >> 32), b_lo = uint32_t(bd);
            const uint32_t a_lo0 = uint32_t(ad) + 2u + uint32_t(mw * 32), a_lo1 = a_lo0 + uint32_t(kMmaWarps * 32);
            const uint32_t acc0 = acc + uint32_t(mw * 64), acc1 = acc + uint32_t((mw + kMmaWarps) * 64);
            const uint32_t row_step = uint32_t(b2_P * 4);
            const uint32_t first = c ? 1u : 0u;
#pragma unroll
            for (int i = 0; i < b2::kMainBlks; ++i) {  // one MMA per run of non-zero pixel blocks (b2_blocks.h); block 0 covers all 64 columns
              const b2::Blk blk = b2::main_blk(i);
              const uint32_t ao = uint32_t(blk.pos >> 2) * row_step + uint32_t((blk.pos & 3) * 2);
              const uint32_t nb = uint32_t(blk.nblk);
              const uint64_t bdesc = (uint64_t(b_hi) << 32) | (b_lo + (uint32_t(blk.cum * b2::kUnitBytes) >> 4) + (nb << 20));  // LBO = 16 * nblk * 16 B
              const uint32_t idn = nb == 4 ? idesc : nb == 2 ? idesc32 : idesc16, dcol = uint32_t(blk.b0 * 16);
              const uint32_t accum = i ? 1u : first;
              if (has0) umma_bf16(acc0 + dcol, (uint64_t(a_hi) << 32) | (a_lo0 + ao), bdesc, idn, accum);
              if constexpr (kMmaWarps < 4) {
                if (has1) umma_bf16(acc1 + dcol, (uint64_t(a_hi) << 32) | (a_lo1 + ao), bdesc, idn, accum);
              }
            }
          }
        } else if (elect_one() && !(dbg & 2)) {
          const int kslabs = span >> 5;
          const uint32_t px_step = uint32_t(span) >> 4, row_step = uint32_t(P * span) >> 4;
          const uint32_t b_tap_step = uint32_t(k8_total * ncols * 16) >> 4, b_ks_step = uint32_t(2 * ncols * 16) >> 4;
          const uint32_t sub_step = uint32_t(8 * span) >> 4;  // next 8-pixel patch column, in 16 B units
          const uint64_t ad = make_sdesc(stage_addr0 + uint32_t(s) * stage_bytes, 16, uint32_t(P * span), layout_type_of_span(span));
          const uint64_t bd = make_sdesc(w_addr + uint32_t(wk8 * ncols * 16), uint32_t(ncols * 16), 128, 0);
          const uint32_t a_hi = uint32_t(ad >> 32), b_hi = uint32_t(bd >> 32), b_lo = uint32_t(bd);
          const uint32_t a_lo0 = uint32_t(ad) + uint32_t(mw) * sub_step, a_lo1 = a_lo0 + uint32_t(kMmaWarps) * sub_step,
                         a_lo2 = a_lo1 + uint32_t(kMmaWarps) * sub_step;
          const uint32_t acc0 = acc + uint32_t(mw * ncols), acc1 = acc + uint32_t((mw + kMmaWarps) * ncols),
                         acc2 = acc + uint32_t((mw + 2 * kMmaWarps) * ncols);
          const uint32_t first = c ? 1u : 0u;
          // every MMA costs its issuing thread a handful of uniform-datapath instructions per (predicated) slot: when no issuer owns a
          // second sub-tile (nsub <= issuers: the deep levels) the loop is issued without the dead slots
          auto issue_tap1 = [&](uint32_t ao, uint32_t bo, bool first_tap) {
#pragma unroll 4
            for (int ks = 0; ks < kslabs; ++ks) {
              const uint64_t bdesc = (uint64_t(b_hi) << 32) | (b_lo + bo + uint32_t(ks) * b_ks_step);
              if (has0) umma_bf16(acc0, (uint64_t(a_hi) << 32) | (a_lo0 + ao + uint32_t(ks * 2)), bdesc, idesc, (first_tap && ks == 0) ? first : 1u);
            }
          };
          auto issue_tap = [&](uint32_t ao, uint32_t bo, bool first_tap) {
#pragma unroll 4
            for (int ks = 0; ks < kslabs; ++ks) {
              const uint32_t aok = ao + uint32_t(ks * 2);
              const uint64_t bdesc = (uint64_t(b_hi) << 32) | (b_lo + bo + uint32_t(ks) * b_ks_step);
              const uint32_t accum = (first_tap && ks == 0) ? first : 1u;
              if (has0) umma_bf16(acc0, (uint64_t(a_hi) << 32) | (a_lo0 + aok), bdesc, idesc, accum);
              if (has1) umma_bf16(acc1, (uint64_t(a_hi) << 32) | (a_lo1 + aok), bdesc, idesc, accum);
              if constexpr (kMmaWarps < 4) {
                if (has2) umma_bf16(acc2, (uint64_t(a_hi) << 32) | (a_lo2 + aok), bdesc, idesc, accum);
              }
            }
          };
          if (taps == 9 && nsub <= kMmaWarps) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
              for (int sft = 0; sft < 3; ++sft) issue_tap1(uint32_t(r) * row_step + uint32_t(sft) * px_step, uint32_t(r * 3 + sft) * b_tap_step, (r | sft) == 0);
            }
          } else if (taps == 9) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
              for (int sft = 0; sft < 3; ++sft) issue_tap(uint32_t(r) * row_step + uint32_t(sft) * px_step, uint32_t(r * 3 + sft) * b_tap_step, (r | sft) == 0);
            }
          } else {
            issue_tap(0u, 0u, true);
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&bar_empty[s]);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar_acc_full[b]);
      __syncwarp();
    }
  } else if (warp < 2 + kEpiWarps) {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quadrant; the (sub-tile, 16-column group) units of a tile alternate
    // between them.  The TMEM load of the next unit is in flight while the current one is processed.
    const int ew = warp - 2;
    const int q = warp & 3;              // TMEM lane quadrant this warp may read
    const int half = ew >> 2;
    const int m = q * 32 + lane;         // accumulator row
    const int pi = m >> 3, pj = m & 7;   // pixel inside the 16x8 patch
    const int ncb = p.ncols >> 4, units = p.nsub * ncb;
    const float relu_floor = p.relu ? 0.f : -INFINITY;
    EpiArgs e;
    e.dbg = p.dbg, e.bias9 = p.bias9, e.pooled = p.pooled;
    e.H = p.H, e.W = p.W, e.cout = p.cout, e.head_classes = p.head_classes, e.out = p.out, e.heat = p.heat, e.logit = p.logit;
    e.drop_mask = p.drop_mask, e.drop_scale = p.drop_scale, e.addend = p.addend, e.relu_mask_src = p.relu_mask_src;
    e.stats_partial = p.stats_partial, e.stats_aux = p.stats_aux, e.aux_mean = p.aux_mean, e.aux_istd = p.aux_istd;
    const int TW = p.TW, ncols = p.ncols, nsub = p.nsub, ntiles = p.ntiles, tiles_x = p.tiles_x, tiles_y = p.tiles_y, dbg = p.dbg, b2 = p.b2, nacc = p.nacc;
    float* const st1 = kHeavy ? &s_stats[kHeavy ? ew : 0][0][0] : nullptr;
    float* const st2 = kHeavy ? &s_stats[kHeavy ? ew : 0][1][0] : nullptr;
    // with <= 2 column groups every unit of this warp has the same 16 channels: keep the statistics in registers
    const bool reg_stats = kHeavy && (ncb <= 2 || b2);  // (2x2 blocking: the four column groups are four pixels of the same 16 channels)
    float sa1[16], sa2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) sa1[k] = 0.f, sa2[k] = 0.f;
    int tile_it = 0;
    if (b2) {
      // 2x2 output blocks: this warp takes pixel row dy = half of every block; one step = two adjacent pixels (32 columns)
      float bias_r[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) bias_r[k] = e.bias9 ? 0.f : s_bias[k];
      const bool relu = p.relu != 0;
      uint32_t wds[16];
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
        const int b = tile_it % nacc, aph = (tile_it / nacc) & 1;
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        if constexpr (!TRAIN && !HEAD) {
          if (e.pooled) {
            // fused 2x2 max pool: this warp takes BOTH pixel rows of the blocks of sub-tiles half, half + 2, ...
            mbar_wait(&bar_acc_full[b], aph);
            tc_fence_after();
            const int y0 = ty * 32 + 2 * pi, xb2 = tx * TW + 2 * pj;
            const size_t rowpix0 = (size_t(n) * e.H + y0) * e.W;
            const uint32_t tb = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * nsub * 64);
            for (int jj = half; jj < nsub && !(dbg & 1); jj += 2) {
              uint32_t A[32], pw[8];
              EpiOperands<2> tops;
              const int x0 = xb2 + jj * 16;
              const bool valid = y0 < e.H && x0 < e.W;  // (even H, W: the whole 2x2 block is inside or outside)
              tmem_ld32(tb + uint32_t(jj * 64), A);
              tmem_ld_wait32(A);
              epi_finish<2, HEAD, TRAIN>(e, A, tops, bias_r, s_bias, s_head, relu, (rowpix0 + x0) * 16, valid, n, y0, x0, sa1, sa2, wds);
#pragma unroll
              for (int k = 0; k < 8; ++k) pw[k] = max_bf16x2(wds[k], wds[8 + k]);
              tmem_ld32(tb + uint32_t(jj * 64 + 32), A);
              tmem_ld_wait32(A);
              epi_finish<2, HEAD, TRAIN>(e, A, tops, bias_r, s_bias, s_head, relu, (rowpix0 + e.W + x0) * 16, valid, n, y0 + 1, x0, sa1, sa2, wds);
#pragma unroll
              for (int k = 0; k < 8; ++k) wds[k] = max_bf16x2(pw[k], max_bf16x2(wds[k], wds[8 + k]));
              if (valid) st_global_v8(e.pooled + ((size_t(n) * (e.H >> 1) + (y0 >> 1)) * (e.W >> 1) + (x0 >> 1)) * 16, wds);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
            continue;
          }
        }
        const int yy = ty * 32 + 2 * pi + half, xb = tx * TW + 2 * pj;
        const bool row_ok = yy < e.H;
        const size_t rowpix = (size_t(n) * e.H + yy) * e.W;
        if constexpr (TRAIN) {
          for (int j = 0; j < nsub; ++j) epi_l2_prefetch<2, HEAD, TRAIN>(e, (rowpix + xb + j * 16) * 16, row_ok && xb + j * 16 < e.W);
        }
        mbar_wait(&bar_acc_full[b], aph);
        tc_fence_after();
        const uint32_t tcol = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * nsub * 64 + half * 32);
        if (!(dbg & 1)) {
          if constexpr (TRAIN || HEAD) {  // one TMEM buffer (registers): operands of the step are in flight while the load is waited for
            uint32_t A[32];
            for (int j = 0; j < nsub; ++j) {
              const int x0 = xb + j * 16;
              const bool valid = row_ok && x0 < e.W;
              EpiOperands<2> tops;
              tmem_ld32(tcol + uint32_t(j * 64), A);
              epi_prefetch<2, HEAD, TRAIN>(e, tops, (rowpix + x0) * 16, valid);
              tmem_ld_wait32(A);
              epi_finish<2, HEAD, TRAIN>(e, A, tops, bias_r, s_bias, s_head, relu, (rowpix + x0) * 16, valid, n, yy, x0, sa1, sa2, wds);
            }
          } else {  // two TMEM buffers: the load of the next step is in flight while this one is finished
            uint32_t A[32], B[32];
            EpiOperands<2> tops;
            tmem_ld32(tcol, A);
            for (int j = 0; j < nsub; j += 2) {
              const int x0 = xb + j * 16;
              tmem_ld_wait32(A);
              if (j + 1 < nsub) tmem_ld32(tcol + uint32_t((j + 1) * 64), B);
              epi_finish<2, HEAD, TRAIN>(e, A, tops, bias_r, s_bias, s_head, relu, (rowpix + x0) * 16, row_ok && x0 < e.W, n, yy, x0, sa1, sa2, wds);
              if (j + 1 < nsub) {
                tmem_ld_wait32(B);
                if (j + 2 < nsub) tmem_ld32(tcol + uint32_t((j + 2) * 64), A);
                epi_finish<2, HEAD, TRAIN>(e, B, tops, bias_r, s_bias, s_head, relu, (rowpix + x0 + 16) * 16, row_ok && x0 + 16 < e.W, n, yy, x0 + 16, sa1, sa2, wds);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
      }
    } else if (!HEAD && !(TRAIN && DECONV) && (ncb == 1 || ncb == 2 || ncb == 4 || ncb == 8)) {
      // classic path, a step = 16 channels of one pixel = 32 contiguous bytes.  With <= 32 GEMM columns per CTA (the
      // 32-channel level, 16-channel dgrads) every unit of this warp has the same 16 channels, so bias and statistics
      // live in registers; wider CTAs re-read the 16 bias values of the step from shared memory.
      // Training variants of the 64 / 128-column layers (deep encoder levels, their dgrads) walk their units GROUP-MAJOR: all
      // sub-tiles of one 16-channel group, then the next group, so that the statistics of a group stay in registers and are
      // folded into the warp's shared-memory slots once per (tile, group) instead of once per unit (64 shuffles each).
      const bool fixed_c = !DECONV && ncb <= 2;
      const bool gm = TRAIN && ncb > 2;
      const int nsub_shift = nsub == 8 ? 3 : nsub == 4 ? 2 : nsub == 2 ? 1 : 0;  // nsub = TW / 8 is a power of two
      const int ncb_shift = ncb == 8 ? 3 : ncb == 4 ? 2 : ncb == 2 ? 1 : 0;  // ncols is 16 << ncb_shift on this path (see the guard below)
      float bias_r[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) bias_r[k] = fixed_c ? s_bias[(half % ncb) * 16 + k] : 0.f;
      const bool relu = p.relu != 0;
      const int nit = (units - half + 1) / 2;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
        const int b = tile_it % nacc, aph = (tile_it / nacc) & 1;
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        const int yy = ty * 16 + pi, xb = tx * TW + pj;
        const bool row_ok = yy < e.H;
        const size_t rowpix = (size_t(n) * e.H + yy) * e.W;
        if constexpr (TRAIN && !DECONV) {
          for (int it = 0; it < nit; ++it) {
            const int u = half + 2 * it;
            const int j = gm ? (it & (nsub - 1)) : (u >> ncb_shift), c0 = gm ? (half + 2 * (it >> nsub_shift)) * 16 : (u & (ncb - 1)) * 16;
            epi_l2_prefetch<1, false, TRAIN>(e, (rowpix + xb + j * 8) * e.cout + ntile_idx * ncols + c0, row_ok && xb + j * 8 < e.W);
          }
        }
        mbar_wait(&bar_acc_full[b], aph);
        tc_fence_after();
        const uint32_t tcol = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * nsub * ncols);
        if (!(dbg & 1)) {
          uint32_t A[16], wds1[8];
          EpiOperands<1> tops;
          for (int it = 0; it < nit; ++it) {
            const int u = half + 2 * it;  // unit = (sub-tile, 16-column group)
            const int j = gm ? (it & (nsub - 1)) : (u >> ncb_shift), c0 = gm ? (half + 2 * (it >> nsub_shift)) * 16 : (u & (ncb - 1)) * 16;
            const int x0 = xb + j * 8;
            const bool valid = row_ok && x0 < e.W;
            size_t eoff;
            int boff = c0;
            if constexpr (DECONV) {  // scatter: GEMM column = (2p + q) * Cout + co -> output pixel (2y + p, 2x + q) of the [N, 2H, 2W, Cout] tensor
              const int gcol = ntile_idx * ncols + c0, pq = gcol / e.cout;
              boff = gcol - pq * e.cout;
              eoff = ((size_t(n) * (2 * e.H) + (2 * yy + (pq >> 1))) * (2 * e.W) + (2 * x0 + (pq & 1))) * e.cout + boff;
            } else {
              eoff = (rowpix + x0) * e.cout + ntile_idx * ncols + c0;
            }
            tmem_ld16(tcol + uint32_t(j * ncols + c0), A);
            epi_prefetch<1, false, TRAIN>(e, tops, eoff, valid);
            if (!fixed_c) {
              const float4* b4 = reinterpret_cast<const float4*>(s_bias + boff);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float4 bb = b4[k];
                bias_r[4 * k] = bb.x, bias_r[4 * k + 1] = bb.y, bias_r[4 * k + 2] = bb.z, bias_r[4 * k + 3] = bb.w;
              }
            }
            tmem_ld_wait16(A);
            epi_finish<1, false, TRAIN>(e, A, tops, bias_r, s_bias, s_head, relu, eoff, valid, n, yy, x0, sa1, sa2, wds1);
            if constexpr (kHeavy) {
              if (gm && j == nsub - 1 && e.stats_partial) {  // last sub-tile of this 16-channel group: fold its statistics into the warp's slots
                const float r1 = warp_reduce16(sa1, lane), r2 = warp_reduce16(sa2, lane);
                if ((lane & 1) == 0) st1[c0 + (lane >> 1)] += r1, st2[c0 + (lane >> 1)] += r2;
#pragma unroll
                for (int k = 0; k < 16; ++k) sa1[k] = 0.f, sa2[k] = 0.f;
              }
            }
            if constexpr (!TRAIN && !DECONV) {
              if (e.pooled) {  // 2x2 max pool across the four lanes that hold the window (pj ^ 1 = lane ^ 1, pi ^ 1 = lane ^ 8)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  wds1[k] = max_bf16x2(wds1[k], __shfl_xor_sync(0xffffffffu, wds1[k], 1));
                  wds1[k] = max_bf16x2(wds1[k], __shfl_xor_sync(0xffffffffu, wds1[k], 8));
                }
                if (valid && !(lane & 9))
                  st_global_v8(e.pooled + ((size_t(n) * (e.H >> 1) + (yy >> 1)) * (e.W >> 1) + (x0 >> 1)) * e.cout + ntile_idx * ncols + c0, wds1);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
      }
    } else
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
      const int b = tile_it % nacc, aph = (tile_it / nacc) & 1;
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
      mbar_wait(&bar_acc_full[b], aph);
      tc_fence_after();
      const int y = b2 ? ty * 32 + 2 * pi : ty * 16 + pi;
      const uint32_t tbase = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * nsub * ncols);
      uint32_t raw[16];
      // classic: this warp takes units half, half+2, ...; 2x2: it takes block row jy = half and walks (sub-tile, jx), so
      // that the two horizontally adjacent pixels of a block (64 contiguous bytes) are read / written back to back
      const bool pair_order = TRAIN && b2;  // (the plain inference epilogue measured slightly faster with the interleaved split)
      const int nit = (dbg & 1) ? 0 : (pair_order ? nsub * 2 : (units - half + 1) / 2);
      auto unit_of = [&](int it) { return pair_order ? ((it >> 1) * 4 + half * 2 + (it & 1)) : half + 2 * it; };
      if (nit > 0 && !(dbg & 16)) tmem_ld16(tbase + uint32_t(unit_of(0) * 16), raw);  // unit u covers columns [16u, 16u+16) of this accumulator set
      for (int it = 0; it < nit; ++it) {
        const int u = unit_of(it);
        // classic: unit = (8-pixel-wide sub-tile j, 16-column group); 2x2: unit = (16-pixel-wide sub-tile, pixel of the block)
        const int j = b2 ? (u >> 2) : u / ncb, c0 = b2 ? 0 : (u % ncb) * 16;
        const int yy = b2 ? y + ((u >> 1) & 1) : y;
        const int x = b2 ? tx * TW + j * 16 + 2 * pj + (u & 1) : tx * TW + j * 8 + pj;
        const bool valid = (yy < e.H) && (x < e.W);
        TrainOperands tops;
        prefetch_train<TRAIN>(e, tops, (size_t(n) * e.H + yy) * e.W + x, ntile_idx * ncols * (b2 ? 0 : 1) + c0, valid);
        if (!(dbg & 16)) tmem_ld_wait16(raw);
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(raw[k]);
        if (it + 1 < nit && !(dbg & 16)) tmem_ld16(tbase + uint32_t(unit_of(it + 1) * 16), raw);
        epilogue_group<DECONV, HEAD, TRAIN>(e, v, tops, n, yy, x, valid, b2 ? 0 : ntile_idx * ncols + c0, c0, s_bias, s_head, relu_floor, reg_stats,
                                            sa1, sa2, st1, st2, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
    }
    if constexpr (kHeavy) {
      if (reg_stats && e.stats_partial) {
        const int c0w = b2 ? 0 : (half % ncb) * 16;  // u = half + 2i  =>  u % ncb is constant for ncb in {1, 2}
        const float r1 = warp_reduce16(sa1, lane), r2 = warp_reduce16(sa2, lane);
        if ((lane & 1) == 0) st1[c0w + (lane >> 1)] += r1, st2[c0w + (lane >> 1)] += r2;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  if constexpr (kHeavy) {
    if (p.stats_partial) {
      // partial layout [cta.x][2][n_total]; the warp slots are summed in a fixed order so the
      // result is deterministic for a given launch geometry.
      const int nch = p.b2 ? p.cout : p.ncols;
      for (int i = threadIdx.x; i < 2 * nch; i += kThreads) {
        const int st = i / nch, ch = i % nch;
        float t = 0.f;
#pragma unroll
        for (int wv = 0; wv < kEpiWarps; ++wv) t += s_stats[kHeavy ? wv : 0][st][kHeavy ? ch : 0];
        if (st == 1 && p.stats_aux) {  // sum dyh*xhat = istd * (sum dyh*z - mean * sum dyh)
          float t1 = 0.f;
#pragma unroll
          for (int wv = 0; wv < kEpiWarps; ++wv) t1 += s_stats[kHeavy ? wv : 0][0][kHeavy ? ch : 0];
          const int gc = ntile_idx * nch + ch;
          t = __ldg(p.aux_istd + gc) * (t - __ldg(p.aux_mean + gc) * t1);
        }
        p.stats_partial[(size_t(blockIdx.x) * 2 + st) * p.cout + ntile_idx * nch + ch] = t;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;  // immutable after first resolution
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct Plan {
  int TW, nsub, nstage, stage_bytes, w_bytes, w_smem_bytes, tmem_cols, smem_total, grid_x, grid_y, tiles_x, tiles_y, ntiles;
  int nchunk, k8_total, b2, b2_P, c4, cw, ncols, nacc, low_on, low_w_off, low_w_bytes;
  int ch_map[kMaxChunks], ch_c0[kMaxChunks], ch_span[kMaxChunks], ch_wk8[kMaxChunks];
};

inline bool is_train(const UnppConvArgs* a) { return a->addend || a->relu_mask_src || a->stats_partial || a->logit || a->drop_mask; }

int make_plan_cw(const UnppConvArgs* a, Plan* pl, int cw);

// K chunks are <= 64 channels wide (128-byte swizzled tile rows).  When the resident weights of a wide n_tile leave no room
// for two 16-column-or-wider stages, 32-channel chunks (64-byte rows, half the stage size) are tried: an N = 64 MMA per
// operand fetch instead of two N = 32 ones for the 128-channel layers of the deep levels.  They are also taken whenever
// they allow a wider tile (more sub-tiles = more MMA-issuing threads busy, fewer stage hand-overs per pixel):
// K64 -> N64 at 64x64, B=128: 67 -> 57 us.
int make_plan(const UnppConvArgs* a, Plan* pl) {
  int rc = make_plan_cw(a, pl, 64);
  static const int prefer32 = [] { const char* e = getenv("UNPP_CW32"); return e ? atoi(e) : 1; }();  // 32-channel chunks whenever they give a wider tile (UNPP_CW32=0: only when 64-channel chunks leave < 16 columns)
  if (a && !a->block2x2 && a->taps == 9 && a->mode == UNPP_MODE_CONV && a->n_tile >= (prefer32 ? 32 : 64) && (rc != UNPP_OK || pl->TW < (prefer32 ? 64 : 16))) {
    bool wide = true;
    for (int i = 0; i < a->nsrc && i < UNPP_MAX_SRC; ++i) wide = wide && a->src_C[i] >= 64;
    Plan alt;
    if (wide && make_plan_cw(a, &alt, 32) == UNPP_OK && alt.TW >= 16 && (rc != UNPP_OK || alt.TW > pl->TW)) {
      *pl = alt;
      return UNPP_OK;
    }
  }
  return rc;
}

int make_plan_cw(const UnppConvArgs* a, Plan* pl, int cw) {
  if (!a || a->nsrc < 1 || a->nsrc > UNPP_MAX_SRC) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: nsrc out of range");
  if (a->taps != 9 && a->taps != 1) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: taps must be 1 or 9");
  if (a->N < 1 || a->H < 1 || a->W < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: empty pixel grid");
  if (a->n_tile < 16 || a->n_tile > 256 || a->n_tile % 16 || a->n_total % a->n_tile)
    return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: n_tile must be a multiple of 16 in [16,256] dividing n_total");
  int nchunk = 0, k8 = 0, max_span = 0;
  for (int i = 0; i < a->nsrc; ++i) {
    const int C = a->src_C[i];
    const bool c4src = C == 4 && a->block2x2 == 2 && a->nsrc == 1;  // first-layer mode: one 4-channel source
    if (C != 16 && C != 32 && C != 64 && C != 128 && !c4src) return unpp::fail(UNPP_ERR_UNSUPPORTED, "conv_tc: source channels must be 16/32/64/128");
    if (!a->src[i] || (reinterpret_cast<uintptr_t>(a->src[i]) & 15)) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: source pointer null or unaligned");
    if (a->src_step[i] < 0 || a->src_step[i] > 2 || (a->src_step[i] == 2 && (a->taps != 1 || (a->src_oy[i] & ~1) || (a->src_ox[i] & ~1))))
      return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: strided sources need taps=1 and offsets in {0,1}");
    for (int c0 = 0; c0 < C; c0 += cw) {
      if (nchunk >= kMaxChunks) return unpp::fail(UNPP_ERR_UNSUPPORTED, "conv_tc: too many K chunks");
      const int w = C - c0 < cw ? C - c0 : cw;
      pl->ch_map[nchunk] = i, pl->ch_c0[nchunk] = c0, pl->ch_span[nchunk] = w * 2, pl->ch_wk8[nchunk] = k8;
      k8 += w / 8;
      if (w * 2 > max_span) max_span = w * 2;
      ++nchunk;
    }
  }
  pl->nchunk = nchunk, pl->k8_total = k8, pl->cw = cw;
  pl->b2 = 0, pl->b2_P = 0, pl->c4 = 0, pl->ncols = a->n_tile, pl->low_on = 0, pl->low_w_off = 0, pl->low_w_bytes = 0;
  if (a->lowres_src && !a->block2x2) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: lowres_src (fused transposed conv) needs block2x2");
  if (a->bias_classes != 0 && a->bias_classes != 1 && a->bias_classes != 9) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: bias_classes must be 0, 1 or 9");
  // static shared memory: 18 KB in the training variants (statistics slots), 2 KB otherwise; 227 KB per CTA in total
  const int smem_budget = ((is_train(a) && !a->head_w) ? 196 : 220) * 1024;
  if (a->block2x2) {
    // 2x2 output blocking: every source is one 16-channel chunk staged as 64-byte pixel-pair rows; N = 4 x 16
    if (a->taps != 9 || a->mode != UNPP_MODE_CONV || a->n_total != 16 || a->n_tile != 16 || (a->H & 1) || (a->W & 1))
      return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: block2x2 needs a 3x3 conv with n_total = n_tile = 16 and even H, W");
    const bool c4 = a->block2x2 == 2;
    if (c4 && (a->nsrc != 1 || a->src_C[0] != 4 || a->src_step[0] == 2 || a->lowres_src || is_train(a)))
      return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: block2x2 = 2 (first layer) needs one dense 4-channel source and the inference epilogue");
    for (int i = 0; i < a->nsrc && !c4; ++i) {
      if (a->src_C[i] != 16 || a->src_step[i] == 2) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: block2x2 needs dense 16-channel sources");
      pl->ch_span[i] = 64, pl->ch_wk8[i] = 2 * i;
    }
    pl->b2 = 1, pl->c4 = c4, pl->ncols = 64;
    pl->w_bytes = c4 ? 16384 : (k8 / 2) * b2::kMainUnits * b2::kUnitBytes;  // only the non-zero (position, pixel) blocks are stored (b2_blocks.h)
    pl->w_smem_bytes = (pl->w_bytes + 1023) / 1024 * 1024;
    if (a->lowres_src) {
      if (a->lowres_C != 32 || !a->lowres_wpacked || (a->H & 3) || (a->W & 3) || is_train(a) || (reinterpret_cast<uintptr_t>(a->lowres_src) & 15))
        return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: fused transposed conv needs a 32-channel low-res source, its composed weights, H and W "
                                            "divisible by 4, and the inference epilogue");
      pl->low_on = 1, pl->low_w_off = pl->w_smem_bytes, pl->low_w_bytes = (a->lowres_C / 16) * b2::kLowUnits * b2::kUnitBytes;
      pl->w_smem_bytes += (pl->low_w_bytes + 1023) / 1024 * 1024;
    }
    // widest tile that leaves room for two stages: fewer tiles (less per-tile handshake), more sub-tiles (more issuers busy)
    int TW = a->W > 32 ? 64 : a->W > 16 ? 32 : 16;
    {
      static const int cap = [] { const char* e = getenv("UNPP_B2_TW"); return e ? atoi(e) : 0; }();  // experiment knob: cap the tile width (read once)
      if ((cap == 16 || cap == 32) && TW > cap) TW = cap;
    }
    for (;; TW >>= 1) {
      pl->b2_P = TW / 2 + (c4 ? 3 : 2);
      pl->stage_bytes = (34 * pl->b2_P * (c4 ? 16 : 64) + 1023) / 1024 * 1024;
      pl->nstage = (smem_budget - pl->w_smem_bytes) / pl->stage_bytes;
      if (pl->nstage >= 2 || TW == 16) break;
    }
    if (pl->nstage < 2) return unpp::fail(UNPP_ERR_UNSUPPORTED, "conv_tc: block2x2 weights leave no room for 2 stages (too many sources)");
    if (pl->nstage > kMaxStages) pl->nstage = kMaxStages;
    pl->TW = TW, pl->nsub = TW / 16;
    pl->nacc = 4 * pl->nsub * 64 <= 512 ? 4 : 2;
    pl->tmem_cols = pl->nacc * pl->nsub * 64;
    pl->smem_total = 1024 + pl->w_smem_bytes + pl->nstage * pl->stage_bytes;
    pl->tiles_x = (a->W + TW - 1) / TW, pl->tiles_y = (a->H + 31) / 32;
    pl->ntiles = pl->tiles_x * pl->tiles_y * a->N;
    pl->grid_y = 1;
    int gx = unpp::num_sms();
    if (gx > pl->ntiles) gx = pl->ntiles;
    pl->grid_x = gx;
    return UNPP_OK;
  }
  const int pad = a->taps == 9 ? 1 : 0;
  pl->w_bytes = a->taps * k8 * a->n_tile * 16;
  pl->w_smem_bytes = (pl->w_bytes + 1023) / 1024 * 1024;
  int TW = 64;
  {
    static const int cap = [] { const char* e = getenv("UNPP_TW"); return e ? atoi(e) : 0; }();  // experiment knob: cap the tile width (read once)
    if (cap == 16 || cap == 32) TW = cap;
  }
  while (TW > 8 && (2 * (TW / 8) * a->n_tile > 512 || TW / 2 >= ((a->W + 7) / 8) * 8)) TW >>= 1;
  for (;; TW >>= 1) {
    pl->stage_bytes = ((16 + 2 * pad) * (TW + 2 * pad) * max_span + 1023) / 1024 * 1024;
    pl->nstage = (smem_budget - pl->w_smem_bytes) / pl->stage_bytes;
    if (pl->nstage >= 2 || TW == 8) break;
  }
  if (pl->nstage < 2) return unpp::fail(UNPP_ERR_UNSUPPORTED, "conv_tc: packed weights of one n_tile do not leave room for 2 stages; lower n_tile");
  if (2 * (TW / 8) * a->n_tile > 512) return unpp::fail(UNPP_ERR_UNSUPPORTED, "conv_tc: n_tile too large for TMEM double buffering");
  if (pl->nstage > kMaxStages) pl->nstage = kMaxStages;
  pl->TW = TW, pl->nsub = TW / 8;
  pl->nacc = 4 * pl->nsub * a->n_tile <= 512 ? 4 : 2;
  int cols = pl->nacc * pl->nsub * a->n_tile, tc = 32;
  while (tc < cols) tc <<= 1;
  pl->tmem_cols = tc;
  pl->smem_total = 1024 + pl->w_smem_bytes + pl->nstage * pl->stage_bytes;
  pl->tiles_x = (a->W + TW - 1) / TW, pl->tiles_y = (a->H + 15) / 16;
  pl->ntiles = pl->tiles_x * pl->tiles_y * a->N;
  pl->grid_y = a->n_total / a->n_tile;
  int gx = unpp::num_sms() / pl->grid_y;
  if (gx < 1) gx = 1;
  if (gx > pl->ntiles) gx = pl->ntiles;
  pl->grid_x = gx;
  return UNPP_OK;
}

}  // namespace

extern "C" int unpp_conv_grid(const UnppConvArgs* a) {
  Plan pl;
  int rc = make_plan(a, &pl);
  return rc ? rc : pl.grid_x;
}

extern "C" int unpp_conv_tc(const UnppConvArgs* a, unpp_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Plan pl;
  if (int rc = make_plan(a, &pl)) return rc;
  if (!a->wpacked) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: wpacked is null");
  if (a->mode != UNPP_MODE_CONV && a->mode != UNPP_MODE_DECONV) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: bad mode");
  if (a->mode == UNPP_MODE_DECONV && (a->taps != 1 || a->n_total % 4 || (a->n_total / 4) % 16 || !a->out))
    return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: deconv mode needs taps=1, n_total=4*Cout with Cout%16==0 and an output");
  if (a->head_w && (a->mode != UNPP_MODE_CONV || a->n_total != 16 || a->n_tile != 16 || !a->heat || !a->head_b || a->head_classes < 1 ||
                    a->head_classes > 8))
    return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: fused head needs conv mode, n_total=n_tile=16, heat/head_b and 1..8 classes");
  if (!a->out && !a->head_w) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: nothing to write");
  if (a->pooled && (a->mode != UNPP_MODE_CONV || !a->relu || !a->out || a->head_w || is_train(a) || (a->H & 1) || (a->W & 1) ||
                    (!a->block2x2 && a->n_tile != 16 && a->n_tile != 32 && a->n_tile != 64 && a->n_tile != 128)))
    return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: the fused max pool needs conv mode with ReLU and an output, even H and W, no head / training operand");
  if (a->stats_partial && a->mode != UNPP_MODE_CONV) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: stats only in conv mode");
  if (a->head_w && (a->addend || a->relu_mask_src || a->stats_partial))
    return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: the fused head (a forward op) takes no addend / ReLU mask / statistics");
  if (a->stats_aux && (!a->aux_mean || !a->aux_istd)) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: stats_aux needs aux_mean/aux_istd");

  EncodeTiledFn enc = get_encode();
  if (!enc) return unpp::fail(UNPP_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled not available from the driver");

  ConvTcParams p;
  memset(&p, 0, sizeof p);
  const int pad = a->taps == 9 ? 1 : 0;
  for (int i = 0; i < a->nsrc; ++i) {
    const cuuint64_t C = a->src_C[i];
    const int box_c = C < cuuint64_t(pl.cw) ? int(C) : pl.cw;
    const cuuint64_t st = a->src_step[i] == 2 ? 2 : 1, fullW = cuuint64_t(a->W) * st, fullH = cuuint64_t(a->H) * st;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a->src[i]);
    if (st == 2) base += (size_t(a->src_oy[i]) * fullW + a->src_ox[i]) * C * 2;
    cuuint64_t gd[4] = {C, cuuint64_t(a->W), cuuint64_t(a->H), cuuint64_t(a->N)};
    cuuint64_t gs[3] = {C * 2 * st, fullW * C * 2 * st, fullH * fullW * C * 2};
    cuuint32_t box[4] = {cuuint32_t(box_c), cuuint32_t(pl.TW + 2 * pad), cuuint32_t(16 + 2 * pad), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMapSwizzle sw = box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    if (pl.b2) {  // the [N,H,W,16] tensor seen as [N,H,W/2,32]: one row = a horizontal pixel pair (64 B)
      gd[0] = 32, gd[1] = cuuint64_t(a->W / 2);
      gs[0] = 64;
      box[0] = 32, box[1] = cuuint32_t(pl.b2_P), box[2] = 34;
      sw = CU_TENSOR_MAP_SWIZZLE_64B;
    }
    if (pl.c4) {  // first layer: [N,H,W,4] seen as [N,H,W/2,8], a pixel pair = 16 B, dense unswizzled rows in shared memory
      gd[0] = 8, gs[0] = 16, box[0] = 8;
      sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    }
    CUresult r = enc(&p.maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<uint8_t*>(base), gd, gs, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return unpp::fail(UNPP_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled failed (CUresult %d) for source %d", int(r), i);
  }
  p.nchunk = pl.nchunk;
  for (int c = 0; c < pl.nchunk; ++c) p.ch_map[c] = pl.ch_map[c], p.ch_c0[c] = pl.ch_c0[c], p.ch_span[c] = pl.ch_span[c], p.ch_wk8[c] = pl.ch_wk8[c];
  p.N = a->N, p.H = a->H, p.W = a->W;
  p.TW = pl.TW, p.nsub = pl.nsub, p.tiles_x = pl.tiles_x, p.tiles_y = pl.tiles_y, p.ntiles = pl.ntiles;
  p.taps = a->taps, p.ncols = pl.ncols, p.k8_total = pl.k8_total;
  p.b2 = pl.b2, p.b2_P = pl.b2_P, p.c4 = pl.c4, p.nacc = pl.nacc;
  p.bias9 = a->bias_classes == 9;
  p.pooled = reinterpret_cast<__nv_bfloat16*>(a->pooled);
  if (pl.low_on) {
    const cuuint64_t C = a->lowres_C, lw = a->W / 2, lh = a->H / 2;
    cuuint64_t gd[4] = {C, lw, lh, cuuint64_t(a->N)};
    cuuint64_t gs[3] = {C * 2, lw * C * 2, lh * lw * C * 2};
    cuuint32_t box[4] = {cuuint32_t(C), cuuint32_t(pl.TW / 2 + 2), 18, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.lowmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->lowres_src), gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return unpp::fail(UNPP_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled failed (CUresult %d) for the low-res source", int(r));
    p.low_on = 1, p.low_P = pl.TW / 2 + 2, p.low_k8 = a->lowres_C / 8, p.low_w_off = pl.low_w_off, p.low_w_bytes = pl.low_w_bytes;
    p.low_wpacked = reinterpret_cast<const __nv_bfloat16*>(a->lowres_wpacked);
  }
  p.stage_bytes = pl.stage_bytes, p.nstage = pl.nstage, p.w_bytes = pl.w_bytes, p.w_smem_bytes = pl.w_smem_bytes;
  p.tmem_cols = pl.tmem_cols;
  p.wpacked = reinterpret_cast<const __nv_bfloat16*>(a->wpacked);
  p.mode = a->mode, p.relu = a->relu;
  p.cout = a->mode == UNPP_MODE_DECONV ? a->n_total / 4 : a->n_total;
  p.bias = a->bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.head_w = a->head_w, p.head_b = a->head_b, p.heat = a->heat, p.logit = a->logit, p.head_classes = a->head_classes;
  p.drop_mask = a->drop_mask, p.drop_scale = a->drop_scale;
  p.addend = reinterpret_cast<const __nv_bfloat16*>(a->addend);
  p.relu_mask_src = reinterpret_cast<const __nv_bfloat16*>(a->relu_mask_src);
  p.stats_partial = a->stats_partial;
  p.stats_aux = reinterpret_cast<const __nv_bfloat16*>(a->stats_aux);
  p.aux_mean = a->aux_mean, p.aux_istd = a->aux_istd;
  p.pf = [] { const char* d = getenv("UNPP_PF"); return d ? atoi(d) : 0; }();  // L2 prefetch distance of the producer (experiment knob)
  p.dbg = [] { const char* d = getenv("UNPP_DBG"); return d ? atoi(d) : 0; }();  // role-disabling experiment bits (scripts/dbg_conv*.py set it per process run)

  const bool deconv = a->mode == UNPP_MODE_DECONV, head = a->head_w != nullptr;
  const bool train = is_train(a);
  if (deconv && train) return unpp::fail(UNPP_ERR_BAD_ARG, "conv_tc: deconv mode has no training epilogue");
  if (deconv && a->n_total / 4 > 256) return unpp::fail(UNPP_ERR_UNSUPPORTED, "conv_tc: deconv Cout > 256");
  const dim3 grid(pl.grid_x, pl.grid_y);
#define UNPP_LAUNCH(D, Hd, T)                                                                                                         \
  do {                                                                                                                                \
    static unsigned char opted_in[64] = {0}; /* per variant and device */                                                             \
    if (cudaError_t e = unpp::opt_in_smem(conv_tc_kernel<D, Hd, T>, (heavy_epilogue(Hd, T) ? 204 : 224) * 1024, opted_in))            \
      return unpp::fail_cuda_err("conv_tc: cudaFuncSetAttribute", e);                                                                 \
    if (cudaError_t e = unpp::launch(conv_tc_kernel<D, Hd, T>, grid, block_threads(heavy_epilogue(Hd, T)), pl.smem_total, stream, p)) \
      return unpp::fail_cuda_err("conv_tc: launch", e);                                                                               \
  } while (0)
  if (deconv) UNPP_LAUNCH(true, false, false);
  else if (head && train) UNPP_LAUNCH(false, true, true);
  else if (head) UNPP_LAUNCH(false, true, false);
  else if (train) UNPP_LAUNCH(false, false, true);
  else UNPP_LAUNCH(false, false, false);
#undef UNPP_LAUNCH
  return UNPP_OK;
}

// ref_kernels.cu — fp32 validation mode of the inference path (UNet_Nested.precision = "fp32").
//
// The product path stores activations in bf16 and multiplies on the tensor cores; its stated bound against the fp32 reference
// is the bf16 bound (heat maps 3e-2).  BASELINE's north_star also names a tolerance for an fp32 / TF32 mode (<= 1e-3 relative):
// these kernels are that mode — the same fused plan (virtual concat through a source list, eval-mode BatchNorm as per-channel
// scale / bias, ReLU, the k2s2 transposed conv, 2x2 max pool, 1x1 head + sigmoid) on NCHW fp32 tensors with plain fp32 FMAs on the
// CUDA cores, straight from the OIHW state_dict weights (no packing, no rounding).  They exist to separate wiring errors from
// rounding when a bf16 result looks suspicious; they are ~30x slower than the tensor-core path and no benchmark runs them.
// Reference arithmetic: models/unet.py:121-156 (unetConv2), 182-202 (unetUp), 242-244 / 283-286 (heads).
#include "common.h"
#include "../../include/unpp.h"
#include <math.h>
#include <stdint.h>

namespace {

struct RefConvParams {
  int N, H, W, nsrc, src_C[UNPP_MAX_SRC], ctot, cout, taps, relu, sigmoid;
  const float* src[UNPP_MAX_SRC];
  const float* weight;  // [cout][ctot][k][k]
  const float* scale;   // per output channel or null
  const float* bias;    // per output channel or null
  float* out;           // [N][cout][H][W]
};

// One thread per output element (n, co, y, x); x is the fastest index: source reads are coalesced along image rows.
__global__ void __launch_bounds__(256) ref_conv_kernel(const RefConvParams p) {
  const long total = long(p.N) * p.cout * p.H * p.W;
  const int k = p.taps == 9 ? 3 : 1, pad = k >> 1;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int x = int(i % p.W), y = int((i / p.W) % p.H), co = int((i / (long(p.W) * p.H)) % p.cout), n = int(i / (long(p.W) * p.H * p.cout));
    float acc = 0.f;
    int cbase = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      const int C = p.src_C[s];
      const float* xs = p.src[s] + size_t(n) * C * p.H * p.W;
      for (int ci = 0; ci < C; ++ci) {
        const float* w = p.weight + (size_t(co) * p.ctot + cbase + ci) * k * k;
        const float* xc = xs + size_t(ci) * p.H * p.W;
        for (int r = 0; r < k; ++r) {
          const int yy = y + r - pad;
          if (yy < 0 || yy >= p.H) continue;
          for (int t = 0; t < k; ++t) {
            const int xx = x + t - pad;
            if (xx < 0 || xx >= p.W) continue;
            acc = fmaf(__ldg(xc + size_t(yy) * p.W + xx), __ldg(w + r * k + t), acc);
          }
        }
      }
      cbase += C;
    }
    if (p.scale) acc *= __ldg(p.scale + co);
    if (p.bias) acc += __ldg(p.bias + co);
    if (p.relu) acc = fmaxf(acc, 0.f);
    if (p.sigmoid) acc = 1.f / (1.f + expf(-acc));
    p.out[i] = acc;
  }
}

// ConvTranspose2d(k=2, s=2): out[n][co][2y+p][2x+q] = b[co] + sum_ci x[n][ci][y][x] * w[ci][co][p][q]   (models/unet.py:187)
__global__ void __launch_bounds__(256) ref_deconv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ out,
                                                         int N, int Cin, int Cout, int H, int W) {
  const int Ho = 2 * H, Wo = 2 * W;
  const long total = long(N) * Cout * Ho * Wo;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int xo = int(i % Wo), yo = int((i / Wo) % Ho), co = int((i / (long(Wo) * Ho)) % Cout), n = int(i / (long(Wo) * Ho * Cout));
    const int y = yo >> 1, pp = yo & 1, xx = xo >> 1, qq = xo & 1;
    float acc = __ldg(b + co);
    for (int ci = 0; ci < Cin; ++ci)
      acc = fmaf(__ldg(x + ((size_t(n) * Cin + ci) * H + y) * W + xx), __ldg(w + ((size_t(ci) * Cout + co) * 2 + pp) * 2 + qq), acc);
    out[i] = acc;
  }
}

__global__ void __launch_bounds__(256) ref_maxpool_kernel(const float* __restrict__ x, float* __restrict__ out, long planes, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long total = planes * Ho * Wo;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int xo = int(i % Wo), yo = int((i / Wo) % Ho);
    const long pl = i / (long(Wo) * Ho);
    const float* r = x + (pl * H + 2 * yo) * W + 2 * xo;
    out[i] = fmaxf(fmaxf(r[0], r[1]), fmaxf(r[W], r[W + 1]));
  }
}

inline int grid_of(long total) {
  long g = (total + 255) / 256;
  const long cap = long(unpp::num_sms()) * 32;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int unpp_ref_conv(const UnppRefConvArgs* a, unpp_stream_t stream) {
  if (!a || a->nsrc < 1 || a->nsrc > UNPP_MAX_SRC || !a->weight || !a->out || a->N < 1 || a->H < 1 || a->W < 1 || a->cout < 1 || (a->taps != 9 && a->taps != 1))
    return unpp::fail(UNPP_ERR_BAD_ARG, "ref_conv: bad argument");
  RefConvParams p;
  memset(&p, 0, sizeof p);
  p.N = a->N, p.H = a->H, p.W = a->W, p.nsrc = a->nsrc, p.cout = a->cout, p.taps = a->taps, p.relu = a->relu, p.sigmoid = a->sigmoid;
  for (int i = 0; i < a->nsrc; ++i) {
    if (!a->src[i] || a->src_C[i] < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "ref_conv: bad source %d", i);
    p.src[i] = static_cast<const float*>(a->src[i]), p.src_C[i] = a->src_C[i], p.ctot += a->src_C[i];
  }
  p.weight = a->weight, p.scale = a->scale, p.bias = a->bias, p.out = a->out;
  unpp::launch(ref_conv_kernel, grid_of(long(a->N) * a->cout * a->H * a->W), 256, 0, reinterpret_cast<cudaStream_t>(stream), p);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("ref_conv: launch");
  return UNPP_OK;
}

extern "C" int unpp_ref_deconv2x2(const float* x, const float* w, const float* b, float* out, int N, int Cin, int Cout, int H, int W, unpp_stream_t stream) {
  if (!x || !w || !b || !out || N < 1 || Cin < 1 || Cout < 1 || H < 1 || W < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "ref_deconv2x2: bad argument");
  unpp::launch(ref_deconv_kernel, grid_of(long(N) * Cout * 4 * H * W), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, w, b, out, N, Cin, Cout, H, W);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("ref_deconv2x2: launch");
  return UNPP_OK;
}

extern "C" int unpp_ref_maxpool2x2(const float* x, float* out, int N, int C, int H, int W, unpp_stream_t stream) {
  if (!x || !out || N < 1 || C < 1 || H < 2 || W < 2 || (H & 1) || (W & 1)) return unpp::fail(UNPP_ERR_BAD_ARG, "ref_maxpool2x2: bad argument");
  unpp::launch(ref_maxpool_kernel, grid_of(long(N) * C * (H / 2) * (W / 2)), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, out, long(N) * C, H, W);
  if (cudaGetLastError() != cudaSuccess) return unpp::fail_cuda("ref_maxpool2x2: launch");
  return UNPP_OK;
}

extern "C" int unpp_sizeof_ref_conv_args(void) { return int(sizeof(UnppRefConvArgs)); }

// umma_probe.cu — standalone hardware-semantics probe for sm_100a (not part of the product).
//
// The production conv kernels feed tcgen05.mma from shared-memory tiles through *shifted*
// matrix descriptors (one halo tile serves all nine filter taps).  This program pins down,
// on a real B200, the descriptor / TMA conventions those kernels rely on, and measures the
// two rates that bound them (TMA fill rate for narrow boxes, SS-mode MMA rate at small N).
//
//   ./umma_probe <test>      test in: m1 m1s m2 m3 m3s m4 m5 m6 m7 t1 t2 t3 t4 p1 p2
// Every test prints "RESULT <name> ... PASS|FAIL" lines; each runs in its own process so a
// trap in one cannot poison the others.
#include "../unet_nested4tiny_objects_keypoints_b200/csrc/sm100.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <functional>

using namespace sm100;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

static inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
  return uint16_t(r >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = uint32_t(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint32_t rng_state = 12345;
static inline float frand() {  // small integers / 8 so that bf16 products and fp32 sums are exact
  rng_state = rng_state * 1664525u + 1013904223u;
  return float(int((rng_state >> 20) % 17) - 8) / 8.0f;
}

// ------------------------------------------------------------------------------------------
// MMA executor: copy a host-prepared smem image in, issue the listed MMAs, dump the accumulator.
struct MmaJob {
  uint64_t adesc, bdesc;  // start-address field is relative to the image base; kernel adds base
  uint32_t accumulate;
  uint32_t pad;
};

__global__ void __launch_bounds__(128, 1) k_mma(const uint8_t* __restrict__ image, int image_bytes, const MmaJob* __restrict__ jobs,
                                                int njobs, uint32_t idesc, int ncols, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x * 16; i < image_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  const uint32_t base = smem_u32(smem);
  if (threadIdx.x == 0) {
    if (base & 1023) printf("NOTE dynamic smem base %u is not 1024-aligned\n", base);
    for (int j = 0; j < njobs; ++j) {
      uint64_t a = jobs[j].adesc + uint64_t((base >> 4) & 0x3FFF);
      uint64_t b = jobs[j].bdesc + uint64_t((base >> 4) & 0x3FFF);
      umma_bf16(tb, a, b, idesc, jobs[j].accumulate);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tb + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * ncols + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

static bool run_mma(const char* name, const std::vector<uint8_t>& image, const std::vector<MmaJob>& jobs, uint32_t idesc, int ncols,
                    const std::vector<float>& expect /*[128][ncols]*/) {
  uint8_t* dimg;
  MmaJob* djobs;
  float* dout;
  size_t ib = (image.size() + 15) / 16 * 16;
  CK(cudaMalloc(&dimg, ib));
  CK(cudaMemset(dimg, 0, ib));
  CK(cudaMemcpy(dimg, image.data(), image.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&djobs, jobs.size() * sizeof(MmaJob)));
  CK(cudaMemcpy(djobs, jobs.data(), jobs.size() * sizeof(MmaJob), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dout, 128 * ncols * 4));
  CK(cudaMemset(dout, 0xFF, 128 * ncols * 4));
  CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k_mma<<<1, 128, 200 * 1024>>>(dimg, int(ib), djobs, int(jobs.size()), idesc, ncols, dout);
  CK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("RESULT %s launch error: %s FAIL\n", name, cudaGetErrorString(e));
    return false;
  }
  std::vector<float> got(128 * ncols);
  CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    double d = fabs(double(got[i]) - double(expect[i]));
    if (!(d <= 1e-3)) {
      if (bad < 4) printf("  %s mismatch at row %zu col %zu: got %g expect %g\n", name, i / ncols, i % ncols, got[i], expect[i]);
      ++bad;
    }
    if (d > maxerr || d != d) maxerr = d;
  }
  printf("RESULT %s maxerr=%g bad=%d/%zu %s\n", name, maxerr, bad, got.size(), bad == 0 ? "PASS" : "FAIL");
  cudaFree(dimg);
  cudaFree(djobs);
  cudaFree(dout);
  return bad == 0;
}

static inline void put_bf(std::vector<uint8_t>& img, size_t byte_off, float v) {
  if (img.size() < byte_off + 2) img.resize(byte_off + 2, 0);
  uint16_t h = f2bf(v);
  memcpy(&img[byte_off], &h, 2);
}

// m1: one 128xNx16 MMA, K-major no-swizzle, dense core matrices.  swap=true exchanges LBO/SBO.
static void test_m1(bool swap) {
  const int N = 16, K = 16;
  std::vector<float> A(128 * K), B(N * K);
  for (auto& v : A) v = frand();
  for (auto& v : B) v = frand();
  std::vector<uint8_t> img;
  // A at 0: addr(m,k) = (m%8)*16 + (m/8)*128 + (k/8)*2048 + (k%8)*2  -> SBO=128, LBO=2048
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < K; ++k) put_bf(img, (m % 8) * 16 + (m / 8) * 128 + (k / 8) * 2048 + (k % 8) * 2, A[m * K + k]);
  // B at 4096: addr(n,k) = (n%8)*16 + (n/8)*128 + (k/8)*256 + (k%8)*2 -> SBO=128, LBO=256
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) put_bf(img, 4096 + (n % 8) * 16 + (n / 8) * 128 + (k / 8) * 256 + (k % 8) * 2, B[n * K + k]);
  std::vector<float> exp(128 * N);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += bf2f(f2bf(A[m * K + k])) * bf2f(f2bf(B[n * K + k]));
      exp[m * N + n] = s;
    }
  MmaJob j{};
  j.adesc = swap ? make_sdesc(0, 128, 2048) : make_sdesc(0, 2048, 128);
  j.bdesc = swap ? make_sdesc(4096, 128, 256) : make_sdesc(4096, 256, 128);
  j.accumulate = 0;
  run_mma(swap ? "m1s(kmajor,LBO/SBO swapped)" : "m1(kmajor,LBO=K-chunk stride,SBO=8-row stride)", img, {j}, make_idesc_bf16(128, N), N, exp);
}

// m2: the production scheme. Planar halo tile [plane][row][col][8ch] (16 B per pixel per plane),
// output patch 16 rows x 8 cols, nine taps selected purely through the A start address,
// SBO = one tile row, LBO = one plane.  Cin = 32 (two K=16 slabs), Cout = 32.
static void test_m2() {
  const int Cin = 32, Cout = 32, R = 16, TW = 16, P = TW + 2, ROWS = R + 2;
  const int plane_bytes = ROWS * P * 16 + 16;  // odd multiple of 16 on purpose
  std::vector<float> X(ROWS * P * Cin), Wt(9 * Cout * Cin);
  for (auto& v : X) v = frand();
  for (auto& v : Wt) v = frand();
  std::vector<uint8_t> img;
  for (int y = 0; y < ROWS; ++y)
    for (int x = 0; x < P; ++x)
      for (int c = 0; c < Cin; ++c) put_bf(img, (c / 8) * plane_bytes + (y * P + x) * 16 + (c % 8) * 2, X[(y * P + x) * Cin + c]);
  const int wbase = ((Cin / 8) * plane_bytes + 127) / 128 * 128;
  // weights: [tap][kchunk8][Cout][8] -> SBO=128 (8 couts), LBO = Cout*16
  for (int t = 0; t < 9; ++t)
    for (int co = 0; co < Cout; ++co)
      for (int c = 0; c < Cin; ++c)
        put_bf(img, wbase + ((t * (Cin / 8) + c / 8) * Cout + co) * 16 + (c % 8) * 2, Wt[(t * Cout + co) * Cin + c]);
  const int px0 = 8, py0 = 0;  // second 8-wide patch of the tile
  std::vector<MmaJob> jobs;
  for (int t = 0; t < 9; ++t)
    for (int ks = 0; ks < Cin / 16; ++ks) {
      int r = t / 3, s = t % 3;
      MmaJob j{};
      j.adesc = make_sdesc((2 * ks) * plane_bytes + ((py0 + r) * P + px0 + s) * 16, plane_bytes, P * 16);
      j.bdesc = make_sdesc(wbase + (t * (Cin / 8) + 2 * ks) * Cout * 16, Cout * 16, 128);
      j.accumulate = jobs.empty() ? 0 : 1;
      jobs.push_back(j);
    }
  std::vector<float> exp(128 * Cout);
  for (int m = 0; m < 128; ++m)
    for (int co = 0; co < Cout; ++co) {
      int oy = py0 + m / 8, ox = px0 + m % 8;
      float sum = 0;
      for (int t = 0; t < 9; ++t)
        for (int c = 0; c < Cin; ++c) sum += X[((oy + t / 3) * P + ox + t % 3) * Cin + c] * Wt[(t * Cout + co) * Cin + c];
      exp[m * Cout + co] = sum;
    }
  run_mma("m2(planar halo tile, 16x8 patch, 9 shifted taps)", img, jobs, make_idesc_bf16(128, Cout), Cout, exp);
}

// m3: MN-major, no swizzle, both operands (the wgrad scheme): D[ci,co] = sum_pix X[pix,ci]*G[pix,co].
// X planar [16 planes][pixels][8], G planar [2 planes][pixels][8]; K = 16 pixels = 2 groups of 8.
static void test_m3(bool swap) {
  const int M = 128, N = 16, NPIX = 64;
  const int xplane = NPIX * 16 + 16, gplane = NPIX * 16 + 48;
  std::vector<float> X(NPIX * M), G(NPIX * N);
  for (auto& v : X) v = frand();
  for (auto& v : G) v = frand();
  std::vector<uint8_t> img;
  for (int p = 0; p < NPIX; ++p)
    for (int c = 0; c < M; ++c) put_bf(img, (c / 8) * xplane + p * 16 + (c % 8) * 2, X[p * M + c]);
  const int gbase = (16 * xplane + 127) / 128 * 128;
  for (int p = 0; p < NPIX; ++p)
    for (int c = 0; c < N; ++c) put_bf(img, gbase + (c / 8) * gplane + p * 16 + (c % 8) * 2, G[p * N + c]);
  std::vector<MmaJob> jobs;
  const int p0 = 3;  // unaligned pixel start on purpose
  for (int ks = 0; ks < 2; ++ks) {  // 32 pixels
    MmaJob j{};
    uint32_t aoff = (p0 + 16 * ks) * 16, boff = gbase + (p0 + 16 * ks) * 16;
    // hypothesis: LBO = stride between 8-K-row groups (128 B), SBO = stride between 8-element MN chunks (plane)
    j.adesc = swap ? make_sdesc(aoff, xplane, 128) : make_sdesc(aoff, 128, xplane);
    j.bdesc = swap ? make_sdesc(boff, gplane, 128) : make_sdesc(boff, 128, gplane);
    j.accumulate = ks;
    jobs.push_back(j);
  }
  std::vector<float> exp(M * N);
  for (int ci = 0; ci < M; ++ci)
    for (int co = 0; co < N; ++co) {
      float s = 0;
      for (int p = p0; p < p0 + 32; ++p) s += X[p * M + ci] * G[p * N + co];
      exp[ci * N + co] = s;
    }
  run_mma(swap ? "m3s(mnmajor,LBO/SBO swapped)" : "m3(mnmajor,LBO=8-K-row group stride,SBO=MN chunk stride)", img, jobs,
          make_idesc_bf16(M, N, 1, 1), N, exp);
}

// Swizzled K-major tiles, pixel-major rows.  span = bytes per row (32 / 128); XOR pattern applied
// on the absolute byte offset inside a 1024-aligned smem image: 128B: bits[4,7) ^= bits[7,10);
// 32B: bit4 ^= bit7.
static inline uint32_t swz(uint32_t a, int span) {
  if (span == 128) return a ^ (((a >> 7) & 7) << 4);
  if (span == 64) return a ^ (((a >> 7) & 3) << 4);
  if (span == 32) return a ^ (((a >> 7) & 1) << 4);
  return a;
}
// m4..m7: rows = pixels of a [ROWS x P] tile, each row `span` bytes (span/2 channels).
// mode 0: dense, aligned start, SBO = 8*span (the textbook case)
// mode 1: start shifted by `shift` rows, SBO = 8*span, base_offset = bo
// mode 2: 16x8 patch mapping: SBO = P*span with P=34, start shifted, base_offset = bo
static void test_swz(const char* name, int span, int mode, int shift, int bo_mode) {
  const int C = span / 2, N = 16, P = 34, ROWS = 18;
  const int npix = ROWS * P;
  std::vector<float> X(npix * C), Wt(N * C);
  for (auto& v : X) v = frand();
  for (auto& v : Wt) v = frand();
  std::vector<uint8_t> img;
  for (int p = 0; p < npix; ++p)
    for (int c = 0; c < C; ++c) put_bf(img, swz(p * span + c * 2, span), X[p * C + c]);
  const int wbase = (npix * span + 1023) / 1024 * 1024;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) put_bf(img, wbase + swz(n * span + c * 2, span), Wt[n * C + c]);
  const uint32_t lt = span == 128 ? 2 : span == 64 ? 4 : 6;
  std::vector<MmaJob> jobs;
  uint32_t start = (mode == 0 ? 0 : shift) * span;
  uint32_t sbo = (mode == 2 ? P : 8) * span;
  uint32_t bo = bo_mode ? ((start >> 7) & 7) : 0;
  for (int ks = 0; ks < C / 16; ++ks) {
    MmaJob j{};
    j.adesc = make_sdesc(start + ks * 32, 16, sbo, lt, bo);
    j.bdesc = make_sdesc(wbase + ks * 32, 16, 8 * span, lt, 0);
    j.accumulate = ks;
    jobs.push_back(j);
  }
  std::vector<float> exp(128 * N);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      int p = (mode == 2) ? shift + (m / 8) * P + (m % 8) : (mode == 1 ? shift + m : m);
      float s = 0;
      for (int c = 0; c < C; ++c) s += X[p * C + c] * Wt[n * C + c];
      exp[m * N + n] = s;
    }
  run_mma(name, img, jobs, make_idesc_bf16(128, N), N, exp);
}

// ------------------------------------------------------------------------------------------
// TMA tests
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) {
    printf("no cuTensorMapEncodeTiled\n");
    exit(2);
  }
  return (EncodeTiledFn)fn;
}

__global__ void __launch_bounds__(128, 1) k_tma(const __grid_constant__ CUtensorMap map, int rank, int c0, int c1, int c2, int c3, int c4,
                                                int dst_off, int bytes, uint8_t* __restrict__ out, int dump_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < dump_bytes; i += blockDim.x) smem[i] = 0xEE;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    if (smem_u32(smem) & 1023) printf("NOTE dynamic smem base %u is not 1024-aligned\n", smem_u32(smem));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    if (rank == 4)
      tma_load_4d(&map, &bar, smem + dst_off, c0, c1, c2, c3);
    else
      tma_load_5d(&map, &bar, smem + dst_off, c0, c1, c2, c3, c4);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < dump_bytes; i += blockDim.x) out[i] = smem[i];
}

struct Tens {  // NHWC bf16 host+device tensor with value = f(n,y,x,c) exactly representable
  int N, H, W, C;
  std::vector<uint16_t> h;
  uint16_t* d;
  float at(int n, int y, int x, int c) const {
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.f;
    return bf2f(h[((size_t(n) * H + y) * W + x) * C + c]);
  }
};
static Tens make_tens(int N, int H, int W, int C) {
  Tens t{N, H, W, C, {}, nullptr};
  t.h.resize(size_t(N) * H * W * C);
  for (auto& v : t.h) v = f2bf(frand() * 8 + 0.5f);
  CK(cudaMalloc(&t.d, t.h.size() * 2));
  CK(cudaMemcpy(t.d, t.h.data(), t.h.size() * 2, cudaMemcpyHostToDevice));
  return t;
}

static void test_tma(const char* name, int variant) {
  EncodeTiledFn enc = get_encode();
  const int BX = 10, BY = 6;
  CUtensorMap map;
  int rank, bytes, dst_off = 0;
  int c[5] = {0, 0, 0, 0, 0};
  Tens t;
  std::function<float(int)> expect;  // expected bf16 value at smem element index (relative to dst), before any swizzle
  int span = 0;                      // swizzle span in bytes, 0 = none
  CUresult r;
  if (variant == 1) {  // 4-D, box inner 8 channels, no swizzle
    t = make_tens(2, 20, 24, 16);
    cuuint64_t gd[4] = {16, 24, 20, 2};
    cuuint64_t gs[3] = {16 * 2, 24 * 16 * 2, 20 * 24 * 16 * 2};
    cuuint32_t box[4] = {8, BX, BY, 1}, es[4] = {1, 1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    rank = 4;
    c[0] = 8, c[1] = -1, c[2] = -1, c[3] = 1;
    bytes = 8 * BX * BY * 2;
    expect = [=](int e) { int ch = e % 8, x = (e / 8) % BX, y = e / 8 / BX; return t.at(1, y - 1, x - 1, 8 + ch); };
  } else if (variant == 2) {  // 5-D with the channel-chunk as dim 3 (stride 16 B)
    t = make_tens(2, 20, 24, 16);
    cuuint64_t gd[5] = {8, 24, 20, 2, 2};
    cuuint64_t gs[4] = {16 * 2, 24 * 16 * 2, 16, 20 * 24 * 16 * 2};
    cuuint32_t box[5] = {8, BX, BY, 2, 1}, es[5] = {1, 1, 1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, t.d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    rank = 5;
    c[0] = 0, c[1] = 16, c[2] = 16, c[3] = 0, c[4] = 1;  // bottom-right corner: OOB on the high side
    bytes = 8 * BX * BY * 2 * 2;
    expect = [=](int e) {
      int ch = e % 8, x = (e / 8) % BX, y = (e / 8 / BX) % BY, k = e / 8 / BX / BY;
      return t.at(1, 16 + y, 16 + x, k * 8 + ch);
    };
  } else {  // 3: C=64 swizzle 128B ; 4: C=16 swizzle 32B ; 5/6: the same with dst shifted off the pattern period
    const int C = (variant == 3 || variant == 5) ? 64 : 16;
    span = C * 2;
    t = make_tens(2, 20, 24, C);
    cuuint64_t gd[4] = {cuuint64_t(C), 24, 20, 2};
    cuuint64_t gs[3] = {cuuint64_t(C) * 2, 24ull * C * 2, 20ull * 24 * C * 2};
    cuuint32_t box[4] = {cuuint32_t(C), BX, BY, 1}, es[4] = {1, 1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    rank = 4;
    c[0] = 0, c[1] = -1, c[2] = 3, c[3] = 0;
    bytes = C * BX * BY * 2;
    if (variant >= 5) dst_off = (variant == 5) ? 3 * 128 : 128;  // 128 B aligned but not pattern aligned
    expect = [=](int e) { int ch = e % C, x = (e / C) % BX, y = e / C / BX; return t.at(0, 3 + y, x - 1, ch); };
  }
  if (r != CUDA_SUCCESS) {
    printf("RESULT %s encode failed CUresult=%d FAIL\n", name, int(r));
    return;
  }
  const int dump = dst_off + bytes + 256;
  uint8_t* dout;
  CK(cudaMalloc(&dout, dump));
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  k_tma<<<1, 128, 64 * 1024>>>(map, rank, c[0], c[1], c[2], c[3], c[4], dst_off, bytes, dout, dump);
  CK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("RESULT %s launch error %s FAIL\n", name, cudaGetErrorString(e));
    return;
  }
  std::vector<uint8_t> got(dump);
  CK(cudaMemcpy(got.data(), dout, dump, cudaMemcpyDeviceToHost));
  // hypothesis A: swizzle on absolute smem offset (dynamic smem base is 1024-aligned); B: relative to dst
  int badA = 0, badB = 0;
  for (int i = 0; i < bytes / 2; ++i) {
    float ex = expect(i);
    uint32_t la = swz(dst_off + i * 2, span), lb = dst_off + swz(i * 2, span);
    uint16_t ga, gb;
    memcpy(&ga, &got[la], 2);
    memcpy(&gb, &got[lb], 2);
    if (bf2f(ga) != ex) ++badA;
    if (bf2f(gb) != ex) ++badB;
  }
  int canary = 0;
  for (int i = dst_off + bytes; i < dump; ++i) canary += got[i] != 0xEE;
  printf("RESULT %s absolute-swizzle-bad=%d dst-relative-bad=%d of %d overrun=%d %s\n", name, badA, badB, bytes / 2, canary,
         (badA == 0 || badB == 0) ? "PASS" : "FAIL");
}

// ------------------------------------------------------------------------------------------
// p1: TMA fill rate.  Persistent CTAs stream 18 x 34 pixel halo tiles of an NHWC tensor into a
// 4-stage smem ring; nothing consumes them (the waiter just recycles the stage).
__global__ void __launch_bounds__(64, 1) k_tma_bw(const __grid_constant__ CUtensorMap map, int planes_per_tile, int inner_ch,
                                                  int stage_bytes, int tx_bytes, int tiles_x, int tiles_y, int nimg, int nstage) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], empty[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstage; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int ntiles = tiles_x * tiles_y * nimg;
  const int plane_bytes = stage_bytes / planes_per_tile;
  if (threadIdx.x == 0) {
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      int s = it % nstage, ph = (it / nstage) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&full[s], tx_bytes);
      int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / (tiles_x * tiles_y);
      for (int p = 0; p < planes_per_tile; ++p)
        tma_load_4d(&map, &full[s], smem + s * stage_bytes + p * plane_bytes, p * inner_ch, tx * 32 - 1, ty * 16 - 1, n);
    }
  } else if (threadIdx.x == 32) {
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      int s = it % nstage, ph = (it / nstage) & 1;
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
    }
  }
}

static void test_p1() {
  EncodeTiledFn enc = get_encode();
  struct V {
    int C, inner;
    CUtensorMapSwizzle sw;
    const char* nm;
  } vs[] = {{16, 8, CU_TENSOR_MAP_SWIZZLE_NONE, "C16 inner 8ch(16B) x2 planes"},
            {16, 16, CU_TENSOR_MAP_SWIZZLE_NONE, "C16 inner 16ch(32B) noswz"},
            {16, 16, CU_TENSOR_MAP_SWIZZLE_32B, "C16 inner 16ch(32B) swz32"},
            {64, 8, CU_TENSOR_MAP_SWIZZLE_NONE, "C64 inner 8ch(16B) x8 planes"},
            {64, 64, CU_TENSOR_MAP_SWIZZLE_128B, "C64 inner 64ch(128B) swz128"},
            {32, 8, CU_TENSOR_MAP_SWIZZLE_NONE, "C32 inner 8ch(16B) x4 planes"},
            {32, 32, CU_TENSOR_MAP_SWIZZLE_64B, "C32 inner 32ch(64B) swz64"}};
  for (auto& v : vs) {
    const int H = 256, W = 256, NI = v.C == 64 ? 32 : 64;
    size_t elems = size_t(NI) * H * W * v.C;
    uint16_t* d;
    CK(cudaMalloc(&d, elems * 2));
    CK(cudaMemset(d, 0, elems * 2));
    CUtensorMap map;
    cuuint64_t gd[4] = {cuuint64_t(v.C), W, H, cuuint64_t(NI)};
    cuuint64_t gs[3] = {cuuint64_t(v.C) * 2, cuuint64_t(W) * v.C * 2, cuuint64_t(H) * W * v.C * 2};
    cuuint32_t box[4] = {cuuint32_t(v.inner), 34, 18, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("RESULT p1 %s encode failed %d FAIL\n", v.nm, int(r));
      continue;
    }
    const int planes = v.C / v.inner;
    int plane_bytes = 34 * 18 * v.inner * 2;
    plane_bytes = (plane_bytes + 1023) / 1024 * 1024;  // keep swizzled planes pattern-aligned
    const int stage_bytes = planes * plane_bytes;
    // expect_tx must equal the bytes TMA really writes: full boxes even when partly OOB
    const int tx_bytes = planes * 34 * 18 * v.inner * 2;
    const int nstage = stage_bytes * 4 <= 200 * 1024 ? 4 : 2;
    CK(cudaFuncSetAttribute(k_tma_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, nstage * stage_bytes));
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
      if (ctas_per_sm * nstage * stage_bytes > 200 * 1024) continue;
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        k_tma_bw<<<148 * ctas_per_sm, 64, nstage * stage_bytes>>>(map, planes, v.inner, stage_bytes, tx_bytes, W / 32, H / 16, NI,
                                                                   nstage);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
          printf("RESULT p1 %s launch error %s FAIL\n", v.nm, cudaGetErrorString(e));
          exit(1);
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      double useful = double(elems) * 2;                                        // each pixel once
      double moved = double(NI) * (H / 16) * (W / 32) * 34 * 18 * v.C * 2;      // incl. halo
      printf("RESULT p1 %-32s ctas/sm=%d  %.3f ms  useful %.0f GB/s  with-halo %.0f GB/s PASS\n", v.nm, ctas_per_sm, best,
             useful / best * 1e-6, moved / best * 1e-6);
    }
    cudaFree(d);
  }
}

// p2: SS-mode MMA rate at small N.  Issued the way the production kernel does it: a converged
// warp runs the loop, one elected lane issues, descriptor updates are adds of compile-time
// constants, NACC independent accumulators are visited round-robin.
template <int NACC>
__global__ void __launch_bounds__(128, 1) k_mma_rate(uint32_t idesc, int n_iter, uint32_t lbo, uint32_t sbo, uint32_t lt,
                                                     int ncol, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x * 16; i < 128 * 1024; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base, base = smem_u32(smem);
  if (threadIdx.x < 32) {
    const uint64_t bdesc = make_sdesc(base + 96 * 1024, lt ? 16 : 2048, lt ? 1024 : 128, lt);
    const uint64_t adesc0 = make_sdesc(base, lbo, sbo, lt);
    long long t0 = clock64();
    for (int it = 0; it < n_iter; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
          umma_bf16(tb + (u % NACC) * ncol, adesc0 + uint64_t(u * (lt ? 64 : 1)), bdesc, idesc, (it > 0 || u >= NACC) ? 1u : 0u);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int NACC>
static void p2_case(int lt, int N, long long* dc) {
  if (NACC * N > 512) return;
  const int NIT = 512;
  CK(cudaFuncSetAttribute(k_mma_rate<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  k_mma_rate<NACC><<<148, 128, 128 * 1024>>>(make_idesc_bf16(128, N), NIT, lt ? 16 : 9808, lt ? 1024 : 34 * 16, lt, N, dc);
  CK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("RESULT p2 launch error %s FAIL\n", cudaGetErrorString(e));
    exit(1);
  }
  long long c;
  CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
  printf("RESULT p2 layout=%s N=%3d independent-accumulators=%2d  %.1f cycles/MMA (128xNx16)  PASS\n", lt ? "swz128" : "planar-noswz", N,
         NACC, double(c) / (NIT * 16));
}


// p3: does the SS-mode MMA rate depend on the ALIGNMENT of the A start address relative to the
// swizzle atom (8 rows x span)?  The production conv shifts the start by whole pixels (= span bytes)
// and whole tile rows (P*span bytes) to select a filter tap; B is the no-swizzle packed weight tile.
__global__ void __launch_bounds__(128, 1) k_mma_rate3(uint32_t idesc, int n_iter, uint32_t a_off, uint32_t sbo, uint32_t lt, uint32_t j_step,
                                                      int ncol, long long* cycles, uint32_t b_step = 0, int nacc = 8, uint32_t b_off = 64 * 1024) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x * 16; i < 128 * 1024; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base, base = smem_u32(smem);
  if (threadIdx.x < 32) {
    const uint64_t bdesc = make_sdesc(base + b_off, uint32_t(ncol * 16), 128, 0);
    const uint64_t adesc0 = make_sdesc(base + a_off, 16, sbo, lt);
    long long t0 = clock64();
    for (int it = 0; it < n_iter; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          umma_bf16(tb + (u % nacc) * ncol, adesc0 + uint64_t((u * j_step) >> 4), bdesc + uint64_t((u * b_step) >> 4), idesc, (it > 0 || u >= nacc) ? 1u : 0u);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

static void test_p4() {
  long long* dc;
  CK(cudaMalloc(&dc, 8));
  CK(cudaFuncSetAttribute(k_mma_rate3, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int NIT = 256;
  for (int N : {16, 64})
    for (uint32_t a_off : {0u, 40u * 1024u, 100u * 1024u})
      for (uint32_t b_off : {48u * 1024u, 64u * 1024u, 80u * 1024u, 96u * 1024u, 97u * 1024u, 112u * 1024u, 160u * 1024u}) {
        const int span = N == 16 ? 32 : 64;
        k_mma_rate3<<<148, 128, 200 * 1024>>>(make_idesc_bf16(128, N), NIT, a_off, span == 32 ? 66 * 32 : 2304, span == 32 ? 6 : 4, 8 * span, N, dc, N == 64 ? 2048u : 512u, N == 64 ? 2 : 8, b_off);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        long long cy;
        CK(cudaMemcpy(&cy, dc, 8, cudaMemcpyDeviceToHost));
        printf("RESULT p4 N=%2d A at %3u KB, B at %3u KB (B steps per MMA): %.1f cycles/MMA PASS\n", N, a_off / 1024, b_off / 1024, double(cy) / (NIT * 8));
      }
}

static void test_p3() {
  long long* dc;
  CK(cudaMalloc(&dc, 8));
  CK(cudaFuncSetAttribute(k_mma_rate3, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  const int NIT = 512;
  struct Case { const char* name; int span; uint32_t a_off, sbo; int N; };
  const Case cases[] = {
      {"swz32 aligned dense (sbo=256)", 32, 0, 256, 16},        {"swz32 start+32B dense", 32, 32, 256, 16},
      {"swz32 start+128B dense", 32, 128, 256, 16},              {"swz32 aligned sbo=66px", 32, 0, 66 * 32, 16},
      {"swz32 start+1px sbo=66px", 32, 32, 66 * 32, 16},         {"swz32 aligned sbo=72px", 32, 0, 72 * 32, 16},
      {"swz32 start+1px sbo=72px", 32, 32, 72 * 32, 16},         {"swz32 aligned sbo=72px N=48", 32, 0, 72 * 32, 48},
      {"swz64 aligned dense (sbo=512)", 64, 0, 512, 32},         {"swz64 start+1px dense", 64, 64, 512, 32},
      {"swz64 start+1px sbo=66px", 64, 64, 66 * 64, 32},         {"swz64 aligned sbo=72px", 64, 0, 72 * 64, 32},
      {"swz128 aligned dense (sbo=1024)", 128, 0, 1024, 64},     {"swz128 start+1px dense", 128, 128, 1024, 64},
      {"swz128 start+1px sbo=34px", 128, 128, 34 * 128, 64},     {"swz128 aligned sbo=40px", 128, 0, 40 * 128, 64},
      {"swz64 aligned dense N=64", 64, 0, 512, 64},              {"swz64 half-row start dense N=64", 64, 32, 512, 64},
      {"swz64 aligned sbo=2304 N=64", 64, 0, 2304, 64},          {"swz64 half-row sbo=2304 N=64", 64, 32, 2304, 64},
      {"swz64 half-row sbo=2368 N=64", 64, 32, 2368, 64},        {"swz64 half-row sbo=2304 N=32", 64, 32, 2304, 32},
      {"swz32 aligned dense N=64", 32, 0, 256, 64},              {"swz32 sbo=66px N=64", 32, 32, 66 * 32, 64},
  };
  for (const Case& c : cases) {
    const uint32_t lt = c.span == 128 ? 2 : c.span == 64 ? 4 : 6;
    if (8 * c.N > 512) {
      continue;
    }
    k_mma_rate3<<<148, 128, 128 * 1024>>>(make_idesc_bf16(128, c.N), NIT, c.a_off, c.sbo, lt, 8 * c.span, c.N, dc);
    if (c.N == 64 && c.span == 64 && c.a_off == 32 && c.sbo == 2304) {  // the 2x2-blocked production pattern: B changes every MMA, 2 accumulators
      for (int variant = 0; variant < 3; ++variant) {
        CK(cudaDeviceSynchronize());
        k_mma_rate3<<<148, 128, 128 * 1024>>>(make_idesc_bf16(128, 64), NIT, 32, 2304, lt, variant == 2 ? 32u : 8 * 64u, 64, dc, variant >= 1 ? 2048u : 0u, 2);
        CK(cudaDeviceSynchronize());
        long long cy;
        CK(cudaMemcpy(&cy, dc, 8, cudaMemcpyDeviceToHost));
        printf("RESULT p3   variant %d (B step %s, 2 accumulators, A step %s)  %.1f cycles/MMA  PASS\n", variant, variant >= 1 ? "2KB" : "0", variant == 2 ? "32B" : "512B",
               double(cy) / (NIT * 8));
      }
    }
    CK(cudaGetLastError());
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("RESULT p3 %s launch error %s FAIL\n", c.name, cudaGetErrorString(e));
      exit(1);
    }
    long long cyc;
    CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    printf("RESULT p3 %-36s N=%3d  %.1f cycles/MMA (128xNx16)  PASS\n", c.name, c.N, double(cyc) / (NIT * 8));
  }
}


// m9: the wgrad scheme on swizzled NHWC tiles.  Both operands MN-major (the contiguous dimension of a [pixel][channel]
// tile is the GEMM M/N dimension, pixels are K).  A atoms (C channels each) are x-SHIFTED copies of the same tile:
// LBO = one pixel (= span bytes), SBO = 8 pixels.  D[(j, ci), co] = sum_{p<16} X[p0 + p + j][ci] * G[p0 + p][co].
static void test_m9(int C, int M) {
  const int span = C * 2, N = C < 64 ? C : 64, NPIX = 96, atoms = M / C;
  std::vector<float> X(NPIX * C), G(NPIX * N);
  for (auto& v : X) v = frand();
  for (auto& v : G) v = frand();
  std::vector<uint8_t> img;
  for (int p = 0; p < NPIX; ++p)
    for (int c = 0; c < C; ++c) put_bf(img, swz(p * span + c * 2, span), X[p * C + c]);
  const int gbase = (NPIX * span + 1023) / 1024 * 1024;
  const int gspan = N * 2;
  for (int p = 0; p < NPIX; ++p)
    for (int c = 0; c < N; ++c) put_bf(img, gbase + swz(p * gspan + c * 2, gspan), G[p * N + c]);
  const uint32_t lt = span == 128 ? 2 : span == 64 ? 4 : 6, glt = gspan == 128 ? 2 : gspan == 64 ? 4 : 6;
  const int p0 = 5;  // unaligned pixel start on purpose (taps shift the start by whole pixels)
  MmaJob j{};
  j.adesc = make_sdesc(p0 * span, span /*LBO: next atom = next pixel*/, 8 * span, lt);
  j.bdesc = make_sdesc(gbase + p0 * gspan, gspan, 8 * gspan, glt);
  j.accumulate = 0;
  std::vector<float> exp(128 * N, 0.f);
  for (int a = 0; a < atoms; ++a)
    for (int ci = 0; ci < C; ++ci)
      for (int co = 0; co < N; ++co) {
        float s = 0;
        for (int p = 0; p < 16; ++p) s += X[(p0 + p + a) * C + ci] * G[(p0 + p) * N + co];
        exp[(a * C + ci) * N + co] = s;
      }
  char name[128];
  snprintf(name, sizeof name, "m9(mn-major swz%d, %d x-shifted atoms via LBO=1px, M=%d N=%d)", span, atoms, M, N);
  if (M == 128) {
    run_mma(name, img, {j}, make_idesc_bf16(128, N, 1, 1), N, exp);
  } else {
    // M = 64: find out which TMEM lanes hold rows 0..63 (rows beyond M were poisoned with NaN bytes by run_mma)
    std::vector<float> e2(128 * N, 0.f);
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < N; ++n) e2[m * N + n] = exp[m * N + n];
    run_mma(name, img, {j}, make_idesc_bf16(64, N, 1, 1), N, e2);
  }
}

static void test_p2() {
  long long* dc;
  CK(cudaMalloc(&dc, 8));
  for (int lt : {0, 2})
    for (int N : {16, 32, 64, 128, 256}) {
      p2_case<1>(lt, N, dc);
      p2_case<2>(lt, N, dc);
      p2_case<4>(lt, N, dc);
      p2_case<8>(lt, N, dc);
      p2_case<16>(lt, N, dc);
    }
}

// m8: mixed layouts in one instruction — A = swizzle-32B pixel-major halo tile (16 channels per
// pixel row), 16x8 patch with SBO = one tile row, nine taps through the start address;
// B = weights in the no-swizzle core-matrix layout.  This is the production conv configuration.
static void test_m8(int span) {
  const int C = span / 2, Cout = 32, TW = 32, P = TW + 2, ROWS = 18;
  std::vector<float> X(ROWS * P * C), Wt(9 * Cout * C);
  for (auto& v : X) v = frand();
  for (auto& v : Wt) v = frand();
  std::vector<uint8_t> img;
  for (int p = 0; p < ROWS * P; ++p)
    for (int c = 0; c < C; ++c) put_bf(img, swz(p * span + c * 2, span), X[p * C + c]);
  const int wbase = (ROWS * P * span + 1023) / 1024 * 1024;
  for (int t = 0; t < 9; ++t)
    for (int co = 0; co < Cout; ++co)
      for (int c = 0; c < C; ++c) put_bf(img, wbase + ((t * (C / 8) + c / 8) * Cout + co) * 16 + (c % 8) * 2, Wt[(t * Cout + co) * C + c]);
  const uint32_t lt = span == 128 ? 2 : span == 64 ? 4 : 6;
  const int px0 = 24, py0 = 0;
  std::vector<MmaJob> jobs;
  for (int t = 0; t < 9; ++t)
    for (int ks = 0; ks < C / 16; ++ks) {
      int r = t / 3, s = t % 3;
      MmaJob j{};
      j.adesc = make_sdesc(((py0 + r) * P + px0 + s) * span + ks * 32, 16, P * span, lt, 0);
      j.bdesc = make_sdesc(wbase + (t * (C / 8) + 2 * ks) * Cout * 16, Cout * 16, 128, 0, 0);
      j.accumulate = jobs.empty() ? 0 : 1;
      jobs.push_back(j);
    }
  std::vector<float> exp(128 * Cout);
  for (int m = 0; m < 128; ++m)
    for (int co = 0; co < Cout; ++co) {
      int oy = py0 + m / 8, ox = px0 + m % 8;
      float sum = 0;
      for (int t = 0; t < 9; ++t)
        for (int c = 0; c < C; ++c) sum += X[((oy + t / 3) * P + ox + t % 3) * C + c] * Wt[(t * Cout + co) * C + c];
      exp[m * Cout + co] = sum;
    }
  char nm[128];
  snprintf(nm, sizeof nm, "m8(A swizzle-%dB halo tile 16x8 patch 9 taps, B no-swizzle)", span);
  run_mma(nm, img, jobs, make_idesc_bf16(128, Cout), Cout, exp);
}

int main(int argc, char** argv) {
  std::string t = argc > 1 ? argv[1] : "m1";
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("# device %s sm_%d%d SMs=%d test=%s\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, t.c_str());
  if (t == "m1") test_m1(false);
  else if (t == "m1s") test_m1(true);
  else if (t == "m2") test_m2();
  else if (t == "m3") test_m3(false);
  else if (t == "m3s") test_m3(true);
  else if (t == "m4") {
    test_swz("m4a(swz128 aligned dense)", 128, 0, 0, 0);
    test_swz("m4b(swz32 aligned dense)", 32, 0, 0, 0);
  } else if (t == "m5") {
    test_swz("m5a(swz128 shift 3 rows, base_offset=0)", 128, 1, 3, 0);
    test_swz("m5b(swz128 shift 3 rows, base_offset=phase)", 128, 1, 3, 1);
  } else if (t == "m6") {
    test_swz("m6a(swz128 16x8 patch SBO=34 rows shift 37, bo=0)", 128, 2, 37, 0);
    test_swz("m6b(swz128 16x8 patch SBO=34 rows shift 37, bo=phase)", 128, 2, 37, 1);
  } else if (t == "m7") {
    test_swz("m7a(swz32 shift 5 rows, bo=0)", 32, 1, 5, 0);
    test_swz("m7b(swz32 16x8 patch SBO=34 rows shift 37, bo=0)", 32, 2, 37, 0);
    test_swz("m7c(swz32 16x8 patch SBO=34 rows shift 37, bo=phase)", 32, 2, 37, 1);
  } else if (t == "t1") test_tma("t1(4d box 8ch no swizzle, negative coords)", 1);
  else if (t == "t2") test_tma("t2(5d chunk-dim stride 16B)", 2);
  else if (t == "t3") test_tma("t3(4d C=64 swizzle128)", 3);
  else if (t == "t4") test_tma("t4(4d C=16 swizzle32)", 4);
  else if (t == "t5") test_tma("t5(4d C=64 swizzle128 dst+384)", 5);
  else if (t == "t6") test_tma("t6(4d C=16 swizzle32 dst+128)", 6);
  else if (t == "p1") test_p1();
  else if (t == "p2") test_p2();
  else if (t == "m9") {
    test_m9(16, 128);
    test_m9(32, 128);
    test_m9(64, 128);
    test_m9(16, 64);
  } else if (t == "p3") test_p3();
  else if (t == "p4") test_p4();
  else if (t == "m8") {
    test_m8(32);
    test_m8(64);
    test_m8(128);
  }
  else printf("unknown test\n");
  return 0;
}

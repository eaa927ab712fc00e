// wgrad.cu — unpp_wgrad entry point + the POINTWISE weight-gradient kernel.
//
// Every 3x3 conv and every transposed conv of the network goes to the tcgen05 kernel (wgrad_tc.cu).  The kernel of this file serves
// what is left: taps = 1 shapes — the 1x1 conv of the is_deconv=False variant (models/unet.py:189-191) and single taps of a
// transposed conv through a strided dZ view — HBM-bound GEMMs whose few hundred outputs live in registers; a 3x3 shape the
// tcgen05 kernel cannot take is rejected (UNPP_ERR_UNSUPPORTED), never re-routed.  (The template still carries its 3x3
// instantiation: tests/test_kernels_gpu.py uses none of it through the C ABI.)  The arithmetic, as an implicit GEMM whose
// reduction dimension is the pixel grid:
//
//     dW[tap][ci][co] = sum_{n,y,x} X[n, y + r - 1, x + s - 1, ci] * dZ[n, y, x, co]
//
// Both operands are NHWC bf16 tiles brought in by TMA (zero-filled halo / out-of-image pixels, so
// no masking anywhere).  A job = one <=64-channel chunk of one source tensor x one <=64-channel
// slice of dZ; blockIdx.y enumerates the jobs, blockIdx.x strides over 8x32-pixel tiles.  Every
// warp keeps its (tap, 16ci, 16co) accumulator fragments in registers across ALL the tiles of the
// CTA (the output is tiny: 9*Cin*Cout), so HBM traffic is one read of X and dZ; the per-CTA result
// goes to a partial buffer reduced in fixed order by unpp_wgrad_reduce (deterministic).
//
// Tensor-core path: warp-level mma.sync m16n8k16 (bf16 -> fp32).  The pixel dimension is K, so both
// operands are needed 'transposed' relative to the NHWC tiles: ldmatrix.trans delivers the X tile
// [pixel][ci] as the A fragment and the dZ tile [pixel][co] as the B fragment straight from the
// swizzled TMA tiles — no transposes are materialised and the loads are bank-conflict free.
#include "sm100.cuh"
#include "common.h"
#include "../../include/unpp.h"
#include "wgrad_tc.h"

using namespace sm100;

namespace {

constexpr int kMaxJobs = 16;
constexpr int TR = 8, TC = 32;  // tile rows / cols (pixels)
constexpr int kThreads = 256;
constexpr int kMaxStages = 8;  // TMA tiles in flight per CTA: the loop is latency-bound with fewer (8x32-pixel tiles are small)

struct WgradParams {
  CUtensorMap xmap[UNPP_MAX_SRC];
  CUtensorMap zmap[4];  // dense dZ: [0]; transposed-conv taps: one strided view per (p, q), selected by blockIdx.z
  int njobs;
  int job_map[kMaxJobs], job_c0[kMaxJobs], job_cs[kMaxJobs], job_cioff[kMaxJobs], job_co0[kMaxJobs], job_con[kMaxJobs];
  int tiles_x, tiles_y, ntiles;
  int pad, cin_total, cout;
  int xstage_bytes, zstage_bytes, nstage;
  float* partial;
};

__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// TMA writes the tiles with the hardware swizzle of their row width (conflict-free ldmatrix):
// the 16-byte chunk index inside a row is XORed with address bits [7, 7+log2(span/16)).
__device__ __forceinline__ uint32_t swz(uint32_t addr, uint32_t mask) { return addr ^ (((addr >> 7) & mask) << 4); }

template <int TAPS, int PPW>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint64_t bar_full[kMaxStages];
  uint8_t* const smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job = blockIdx.y;
  const int cs = p.job_cs[job], con = p.job_con[job];
  const int nco = con >> 4, npairs = (cs >> 4) * nco;
  const int pad = p.pad, PX = TC + 2 * pad, PY = TR + 2 * pad;
  const int stage_bytes = p.xstage_bytes + p.zstage_bytes, xstage_bytes = p.xstage_bytes;
  const int ntiles = p.ntiles, tiles_x = p.tiles_x, tiles_y = p.tiles_y, nstage = p.nstage;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) mbar_init(&bar_full[i], 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.xmap[p.job_map[job]]);
    tma_prefetch_desc(&p.zmap[blockIdx.z]);
  }
  __syncthreads();
  unpp::pdl_wait();  // see common.h: the set-up above overlaps the tail of the preceding kernel
  unpp::pdl_trigger();

  auto issue = [&](int tile, int buf) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    uint8_t* xs = smem + size_t(buf) * stage_bytes;
    uint8_t* zs = xs + xstage_bytes;
    mbar_arrive_expect_tx(&bar_full[buf], uint32_t(PY * PX * cs * 2 + TR * TC * con * 2));
    tma_load_4d(&p.xmap[p.job_map[job]], &bar_full[buf], xs, p.job_c0[job], tx * TC - pad, ty * TR - pad, n);
    tma_load_4d(&p.zmap[blockIdx.z], &bar_full[buf], zs, p.job_co0[job], tx * TC, ty * TR, n);
  };

  // fragment ownership: a warp owns ppw_job (tap x 16ci x 16co) fragment sets that share the ci block
  int pair0, row0, row_step;
  const int ppw_job = npairs >= 8 ? npairs / 8 : 1;  // <= PPW
  if (npairs >= 8) {
    pair0 = warp * ppw_job, row0 = 0, row_step = 1;
  } else {
    pair0 = warp % npairs, row0 = warp / npairs, row_step = 8 / npairs;
  }
  const int cb = pair0 / nco, ob0 = pair0 % nco;  // pairs pair0 .. pair0+ppw_job-1 have the same cb (nco is even when ppw_job = 2)

  // per-lane ldmatrix row addresses (bytes inside a tile, before the row / tap / k-step offsets)
  const int lt = lane >> 3, li = lane & 7;
  const uint32_t a_lane = uint32_t((li + 8 * (lt >> 1)) * cs + cb * 16 + 8 * (lt & 1)) * 2;      // A = X^T: k = pixel, m = ci
  const uint32_t b_lane = uint32_t((li + 8 * (lt & 1)) * con + ob0 * 16 + 8 * (lt >> 1)) * 2;   // B = dZ:  k = pixel, n = co
  const uint32_t xmask = uint32_t(cs * 2 / 16 - 1) & 7u, zmask = uint32_t(con * 2 / 16 - 1) & 7u;
  const uint32_t x_px = uint32_t(cs * 2), z_px = uint32_t(con * 2);
  const uint32_t smem_base = smem_u32(smem);

  float acc[PPW][TAPS][2][4];
#pragma unroll
  for (int j = 0; j < PPW; ++j)
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][t][h][e] = 0.f;

  // prologue: fill the ring (tile k of this CTA goes to buffer k % nstage)
  if (threadIdx.x == 0)
    for (int k = 0; k < nstage; ++k) {
      const int tile = blockIdx.x + k * int(gridDim.x);
      if (tile < ntiles) issue(tile, k);
    }
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it % nstage;
    mbar_wait(&bar_full[buf], (it / nstage) & 1);
    const uint32_t xs = smem_base + uint32_t(buf) * stage_bytes, zs = xs + xstage_bytes;
    for (int row = row0; row < TR; row += row_step) {
#pragma unroll
      for (int kk = 0; kk < TC / 16; ++kk) {
        const int x0 = kk * 16;
        uint32_t fb[PPW][4];
#pragma unroll
        for (int j = 0; j < PPW; ++j)
          if (j < ppw_job) ldsm_x4_trans(swz(zs + uint32_t(row * TC + x0) * z_px + b_lane + uint32_t(j * 32), zmask), fb[j]);
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
          const int r = TAPS == 9 ? t / 3 : 0, s = TAPS == 9 ? t % 3 : 0;
          uint32_t fa[4];
          ldsm_x4_trans(swz(xs + uint32_t((row + r) * PX + x0 + s) * x_px + a_lane, xmask), fa);
#pragma unroll
          for (int j = 0; j < PPW; ++j) {
            if (j < ppw_job) {
              mma_16816(acc[j][t][0], fa, fb[j][0], fb[j][1]);
              mma_16816(acc[j][t][1], fa, fb[j][2], fb[j][3]);
            }
          }
        }
      }
    }
    __syncthreads();  // everyone is done with `buf`: refill it with the tile nstage iterations ahead
    if (threadIdx.x == 0) {
      const int nxt = tile + nstage * int(gridDim.x);
      if (nxt < ntiles) issue(nxt, buf);
    }
  }

  // stage the fragments in shared memory ([frag][16 ci][16 co] fp32), then sum the warps that share a pair in a fixed order
  float* stage_f = reinterpret_cast<float*>(smem);
  {
    const int g = lane >> 2, c2 = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < PPW; ++j)
#pragma unroll
      for (int t = 0; t < TAPS; ++t) {
        float* f = stage_f + ((warp * PPW + j) * TAPS + t) * 256;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          *reinterpret_cast<float2*>(f + g * 16 + h * 8 + c2) = make_float2(acc[j][t][h][0], acc[j][t][h][1]);
          *reinterpret_cast<float2*>(f + (g + 8) * 16 + h * 8 + c2) = make_float2(acc[j][t][h][2], acc[j][t][h][3]);
        }
      }
  }
  __syncthreads();
  const int nout = TAPS * cs * con;
  const int wpp = npairs >= 8 ? 1 : 8 / npairs;
  for (int i = threadIdx.x; i < nout; i += kThreads) {
    const int col = i % con, cil = (i / con) % cs, t = i / (con * cs);
    const int pair = (cil >> 4) * nco + (col >> 4);
    const int e = (cil & 15) * 16 + (col & 15);
    float s = 0.f;
    if (npairs >= 8) {
      const int w = pair / ppw_job, j = pair % ppw_job;
      s = stage_f[((w * PPW + j) * TAPS + t) * 256 + e];
    } else {
      for (int k = 0; k < wpp; ++k) s += stage_f[(((pair + k * npairs) * PPW) * TAPS + t) * 256 + e];
    }
    p.partial[((size_t(blockIdx.z * gridDim.x + blockIdx.x) * TAPS + t) * p.cin_total + p.job_cioff[job] + cil) * p.cout + p.job_co0[job] + col] = s;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) != cudaSuccess || !q) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  return fn;
}

struct Plan {
  int njobs, ppw, pad, cin_total, tiles_x, tiles_y, ntiles, grid_x, grid_z, xstage_bytes, zstage_bytes, smem_total, nstage;
  int job_map[kMaxJobs], job_c0[kMaxJobs], job_cs[kMaxJobs], job_cioff[kMaxJobs], job_co0[kMaxJobs], job_con[kMaxJobs];
};

int make_plan(const UnppWgradArgs* a, Plan* pl) {
  if (!a || a->nsrc < 1 || a->nsrc > UNPP_MAX_SRC) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: nsrc out of range");
  if (a->taps != 9 && a->taps != 1) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: taps must be 1 or 9");
  if (a->N < 1 || a->H < 1 || a->W < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: empty pixel grid");
  if (a->cout != 16 && a->cout != 32 && a->cout != 64 && a->cout != 128) return unpp::fail(UNPP_ERR_UNSUPPORTED, "wgrad: cout must be 16/32/64/128");
  if (a->dz_step != 1 && a->dz_step != 2) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: dz_step must be 1 or 2");
  const bool all4 = a->dz_step == 2 && a->dz_oy < 0;
  if (all4 && a->taps != 1) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: the four-tap transposed-conv form needs taps = 1");
  pl->grid_z = all4 ? 4 : 1;
  int nj = 0, ci_off = 0, max_cs = 0, max_con = 0, max_pairs = 0;
  for (int i = 0; i < a->nsrc; ++i) {
    const int C = a->src_C[i];
    if (C != 16 && C != 32 && C != 64 && C != 128) return unpp::fail(UNPP_ERR_UNSUPPORTED, "wgrad: source channels must be 16/32/64/128");
    if (!a->src[i] || (reinterpret_cast<uintptr_t>(a->src[i]) & 15)) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: source pointer null or unaligned");
    for (int c0 = 0; c0 < C; c0 += 64) {
      const int cs = C - c0 < 64 ? C - c0 : 64;
      for (int co0 = 0; co0 < a->cout; co0 += 64) {
        const int con = a->cout - co0 < 64 ? a->cout - co0 : 64;
        if (nj >= kMaxJobs) return unpp::fail(UNPP_ERR_UNSUPPORTED, "wgrad: too many jobs");
        pl->job_map[nj] = i, pl->job_c0[nj] = c0, pl->job_cs[nj] = cs, pl->job_cioff[nj] = ci_off + c0, pl->job_co0[nj] = co0, pl->job_con[nj] = con;
        const int pairs = (cs / 16) * (con / 16);
        if (pairs != 1 && pairs != 2 && pairs != 4 && pairs != 8 && pairs != 16)
          return unpp::fail(UNPP_ERR_UNSUPPORTED, "wgrad: (chunk/16)*(cout slice/16) must be a power of two <= 16");
        if (cs > max_cs) max_cs = cs;
        if (con > max_con) max_con = con;
        if (pairs > max_pairs) max_pairs = pairs;
        ++nj;
      }
    }
    ci_off += C;
  }
  pl->njobs = nj, pl->cin_total = ci_off;
  pl->ppw = max_pairs > 8 ? 2 : 1;
  pl->pad = a->taps == 9 ? 1 : 0;
  const int PX = TC + 2 * pl->pad, PY = TR + 2 * pl->pad;
  pl->xstage_bytes = (PY * PX * max_cs * 2 + 1023) / 1024 * 1024;  // 1 KB granules keep every tile on the swizzle pattern period
  pl->zstage_bytes = (TR * TC * max_con * 2 + 1023) / 1024 * 1024;
  const int stage = pl->xstage_bytes + pl->zstage_bytes;
  pl->nstage = (200 * 1024) / stage;
  if (pl->nstage > kMaxStages) pl->nstage = kMaxStages;
  if (pl->nstage < 2) return unpp::fail(UNPP_ERR_UNSUPPORTED, "wgrad: tile does not fit twice in shared memory");
  const int pipe = pl->nstage * stage;
  const int staging = 8 * pl->ppw * a->taps * 256 * 4;
  pl->smem_total = 1024 + (pipe > staging ? pipe : staging);
  pl->tiles_x = (a->W + TC - 1) / TC, pl->tiles_y = (a->H + TR - 1) / TR;
  pl->ntiles = pl->tiles_x * pl->tiles_y * a->N;
  int gx = unpp::num_sms() / (nj * pl->grid_z);
  if (gx < 1) gx = 1;
  if (gx > pl->ntiles) gx = pl->ntiles;
  pl->grid_x = gx;
  return UNPP_OK;
}

template <int TAPS, int PPW>
int launch(const WgradParams& p, const Plan& pl, cudaStream_t stream) {
  static unsigned char opted_in[64] = {0};  // per instantiation and device
  if (cudaError_t e = unpp::opt_in_smem(wgrad_kernel<TAPS, PPW>, 220 * 1024, opted_in)) return unpp::fail_cuda_err("wgrad: cudaFuncSetAttribute", e);
  if (cudaError_t e = unpp::launch(wgrad_kernel<TAPS, PPW>, dim3(pl.grid_x, pl.njobs, pl.grid_z), kThreads, pl.smem_total, stream, p))
    return unpp::fail_cuda_err("wgrad: launch", e);
  return UNPP_OK;
}

}  // namespace

extern "C" int unpp_wgrad_grid(const UnppWgradArgs* a) {
  if (a && a->nsrc >= 1 && a->nsrc <= UNPP_MAX_SRC && a->N >= 1 && a->H >= 1 && a->W >= 1 && unpp::wgrad_tc_eligible(a)) return unpp::wgrad_tc_grid(a);
  if (a && a->taps != 1) return unpp::fail(UNPP_ERR_UNSUPPORTED, "wgrad: 3x3 shape not supported by the tcgen05 weight-gradient kernel (channels must be 16/32/64/128)");
  Plan pl;
  int rc = make_plan(a, &pl);
  return rc ? rc : pl.grid_x;
}

extern "C" int unpp_wgrad(const UnppWgradArgs* a, unpp_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || a->nsrc < 1 || a->nsrc > UNPP_MAX_SRC) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: nsrc out of range");
  if (a->N < 1 || a->H < 1 || a->W < 1) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: empty pixel grid");
  if (!a->dz || !a->partial) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: dz / partial is null");
  if ((reinterpret_cast<uintptr_t>(a->dz) & 15)) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: dz pointer unaligned");
  for (int i = 0; i < a->nsrc; ++i)
    if (!a->src[i] || (reinterpret_cast<uintptr_t>(a->src[i]) & 15)) return unpp::fail(UNPP_ERR_BAD_ARG, "wgrad: source pointer null or unaligned");
  if (unpp::wgrad_tc_eligible(a)) return unpp::wgrad_tc_launch(a, stream);  // tcgen05 path: every 3x3 conv and every transposed conv of the network
  // What is left for the kernel of this file: POINTWISE weight gradients (taps = 1: the 1x1 conv of the is_deconv=False variant,
  // models/unet.py:189-191, and single taps of a transposed conv) — HBM-bound GEMMs with K = pixels and a few hundred outputs.
  // A 3x3 shape the tcgen05 kernel rejects is an error, not a silent change of kernel.
  if (a->taps != 1) return unpp::fail(UNPP_ERR_UNSUPPORTED, "wgrad: 3x3 shape not supported by the tcgen05 weight-gradient kernel (channels must be 16/32/64/128)");
  Plan pl;
  if (int rc = make_plan(a, &pl)) return rc;
  EncodeTiledFn enc = get_encode();
  if (!enc) return unpp::fail(UNPP_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled not available from the driver");
  WgradParams p;
  memset(&p, 0, sizeof p);
  const int PX = TC + 2 * pl.pad, PY = TR + 2 * pl.pad;
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < a->nsrc; ++i) {
    const cuuint64_t C = a->src_C[i];
    cuuint64_t gd[4] = {C, cuuint64_t(a->W), cuuint64_t(a->H), cuuint64_t(a->N)};
    cuuint64_t gs[3] = {C * 2, cuuint64_t(a->W) * C * 2, cuuint64_t(a->H) * a->W * C * 2};
    const int bc = C < 64 ? int(C) : 64;
    cuuint32_t box[4] = {cuuint32_t(bc), cuuint32_t(PX), cuuint32_t(PY), 1};
    CUresult r = enc(&p.xmap[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->src[i]), gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     bc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : bc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return unpp::fail(UNPP_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled failed (CUresult %d) for source %d", int(r), i);
  }
  for (int z = 0; z < pl.grid_z; ++z) {
    // dZ may be a stride-2 view of a [N, 2H, 2W, cout] tensor (transposed-conv weight gradient): tap (p, q) = (oy, ox)
    const int oy = pl.grid_z == 4 ? (z >> 1) : a->dz_oy, ox = pl.grid_z == 4 ? (z & 1) : a->dz_ox;
    const cuuint64_t C = a->cout, st = a->dz_step, fullW = cuuint64_t(a->W) * st, fullH = cuuint64_t(a->H) * st;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a->dz) + (size_t(oy) * fullW + ox) * C * 2;
    cuuint64_t gd[4] = {C, cuuint64_t(a->W), cuuint64_t(a->H), cuuint64_t(a->N)};
    cuuint64_t gs[3] = {C * 2 * st, fullW * C * 2 * st, fullH * fullW * C * 2};
    const int con = a->cout < 64 ? a->cout : 64;
    cuuint32_t box[4] = {cuuint32_t(con), cuuint32_t(TC), cuuint32_t(TR), 1};
    CUresult r = enc(&p.zmap[z], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<uint8_t*>(base), gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     con == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : con == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return unpp::fail(UNPP_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled failed (CUresult %d) for dZ", int(r));
  }
  p.njobs = pl.njobs;
  for (int j = 0; j < pl.njobs; ++j) {
    p.job_map[j] = pl.job_map[j], p.job_c0[j] = pl.job_c0[j], p.job_cs[j] = pl.job_cs[j];
    p.job_cioff[j] = pl.job_cioff[j], p.job_co0[j] = pl.job_co0[j], p.job_con[j] = pl.job_con[j];
  }
  p.tiles_x = pl.tiles_x, p.tiles_y = pl.tiles_y, p.ntiles = pl.ntiles;
  p.pad = pl.pad, p.cin_total = pl.cin_total, p.cout = a->cout;
  p.xstage_bytes = pl.xstage_bytes, p.zstage_bytes = pl.zstage_bytes, p.nstage = pl.nstage;
  p.partial = a->partial;
  if (a->taps == 9) return pl.ppw == 2 ? launch<9, 2>(p, pl, stream) : launch<9, 1>(p, pl, stream);
  return pl.ppw == 2 ? launch<1, 2>(p, pl, stream) : launch<1, 1>(p, pl, stream);
}

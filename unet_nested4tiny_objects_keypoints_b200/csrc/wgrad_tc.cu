// wgrad_tc.cu — weight gradient of the 3x3 convs of the 16/32-channel levels on tcgen05 tensor cores.
//
//     dW[r][s][ci][co] = sum_{n,y,x} X[n, y + r - 1, x + s - 1, ci] * dZ[n, y, x, co]
//
// The reduction dimension of this GEMM is the pixel grid, so BOTH operands are "MN-major" for the
// tensor core: in an NHWC tile [pixel][channel] the contiguous dimension is the GEMM M (ci) / N (co)
// dimension and the pixels are K.  tcgen05.mma reads such tiles directly (descriptor major bits), with
// the hardware swizzle TMA wrote them with, so nothing is transposed or copied:
//
//   * A = one row of the X tile.  M = 128 = (128 / C) "atoms" of C channels; the leading-dimension
//     byte offset of the descriptor is ONE PIXEL, so atom j is the same row shifted by j pixels:
//     rows (j, ci) of the accumulator are the filter column s = j (j < 3 is used, the rest of the
//     M = 128 the instruction insists on is discarded — the tensor pipe is idle anyway).
//   * B = three rows of the dZ tile.  N = 3 * Cout; the leading-dimension byte offset is one tile
//     row, so atom i is dZ one row further down: columns (i, co) are the filter row r = 2 - i.
//   * one MMA (K = 16 pixels) therefore accumulates all nine taps of a 16-pixel row segment:
//     D[(j, ci)][(i, co)] += sum_p X[Y][x0 - 1 + p + j][ci] * dZ[Y - 1 + i][x0 + p][co].
//     (probe/umma_probe.cu m9 verifies the descriptor form on B200.)
//
// Work is partitioned by X rows / dZ columns: a tile is TR rows x TW columns; its X box has a one
// pixel halo in x only, its dZ box a one-row halo in y only, and TMA zero-fills everything outside
// the image, so every (pixel, tap) pair is counted exactly once and nothing is masked.  The
// accumulators ([128 lanes] x [3 * Cout] fp32 per source, up to four row-interleaved copies so that
// four issuing threads work in parallel) stay in TMEM for the whole kernel; each CTA writes one
// partial [9][Cin_total][Cout] at the end, reduced in fixed order by unpp_wgrad_reduce (deterministic).
//
// Pixel-pair view (all sources and the output 16 channels wide — the full-resolution level): the
// [N,H,W,16] tensors are read as [N,H,W/2,32], i.e. a K row is a horizontal pixel PAIR (64 B TMA rows
// move at ~5.8 TB/s, 32 B rows only at ~4.4) and the M / N element inside an atom is (parity e, channel).
// With the X box starting one pair left of the dZ box, D[(j, e, ci)][(i, e', co)] is the tap with
// x-offset e - e' + 2j - 1 (in {0,1,2} for six of the sixteen (j, e, e') combinations): every tap is the
// sum of exactly two accumulator entries, which the read-out writes to two partial "slots" (the grid
// query reports 2 partials per CTA), so the fixed-order reduction stays as it is.
//
// Wide layers (a source or the output 64 / 128 channels wide): sources are cut into <= 64-channel chunks and
// the output into <= 64-channel slices; blockIdx.y enumerates the (chunk, slice) jobs, each CTA accumulating one
// of them over its share of the tiles (8-row tiles there: a 64-channel tile row is 128 B per pixel).  A
// 64-channel chunk has only two atoms in M = 128, so it takes two MMAs per K step: start shifts 0 (filter
// columns 0, 1) and +2 pixels (column 2).
//
// Transposed conv k2s2 (taps = 1, dz = [N, 2H, 2W, Cout] read through a stride-2 relation): dW[ci][co][p][q] =
// sum X[n,y,x,ci] * dU[n, 2y+p, 2x+q, co].  dU is viewed as [N, 2H, W, 2*Cout]: the K row of low-resolution pixel x is
// the hi-res pixel pair (2x, 2x+1), so the N element inside an atom is (q, co) and the two atoms (leading offset =
// one hi-res tile row) are p = 0, 1: one MMA per K step yields all four taps (Cout = 64: two boxes, one per q).
//
// Shared-memory operand traffic per MMA: 4 KB (A) + 96 * Cout B (B) for 16 K rows, i.e. 224 B per
// pixel and 16-channel source in the pair view (352 B without it) against 64 B of HBM traffic.
#include <cstdlib>
#include "sm100.cuh"
#include "common.h"
#include "wgrad_tc.h"

using namespace sm100;

namespace {

constexpr int TW = 32, PX = TW + 2;  // K rows (pixels or pixel pairs) per tile row; tile rows TR = 8 or 16 (plan)
constexpr int kIssuers = 4;
constexpr int kThreads = 256;  // warp 0: TMA producer, warps 1..4: MMA issuers, all eight: final TMEM read-out
constexpr int kMaxX = 8, kMaxZ = 4;
constexpr int kMaxChunks = 2 * UNPP_MAX_SRC, kMaxJobs = 16;

struct Params {
  CUtensorMap xmap[UNPP_MAX_SRC];
  CUtensorMap zmap;
  // K chunks (<= 64 channels of one source) and jobs (a run of chunks x one output slice); blockIdx.y = job
  int ch_map[kMaxChunks], ch_c0[kMaxChunks], ch_C[kMaxChunks], ch_cioff[kMaxChunks];  // ch_C: EFFECTIVE atom width (32 in the pair view)
  int job_ch0[kMaxJobs], job_nch[kMaxJobs], job_co0[kMaxJobs];
  int tiles_x, tiles_y, ntiles;
  int cin_total, cout;       // real widths of the partial [9][cin_total][cout]
  int TR;                    // tile rows
  int pair;                  // pixel-pair view
  int zspan;                 // bytes per K row of the dZ tile (= effective slice width * 2)
  int xslot, zslot, nx, nz;  // ring geometry: bytes per slot, slots
  int nsplit, ncols, tmem_cols;
  // geometry that differs between the 3x3 conv and the transposed-conv form
  int dc;                    // transposed-conv form
  int xpx, xorg;             // X box: K rows per tile row (34 / 32), x origin relative to the tile (-1 / 0)
  int zrows, zy_mul, zy_org; // dZ box: rows, y origin = zy_mul * ty * TR + zy_org
  int zrow_mul;              // B start row of X row r = zrow_mul * r
  int nq, zbox_bytes;        // dZ boxes per tile (q halves when 2*Cout > 64) and bytes per box
  float* partial;
};

__device__ __forceinline__ uint32_t layout_of_span(int span) { return span == 128 ? 2u : span == 64 ? 4u : 6u; }

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t x_full[kMaxX], x_empty[kMaxX], z_full[kMaxZ], z_empty[kMaxZ], bar_done;
  __shared__ uint32_t tmem_slot;
  uint8_t* const zring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* const xring = zring + p.nz * p.zslot;
  // (the shuffle tells the compiler that the warp index is warp-uniform: the issuing threads' descriptor arithmetic then stays in
  // uniform registers instead of one R2UR per MMA — the line where ncu showed their stalls; launch times did not change, though:
  // the kernel is bound by the MN-major operand fetch / HBM, not by the issue rate)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int job = blockIdx.y;
  const int ch0 = p.job_ch0[job], nch = p.job_nch[job], co0 = p.job_co0[job];

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxX; ++i) mbar_init(&x_full[i], 1), mbar_init(&x_empty[i], kIssuers);
    for (int i = 0; i < kMaxZ; ++i) mbar_init(&z_full[i], 1), mbar_init(&z_empty[i], kIssuers);
    mbar_init(&bar_done, kIssuers);
    fence_mbar_init();
    for (int k = 0; k < nch; ++k) tma_prefetch_desc(&p.xmap[p.ch_map[ch0 + k]]);
    tma_prefetch_desc(&p.zmap);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  unpp::pdl_wait();  // see common.h: the set-up above overlaps the tail of the preceding kernel
  unpp::pdl_trigger();
  const uint32_t tmem_base = tmem_slot;
  const int ntiles = p.ntiles, nx = p.nx, nz = p.nz, nsplit = p.nsplit, ncols = p.ncols;
  const int zspan = p.zspan, TR = p.TR;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int xi = 0, zi = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, n = tile / (p.tiles_x * p.tiles_y);
        {
          const int s = zi % nz;
          mbar_wait(&z_empty[s], ((zi / nz) & 1) ^ 1);
          mbar_arrive_expect_tx(&z_full[s], uint32_t(p.nq * p.zrows * TW * zspan));
          for (int qb = 0; qb < p.nq; ++qb)
            tma_load_4d(&p.zmap, &z_full[s], zring + size_t(s) * p.zslot + size_t(qb) * p.zbox_bytes, p.dc ? qb * (zspan >> 1) : (p.pair ? 0 : co0),
                        tx * TW, p.zy_mul * ty * TR + p.zy_org, n);
          ++zi;
        }
        for (int k = 0; k < nch; ++k, ++xi) {
          const int s = xi % nx, c = ch0 + k;
          mbar_wait(&x_empty[s], ((xi / nx) & 1) ^ 1);
          mbar_arrive_expect_tx(&x_full[s], uint32_t(TR * p.xpx * p.ch_C[c] * 2));
          tma_load_4d(&p.xmap[p.ch_map[c]], &x_full[s], xring + size_t(s) * p.xslot, p.ch_c0[c], tx * TW + p.xorg, ty * TR, n);
        }
      }
    }
  } else if (warp <= kIssuers) {
    // ------------------------------------------------------------------ MMA issuers
    // Accumulator a = (acc0(chunk) + mi) * nsplit + h (chunk, MMA of the K step, rows h, h + nsplit, ... of every tile)
    // belongs to issuer a % kIssuers.
    const int w = warp - 1;
    const uint32_t idesc = make_idesc_bf16(128, ncols, 1, 1);
    const uint32_t zring_a = smem_u32(zring), xring_a = smem_u32(xring);
    const uint32_t zslot = uint32_t(p.zslot), xslot = uint32_t(p.xslot);
    int spans[kMaxChunks];
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) spans[k] = ch0 + k < kMaxChunks ? p.ch_C[ch0 + k] * 2 : 0;
    const uint32_t z_row = uint32_t(p.zrow_mul * TW * zspan) >> 4, z_seg = uint32_t(16 * zspan) >> 4;
    const int dc = p.dc, xpx = p.xpx, nq = p.nq;
    const uint32_t zbox16 = uint32_t(p.zbox_bytes) >> 4;
    int xi = 0, zi = 0, tile_it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it, ++zi) {
      const int zs = zi % nz;
      mbar_wait(&z_full[zs], (zi / nz) & 1);
      const uint64_t bd = make_sdesc(zring_a + uint32_t(zs) * zslot, uint32_t(TW * zspan), uint32_t(8 * zspan), layout_of_span(zspan));
      const uint32_t b_hi = uint32_t(bd >> 32), b_lo = uint32_t(bd);
      int acc0 = 0;
#pragma unroll 1
      for (int k = 0; k < nch; ++k, ++xi) {
        const int xs = xi % nx;
        int span = spans[0];
#pragma unroll
        for (int kk = 1; kk < kMaxChunks; ++kk)
          if (kk == k) span = spans[kk];
        // MMAs per K step: 3x3 with a 64-channel chunk = two atoms per MMA, so two start shifts; transposed conv = one per dZ box
        const int nm = dc ? nq : (span == 128 ? 2 : 1);
        mbar_wait(&x_full[xs], (xi / nx) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = make_sdesc(xring_a + uint32_t(xs) * xslot, uint32_t(span), uint32_t(8 * span), layout_of_span(span));
          const uint32_t a_hi = uint32_t(ad >> 32), a_lo = uint32_t(ad);
          const uint32_t x_row = uint32_t(xpx * span) >> 4, x_seg = uint32_t(16 * span) >> 4;
          for (int mi = 0; mi < nm; ++mi) {
            for (int h = 0; h < nsplit; ++h) {
              const int a = (acc0 + mi) * nsplit + h;
              if (a % kIssuers != w) continue;
              const uint32_t acc = tmem_base + uint32_t(a * ncols);
              const uint32_t a_mi = a_lo + (dc ? 0u : (uint32_t(mi * 2 * span) >> 4));  // 3x3, second MMA: the row shifted by two more pixels
              const uint32_t b_mi = b_lo + (dc ? uint32_t(mi) * zbox16 : 0u);           // transposed conv: the dZ box of q = mi
              uint32_t accum = tile_it ? 1u : 0u;
              uint32_t ar = a_mi + uint32_t(h) * x_row, br = b_mi + uint32_t(h) * z_row;
              const uint32_t a_step = uint32_t(nsplit) * x_row, b_step = uint32_t(nsplit) * z_row;
#pragma unroll 4
              for (int row = h; row < TR; row += nsplit, ar += a_step, br += b_step) {
#pragma unroll
                for (int seg = 0; seg < TW / 16; ++seg) {
                  umma_bf16(acc, (uint64_t(a_hi) << 32) | (ar + uint32_t(seg) * x_seg), (uint64_t(b_hi) << 32) | (br + uint32_t(seg) * z_seg), idesc,
                            accum);
                  accum = 1u;
                }
              }
            }
          }
        }
        acc0 += nm;
        __syncwarp();
        if (elect_one()) umma_commit(&x_empty[xs]);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&z_empty[zs]);
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar_done);
    __syncwarp();
  }

  // ------------------------------------------------------------------ read-out: TMEM -> this CTA's partial
  __syncthreads();  // idle warps sleep here instead of polling the barrier
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  {
    const int q = warp & 3, half = warp >> 2;
    const int m = q * 32 + lane;
    const int ngroups = ncols >> 4;
    const int cout = p.cout, pair = p.pair;
    const int ecout = zspan >> 1;  // effective slice width
    const size_t part_elems = size_t(9) * p.cin_total * cout;
    float* const part = p.partial + size_t(blockIdx.x) * (pair ? 2 : 1) * part_elems;
    int acc0 = 0;
    for (int k = 0; k < nch; ++k) {
      const int c = ch0 + k;
      const int C = p.ch_C[c];  // effective width of an atom
      const int nm = p.dc ? p.nq : (C == 64 ? 2 : 1);
      for (int mi = 0; mi < nm; ++mi) {
        const int a0 = (acc0 + mi) * nsplit;
        const int j = p.dc ? (m / C ? 3 : 0) : m / C + 2 * mi;   // filter column (pair view: pair shift); transposed conv: atom 0 only
        const int lanes_used = p.dc ? C : (nm == 2 ? (mi ? 64 : 128) : 3 * C);
        if (q * 32 >= lanes_used) continue;  // warp-uniform: this lane quadrant only holds discarded shifts
        const int e = pair ? ((m >> 4) & 1) : 0, ci = pair ? (m & 15) : m % C;
        for (int g = half; g < ngroups; g += 2) {
          float v[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = 0.f;
          for (int h = 0; h < nsplit; ++h) {  // fixed order: deterministic
            uint32_t raw[16];
            tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t((a0 + h) * ncols + g * 16), raw);
            tmem_ld_wait16(raw);
#pragma unroll
            for (int t = 0; t < 16; ++t) v[t] += __uint_as_float(raw[t]);
          }
          if (p.dc) {  // column = (p, q, co) [one box] or (p, co) with q = mi [two boxes]; partial is [4 (2p+q)][grid][cin][cout]
            if (j == 0) {
              const int pp = (g * 16) / ecout, within = (g * 16) % ecout;
              const int qq = p.nq == 2 ? mi : within / cout, cc0 = p.nq == 2 ? within : within % cout;
              float4* dst = reinterpret_cast<float4*>(p.partial + (size_t(2 * pp + qq) * gridDim.x + blockIdx.x) * p.cin_total * cout +
                                                      size_t(p.ch_cioff[c] + ci) * cout + cc0);
#pragma unroll
              for (int t = 0; t < 4; ++t) dst[t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
            }
            continue;
          }
          int i, col0, sft, slot = 0;
          if (pair) {  // column group = (dZ row shift i, parity e'); tap column = e - e' + 2j - 1
            i = g >> 1;
            col0 = 0;
            sft = e - (g & 1) + 2 * j - 1;
            slot = sft == 1 ? e : 1 - e;
          } else {
            i = (g * 16) / ecout, col0 = (g * 16) % ecout, sft = j;
          }
          if (j < 3 && sft >= 0 && sft < 3) {
            const int tap = (2 - i) * 3 + sft;
            float4* dst = reinterpret_cast<float4*>(part + slot * part_elems + (size_t(tap) * p.cin_total + p.ch_cioff[c] + ci) * cout + co0 + col0);
#pragma unroll
            for (int t = 0; t < 4; ++t) dst[t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
          }
        }
      }
      acc0 += nm;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
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
  bool ok;
  int cin_total, xslot, zslot, nx, nz, nsplit, ncols, tmem_cols, smem_total, tiles_x, tiles_y, ntiles, grid_x;
  int TR, pair, ecout, We;  // tile rows; pixel-pair view; effective output slice width; K rows per image row
  int dc, nq;               // transposed-conv form; dZ boxes per tile
  int nchunks, njobs;
  int box_C[UNPP_MAX_SRC];  // channels per TMA box of each source (effective)
  int ch_map[kMaxChunks], ch_c0[kMaxChunks], ch_C[kMaxChunks], ch_cioff[kMaxChunks];
  int job_ch0[kMaxJobs], job_nch[kMaxJobs], job_co0[kMaxJobs];
};

CUtensorMapSwizzle swizzle_of(int channels) { return channels == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : channels == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B; }

void make_plan(const UnppWgradArgs* a, Plan* pl) {
  pl->ok = false;
  const bool dc = a->taps == 1 && a->dz_step == 2 && a->dz_oy < 0;  // the four taps of a k2s2 transposed conv in one launch
  if (!dc && (a->taps != 9 || a->dz_step != 1)) return;
  if (a->cout != 16 && a->cout != 32 && a->cout != 64 && a->cout != 128) return;
  if (dc && (a->cout > 64 || a->nsrc != 1)) return;
  bool all16 = a->cout == 16, wide = a->cout > 32;
  for (int i = 0; i < a->nsrc; ++i) {
    const int C = a->src_C[i];
    if (C != 16 && C != 32 && C != 64 && C != 128) return;
    all16 = all16 && C == 16;
    wide = wide || C > 32;
  }
  pl->dc = dc ? 1 : 0, pl->nq = 1;
  pl->pair = (!dc && all16 && !(a->W & 1) && a->nsrc * 96 <= 512) ? 1 : 0;
  if (dc) {  // dZ rows are hi-res pixel PAIRS: 2*Cout elements, at most 64 per box
    pl->ecout = 2 * a->cout > 64 ? 64 : 2 * a->cout;
    pl->nq = 2 * a->cout / pl->ecout;
    pl->ncols = 2 * pl->ecout;
  } else {
    pl->ecout = pl->pair ? 32 : (a->cout > 64 ? 64 : a->cout);
    pl->ncols = 3 * pl->ecout;
  }
  pl->We = pl->pair ? a->W / 2 : a->W;
  // chunks
  int off = 0, nchunks = 0, maxc = 0;
  for (int i = 0; i < a->nsrc; ++i) {
    const int C = a->src_C[i];
    pl->box_C[i] = pl->pair ? 32 : (C > 64 ? 64 : C);
    for (int c0 = 0; c0 < C; c0 += 64) {
      if (nchunks >= kMaxChunks) return;
      pl->ch_map[nchunks] = i, pl->ch_c0[nchunks] = c0, pl->ch_C[nchunks] = pl->box_C[i], pl->ch_cioff[nchunks] = off + c0;
      if (pl->box_C[i] > maxc) maxc = pl->box_C[i];
      ++nchunks;
    }
    off += C;
  }
  pl->cin_total = off, pl->nchunks = nchunks;
  // jobs: narrow 3x3 layers = one job with every chunk; wide layers = one job per (chunk, output slice); transposed conv = one per chunk
  int njobs = 0, max_acc = 0;
  if (dc) {
    for (int c = 0; c < nchunks; ++c) pl->job_ch0[njobs] = c, pl->job_nch[njobs] = 1, pl->job_co0[njobs] = 0, ++njobs;
    max_acc = pl->nq;
  } else if (!wide) {
    pl->job_ch0[0] = 0, pl->job_nch[0] = nchunks, pl->job_co0[0] = 0;
    njobs = 1, max_acc = nchunks;
  } else {
    for (int c = 0; c < nchunks; ++c)
      for (int co0 = 0; co0 < a->cout; co0 += 64) {
        if (njobs >= kMaxJobs) return;
        pl->job_ch0[njobs] = c, pl->job_nch[njobs] = 1, pl->job_co0[njobs] = co0;
        ++njobs;
        const int nm = pl->ch_C[c] == 64 ? 2 : 1;
        if (nm > max_acc) max_acc = nm;
      }
  }
  pl->njobs = njobs;
  if (max_acc * pl->ncols > 512) return;
  int ns = 512 / (max_acc * pl->ncols);
  pl->nsplit = ns > kIssuers ? kIssuers : ns;
  int need = max_acc * pl->nsplit * pl->ncols, tc = 32;
  while (tc < need) tc <<= 1;
  pl->tmem_cols = tc;
  int TR = (maxc == 64 || pl->ecout == 64) ? 8 : 16;
  if (dc) TR = 16 / (pl->ecout / 32) / pl->nq;  // keeps the dZ slot (2*TR hi-res rows per box) at 64 KB
  pl->TR = TR;
  const int xpx = dc ? TW : PX, zrows = dc ? 2 * TR : TR + 2;
  pl->xslot = (TR * xpx * maxc * 2 + 1023) / 1024 * 1024;
  pl->zslot = pl->nq * ((zrows * TW * pl->ecout * 2 + 1023) / 1024 * 1024);
  const int budget = 212 * 1024;
  pl->nz = pl->zslot <= 20 * 1024 ? 3 : 2;
  pl->nx = (budget - pl->nz * pl->zslot) / pl->xslot;
  if (pl->nx > kMaxX) pl->nx = kMaxX;
  if (pl->nx < 2) return;
  pl->smem_total = 1024 + pl->nz * pl->zslot + pl->nx * pl->xslot + 1024;  // tail: the discarded shifts of the last row read past the slot
  pl->tiles_x = (pl->We + TW - 1) / TW, pl->tiles_y = (a->H + TR - 1) / TR;
  pl->ntiles = pl->tiles_x * pl->tiles_y * a->N;
  int gx = unpp::num_sms() / njobs;
  if (gx < 1) gx = 1;
  pl->grid_x = pl->ntiles < gx ? pl->ntiles : gx;
  pl->ok = true;
}

}  // namespace

namespace unpp {

bool wgrad_tc_eligible(const UnppWgradArgs* a) {
  Plan pl;
  make_plan(a, &pl);
  return pl.ok;
}

int wgrad_tc_grid(const UnppWgradArgs* a) {
  Plan pl;
  make_plan(a, &pl);
  return pl.grid_x * (pl.pair ? 2 : 1);  // partials written (the pair view writes two per CTA)
}

int wgrad_tc_launch(const UnppWgradArgs* a, cudaStream_t stream) {
  Plan pl;
  make_plan(a, &pl);
  if (!pl.ok) return unpp::fail(UNPP_ERR_UNSUPPORTED, "wgrad_tc: shape not supported by the tcgen05 path");
  EncodeTiledFn enc = get_encode();
  if (!enc) return unpp::fail(UNPP_ERR_CUDA, "wgrad_tc: cuTensorMapEncodeTiled not available from the driver");
  Params p;
  memset(&p, 0, sizeof p);
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < a->nsrc; ++i) {
    const cuuint64_t C = pl.pair ? 32 : a->src_C[i];  // channels per K row of the tensor as TMA sees it
    cuuint64_t gd[4] = {C, cuuint64_t(pl.We), cuuint64_t(a->H), cuuint64_t(a->N)};
    cuuint64_t gs[3] = {C * 2, cuuint64_t(pl.We) * C * 2, cuuint64_t(a->H) * pl.We * C * 2};
    cuuint32_t box[4] = {cuuint32_t(pl.box_C[i]), cuuint32_t(pl.dc ? TW : PX), cuuint32_t(pl.TR), 1};
    CUresult r = enc(&p.xmap[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->src[i]), gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_of(pl.box_C[i]), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return unpp::fail(UNPP_ERR_CUDA, "wgrad_tc: cuTensorMapEncodeTiled failed (CUresult %d) for source %d", int(r), i);
  }
  {
    // 3x3: dZ [N,H,W,Cout] (pair view: [N,H,W/2,32]); transposed conv: dU [N,2H,2W,Cout] viewed as [N,2H,W,2*Cout]
    const cuuint64_t C = pl.dc ? 2 * a->cout : (pl.pair ? 32 : a->cout);
    const cuuint64_t Hz = pl.dc ? 2 * a->H : a->H;
    cuuint64_t gd[4] = {C, cuuint64_t(pl.We), Hz, cuuint64_t(a->N)};
    cuuint64_t gs[3] = {C * 2, cuuint64_t(pl.We) * C * 2, Hz * pl.We * C * 2};
    cuuint32_t box[4] = {cuuint32_t(pl.ecout), cuuint32_t(TW), cuuint32_t(pl.dc ? 2 * pl.TR : pl.TR + 2), 1};
    CUresult r = enc(&p.zmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->dz), gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_of(pl.ecout), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return unpp::fail(UNPP_ERR_CUDA, "wgrad_tc: cuTensorMapEncodeTiled failed (CUresult %d) for dZ", int(r));
  }
  for (int c = 0; c < pl.nchunks; ++c) p.ch_map[c] = pl.ch_map[c], p.ch_c0[c] = pl.ch_c0[c], p.ch_C[c] = pl.ch_C[c], p.ch_cioff[c] = pl.ch_cioff[c];
  for (int j = 0; j < pl.njobs; ++j) p.job_ch0[j] = pl.job_ch0[j], p.job_nch[j] = pl.job_nch[j], p.job_co0[j] = pl.job_co0[j];
  p.tiles_x = pl.tiles_x, p.tiles_y = pl.tiles_y, p.ntiles = pl.ntiles;
  p.cin_total = pl.cin_total, p.cout = a->cout;
  p.pair = pl.pair, p.zspan = pl.ecout * 2, p.TR = pl.TR;
  p.xslot = pl.xslot, p.zslot = pl.zslot, p.nx = pl.nx, p.nz = pl.nz;
  p.nsplit = pl.nsplit, p.ncols = pl.ncols, p.tmem_cols = pl.tmem_cols;
  p.dc = pl.dc, p.nq = pl.nq;
  p.xpx = pl.dc ? TW : PX, p.xorg = pl.dc ? 0 : -1;
  p.zrows = pl.dc ? 2 * pl.TR : pl.TR + 2, p.zy_mul = pl.dc ? 2 : 1, p.zy_org = pl.dc ? 0 : -1, p.zrow_mul = pl.dc ? 2 : 1;
  p.zbox_bytes = pl.zslot / pl.nq;
  p.partial = a->partial;
  static unsigned char opted_in[64] = {0};  // per device
  if (cudaError_t e = unpp::opt_in_smem(wgrad_tc_kernel, 224 * 1024, opted_in)) return unpp::fail_cuda_err("wgrad_tc: cudaFuncSetAttribute", e);
  if (cudaError_t e = unpp::launch(wgrad_tc_kernel, dim3(pl.grid_x, pl.njobs), kThreads, pl.smem_total, stream, p)) return unpp::fail_cuda_err("wgrad_tc: launch", e);
  return UNPP_OK;
}

}  // namespace unpp

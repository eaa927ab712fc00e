// b2_blocks.h — the non-zero weight blocks of the 2x2 output-blocked convolution, shared by the weight packer
// (aux_kernels.cu) and the MMA issue loops (conv_tc.cu).
//
// With 2x2 output blocking the GEMM columns are (pixel q = 2*py + px of the block, output channel) and K walks the 4x4
// input window; window position (dy, dx) only reaches the pixels with 0 <= dy - py <= 2 and 0 <= dx - px <= 2, so 28 of
// the 64 (position, pixel) weight blocks are identically zero.  Instead of one N = 64 MMA per position the kernel issues
// one MMA per RUN of adjacent non-zero pixel blocks (N = 16, 32 or 64 columns, written at column 16 * b0 of the
// accumulator): 20 MMAs / 36 blocks per 16 input channels instead of 16 MMAs / 64 blocks — 44 % fewer weight bytes in
// shared memory (room for more TMA stages) at the same tensor-pipe time.  The fused transposed conv (3x3 taps over the
// low-resolution tensor, composed weights) has the same structure: 11 MMAs / 16 blocks per 16 channels instead of 9 / 36.
// The first entry of each table covers all four pixels, so it is the one MMA that may overwrite the accumulator.
#pragma once

namespace b2 {

struct Blk {
  int pos;   // window position dy * 4 + dx (main) or low-resolution tap r * 3 + s (fused transposed conv)
  int b0;    // first 16-column block of the accumulator the MMA writes (= first pixel q of the run)
  int nblk;  // pixels in the run: N = 16 * nblk
  int cum;   // 16-column blocks before this one in the packed layout
};

constexpr int kMainBlks = 20, kMainUnits = 36, kLowBlks = 11, kLowUnits = 16;
constexpr int kUnitBytes = 512;  // one 16-column block of one K = 16 slab: [2 k8][16 columns][8 channels] bf16

__host__ __device__ constexpr Blk main_blk(int i) {
  constexpr Blk t[kMainBlks] = {{5, 0, 4, 0},   {0, 0, 1, 4},   {1, 0, 2, 5},   {2, 0, 2, 7},   {3, 1, 1, 9},   {4, 0, 1, 10},  {4, 2, 1, 11},
                                {6, 0, 4, 12},  {7, 1, 1, 16},  {7, 3, 1, 17},  {8, 0, 1, 18},  {8, 2, 1, 19},  {9, 0, 4, 20},  {10, 0, 4, 24},
                                {11, 1, 1, 28}, {11, 3, 1, 29}, {12, 2, 1, 30}, {13, 2, 2, 31}, {14, 2, 2, 33}, {15, 3, 1, 35}};
  return t[i];
}
__host__ __device__ constexpr Blk low_blk(int i) {
  constexpr Blk t[kLowBlks] = {{4, 0, 4, 0}, {0, 0, 1, 4},  {1, 0, 2, 5},  {2, 1, 1, 7},  {3, 0, 1, 8}, {3, 2, 1, 9},
                               {5, 1, 1, 10}, {5, 3, 1, 11}, {6, 2, 1, 12}, {7, 2, 2, 13}, {8, 3, 1, 15}};
  return t[i];
}

// Per 16-column unit of the packed layout: the block it belongs to (for the weight packer, which walks units).
template <int UNITS>
struct UnitTable {
  Blk of[UNITS];
};
__host__ __device__ constexpr UnitTable<kMainUnits> main_units() {
  UnitTable<kMainUnits> t{};
  for (int i = 0; i < kMainBlks; ++i)
    for (int j = 0; j < main_blk(i).nblk; ++j) t.of[main_blk(i).cum + j] = main_blk(i);
  return t;
}
__host__ __device__ constexpr UnitTable<kLowUnits> low_units() {
  UnitTable<kLowUnits> t{};
  for (int i = 0; i < kLowBlks; ++i)
    for (int j = 0; j < low_blk(i).nblk; ++j) t.of[low_blk(i).cum + j] = low_blk(i);
  return t;
}

}  // namespace b2

// Tensor-core weight gradient of the 3x3 convolutions and (taps = 1) of the 1x1 heads (fit(),
// base_model.py:153-162 applied to the conv kernels of simple_fcn.py:39-79):
//
//     dW[tap][ci][co] = sum over pixels p   x[p + shift(tap)][ci] * dy[p][co]
//
// GEMM view with K = pixels:  D[co, (tap, ci)] += dY^T[co, p] * Xshift[p, (tap, ci)].  Both
// operands are stored pixel-major with the channels contiguous (NHWC), i.e. "MN-major" for this
// product, which tcgen05.mma reads directly (a_major = b_major = 1): a TMA box
// {64 ch, TW, TH, 1} lands as 128 pixel rows of 128 B, exactly the 128B-swizzled MN-major atom
// (64 channels x 8 pixel rows per 1024 B).  The tap shift and the zero padding are again done by
// the TMA coordinates.  One CTA accumulates a [128 co x up to 256 (tap,ci)] block over its share of
// the pixel tiles in TMEM and adds it to the fp32 gradient with red.global.add.
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kBoxBytes = 128 * 128;                 // one {64 ch x 128 px} box
constexpr int kStages = 2;
constexpr int kStageBytes = 6 * kBoxBytes;           // 2 dy atoms + 4 x atoms
constexpr int kThreads = 192;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;

// MN-major, 128-byte-swizzled operand: LBO = distance between 64-channel atoms, SBO = distance
// between groups of 8 K rows (pixels).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes,
                                                        uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ ConvWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* done_bar = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // task decode: blockIdx.x = (m_block, n_group, k_split)
  int task = blockIdx.x;
  const int split = task % p.k_splits;
  task /= p.k_splits;
  const int n_group = task % p.n_groups;
  const int m_block = task / p.n_groups;
  const int atom0 = n_group * 4;
  const int n_atoms = min(4, p.total_atoms - atom0);
  const int cin_chunks = p.cin / 64;
  const int total_tiles = p.N * p.tiles_y * p.tiles_x;
  const int my_tiles = (total_tiles - split + p.k_splits - 1) / p.k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_dy);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    uint32_t stage = 0, phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = split + i * p.k_splits;
      const int tx = tile % p.tiles_x;
      const int rest = tile / p.tiles_x;
      const int ty = rest % p.tiles_y;
      const int img = rest / p.tiles_y;
      const int y0 = ty * p.th, x0 = tx * p.tw;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one_sync()) {
        uint8_t* base = smem + stage * kStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], (2 + n_atoms) * kBoxBytes);
        for (int a = 0; a < 2; ++a)      // dy atoms: channels m0 + 64 a (beyond Cout: zero fill)
          tma_load_4d(base + a * kBoxBytes, &p.tmap_dy, &full_bar[stage],
                      m_block * 128 + a * 64, x0, y0, img);
        for (int a = 0; a < n_atoms; ++a) {
          const int atom = atom0 + a;
          const int tap = atom / cin_chunks, cc = atom - tap * cin_chunks;
          const int sx = p.taps == 1 ? 0 : tap % 3 - 1, sy = p.taps == 1 ? 0 : tap / 3 - 1;
          tma_load_4d(base + (2 + a) * kBoxBytes, &p.tmap_x, &full_bar[stage], cc * 64, x0 + sx,
                      y0 + sy, img);
        }
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    // bf16 x bf16 -> fp32, A and B MN-major, M = 128, N = 64 * n_atoms
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           (static_cast<uint32_t>((64 * n_atoms) >> 3) << 17) |
                           (static_cast<uint32_t>(128 >> 4) << 24);
    uint32_t stage = 0, phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
        const uint32_t b_addr = a_addr + 2 * kBoxBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {     // 128 pixels per stage = 8 steps of K = 16
          umma_bf16(tmem_base, umma_desc_sw128_mn(a_addr + k * 2048, kBoxBytes, 1024),
                    umma_desc_sw128_mn(b_addr + k * 2048, kBoxBytes, 1024), idesc,
                    (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (i == my_tiles - 1) umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (my_tiles > 0) {
    // ------------------------------------------------------------- epilogue: TMEM -> red.add
    const int q = warp & 3;
    const int co = m_block * 128 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int a = 0; a < n_atoms; ++a) {
      const int atom = atom0 + a;
      const int tap = atom / cin_chunks, cc = atom - tap * cin_chunks;
      float* dst = p.dw + (static_cast<size_t>(tap) * p.cin + cc * 64) * p.cout + co;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + a * 64 + half * 32, r);
        tmem_ld_wait();
        if (co < p.cout) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            atomicAdd(dst + static_cast<size_t>(half * 32 + j) * p.cout, __uint_as_float(r[j]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

int launch_conv_wgrad_tc(const ConvWgradParams& p, cudaStream_t stream) {
  XV_CHECK(p.cin % 64 == 0, "conv_wgrad_tc: Cin must be a multiple of 64");
  XV_CHECK(p.th * p.tw == 128, "conv_wgrad_tc: tile must hold 128 pixels");
  XV_CHECK(p.taps == 9 || p.taps == 1, "conv_wgrad_tc: 3x3 or 1x1 filters");
  static bool configured = false;
  if (!configured) {
    XV_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmemBytes));
    configured = true;
  }
  const int grid = p.m_blocks * p.n_groups * p.k_splits;
  conv_wgrad_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace xv

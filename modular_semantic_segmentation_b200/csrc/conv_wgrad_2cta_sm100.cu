// CTA-pair (tcgen05 cta_group::2) weight gradient for the Cout % 256 == 0, Cin % 128 == 0 layers
// (conv3_1 .. conv5_3 of simple_fcn.py:45-67 in fit()): same GEMM as conv_wgrad_sm100.cu,
//
//     D[co, (tap, ci)] += dY^T[co, p] * Xshift[p, (tap, ci)]        (K = pixels, MN-major operands)
//
// but one MMA covers M = 256 output channels: each CTA of the pair loads the dy operand of ITS
// 128 channels and only HALF of the x operand (two of the four (tap, ci) atoms); the tensor cores
// read the other half from the peer's shared memory.
//
// Why: ncu on the single-CTA kernel shows the tensor pipe 64-68 % active with L2 and DRAM far
// from their peaks - a CTA needs 96 KB per 128-pixel step (1120 cycles of MMA work, 86 B / cycle),
// more than one SM takes in.  Here a step is 64 KB per CTA (58 B / cycle) and three steps fit in
// the ring instead of two.
//
// Protocol (barriers at the same shared-memory offsets in both CTAs, as in
// conv_igemm_2cta_sm100.cu): full[s] is the leader's, armed with the bytes of BOTH CTAs' loads
// (.cta_group::2 TMA loads route their completion to the leader); empty[s] and done exist per CTA
// and are released by one multicast tcgen05.commit of the leader's MMA warp.
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kBoxBytes = 128 * 128;                 // one {64 ch x 128 px} box
constexpr int kStages = 3;
constexpr int kStageBytes = 4 * kBoxBytes;           // 2 dy atoms + 2 of the 4 x atoms
constexpr int kThreads = 192;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;       // shared::cluster address -> same offset in CTA 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint64_t* bar,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// MN-major, 128-byte-swizzled operand (see conv_wgrad_sm100.cu)
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_wgrad_2cta_kernel(const __grid_constant__ ConvWgradParams p) {
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
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  // task decode: pair index = (m_pair, n_group, k_split); m_blocks counts PAIRS of 128 channels
  int task = blockIdx.x >> 1;
  const int split = task % p.k_splits;
  task /= p.k_splits;
  const int n_group = task % p.n_groups;
  const int m_pair = task / p.n_groups;
  const int atom0 = n_group * 4;
  const int n_atoms = min(4, p.total_atoms - atom0);         // even (Cin % 128 == 0)
  const int my_atoms = n_atoms / 2;                          // x atoms this CTA loads
  const int cin_chunks = p.cin / 64;
  const int total_tiles = p.N * p.tiles_y * p.tiles_x;
  const int my_tiles = (total_tiles - split + p.k_splits - 1) / p.k_splits;
  const int co0 = (2 * m_pair + rank) * 128;                 // this CTA's output channels

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
  if (warp == 1) tmem_alloc_pair(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (both CTAs)
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
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (2 + my_atoms) * kBoxBytes);
        for (int a = 0; a < 2; ++a)
          tma_load_4d_pair(base + a * kBoxBytes, &p.tmap_dy, &full_bar[stage], co0 + a * 64, x0,
                           y0, img);
        for (int a = 0; a < my_atoms; ++a) {
          const int atom = atom0 + rank * my_atoms + a;
          const int tap = atom / cin_chunks, cc = atom - tap * cin_chunks;
          const int sx = tap % 3 - 1, sy = tap / 3 - 1;
          tma_load_4d_pair(base + (2 + a) * kBoxBytes, &p.tmap_x, &full_bar[stage], cc * 64,
                           x0 + sx, y0 + sy, img);
        }
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (leader only)
    if (leader) {
      // bf16 x bf16 -> fp32, A and B MN-major, M = 256 (128 per CTA), N = 64 * n_atoms
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             (static_cast<uint32_t>((64 * n_atoms) >> 3) << 17) |
                             (static_cast<uint32_t>(256 >> 4) << 24);
      uint32_t stage = 0, phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_addr = a_addr + 2 * kBoxBytes;
#pragma unroll
          for (int k = 0; k < 8; ++k) {     // 128 pixels per stage = 8 steps of K = 16
            umma_bf16_pair(tmem_base, umma_desc_sw128_mn(a_addr + k * 2048, kBoxBytes, 1024),
                           umma_desc_sw128_mn(b_addr + k * 2048, kBoxBytes, 1024), idesc,
                           (i | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage]);
          if (i == my_tiles - 1) umma_commit_pair(done_bar);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (my_tiles > 0) {
    // ------------------------------------------------------------- epilogue: TMEM -> red.add
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int a = 0; a < n_atoms; ++a) {
      // accumulator columns: the leader's x atoms first, then the peer's
      const int atom = atom0 + a;
      const int tap = atom / cin_chunks, cc = atom - tap * cin_chunks;
      float* dst = p.dw + (static_cast<size_t>(tap) * p.cin + cc * 64) * p.cout + co;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + a * 64 + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          atomicAdd(dst + static_cast<size_t>(half * 32 + j) * p.cout, __uint_as_float(r[j]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();             // nobody leaves while the peer may still signal or read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 256);
  }
}

}  // namespace

// Same parameters as launch_conv_wgrad_tc, except m_blocks = Cout / 256 (pairs of 128-channel
// blocks); needs Cout % 256 == 0, Cin % 128 == 0, 3x3 filters.
int launch_conv_wgrad_2cta(const ConvWgradParams& p, cudaStream_t stream) {
  XV_CHECK(p.cin % 128 == 0 && p.cout % 256 == 0 && p.taps == 9,
           "conv_wgrad_2cta: Cin % 128, Cout % 256 and 3x3 filters required");
  XV_CHECK(p.th * p.tw == 128, "conv_wgrad_2cta: tile must hold 128 pixels");
  static bool configured = false;
  if (!configured) {
    XV_CUDA(cudaFuncSetAttribute(conv_wgrad_2cta_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  const int grid = 2 * p.m_blocks * p.n_groups * p.k_splits;
  conv_wgrad_2cta_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace xv

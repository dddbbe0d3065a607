// 2-CTA-cluster variant of the pixel-major implicit-GEMM convolution (BLOCK_N = 256, 3x3 / 1x1,
// bf16 epilogue): the two CTAs of a cluster work on two different 128-pixel tiles against the SAME
// 256 output channels, and each loads only HALF of the 32 KB weight tile of a K block, multicast
// into both CTAs' shared memory (cp.async.bulk.tensor ... .multicast::cluster).  L2 -> SM operand
// traffic per CTA and K block drops from 48 KB to 32 KB; measured on B200 the Cout >= 256 layers
// were limited by exactly that traffic (~87 B/clk/SM).
//
// Protocol per smem stage: full barrier (1 arrival = own producer, 48 KB of transactions: own
// activation box + both weight halves), empty barrier (2 arrivals = tcgen05.commit of BOTH CTAs,
// multicast), because a stage's weight buffer is written by both producers.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;     // 16 KB
constexpr int kBBytes = kBlockN * kBlockK * 2;     // 32 KB (two 16 KB halves)
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kStages = 4;
constexpr int kOutBufBytes = kBlockM * 128;
constexpr int kThreads = 192;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 2 * kOutBufBytes + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                               int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <int TAPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_igemm_mc_kernel(const __grid_constant__ ConvIgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint8_t* smem_out = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + 2 * kOutBufBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int m_tiles = p.N * p.tiles_y * p.tiles_x;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int total_units = m_pairs * p.n_blocks;
  const int cin_chunks = p.cin / kBlockK;
  const int num_kb = TAPS * cin_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_in);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_out);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 2);      // MMA commits of both CTAs of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();                   // barriers of BOTH CTAs are initialised before any traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (image, y0, x0, n0, valid): CTA `rank` takes M tile 2*pair + rank
  auto decode = [&](int unit, int& img, int& y0, int& x0, int& n0, bool& valid) {
    const int nb = unit % p.n_blocks;
    int mt = 2 * (unit / p.n_blocks) + rank;
    valid = mt < m_tiles;
    if (!valid) mt = m_tiles - 1;       // odd tile count: the partner recomputes the last tile
    const int tx = mt % p.tiles_x;
    const int rest = mt / p.tiles_x;
    const int ty = rest % p.tiles_y;
    img = rest / p.tiles_y;
    y0 = ty * p.th;
    x0 = tx * p.tw;
    n0 = nb * kBlockN;
  };

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    uint32_t stage = 0, phase = 0;
    for (int unit = cluster_id; unit < total_units; unit += num_clusters) {
      int img, y0, x0, n0;
      bool valid;
      decode(unit, img, y0, x0, n0, valid);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / cin_chunks;
        const int cc = kb - tap * cin_chunks;
        const int dy = (TAPS == 9) ? tap / 3 - 1 : 0;
        const int dx = (TAPS == 9) ? tap % 3 - 1 : 0;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_4d(smem_a + stage * kABytes, &p.tmap_in, &full_bar[stage], cc * kBlockK, x0 + dx,
                      y0 + dy, img);
          // my half of the weight tile (128 of the 256 output channels) goes to both CTAs
          tma_load_2d_mc(smem_b + stage * kBBytes + rank * (kBBytes / 2), &p.tmap_w,
                         &full_bar[stage], tap * p.cin + cc * kBlockK, n0 + rank * 128,
                         static_cast<uint16_t>(3));
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kBlockN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int unit = cluster_id; unit < total_units; unit += num_clusters) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kBlockN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32, 1024, 0),
                      umma_desc_sw128(b_addr + k * 32, 1024, 0), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_mc(&empty_bar[stage], static_cast<uint16_t>(3));   // frees the stage in both CTAs
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------------------------------------- epilogue (128 threads)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool issuer = (threadIdx.x == 64);
    uint32_t acc = 0, acc_phase = 0, gchunk = 0;
    for (int unit = cluster_id; unit < total_units; unit += num_clusters) {
      int img, y0, x0, n0;
      bool valid;
      decode(unit, img, y0, x0, n0, valid);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBlockN;
#pragma unroll 1
      for (int chunk = 0; chunk < kBlockN / 64; ++chunk, ++gchunk) {
        uint8_t* buf = smem_out + (gchunk & 1) * kOutBufBytes;
        if (issuer) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + chunk * 64 + half * 32, r);
          tmem_ld_wait();
          const float* bias = p.bias + n0 + chunk * 64 + half * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t packed[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v0 = __uint_as_float(r[j * 8 + e * 2]) + __ldg(bias + j * 8 + e * 2);
              float v1 = __uint_as_float(r[j * 8 + e * 2 + 1]) + __ldg(bias + j * 8 + e * 2 + 1);
              if (p.relu) {
                v0 = fmaxf(v0, 0.f);
                v1 = fmaxf(v1, 0.f);
              }
              packed[e] = pack_bf16x2(v0, v1);
            }
            const int piece = (half * 4 + j) ^ (row & 7);
            *reinterpret_cast<uint4*>(buf + row * 128 + piece * 16) =
                make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (issuer && valid) {
          tma_store_4d(&p.tmap_out, buf, n0 + chunk * 64, x0, y0, img);
          tma_store_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();                   // the partner may still multicast into / arrive on my smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int TAPS>
int launch_mc(const ConvIgemmParams& p, cudaStream_t stream) {
  auto kernel = conv_igemm_mc_kernel<TAPS>;
  static bool configured = false;
  if (!configured) {
    XV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  const int m_tiles = p.N * p.tiles_y * p.tiles_x;
  const int units = ((m_tiles + 1) / 2) * p.n_blocks;
  int clusters = device_info().num_sms / 2;
  if (units < clusters) clusters = units;
  kernel<<<2 * clusters, kThreads, kSmemBytes, stream>>>(p);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

// Same parameters as launch_conv_igemm with BLOCK_N = 256, except tmap_w has box {64, 128}.
int launch_conv_igemm_mc(const ConvIgemmParams& p, int taps, cudaStream_t stream) {
  XV_CHECK(p.th * p.tw == kBlockM, "conv_igemm_mc: tile must hold 128 pixels");
  XV_CHECK(p.cin % kBlockK == 0, "conv_igemm_mc: Cin must be a multiple of 64");
  XV_CHECK(p.cout % 64 == 0, "conv_igemm_mc: Cout must be a multiple of 64");
  if (taps == 9) return launch_mc<9>(p, stream);
  if (taps == 1) return launch_mc<1>(p, stream);
  return fail("conv_igemm_mc: only 1x1 and 3x3 kernels");
}

}  // namespace xv

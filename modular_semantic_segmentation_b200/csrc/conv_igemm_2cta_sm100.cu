// CTA-pair (tcgen05 cta_group::2) 3x3 convolution for the Cout >= 256 layers (conv3_x .. conv5_x of
// simple_fcn.py:45-67), bf16 NHWC in / out, + bias + ReLU.
//
// Why: the single-CTA kernel is bound by shared-memory bandwidth - per K block the TMA writes the
// 32 KB weight slice and the four MMAs read 48 KB of operands against 128 B / cycle.  Here the two
// CTAs of a cluster (one TPC) issue ONE tcgen05.mma.cta_group::2 of shape M = 256 (two 128-pixel
// tiles, one per CTA) x N = 256 output channels x K = 16: each CTA loads its own pixel operand and
// only HALF of the weight slice (16 KB per K block); the tensor cores read the other half from the
// peer's shared memory.  Pixel operands use the halo scheme of conv_igemm_sm100.cu (three
// column-shifted (TH + 2) x TW patch copies per input-channel chunk).
//
// Protocol (all barriers live at the same shared-memory offsets in both CTAs):
//   a_full / b_full    only the leader's (cluster rank 0) are used: the leader's producer arrives
//                      with expect_tx for BOTH CTAs' bytes, and both CTAs' TMA loads are the
//                      .cta_group::2 form whose completion is routed to the leader's barrier.
//   a_empty / b_empty  per CTA; the leader's MMA warp releases a stage in both CTAs with one
//                      multicast tcgen05.commit.
//   tmem_full          per CTA, multicast commit after the last K block of a tile pair.
//   tmem_empty         the leader's; 8 arrivals = 4 epilogue warps of each CTA (the peer arrives
//                      remotely through mapa).
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kBlockM = 128;                      // pixels per CTA
constexpr int kBlockK = 64;
constexpr int kACopies = 3;
constexpr int kACopyBytes = 20480;                // (8 + 2) x 16 or (16 + 2) x 8 pixel rows
constexpr int kBStages = 6;
constexpr int kOutBufBytes = kBlockM * 128;
constexpr int kPoolBufBytes = (kBlockM / 4) * 128;  // 32 pooled pixels x 64 channels
constexpr int kThreads = 192;
// BLOCK_N = output channels per tile pair (256, or 128 / 64 for the narrow layers); every CTA holds
// BLOCK_N / 2 weight rows of a K block
template <int BLOCK_N>
struct PairCfg {
  static constexpr int kBHalfBytes = (BLOCK_N / 2) * 128;
  static constexpr int kSmemBytes = 1024 + kACopies * kACopyBytes + kBStages * kBHalfBytes +
                                    2 * kOutBufBytes + 2 * kPoolBufBytes + 256;
};
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;    // shared::cluster address -> same offset in CTA 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of a CTA pair: data into the executing CTA's shared memory, transaction bytes onto the
// LEADER's mbarrier (peer bit of the barrier address cleared).
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint64_t* bar,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint64_t* bar,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
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
// D[tmem, both CTAs] (+)= A[smem, per CTA] * B[smem, half per CTA]; issued by the leader only.
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all prior MMAs of the pair have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// arrive on the barrier at this offset in the LEADER CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

struct PairTile {
  int img, y0, x0, n0;
};

__device__ __forceinline__ PairTile decode_pair(const ConvIgemmParams& p, int pt, int rank,
                                                int block_n) {
  PairTile c;
  const int nb = pt % p.n_blocks;
  const int mt = 2 * (pt / p.n_blocks) + rank;           // this CTA's 128-pixel tile
  const int tx = mt % p.tiles_x;
  const int rest = mt / p.tiles_x;
  const int ty = rest % p.tiles_y;
  c.img = rest / p.tiles_y;       // == N for the padding half of an odd last pair: every TMA box is
  c.y0 = ty * p.th;               // then out of bounds (zero-filled loads, clipped stores)
  c.x0 = tx * p.tw;
  c.n0 = nb * block_n;
  return c;
}

// POOL: 0 = plain convolution; 1 = only the 2x2 max-pooled output is stored (conv3_3 -> pool3,
// simple_fcn.py:48); 2 = both the full and the pooled output (conv4_3 feeds score_conv4 AND pool4,
// simple_fcn.py:58,74).  The pool is taken on the packed bf16 results with two warp shuffles: a
// warp's 32 TMEM lanes are 4 x 8 or 2 x 16 pixels of the tile, so the 2x2 partners of a pixel are
// the lanes at xor 1 (x) and xor tw (y).
template <int BLOCK_N, int POOL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_igemm_2cta_kernel(const __grid_constant__ ConvIgemmParams p) {
  constexpr int kBlockN = BLOCK_N;
  constexpr int kBHalfBytes = PairCfg<BLOCK_N>::kBHalfBytes;
  constexpr int kTmemCols = 2 * BLOCK_N;        // two accumulator stages (128 / 256 / 512)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + kACopies * kACopyBytes;
  uint8_t* smem_out = smem_b + kBStages * kBHalfBytes;
  uint8_t* smem_pool = smem_out + 2 * kOutBufBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_pool + 2 * kPoolBufBytes);
  uint64_t* b_full = bars;
  uint64_t* b_empty = b_full + kBStages;
  uint64_t* a_full = b_empty + kBStages;
  uint64_t* a_empty = a_full + kACopies;
  uint64_t* tmem_full_bar = a_empty + kACopies;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int pixel_tiles = p.N * p.tiles_y * p.tiles_x;
  const int total_pairs = ((pixel_tiles + 1) / 2) * p.n_blocks;
  const int cin_chunks = p.cin / kBlockK;
  const uint32_t copy_bytes = static_cast<uint32_t>((p.th + 2) * p.tw) * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_in);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_out);
    if (POOL) tma_prefetch_desc(&p.tmap_pool);
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < kACopies; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();             // the peer's barriers exist before anything is signalled remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (both CTAs)
    uint32_t bs = 0, bphase = 0, as = 0, aphase = 0;
    for (int pt = pair_id; pt < total_pairs; pt += num_pairs) {
      const PairTile c = decode_pair(p, pt, rank, kBlockN);
      for (int cc = 0; cc < cin_chunks; ++cc) {
        for (int dxi = 0; dxi < 3; ++dxi) {
          mbar_wait(&a_empty[as], aphase ^ 1);
          if (elect_one_sync()) {
            if (leader) mbar_arrive_expect_tx(&a_full[as], 2 * copy_bytes);
            tma_load_4d_pair(smem_a + as * kACopyBytes, &p.tmap_in, &a_full[as], cc * kBlockK,
                             c.x0 + dxi - 1, c.y0 - 1, c.img);
          }
          __syncwarp();
          if (++as == kACopies) {
            as = 0;
            aphase ^= 1;
          }
          for (int dyi = 0; dyi < 3; ++dyi) {
            mbar_wait(&b_empty[bs], bphase ^ 1);
            if (elect_one_sync()) {
              if (leader) mbar_arrive_expect_tx(&b_full[bs], 2 * kBHalfBytes);
              tma_load_2d_pair(smem_b + bs * kBHalfBytes, &p.tmap_w, &b_full[bs],
                               (dyi * 3 + dxi) * p.cin + cc * kBlockK, c.n0 + rank * (kBlockN / 2));
            }
            __syncwarp();
            if (++bs == kBStages) {
              bs = 0;
              bphase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (leader only)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBlockM, kBlockN);
      uint32_t bs = 0, bphase = 0, as = 0, aphase = 0, acc = 0, acc_phase = 0;
      for (int pt = pair_id; pt < total_pairs; pt += num_pairs) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlockN;
        int first = 1;
        for (int cc = 0; cc < cin_chunks; ++cc) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&a_full[as], aphase);
            for (int dyi = 0; dyi < 3; ++dyi) {
              mbar_wait(&b_full[bs], bphase);
              tc_fence_after();
              if (elect_one_sync()) {
                const uint32_t a_addr =
                    smem_u32(smem_a + as * kACopyBytes) + static_cast<uint32_t>(dyi * p.tw) * 128;
                const uint32_t b_addr = smem_u32(smem_b + bs * kBHalfBytes);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  umma_bf16_pair(d_tmem, umma_desc_sw128(a_addr + k * 32, 1024, 0),
                                 umma_desc_sw128(b_addr + k * 32, 1024, 0), idesc,
                                 (first && k == 0) ? 0u : 1u);
                }
                umma_commit_pair(&b_empty[bs]);
                if (dyi == 2) umma_commit_pair(&a_empty[as]);
                if (cc == cin_chunks - 1 && dxi == 2 && dyi == 2)
                  umma_commit_pair(&tmem_full_bar[acc]);
              }
              __syncwarp();
              first = 0;
              if (++bs == kBStages) {
                bs = 0;
                bphase ^= 1;
              }
            }
            if (++as == kACopies) {
              as = 0;
              aphase ^= 1;
            }
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (both CTAs)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool issuer = (threadIdx.x == 64);
    uint32_t acc = 0, acc_phase = 0, gchunk = 0;
    const uint32_t zero2 = 0u;
    // pooled pixel this thread stores (the even-x, even-y pixel of every 2x2 block keeps it)
    const int py = row / p.tw, px = row - py * p.tw;
    const bool keeper = POOL && ((px | py) & 1) == 0;
    const uint32_t prow = static_cast<uint32_t>((py >> 1) * (p.tw >> 1) + (px >> 1));
    for (int pt = pair_id; pt < total_pairs; pt += num_pairs) {
      const PairTile c = decode_pair(p, pt, rank, kBlockN);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBlockN;
#pragma unroll 1
      for (int chunk = 0; chunk < kBlockN / 64; ++chunk, ++gchunk) {
        uint8_t* buf = smem_out + (gchunk & 1) * kOutBufBytes;
        uint8_t* pbuf = smem_pool + (gchunk & 1) * kPoolBufBytes;
        if (issuer) tma_store_wait_read<1>();   // the store(s) that last used `buf` have drained
        named_bar_sync(1, 128);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + chunk * 64 + half * 32, r);
          tmem_ld_wait();
          const float4* bias4 =
              reinterpret_cast<const float4*>(p.bias + c.n0 + chunk * 64 + half * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t packed[4];
            const float4 b_lo = __ldg(bias4 + j * 2), b_hi = __ldg(bias4 + j * 2 + 1);
            const float bb[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v0, v1;
              asm("{\n\t.reg .b64 a, b;\n\t"
                  "mov.b64 a, {%2, %3};\n\t"
                  "mov.b64 b, {%4, %5};\n\t"
                  "add.rn.f32x2 a, a, b;\n\t"
                  "mov.b64 {%0, %1}, a;\n\t}"
                  : "=f"(v0), "=f"(v1)
                  : "r"(r[j * 8 + e * 2]), "r"(r[j * 8 + e * 2 + 1]), "f"(bb[e * 2]),
                    "f"(bb[e * 2 + 1]));
              __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
              if (p.relu) h = __hmax2(h, *reinterpret_cast<const __nv_bfloat162*>(&zero2));
              packed[e] = *reinterpret_cast<uint32_t*>(&h);
            }
            if (POOL != 1) {
              const int piece = (half * 4 + j) ^ (row & 7);   // 128B swizzle
              *reinterpret_cast<uint4*>(buf + row * 128 + piece * 16) =
                  make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
            if (POOL) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&packed[e]);
                uint32_t o = __shfl_xor_sync(0xffffffffu, packed[e], 1);
                v = __hmax2(v, *reinterpret_cast<__nv_bfloat162*>(&o));
                uint32_t vv = *reinterpret_cast<uint32_t*>(&v);
                o = __shfl_xor_sync(0xffffffffu, vv, p.tw);
                v = __hmax2(v, *reinterpret_cast<__nv_bfloat162*>(&o));
                packed[e] = *reinterpret_cast<uint32_t*>(&v);
              }
              if (keeper) {
                const uint32_t piece = static_cast<uint32_t>(half * 4 + j) ^ (prow & 7);
                *reinterpret_cast<uint4*>(pbuf + prow * 128 + piece * 16) =
                    make_uint4(packed[0], packed[1], packed[2], packed[3]);
              }
            }
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (issuer) {
          if (POOL != 1) tma_store_4d(&p.tmap_out, buf, c.n0 + chunk * 64, c.x0, c.y0, c.img);
          if (POOL)
            tma_store_4d(&p.tmap_pool, pbuf, c.n0 + chunk * 64, c.x0 >> 1, c.y0 >> 1, c.img);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                          // one arrival per epilogue warp on the leader
        if (leader) {
          mbar_arrive(&tmem_empty_bar[acc]);
        } else {
          mbar_arrive_leader(&tmem_empty_bar[acc]);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();             // nobody leaves while the peer may still signal or read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace

// p.tmap_in box {64, tw, th + 2, 1}; p.tmap_w box {64, BLOCK_N / 2}; p.tmap_out box
// {64, tw, th, 1}; p.n_blocks = ceil(Cout / BLOCK_N); p.pool_mode != 0: p.tmap_pool describes the
// pooled tensor [N, H/2, W/2, Cout] with box {64, tw/2, th/2, 1}.
template <int BLOCK_N, int POOL = 0>
static int launch_pair(const ConvIgemmParams& p, cudaStream_t stream) {
  auto kernel = conv_igemm_2cta_kernel<BLOCK_N, POOL>;
  static bool configured = false;
  if (!configured) {
    XV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 PairCfg<BLOCK_N>::kSmemBytes));
    configured = true;
  }
  const int pixel_tiles = p.N * p.tiles_y * p.tiles_x;
  const int total_pairs = ((pixel_tiles + 1) / 2) * p.n_blocks;
  const int max_pairs = device_info().num_sms / 2;
  const int pairs = total_pairs < max_pairs ? total_pairs : max_pairs;
  kernel<<<2 * pairs, kThreads, PairCfg<BLOCK_N>::kSmemBytes, stream>>>(p);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_conv_igemm_2cta(const ConvIgemmParams& p, int block_n, cudaStream_t stream) {
  XV_CHECK(p.th * p.tw == kBlockM && (p.tw == 8 || p.tw == 16) &&
               (p.th + 2) * p.tw * 128 <= kACopyBytes,
           "conv_igemm_2cta: needs an 8x16 or 16x8 pixel tile");
  XV_CHECK(p.cin % kBlockK == 0, "conv_igemm_2cta: Cin must be a multiple of 64");
  XV_CHECK(p.pool_mode == 0 || (block_n == 256 && p.H % 2 == 0 && p.W % 2 == 0),
           "conv_igemm_2cta: the fused pool needs BLOCK_N = 256 and even H, W");
  if (block_n == 256 && p.pool_mode == 1) return launch_pair<256, 1>(p, stream);
  if (block_n == 256 && p.pool_mode == 2) return launch_pair<256, 2>(p, stream);
  if (block_n == 256) return launch_pair<256>(p, stream);
  if (block_n == 128) return launch_pair<128>(p, stream);
  if (block_n == 64) return launch_pair<64>(p, stream);
  return fail("conv_igemm_2cta: BLOCK_N must be 64, 128 or 256");
}

}  // namespace xv

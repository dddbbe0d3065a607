// First convolution of an expert: 3x3, Cin <= 3 raw fp32 channels -> 64 channels (conv1_1 of
// simple_fcn.py:39, block_0_1 of adapnet.py:120), + bias + ReLU, bf16 NHWC output.
//
// GEMM view: D[128 pixels, 64] = A[128, K = 64] * W[64, K]^T with the K axis holding the 3x3xCin
// neighbourhood twice - once rounded to bf16 ("hi") and once as the bf16 of the rounding error
// ("lo") - against the same weights, so the product carries ~16 mantissa bits of the raw input
// (depth values reach 65535; a single bf16 would lose the low 8 bits).
//
// Data flow per 128-pixel tile (one CTA per SM, persistent):
//   warp 0        TMA: the fp32 input patch (tile + 1-pixel halo, all channels) -> shared memory.
//                 The image is viewed as [N, H, W*Cin] so that a patch row is one contiguous box;
//                 out-of-image rows / columns are zero-filled by TMA = 'same' padding.  The weight
//                 tile (8 KB) is loaded once and stays resident.
//   warps 10..25  four groups of 128 "packers": thread = pixel, reads its 9*Cin neighbours from the
//                 patch, splits hi / lo and writes its 128-byte A row (SWIZZLE_128B layout).
//   warp 1        4 x tcgen05.mma (M=128, N=64, K=16) into one of two TMEM accumulators.
//   warps 2..9    epilogue (two warps per TMEM lane quarter, 32 channels each): tcgen05.ld ->
//                 bf16, ReLU -> swizzled staging -> TMA store.  A single epilogue warp per quarter
//                 is latency-bound on its own instruction stream (measured), hence two.  The bias is
//                 added by the MMA itself: K columns 62 / 63 of every A row hold 1.0 and the same
//                 columns of the resident weight tile hold the bias split into hi + lo bf16 halves
//                 (exact to 2^-17 of the bias), so the epilogue has no loads at all.
// Every hand-off is an mbarrier ring: patch (8 deep), A rows (6 deep), accumulators (2 deep).
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kBlockM = 128;
constexpr int kCout = 64;
constexpr int kABytes = kBlockM * 128;            // 16 KB operand rows of one tile
constexpr int kWBytes = kCout * 128;              // 8 KB weights, resident
constexpr int kOutBufBytes = kBlockM * 128;       // 16 KB output staging
constexpr int kAStages = 6;
constexpr int kPatchStages = 8;
constexpr int kPatchBytes = 8192;                 // upper bound of one input patch
constexpr int kGroups = 4;                        // packer groups of 128 threads
constexpr int kEpiThreads = 256;                  // warps 2..9
constexpr int kCtrlThreads = 64 + kEpiThreads;    // warps 0..9
constexpr int kThreads = kCtrlThreads + 128 * kGroups;
constexpr int kOutBufs = 3;                       // staging ring: one barrier per tile suffices
constexpr int kSmemBytes = 1024 + kAStages * kABytes + kWBytes + kOutBufs * kOutBufBytes +
                           kPatchStages * kPatchBytes + 512;

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int CIN>
__global__ void __launch_bounds__(kThreads, 1)
conv_c1_kernel(const __grid_constant__ ConvIgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem_a + kAStages * kABytes;
  uint8_t* smem_out = smem_w + kWBytes;
  uint8_t* smem_patch = smem_out + kOutBufs * kOutBufBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_patch + kPatchStages * kPatchBytes);
  uint64_t* full_bar = bars;                                   // A rows written
  uint64_t* empty_bar = full_bar + kAStages;                   // A rows consumed by the MMAs
  uint64_t* patch_full = empty_bar + kAStages;
  uint64_t* patch_empty = patch_full + kPatchStages;
  uint64_t* tmem_full_bar = patch_empty + kPatchStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.N * p.tiles_y * p.tiles_x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_in);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_out);
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kPatchStages; ++s) {
      mbar_init(&patch_full[s], 1);
      mbar_init(&patch_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kEpiThreads / 32);   // one arrival per epilogue warp
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kCout);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int& img, int& y0, int& x0) {
    const int tx = tile % p.tiles_x;
    const int rest = tile / p.tiles_x;
    img = rest / p.tiles_y;
    y0 = (rest - img * p.tiles_y) * p.th;
    x0 = tx * p.tw;
  };

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_bar, kWBytes);
      tma_load_2d(smem_w, &p.tmap_w, w_bar, 0, 0);
    }
    __syncwarp();
    const uint32_t patch_bytes = static_cast<uint32_t>(p.patch_w) * (p.th + 2) * 4;
    uint32_t ps = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int img, y0, x0;
      decode(tile, img, y0, x0);
      mbar_wait(&patch_empty[ps], phase ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&patch_full[ps], patch_bytes);
        // the innermost TMA coordinate must sit on a 16-byte boundary: round the first float of
        // the patch down to a multiple of 4, the packers skip the 0..3 extra floats
        tma_load_3d(smem_patch + ps * kPatchBytes, &p.tmap_in, &patch_full[ps],
                    ((x0 - 1) * CIN) & ~3, y0 - 1, img);
      }
      __syncwarp();
      if (++ps == kPatchStages) {
        ps = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kCout);
    mbar_wait(w_bar, 0);
    // bias -> K columns 62 (hi) and 63 (lo) of the weight rows; element k of row r sits at
    // r * 128 + ((k / 8) ^ (r & 7)) * 16 + (k % 8) * 2 in the 128-byte-swizzled tile
    for (int r = lane; r < kCout; r += 32) {
      const float b = __ldg(p.bias + r);
      const float b_hi = __bfloat162float(__float2bfloat16_rn(b));
      *reinterpret_cast<uint32_t*>(smem_w + r * 128 + ((7 ^ (r & 7)) << 4) + 12) =
          pack_bf16x2(b_hi, b - b_hi);
    }
    fence_proxy_async_smem();
    __syncwarp();
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    const uint32_t w_addr = smem_u32(smem_w);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + acc * kCout, umma_desc_sw128(a_addr + k * 32, 1024, 0),
                    umma_desc_sw128(w_addr + k * 32, 1024, 0), idesc, k != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full_bar[acc]);
      }
      __syncwarp();
      if (++stage == kAStages) {
        stage = 0;
        phase ^= 1;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= kCtrlThreads / 32) {
    // ------------------------------------------------------------- operand packers
    constexpr int K9 = 9 * CIN;
    const int group = (threadIdx.x - kCtrlThreads) >> 7;
    const int row = (threadIdx.x - kCtrlThreads) & 127;      // pixel inside the tile
    const int py = row / p.tw;
    const int px = row - py * p.tw;
    const int pw = p.patch_w;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      if ((local % kGroups) != group) continue;
      const uint32_t ps = static_cast<uint32_t>(local) % kPatchStages;
      const uint32_t pphase = (static_cast<uint32_t>(local) / kPatchStages) & 1u;
      const uint32_t stage = static_cast<uint32_t>(local) % kAStages;
      const uint32_t sphase = (static_cast<uint32_t>(local) / kAStages) & 1u;
      const int x_first = ((tile % p.tiles_x) * p.tw - 1) * CIN;     // float column of pixel x0 - 1
      mbar_wait(&patch_full[ps], pphase);
      if (p.debug_flags & 1) {                  // experiment: no operand packing work
        mbar_wait(&empty_bar[stage], sphase ^ 1);
        named_bar_sync(2 + group, 128);
        if (row == 0) {
          mbar_arrive(&full_bar[stage]);
          mbar_arrive(&patch_empty[ps]);
        }
        continue;
      }
      // patch origin = (y0 - 1, x_first rounded down to 4 floats): neighbour (dy, dx) of pixel
      // (py, px) sits at row py + dy + 1, float column skip + (px + dx + 1) * CIN + ci
      const float* patch = reinterpret_cast<const float*>(smem_patch + ps * kPatchBytes) +
                           py * pw + px * CIN + (x_first - (x_first & ~3));
      float hi[K9], lo[K9];
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
#pragma unroll
        for (int j = 0; j < 3 * CIN; ++j) {               // 3 taps x CIN are contiguous floats
          const float raw = patch[ty * pw + j];
          const float h = __bfloat162float(__float2bfloat16_rn(raw));
          hi[ty * 3 * CIN + j] = h;
          lo[ty * 3 * CIN + j] = raw - h;
        }
      }
      mbar_wait(&empty_bar[stage], sphase ^ 1);
      uint8_t* dst = smem_a + stage * kABytes + row * 128;
#pragma unroll
      for (int piece = 0; piece < 8; ++piece) {
        uint32_t packed[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          float v[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = piece * 8 + e2 * 2 + h;          // compile-time after unrolling
            v[h] = k < K9 ? hi[k < K9 ? k : 0]
                          : (k < 2 * K9 ? lo[k < 2 * K9 ? k - K9 : 0] : (k >= 62 ? 1.f : 0.f));
          }
          packed[e2] = pack_bf16x2(v[0], v[1]);
        }
        *reinterpret_cast<uint4*>(dst + ((piece ^ (row & 7)) << 4)) =
            make_uint4(packed[0], packed[1], packed[2], packed[3]);
      }
      fence_proxy_async_smem();                 // generic-proxy writes -> visible to the MMA
      named_bar_sync(2 + group, 128);           // whole group done: one arrival per barrier
      if (row == 0) {
        mbar_arrive(&full_bar[stage]);
        mbar_arrive(&patch_empty[ps]);
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (256 threads)
    // Per tile: both TMEM loads in flight at once, accumulator handed back before the math,
    // ONE named barrier (the staging ring is 3 deep: when a thread
    // passes the barrier of tile i-1 the issuer has already seen the store of tile i-3 drain).
    const int q = warp & 3;                     // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;           // which 32 of the 64 channels
    const int row = q * 32 + lane;
    const bool issuer = (threadIdx.x == 64);
    uint32_t acc = 0, acc_phase = 0, local = 0;
    // tile coordinates are only needed by the storing thread; they advance by gridDim.x tiles per
    // iteration with add-and-carry instead of divisions
    int tx = 0, ty = 0, img = 0, step_x = 0, step_y = 0, step_img = 0;
    if (warp == 2) {
      int t = blockIdx.x;
      tx = t % p.tiles_x;
      t /= p.tiles_x;
      ty = t % p.tiles_y;
      img = t / p.tiles_y;
      int g = gridDim.x;
      step_x = g % p.tiles_x;
      g /= p.tiles_x;
      step_y = g % p.tiles_y;
      step_img = g / p.tiles_y;
    }
    const uint32_t out_base = smem_u32(smem_out) + row * 128;
    const uint32_t zero2 = 0u;                  // bf16x2 zeros for the packed ReLU
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kCout;
      if (p.debug_flags & 8) {                  // experiment: no epilogue work (nothing is stored)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + half * 32, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);   // drained: next tile's MMAs may start
      const uint32_t buf_off = (local % kOutBufs) * kOutBufBytes;
      uint8_t* buf = smem_out + buf_off;
      {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t packed[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[j * 8 + e * 2]),
                                                     __uint_as_float(r[j * 8 + e * 2 + 1]));
            if (p.relu) h = __hmax2(h, *reinterpret_cast<const __nv_bfloat162*>(&zero2));
            packed[e] = *reinterpret_cast<uint32_t*>(&h);
          }
          const int piece = (half * 4 + j) ^ (row & 7);     // 128B swizzle
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(out_base + buf_off +
                                                                      piece * 16),
                       "r"(packed[0]), "r"(packed[1]), "r"(packed[2]), "r"(packed[3])
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, kEpiThreads);
      if (issuer) {
        if (!(p.debug_flags & 4)) {
          tma_store_4d(&p.tmap_out, buf, 0, tx * p.tw, ty * p.th, img);
          tma_store_commit();
        }
        tma_store_wait_read<1>();               // the store of the previous tile has drained
      }
      if (warp == 2) {                          // next tile of this CTA
        tx += step_x;
        int carry = tx >= p.tiles_x ? 1 : 0;
        tx -= carry * p.tiles_x;
        ty += step_y + carry;
        carry = ty >= p.tiles_y ? 1 : 0;
        ty -= carry * p.tiles_y;
        img += step_img + carry;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kCout);
  }
}

template <int CIN>
int launch(const ConvIgemmParams& p, cudaStream_t stream) {
  auto kernel = conv_c1_kernel<CIN>;
  static bool configured = false;
  if (!configured) {
    XV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  const int total_tiles = p.N * p.tiles_y * p.tiles_x;
  const int grid = total_tiles < device_info().num_sms ? total_tiles : device_info().num_sms;
  kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

bool conv_c1_patch_fits(int th, int tw, int cin_raw, int* patch_w) {
  // TMA box rows start on a 16-byte boundary (up to 3 floats early) and are multiples of 16 bytes
  const int w = ((tw + 2) * cin_raw + 3 + 3) / 4 * 4;
  *patch_w = w;
  return w <= 256 && th + 2 <= 256 && w * (th + 2) * 4 <= kPatchBytes;
}

// p.tmap_in: fp32 input viewed as (W*Cin, H, N), box {patch_w, th + 2, 1}, no swizzle;
// p.tmap_w: bf16 [64, 64] box {64, 64}; p.tmap_out: bf16 output box {64, tw, th, 1}.
int launch_conv_c1(const ConvIgemmParams& p, int cin_raw, cudaStream_t stream) {
  XV_CHECK(p.th * p.tw == kBlockM, "conv_c1: tile must hold 128 pixels");
  XV_CHECK(p.cout == kCout, "conv_c1: the first convolution has 64 output channels");
  int patch_w = 0;
  XV_CHECK(conv_c1_patch_fits(p.th, p.tw, cin_raw, &patch_w) && patch_w == p.patch_w,
           "conv_c1: input patch does not fit its shared-memory slot");
  if (cin_raw == 1) return launch<1>(p, stream);
  if (cin_raw == 2) return launch<2>(p, stream);
  if (cin_raw == 3) return launch<3>(p, stream);
  return fail("conv_c1: Cin must be 1, 2 or 3");
}

}  // namespace xv

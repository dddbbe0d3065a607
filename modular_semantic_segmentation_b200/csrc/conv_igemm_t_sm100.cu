// "Transposed" bf16 implicit-GEMM 3x3 convolution for layers with few output channels
// (Cout <= 128: conv1_2, conv2_1, conv2_2 of simple_fcn.py:40-43), optional fused 2x2 max pool.
//
// Why a second kernel: measured on B200, one tcgen05.mma (cta_group::1, M=128, operands in
// shared memory) costs ~110 cycles + ~0.14*N regardless of how little work it carries
// (N=64: ~110 cyc for a 32-cycle MMA, N=128: ~120 for 64, N=256: ~140 for 128).  With the
// output channels on the N axis a Cout=64/128 layer therefore runs the tensor pipe at 29% / 53%.
// Here the roles are swapped:
//     D^T[cout, pixel] = sum_k  W[cout, k] * X[pixel, k]
//   M = 128 output channels (rows >= Cout are zero-filled by TMA), N = 256 pixels (a 16x16
//   patch), K block = 64 input channels of one filter tap,
// so every instruction is the efficient 128x256x16 shape.  The accumulator holds channels on
// TMEM lanes and pixels on columns; an epilogue thread owns ONE channel of 256 pixels, which
// makes the 2x2 max pool of simple_fcn.py:41,44 a register-local max of four columns.
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kTile = 16;                          // 16 x 16 output pixels
constexpr int kPixels = kTile * kTile;             // N = 256
constexpr int kBlockM = 128;                       // output channels per block
constexpr int kBlockK = 64;
constexpr int kWBytes = kBlockM * kBlockK * 2;     // 16 KB weights (A operand)
constexpr int kXBytes = kPixels * kBlockK * 2;     // 32 KB activations (B operand)
constexpr int kStageBytes = kWBytes + kXBytes;
constexpr int kStages = 4;
constexpr int kStagingBytes = 32768;               // 2 channel chunks x 128 rows x 128 B
constexpr int kThreads = 192;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kStagingBytes + 256;
// HALO variant (3x3, dilation 1): the producer loads three column-shifted copies of the
// (16 + 2) x 16 pixel patch per input-channel chunk instead of nine shifted 16 x 16 tiles; the
// three row shifts are a 2 KB offset in the operand descriptor.  Shared-memory fill traffic per
// tile drops from 9 x 32 KB to 3 x 36 KB - this kernel is bound by shared-memory bandwidth
// (TMA writes + MMA operand reads = 96 KB per K block against 128 B / cycle).
constexpr int kCopyRows = kTile + 2;
constexpr int kCopyBytes = kCopyRows * kTile * 128;        // 36 KB
constexpr int kXCopies = 3;
constexpr int kWStages = 4;
constexpr int kHaloSmemBytes =
    1024 + kXCopies * kCopyBytes + kWStages * kWBytes + kStagingBytes + 256;

template <bool POOL, bool GEN = false, bool HALO = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_t_kernel(const __grid_constant__ ConvIgemmParams p) {
  static_assert(!(GEN && HALO), "the halo variant is for plain 3x3 filters");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // plain: kStages x (W 16 KB) | kStages x (X 32 KB); halo: kXCopies x 36 KB | kWStages x 16 KB
  uint8_t* smem_w = HALO ? smem + kXCopies * kCopyBytes : smem;
  uint8_t* smem_x = HALO ? smem : smem + kStages * kWBytes;
  uint8_t* staging = HALO ? smem_w + kWStages * kWBytes : smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* x_full_bar = bars + 2 * kStages + 5;      // HALO only: kXCopies + kXCopies barriers
  uint64_t* x_empty_bar = x_full_bar + kXCopies;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.N * p.tiles_y * p.tiles_x * p.n_blocks;
  const int cin_chunks = p.cin / kBlockK;
  const int num_kb = (GEN ? p.taps : 9) * cin_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_in);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_out);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    if (HALO) {
      for (int s = 0; s < kXCopies; ++s) {
        mbar_init(&x_full_bar[s], 1);
        mbar_init(&x_empty_bar[s], 1);
      }
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // Cout <= 64: the weight box holds 64 rows; rows 64..127 of every weight stage stay zero for
  // the whole kernel instead of being zero-filled by TMA for every K block
  const uint32_t w_bytes = static_cast<uint32_t>(p.w_rows) * 128;
  if (p.w_rows < kBlockM) {
    constexpr int kNumW = HALO ? kWStages : kStages;
    for (int i = threadIdx.x; i < kNumW * (kWBytes / 2 / 16); i += kThreads) {
      const int st = i / (kWBytes / 2 / 16), off = i % (kWBytes / 2 / 16);
      *reinterpret_cast<uint4*>(smem_w + st * kWBytes + kWBytes / 2 + off * 16) =
          make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int& img, int& y0, int& x0, int& n0) {
    const int nb = tile % p.n_blocks;
    const int mt = tile / p.n_blocks;
    const int tx = mt % p.tiles_x;
    const int rest = mt / p.tiles_x;
    const int ty = rest % p.tiles_y;
    img = rest / p.tiles_y;
    y0 = ty * kTile;
    x0 = tx * kTile;
    n0 = nb * kBlockM;
  };

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    uint32_t stage = 0, phase = 0;
    if constexpr (HALO) {
      uint32_t xs = 0, xphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int img, y0, x0, n0;
        decode(tile, img, y0, x0, n0);
        for (int cc = 0; cc < cin_chunks; ++cc) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&x_empty_bar[xs], xphase ^ 1);
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&x_full_bar[xs], kCopyBytes);
              tma_load_4d(smem_x + xs * kCopyBytes, &p.tmap_in, &x_full_bar[xs], cc * kBlockK,
                          x0 + dxi - 1, y0 - 1, img);
            }
            __syncwarp();
            if (++xs == kXCopies) {
              xs = 0;
              xphase ^= 1;
            }
            for (int dyi = 0; dyi < 3; ++dyi) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (elect_one_sync()) {
                mbar_arrive_expect_tx(&full_bar[stage], w_bytes);
                tma_load_2d(smem_w + stage * kWBytes, &p.tmap_w, &full_bar[stage],
                            (dyi * 3 + dxi) * p.cin + cc * kBlockK, n0);
              }
              __syncwarp();
              if (++stage == kWStages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    } else {
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int img, y0, x0, n0;
      decode(tile, img, y0, x0, n0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / cin_chunks;
        const int cc = kb - tap * cin_chunks;
        int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const CUtensorMap* tmap_x = &p.tmap_in;
        if (GEN) {
          const int ty = tap / p.kw;
          dy = ty * p.dil - p.pad;
          dx = (tap - ty * p.kw) * p.dil - p.pad;
          if (p.stride2) {
            tmap_x = &p.tmap_in_par[(dy & 1) * 2 + (dx & 1)];
            dy >>= 1;
            dx >>= 1;
          }
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[stage], kXBytes + w_bytes);
          tma_load_4d(smem_x + stage * kXBytes, tmap_x, &full_bar[stage], cc * kBlockK,
                      x0 + dx, y0 + dy, img);
          tma_load_2d(smem_w + stage * kWBytes, &p.tmap_w, &full_bar[stage],
                      tap * p.cin + cc * kBlockK, n0);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kPixels);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    uint32_t xs = 0, xphase = 0;                 // HALO: position in the ring of patch copies
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kPixels;
      if constexpr (HALO) {
        int first = 1;
        for (int cc = 0; cc < cin_chunks; ++cc) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&x_full_bar[xs], xphase);
            for (int dyi = 0; dyi < 3; ++dyi) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              if (elect_one_sync()) {
                const uint32_t w_addr = smem_u32(smem_w + stage * kWBytes);
                // rows dyi .. dyi + 15 of the 18-row copy: 16 pixel rows of 128 B each = 2 KB step
                const uint32_t x_addr = smem_u32(smem_x + xs * kCopyBytes) + dyi * (kTile * 128);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  umma_bf16(d_tmem, umma_desc_sw128(w_addr + k * 32, 1024, 0),
                            umma_desc_sw128(x_addr + k * 32, 1024, 0), idesc,
                            (first && k == 0) ? 0u : 1u);
                }
                umma_commit(&empty_bar[stage]);
                if (dyi == 2) umma_commit(&x_empty_bar[xs]);
                if (cc == cin_chunks - 1 && dxi == 2 && dyi == 2) umma_commit(&tmem_full_bar[acc]);
              }
              __syncwarp();
              first = 0;
              if (++stage == kWStages) {
                stage = 0;
                phase ^= 1;
              }
            }
            if (++xs == kXCopies) {
              xs = 0;
              xphase ^= 1;
            }
          }
        }
      } else {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t w_addr = smem_u32(smem_w + stage * kWBytes);
          const uint32_t x_addr = smem_u32(smem_x + stage * kXBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_bf16(d_tmem, umma_desc_sw128(w_addr + k * 32, 1024, 0),
                      umma_desc_sw128(x_addr + k * 32, 1024, 0), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------------------------------------- epilogue (128 threads)
    const int q = warp & 3;
    const int ch = q * 32 + lane;               // output channel inside the 128-channel block
    const int ch64 = ch & 63;
    uint8_t* my_chunk = staging + (ch >> 6) * 16384;   // 64-channel chunk buffer
    const uint32_t ch_off = static_cast<uint32_t>(ch64 & 7) * 2;
    const uint32_t ch_piece = static_cast<uint32_t>(ch64 >> 3);
    const bool issuer = (threadIdx.x == 64);
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int img, y0, x0, n0;
      decode(tile, img, y0, x0, n0);
      const bool live = (n0 + ch) < p.cout;       // rows beyond Cout are padding
      const float bias = live ? __ldg(p.bias + n0 + ch) : 0.f;
      const int n_chunks = (p.cout - n0) > 64 ? 2 : 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kPixels;
      if constexpr (POOL) {
        // whole tile -> 8x8 pooled pixels (rows of 128 B per 64-channel chunk)
        if (issuer) tma_store_wait_read<0>();
        named_bar_sync(1, 128);
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {             // 32 columns = tile rows 2j, 2j+1
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + j * 32, r);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int pxp = 0; pxp < 8; ++pxp) {
              float v = fmaxf(fmaxf(__uint_as_float(r[2 * pxp]), __uint_as_float(r[2 * pxp + 1])),
                              fmaxf(__uint_as_float(r[16 + 2 * pxp]),
                                    __uint_as_float(r[17 + 2 * pxp])));
              v += bias;
              if (p.relu) v = fmaxf(v, 0.f);
              const uint32_t row = static_cast<uint32_t>(j * 8 + pxp);   // pooled pixel index
              *reinterpret_cast<__nv_bfloat16*>(my_chunk + row * 128 +
                                                ((ch_piece ^ (row & 7)) << 4) + ch_off) =
                  __float2bfloat16_rn(v);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (issuer) {
          for (int c = 0; c < n_chunks; ++c)
            tma_store_4d(&p.tmap_out, staging + c * 16384, n0 + c * 64, x0 >> 1, y0 >> 1, img);
          tma_store_commit();
        }
      } else {
        // Un-pooled output: the accumulator (lanes = channels, columns = pixels) has to reach the
        // pixel-major staging rows (128 B = 64 channels per pixel).  tcgen05.ld.16x256b hands
        // every thread the mma-style fragment (row t/4 [+8], columns 2(t%4), +1 of each 8-column
        // group), which is exactly what stmatrix.trans wants: one stmatrix.x4 writes four 8x8
        // bf16 blocks as 16-byte pieces of eight pixel rows - 16 store instructions per warp and
        // 128-pixel half instead of the 128 two-byte scatter stores of the first version (the
        // shared-memory port is this kernel's bottleneck: tensor pipe 65 % on conv2_1).
        const bool warp_live = (n0 + q * 32) < p.cout;      // Cout % 64 == 0: uniform per warp
        float fb[2][2];
#pragma unroll
        for (int hs = 0; hs < 2; ++hs)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            const int c = n0 + q * 32 + hs * 16 + (lane >> 2) + rh * 8;
            fb[hs][rh] = c < p.cout ? __ldg(p.bias + c) : 0.f;
          }
        const int mi = lane >> 3, rr = lane & 7;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {    // 128 pixels = 8 tile rows at a time
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(1, 128);
#pragma unroll 1
          for (int blk = 0; blk < 2; ++blk) {     // 64 pixels per fragment load
#pragma unroll
            for (int hs = 0; hs < 2; ++hs) {      // lanes 0-15 / 16-31 of this warp's quarter
              uint32_t r[32];
              tmem_ld_16x256b_x8(t_row + (static_cast<uint32_t>(hs * 16) << 16) + half * 128 +
                                     blk * 64,
                                 r);
              tmem_ld_wait();
              if (warp_live) {
                const int chgrp = q * 32 + hs * 16 + (mi & 1) * 8;      // channels of "my" block
                uint8_t* chunk = staging + (chgrp >> 6) * 16384;
                const uint32_t piece = static_cast<uint32_t>((chgrp & 63) >> 3);
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                  uint32_t m[4];
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    const int rep = i + u;
                    float v0 = __uint_as_float(r[4 * rep]) + fb[hs][0];
                    float v1 = __uint_as_float(r[4 * rep + 1]) + fb[hs][0];
                    float v2 = __uint_as_float(r[4 * rep + 2]) + fb[hs][1];
                    float v3 = __uint_as_float(r[4 * rep + 3]) + fb[hs][1];
                    if (p.relu) {
                      v0 = fmaxf(v0, 0.f);
                      v1 = fmaxf(v1, 0.f);
                      v2 = fmaxf(v2, 0.f);
                      v3 = fmaxf(v3, 0.f);
                    }
                    m[2 * u] = pack_bf16x2(v0, v1);
                    m[2 * u + 1] = pack_bf16x2(v2, v3);
                  }
                  const uint32_t row = static_cast<uint32_t>(blk * 64 + (i + (mi >> 1)) * 8 + rr);
                  const uint32_t addr =
                      smem_u32(chunk) + row * 128 + ((piece ^ (row & 7)) << 4);
                  asm volatile(
                      "stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(
                          addr),
                      "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3])
                      : "memory");
                }
              }
            }
          }
          if (half == 1) {
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (issuer) {
            for (int c = 0; c < n_chunks; ++c)
              tma_store_4d(&p.tmap_out, staging + c * 16384, n0 + c * 64, x0, y0 + half * 8, img);
            tma_store_commit();
          }
        }
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
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool POOL, bool GEN, bool HALO = false>
int launch_t(const ConvIgemmParams& p, cudaStream_t stream) {
  auto kernel = conv_igemm_t_kernel<POOL, GEN, HALO>;
  constexpr int kBytes = HALO ? kHaloSmemBytes : kSmemBytes;
  static bool configured = false;
  if (!configured) {
    XV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
    configured = true;
  }
  const int total_tiles = p.N * p.tiles_y * p.tiles_x * p.n_blocks;
  const int grid = total_tiles < device_info().num_sms ? total_tiles : device_info().num_sms;
  kernel<<<grid, kThreads, kBytes, stream>>>(p);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

// Params: th = tw = 16, tiles_* computed for 16x16 tiles, n_blocks = ceil(cout / 128);
// tmap_in box {64,16,16,1}; tmap_w box {64,128}; tmap_out box {64,16,8,1} on the full-size
// output, or {64,8,8,1} on the pooled output when pool != 0.
int launch_conv_igemm_t(const ConvIgemmParams& p, bool pool, cudaStream_t stream) {
  XV_CHECK(p.cin % kBlockK == 0, "conv_igemm_t: Cin must be a multiple of 64");
  XV_CHECK(p.cout % 64 == 0, "conv_igemm_t: Cout must be a multiple of 64");
  XV_CHECK(!pool || (p.H % 2 == 0 && p.W % 2 == 0), "conv_igemm_t: pooling needs even H, W");
  XV_CHECK(p.w_rows == 64 || p.w_rows == 128, "conv_igemm_t: weight box must hold 64 or 128 rows");
  if (p.halo) {   // tmap_in box {64, 16, 18, 1}
    return pool ? launch_t<true, false, true>(p, stream) : launch_t<false, false, true>(p, stream);
  }
  return pool ? launch_t<true, false>(p, stream) : launch_t<false, false>(p, stream);
}

int launch_conv_igemm_t_generic(const ConvIgemmParams& p, cudaStream_t stream) {
  XV_CHECK(p.cin % kBlockK == 0, "conv_igemm_t: Cin must be a multiple of 64");
  XV_CHECK(p.cout % 64 == 0, "conv_igemm_t: Cout must be a multiple of 64");
  XV_CHECK(p.taps > 0 && p.kw > 0 && p.dil > 0, "conv_igemm_t: generic geometry not set");
  XV_CHECK(p.w_rows == 64 || p.w_rows == 128, "conv_igemm_t: weight box must hold 64 or 128 rows");
  return launch_t<false, true>(p, stream);
}

}  // namespace xv

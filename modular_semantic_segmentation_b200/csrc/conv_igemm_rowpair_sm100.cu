// Row-pair variant of the transposed-role 3x3 convolution for Cin = 64, Cout <= 64: conv1_2 of
// simple_fcn.py:40-41, either with the 2x2 max pool fused (inference: only pool1 is stored) or
// with the full activation as output (fit(): forward and data gradient of conv1_2).
//
// Why: with the output channels on the M axis (conv_igemm_t_sm100.cu) a 64-channel layer fills
// half of the 128 accumulator lanes, so half of every tcgen05.mma multiplies zeros - conv1_2 ran
// at 0.52 of the bf16 peak and was the kernel furthest below its roofline.  Here the lower half
// of M does useful work as well: lanes 0..63 accumulate output row y, lanes 64..127 accumulate
// output row y + 1 of the same pixel column.  For an operand patch shifted by s rows
//     lanes  0..63  need filter row dy = s + 1   (out[y]     += W[dy] * in[y + dy - 1])
//     lanes 64..127 need filter row dy = s       (out[y + 1] += W[dy] * in[y + 1 + dy - 1])
// so the A operand of shift s is the 128-row block [W[s+1]; W[s]] (zero where dy leaves 0..2),
// s = -1..2: FOUR row shifts serve TWO output rows, 12 instead of 18 MMA groups per pixel pair.
// Nothing has to be combined afterwards - the two halves are different outputs, not partial
// sums.
//
// Shared memory (one CTA per SM, persistent over tiles of 32 rows x 16 pixels):
//   * weights, resident for the whole kernel: per filter column dx the 64-row blocks
//     [0, W[2,dx], W[1,dx], W[0,dx]] and one closing zero block = 13 x 8 KB; A(s, dx) starts at
//     block 4 dx + 2 - s and spans two blocks;
//   * a ring of three 17 x 16 pixel patches holding input rows of ONE parity (two tensor maps
//     that see every second row): the N = 256 operand "pixels (y0 + 2 i + s, x)" is then a
//     contiguous slice of the odd-row patch (s = -1, 1) or of the even-row patch (s = 0, 2);
//     the three filter columns are three column-shifted loads as in the halo kernel;
//   * 16 KB staging for the pooled 16 x 8 pixel output tile, 2 KB exchange buffer.
// Epilogue: a thread owns one channel of one row parity (TMEM lane) and 256 columns = 16 row
// pairs x 16 pixels.  It takes the horizontal max of its row, adds the bias, applies ReLU and
// rounds to bf16 - all monotone, so the order against the vertical max does not matter - and
// swaps half of the values with the thread holding the other row of the pair (lane ^ 64) through
// shared memory; each of the two then finishes four of the eight pooled pixels.
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kTileW = 16;
constexpr int kPairs = 16;                          // row pairs per tile: 32 output rows
constexpr int kPixels = kPairs * kTileW;            // N = 256
constexpr int kWBlock = 64 * 128;                   // 64 filter rows x 64 input channels, 8 KB
constexpr int kWBlocks = 13;
constexpr int kPatchRows = kPairs + 1;
constexpr int kPatchBytes = kPatchRows * kTileW * 128;      // 34 KB
constexpr int kRing = 3;
constexpr int kStagingBytes = 16384;                // 128 pooled pixels x 128 B
constexpr int kXchgBytes = 2 * 128 * 8;
constexpr int kThreads = 192;
constexpr int kSmemBytes =
    1024 + kWBlocks * kWBlock + kRing * kPatchBytes + kStagingBytes + kXchgBytes + 256;
static_assert(kSmemBytes <= 227 * 1024, "row-pair conv: shared memory budget");

// POOL: only the 2x2 max-pooled tensor is stored (inference, conv1_2 -> pool1); !POOL: the full
// activation (fit() forward / data gradient, or a caller that wants conv1_2 itself).
template <bool POOL>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_rowpair_kernel(const __grid_constant__ ConvIgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem_w + kWBlocks * kWBlock;
  uint8_t* staging = smem_x + kRing * kPatchBytes;
  uint8_t* xchg = staging + kStagingBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + kXchgBytes);
  uint64_t* x_full_bar = bars;
  uint64_t* x_empty_bar = bars + kRing;
  uint64_t* tmem_full_bar = bars + 2 * kRing;
  uint64_t* tmem_empty_bar = bars + 2 * kRing + 2;
  uint64_t* w_bar = bars + 2 * kRing + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRing + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.N * p.tiles_y * p.tiles_x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_in_par[0]);
    tma_prefetch_desc(&p.tmap_in_par[1]);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_out);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&x_full_bar[s], 1);
      mbar_init(&x_empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // the zero blocks 0, 4, 8, 12 (and, for Cout < 64, nothing else: TMA zero-fills rows >= Cout)
  for (int i = threadIdx.x; i < 4 * (kWBlock / 16); i += kThreads) {
    const int blk = (i / (kWBlock / 16)) * 4, off = i % (kWBlock / 16);
    *reinterpret_cast<uint4*>(smem_w + blk * kWBlock + off * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int& img, int& y0, int& x0) {
    const int tx = tile % p.tiles_x;
    const int rest = tile / p.tiles_x;
    const int ty = rest % p.tiles_y;
    img = rest / p.tiles_y;
    y0 = ty * (2 * kPairs);
    x0 = tx * kTileW;
  };

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_bar, 9 * kWBlock);
      for (int dyi = 0; dyi < 3; ++dyi)
        for (int dxi = 0; dxi < 3; ++dxi)
          tma_load_2d(smem_w + (4 * dxi + 3 - dyi) * kWBlock, &p.tmap_w, w_bar,
                      (dyi * 3 + dxi) * p.cin, 0);
    }
    __syncwarp();
    uint32_t xs = 0, xphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int img, y0, x0;
      decode(tile, img, y0, x0);
      for (int dxi = 0; dxi < 3; ++dxi) {
#pragma unroll
        for (int par = 1; par >= 0; --par) {       // odd rows y0 - 1 + 2 j first, then y0 + 2 j
          mbar_wait(&x_empty_bar[xs], xphase ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&x_full_bar[xs], kPatchBytes);
            tma_load_4d(smem_x + xs * kPatchBytes, &p.tmap_in_par[par], &x_full_bar[xs], 0,
                        x0 + dxi - 1, (y0 >> 1) - par, img);
          }
          __syncwarp();
          if (++xs == kRing) {
            xs = 0;
            xphase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, kPixels);
    uint32_t acc = 0, acc_phase = 0, xs = 0, xphase = 0;
    mbar_wait(w_bar, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kPixels;
      for (int dxi = 0; dxi < 3; ++dxi) {
#pragma unroll
        for (int par = 1; par >= 0; --par) {
          mbar_wait(&x_full_bar[xs], xphase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t patch = smem_u32(smem_x + xs * kPatchBytes);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              // odd patch: s = -1 (rows from j = 0), s = 1 (from j = 1); even patch: s = 0, 2
              const int s = 2 * j - par;
              const uint32_t w_addr = smem_u32(smem_w + (4 * dxi + 2 - s) * kWBlock);
              const uint32_t x_addr = patch + j * (kTileW * 128);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, umma_desc_sw128(w_addr + k * 32, 1024, 0),
                          umma_desc_sw128(x_addr + k * 32, 1024, 0), idesc,
                          (dxi == 0 && par == 1 && j == 0 && k == 0) ? 0u : 1u);
            }
            umma_commit(&x_empty_bar[xs]);
            if (dxi == 2 && par == 0) umma_commit(&tmem_full_bar[acc]);
          }
          __syncwarp();
          if (++xs == kRing) {
            xs = 0;
            xphase ^= 1;
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------------------------------------- epilogue (128 threads)
    const int q = warp & 3;
    const int e = q * 32 + lane;                 // TMEM lane: channel e & 63 of row parity e >> 6
    const int ch = e & 63;
    const int half = e >> 6;
    const uint32_t ch_off = static_cast<uint32_t>(ch & 7) * 2;
    const uint32_t ch_piece = static_cast<uint32_t>(ch >> 3);
    const bool issuer = (threadIdx.x == 64);
    const float bias = ch < p.cout ? __ldg(p.bias + ch) : 0.f;
    uint2* xb = reinterpret_cast<uint2*>(xchg);
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int img, y0, x0;
      decode(tile, img, y0, x0);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kPixels;
      if constexpr (!POOL) {
        // Un-pooled output, 8 image rows (4 row pairs = 64 accumulator columns) per round through
        // the 16 KB staging buffer: tcgen05.ld.16x256b hands every thread the mma-style fragment
        // and stmatrix.x4.trans writes four 8-channel x 8-pixel blocks as 16-byte pieces of eight
        // pixel rows (as in conv_igemm_t_sm100.cu); the staging row of accumulator column
        // (pair ip, pixel x) is (2 ip + half) * 16 + x.
        float fb[2][2];
#pragma unroll
        for (int hs = 0; hs < 2; ++hs)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            const int c = (q & 1) * 32 + hs * 16 + (lane >> 2) + rh * 8;
            fb[hs][rh] = c < p.cout ? __ldg(p.bias + c) : 0.f;
          }
        const int mi = lane >> 3, rr = lane & 7;
#pragma unroll 1
        for (int round = 0; round < 4; ++round) {
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(1, 128);
#pragma unroll
          for (int hs = 0; hs < 2; ++hs) {
            uint32_t r[32];
            tmem_ld_16x256b_x8(t_row + (static_cast<uint32_t>(hs * 16) << 16) + round * 64, r);
            tmem_ld_wait();
            const uint32_t piece =
                static_cast<uint32_t>(((q & 1) * 32 + hs * 16 + (mi & 1) * 8) >> 3);
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
              const int grp = i + (mi >> 1);                     // 8-column group of the round
              const uint32_t srow =
                  static_cast<uint32_t>((2 * (grp >> 1) + half) * 16 + (grp & 1) * 8 + rr);
              const uint32_t addr = smem_u32(staging) + srow * 128 + ((piece ^ (srow & 7)) << 4);
              asm volatile(
                  "stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(
                      addr),
                  "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3])
                  : "memory");
            }
          }
          if (round == 3) {
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (issuer) {
            tma_store_4d(&p.tmap_out, staging, 0, x0, y0 + round * 8, img);
            tma_store_commit();
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      if (issuer) tma_store_wait_read<0>();
      named_bar_sync(1, 128);
#pragma unroll 1
      for (int j = 0; j < kPairs / 2; ++j) {       // 32 columns = row pairs 2 j and 2 j + 1
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + j * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float a = fmaxf(__uint_as_float(r[ii * 16 + 4 * k]),
                            __uint_as_float(r[ii * 16 + 4 * k + 1])) + bias;
            float b = fmaxf(__uint_as_float(r[ii * 16 + 4 * k + 2]),
                            __uint_as_float(r[ii * 16 + 4 * k + 3])) + bias;
            if (p.relu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            w[k] = pack_bf16x2(a, b);
          }
          // the even-row thread finishes pooled pixels 0..3, the odd-row thread 4..7
          const int buf = ii;
          xb[buf * 128 + e] = half ? make_uint2(w[0], w[1]) : make_uint2(w[2], w[3]);
          named_bar_sync(2 + ii, 128);
          const uint2 other = xb[buf * 128 + (e ^ 64)];
          const uint32_t m0 = half ? w[2] : w[0], m1 = half ? w[3] : w[1];
          const __nv_bfloat162 f0 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&m0),
                                            *reinterpret_cast<const __nv_bfloat162*>(&other.x));
          const __nv_bfloat162 f1 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&m1),
                                            *reinterpret_cast<const __nv_bfloat162*>(&other.y));
          if (ch < p.cout) {
            const uint32_t row0 = static_cast<uint32_t>((2 * j + ii) * 8 + 4 * half);
            const __nv_bfloat16 v[4] = {f0.x, f0.y, f1.x, f1.y};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t row = row0 + k;       // pooled pixel (row pair, column pair)
              *reinterpret_cast<__nv_bfloat16*>(staging + row * 128 +
                                                ((ch_piece ^ (row & 7)) << 4) + ch_off) = v[k];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (issuer) {
        tma_store_4d(&p.tmap_out, staging, 0, x0 >> 1, y0 >> 1, img);
        tma_store_commit();
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

}  // namespace

// Params: N, H, W, cin = 64, cout <= 64, tiles_y = ceil(H / 32), tiles_x = ceil(W / 16);
// tmap_in_par[0] / [1]: the even / odd input rows, box {64, 16, 17, 1}; tmap_w box {64, 64};
// pool: H, W even, tmap_out = the pooled output [N, H/2, W/2, Cout], box {64, 8, 16, 1};
// !pool: tmap_out = the full output [N, H, W, Cout], box {64, 16, 8, 1}.
template <bool POOL>
static int launch_rowpair(const ConvIgemmParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    XV_CUDA(cudaFuncSetAttribute(conv_igemm_rowpair_kernel<POOL>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  const int total_tiles = p.N * p.tiles_y * p.tiles_x;
  const int grid = total_tiles < device_info().num_sms ? total_tiles : device_info().num_sms;
  conv_igemm_rowpair_kernel<POOL><<<grid, kThreads, kSmemBytes, stream>>>(p);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_conv_igemm_rowpair(const ConvIgemmParams& p, bool pool, cudaStream_t stream) {
  XV_CHECK(p.cin == 64 && p.cout <= 64, "conv_igemm_rowpair: Cin = 64 and Cout <= 64 only");
  XV_CHECK(!pool || (p.H % 2 == 0 && p.W % 2 == 0), "conv_igemm_rowpair: pooling needs even H, W");
  XV_CHECK(pool || p.cout == 64, "conv_igemm_rowpair: the un-pooled epilogue needs Cout = 64");
  return pool ? launch_rowpair<true>(p, stream) : launch_rowpair<false>(p, stream);
}

}  // namespace xv

// bf16 implicit-GEMM convolution for sm_100a: tcgen05.mma with TMEM accumulators, operands
// staged by TMA, fused bias + ReLU epilogue.
//
// Replaces tf.layers.conv2d as wrapped by xview/models/custom_layers.py:124-139 (3x3 and
// 1x1, stride 1, 'same', + bias, ReLU) for every convolution of the VGG16-FCN expert
// (xview/models/simple_fcn.py:39-79).
//
// GEMM view:  D[pixels, Cout] = sum over (tap, cin) A[pixel shifted by tap, cin] * W[cout, tap, cin]
//   M tile  = 128 output pixels = a TH x TW patch of one image (TH*TW == 128)
//   N tile  = BLOCK_N output channels
//   K block = 64 input channels of one filter tap (64 bf16 = 128 B = one swizzle row)
// The A operand of a K block is ONE 4-D TMA box {64 ch, TW, TH, 1 image} whose start
// coordinate is shifted by the tap offset; out-of-image coordinates are zero-filled by the
// TMA unit, which is exactly the 'same' padding.  The box lands in shared memory as 128
// rows of 128 B in the 128B-swizzled K-major layout tcgen05.mma consumes, so there is no
// im2col buffer anywhere.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over tiles):
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + tcgen05.mma issuer (one lane)
//   warps 2..5  epilogue: tcgen05.ld -> bias/ReLU -> bf16 -> swizzled smem -> TMA store
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                      // bf16 elements = 128 bytes
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kOutBufBytes = kBlockM * 128;      // one 64-channel bf16 output chunk
constexpr int kThreads = 192;
constexpr int kPackGroups = 4;                    // conv1_1 mode: groups of 128 operand packers
constexpr int kPackThreads = 128 * kPackGroups;
constexpr int kPrefetchTiles = 2;               // L2 prefetch distance in tiles per CTA

// RES: two extra 16 KB buffers receive the residual operand tiles (one pipeline stage less).
// HALO (3x3, dilation 1): the pixel operand is loaded as three column-shifted copies of the
// (th + 2) x tw patch per input-channel chunk instead of nine shifted tiles (the row shift is an
// offset of tw rows in the operand descriptor), the weights stream through their own ring.  The
// kernel is bound by shared-memory bandwidth, so 2.4x less pixel fill traffic is time.
constexpr int kACopies = 3;
constexpr int kACopyBytes = 20480;               // (8 + 2) x 16 or (16 + 2) x 8 pixel rows of 128 B
template <int BLOCK_N, bool RES = false, bool HALO = false>
struct IgemmCfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBufs = RES ? 4 : 2;           // 2 output staging (+ 2 residual) buffers
  // plain: 4 / 6 / 8 stages for N = 256 / 128 / 64 (3 for N = 256 with the residual buffers);
  // halo: that many weight stages next to the three patch copies
  static constexpr int kHaloStages =
      (196608 - kACopies * kACopyBytes) / kBBytes > 8 ? 8 : (196608 - kACopies * kACopyBytes) / kBBytes;
  static constexpr int kStages =
      HALO ? kHaloStages : (196608 - (kOutBufs - 2) * kOutBufBytes) / kStageBytes;
  static constexpr int kPipeBytes =
      HALO ? kACopies * kACopyBytes + kStages * kBBytes : kStages * kStageBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;          // two accumulator stages
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes =
      1024 /*align slack*/ + kPipeBytes + kOutBufs * kOutBufBytes + kBarBytes;
};

struct TileCoord {
  int img, y0, x0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvIgemmParams& p, int tile,
                                                 int block_n) {
  TileCoord c;
  int nb = tile % p.n_blocks;
  int mt = tile / p.n_blocks;
  int tx = mt % p.tiles_x;
  int rest = mt / p.tiles_x;
  int ty = rest % p.tiles_y;
  c.img = rest / p.tiles_y;
  c.y0 = ty * p.th;
  c.x0 = tx * p.tw;
  c.n0 = nb * block_n;
  return c;
}

// C1 > 0: conv1_1 mode with C1 raw input channels - warps 6..9 build the A operand rows from the
// fp32 input (3x3 neighbourhood split into hi + lo bf16 halves, see layers.cu), TMA loads only
// the 8 KB weight tile.
template <int BLOCK_N, int TAPS, bool OUT_F32, int C1 = 0, bool RES = false, bool HALO = false>
__global__ void __launch_bounds__(C1 > 0 ? kThreads + kPackThreads : kThreads, 1)
conv_igemm_kernel(const __grid_constant__ ConvIgemmParams p) {
  static_assert(!HALO || (TAPS == 9 && !OUT_F32 && C1 == 0 && !RES), "halo: plain 3x3 bf16 layers");
  using Cfg = IgemmCfg<BLOCK_N, RES, HALO>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;                                  // HALO: the three patch copies
  uint8_t* smem_b = smem + (HALO ? kACopies * kACopyBytes : kStages * kABytes);
  uint8_t* smem_out = smem + Cfg::kPipeBytes;
  uint8_t* smem_res = smem_out + 2 * kOutBufBytes;          // RES only
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + Cfg::kOutBufs * kOutBufBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* res_full_bar = bars + 2 * kStages + 5;          // RES only, two barriers
  uint64_t* a_full_bar = bars + 2 * kStages + 7;            // HALO only, kACopies + kACopies
  uint64_t* a_empty_bar = a_full_bar + kACopies;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.N * p.tiles_y * p.tiles_x * p.n_blocks;
  const int cin_chunks = p.cin / kBlockK;
  const int num_kb = (TAPS == 0 ? p.taps : TAPS) * cin_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_in);
    tma_prefetch_desc(&p.tmap_w);
    if (!OUT_F32) tma_prefetch_desc(&p.tmap_out);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], C1 > 0 ? 2 : 1);     // + one arrival per operand-packing group
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
      if (RES) mbar_init(&res_full_bar[s], 1);
    }
    if (HALO) {
      for (int s = 0; s < kACopies; ++s) {
        mbar_init(&a_full_bar[s], 1);
        mbar_init(&a_empty_bar[s], 1);
      }
    }
    if (RES) tma_prefetch_desc(&p.tmap_res);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps run their loops warp-uniformly and elect one lane only around the
  // issue instructions: TMA / tcgen05 operands live in uniform registers, and a loop nested
  // inside a divergent `lane == 0` branch makes the compiler wrap every UTMALDG / UTCHMMA in
  // a waterfall loop (measured: ~145 cycles per MMA instead of the 32..128 cycle floor).
  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    uint32_t stage = 0, phase = 0;
    if constexpr (HALO) {
      uint32_t as = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord c = decode_tile(p, tile, BLOCK_N);
        for (int cc = 0; cc < cin_chunks; ++cc) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&a_empty_bar[as], aphase ^ 1);
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&a_full_bar[as], static_cast<uint32_t>((p.th + 2) * p.tw) * 128);
              tma_load_4d(smem_a + as * kACopyBytes, &p.tmap_in, &a_full_bar[as], cc * kBlockK,
                          c.x0 + dxi - 1, c.y0 - 1, c.img);
            }
            __syncwarp();
            if (++as == kACopies) {
              as = 0;
              aphase ^= 1;
            }
            for (int dyi = 0; dyi < 3; ++dyi) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (elect_one_sync()) {
                mbar_arrive_expect_tx(&full_bar[stage], Cfg::kBBytes);
                tma_load_2d(smem_b + stage * Cfg::kBBytes, &p.tmap_w, &full_bar[stage],
                            (dyi * 3 + dxi) * p.cin + cc * kBlockK, c.n0);
              }
              __syncwarp();
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
    for (int tile = blockIdx.x; !HALO && tile < total_tiles; tile += gridDim.x) {
      if (p.debug_flags & 256) break;       // experiment: MMA warp free-runs, no smem pipeline
      const TileCoord c = decode_tile(p, tile, BLOCK_N);
      if (p.debug_flags & 128) {            // experiment: L2 prefetch of a later tile
        const int ahead = tile + kPrefetchTiles * static_cast<int>(gridDim.x);
        if (ahead < total_tiles && elect_one_sync()) {
          const TileCoord f = decode_tile(p, ahead, BLOCK_N);
          for (int cc = 0; cc < cin_chunks; ++cc)
            tma_prefetch_l2_4d(&p.tmap_in, cc * kBlockK, f.x0, f.y0, f.img);
        }
        __syncwarp();
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / cin_chunks;
        const int cc = kb - tap * cin_chunks;
        int dy = (TAPS == 9) ? tap / 3 - 1 : 0;
        int dx = (TAPS == 9) ? tap % 3 - 1 : 0;
        if (TAPS == 0) {
          const int ty = tap / p.kw;
          dy = ty * p.dil - p.pad;
          dx = (tap - ty * p.kw) * p.dil - p.pad;
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (C1 > 0) {
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kBBytes);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &p.tmap_w, &full_bar[stage], 0, c.n0);
          }
        } else if (elect_one_sync()) {
          if (p.debug_flags & 64) {         // experiment: no TMA traffic at all
            mbar_arrive(&full_bar[stage]);
          } else if (p.debug_flags & 1) {          // experiment: no weight traffic
            mbar_arrive_expect_tx(&full_bar[stage], kABytes);
            tma_load_4d(smem_a + stage * kABytes, &p.tmap_in, &full_bar[stage], cc * kBlockK,
                        c.x0 + dx, c.y0 + dy, c.img);
          } else if (p.debug_flags & 2) {   // experiment: no activation traffic
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kBBytes);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &p.tmap_w, &full_bar[stage],
                        tap * p.cin + cc * kBlockK, c.n0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_4d(smem_a + stage * kABytes, &p.tmap_in, &full_bar[stage], cc * kBlockK,
                        c.x0 + dx, c.y0 + dy, c.img);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &p.tmap_w, &full_bar[stage],
                        tap * p.cin + cc * kBlockK, c.n0);
          }
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
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    uint32_t as = 0, aphase = 0;                 // HALO: position in the ring of patch copies
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      if constexpr (HALO) {
        int first = 1;
        for (int cc = 0; cc < cin_chunks; ++cc) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&a_full_bar[as], aphase);
            for (int dyi = 0; dyi < 3; ++dyi) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              if (elect_one_sync()) {
                // tile rows dyi .. dyi + th - 1 of the (th + 2)-row copy: tw pixel rows per step
                const uint32_t a_addr =
                    smem_u32(smem_a + as * kACopyBytes) + static_cast<uint32_t>(dyi * p.tw) * 128;
                const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32, 1024, 0),
                            umma_desc_sw128(b_addr + k * 32, 1024, 0), idesc,
                            (first && k == 0) ? 0u : 1u);
                }
                umma_commit(&empty_bar[stage]);
                if (dyi == 2) umma_commit(&a_empty_bar[as]);
                if (cc == cin_chunks - 1 && dxi == 2 && dyi == 2) umma_commit(&tmem_full_bar[acc]);
              }
              __syncwarp();
              first = 0;
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
              }
            }
            if (++as == kACopies) {
              as = 0;
              aphase ^= 1;
            }
          }
        }
      }
      for (int kb = 0; !HALO && kb < num_kb; ++kb) {
        if (!(p.debug_flags & 256)) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
        }
        if (elect_one_sync()) {
          const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32, 1024, 0),
                      umma_desc_sw128(b_addr + k * 32, 1024, 0), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          if (!(p.debug_flags & 256)) umma_commit(&empty_bar[stage]);   // frees the smem stage
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);   // accumulator complete
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
  } else if (C1 > 0 && warp >= 6) {
    // ------------------------------------------------------------- conv1_1 operand packing
    constexpr int CIN = C1 > 0 ? C1 : 1;
    constexpr int K9 = 9 * CIN;
    // kPackGroups groups of 128 threads take tiles round-robin so the global-load latency of one
    // tile overlaps the packing / barrier wait of the others
    const int group = (threadIdx.x - kThreads) >> 7;
    const int row = (threadIdx.x - kThreads) & 127;     // pixel index inside the tile
    const int py = row / p.tw;
    const int px = row - py * p.tw;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      if ((local % kPackGroups) != group) continue;
      const uint32_t stage = static_cast<uint32_t>(local) % kStages;
      const uint32_t phase = (static_cast<uint32_t>(local) / kStages) & 1u;
      const TileCoord c = decode_tile(p, tile, BLOCK_N);
      const int y = c.y0 + py, x = c.x0 + px;
      if (p.debug_flags & 1) {                  // experiment: no operand packing work
        mbar_wait(&empty_bar[stage], phase ^ 1);
        named_bar_sync(2 + group, 128);
        if (row == 0) mbar_arrive(&full_bar[stage]);
        continue;
      }
      float hi[K9], lo[K9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        const bool in = y < p.H && x < p.W && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        const float* src =
            p.x_raw + ((static_cast<size_t>(c.img) * p.H + (in ? yy : 0)) * p.W + (in ? xx : 0)) * CIN;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float raw = in ? __ldg(src + ci) : 0.f;
          const float h = __bfloat162float(__float2bfloat16_rn(raw));
          hi[tap * CIN + ci] = h;
          lo[tap * CIN + ci] = raw - h;
        }
      }
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* dst = smem_a + stage * kABytes + row * 128;
#pragma unroll
      for (int piece = 0; piece < 8; ++piece) {
        uint32_t packed[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          float v[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = piece * 8 + e2 * 2 + h;   // compile-time after unrolling
            v[h] = k < K9 ? hi[k < K9 ? k : 0] : (k < 2 * K9 ? lo[k < 2 * K9 ? k - K9 : 0] : 0.f);
          }
          packed[e2] = pack_bf16x2(v[0], v[1]);
        }
        *reinterpret_cast<uint4*>(dst + ((piece ^ (row & 7)) << 4)) =
            make_uint4(packed[0], packed[1], packed[2], packed[3]);
      }
      fence_proxy_async_smem();                 // generic-proxy writes -> visible to the MMA
      named_bar_sync(2 + group, 128);           // whole group done: one mbarrier arrival, not 128
      if (row == 0) mbar_arrive(&full_bar[stage]);
    }
  } else {
    // ------------------------------------------------------------- epilogue (128 threads)
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;           // pixel index inside the tile
    const int py = row / p.tw;
    const int px = row - py * p.tw;
    const bool issuer = (threadIdx.x == 64);
    uint32_t acc = 0, acc_phase = 0, gchunk = 0;
    // RES: TMA load of the residual tile of this CTA's g-th output chunk into rbuf[g & 1]
    auto request_residual = [&](uint32_t g) {
      constexpr uint32_t kChunks = BLOCK_N / 64;
      const int t = blockIdx.x + static_cast<int>(g / kChunks) * static_cast<int>(gridDim.x);
      if (t >= total_tiles) return;
      const TileCoord rc = decode_tile(p, t, BLOCK_N);
      mbar_arrive_expect_tx(&res_full_bar[g & 1], kOutBufBytes);
      tma_load_4d(smem_res + (g & 1) * kOutBufBytes, &p.tmap_res, &res_full_bar[g & 1],
                  rc.n0 + static_cast<int>(g % kChunks) * 64, rc.x0, rc.y0, rc.img);
    };
    if constexpr (RES) {
      if (issuer) {
        request_residual(0);
        request_residual(1);
      }
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile(p, tile, BLOCK_N);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / 64; ++chunk, ++gchunk) {
        if (p.debug_flags & 8) continue;   // experiment: no epilogue work at all
        if constexpr (!OUT_F32) {
          uint8_t* buf = smem_out + (gchunk & 1) * kOutBufBytes;
          if (issuer) tma_store_wait_read<1>();   // the store that last used `buf` has drained
          named_bar_sync(1, 128);
          if constexpr (RES) {
            // the residual tile of this chunk was requested two chunks ago (or in the prologue)
            mbar_wait(&res_full_bar[gchunk & 1], (gchunk >> 1) & 1);
          }
          const uint8_t* rbuf = smem_res + (gchunk & 1) * kOutBufBytes;
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
              const int piece = (half * 4 + j) ^ (row & 7);   // 128B swizzle
              const float4 b_lo = __ldg(bias4 + j * 2), b_hi = __ldg(bias4 + j * 2 + 1);
              const float bb[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
              if constexpr (RES) {
                const uint4 rv = *reinterpret_cast<const uint4*>(rbuf + row * 128 + piece * 16);
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float v0 = __uint_as_float(r[j * 8 + e * 2]) + bb[e * 2];
                  float v1 = __uint_as_float(r[j * 8 + e * 2 + 1]) + bb[e * 2 + 1];
                  if (p.relu) {
                    v0 = fmaxf(v0, 0.f);
                    v1 = fmaxf(v1, 0.f);
                  }
                  const float2 a = __bfloat1622float2(
                      *reinterpret_cast<const __nv_bfloat162*>(&rw[e]));
                  packed[e] = pack_bf16x2(fmaxf(v0 + a.x, 0.f), fmaxf(v1 + a.y, 0.f));
                }
              } else {
                // two outputs per instruction: packed fp32 add, packed bf16 convert, packed max
                // (ReLU after the rounding gives the same bits as before it)
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
                  if (p.relu) h = __hmax2(h, __floats2bfloat162_rn(0.f, 0.f));
                  packed[e] = *reinterpret_cast<uint32_t*>(&h);
                }
              }
              *reinterpret_cast<uint4*>(buf + row * 128 + piece * 16) =
                  make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (issuer && !(p.debug_flags & 4)) {
            tma_store_4d(&p.tmap_out, buf, c.n0 + chunk * 64, c.x0, c.y0, c.img);
            tma_store_commit();
          }
          if constexpr (RES) {
            if (issuer) request_residual(gchunk + 2);   // everybody is done with rbuf
          }
        } else {
          const int y = c.y0 + py, x = c.x0 + px;
          const bool valid = (y < p.H) && (x < p.W);
          float* dst = p.out_f32 +
                       ((static_cast<size_t>(c.img) * p.H + y) * p.W + x) * p.cout;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + chunk * 64 + half * 32, r);
            tmem_ld_wait();
            const int col0 = c.n0 + chunk * 64 + half * 32;
            if (valid && (p.cout & 3) == 0) {
              // 16-byte stores (Cout % 4 == 0: every pixel row of the output is 16-byte aligned)
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const int col = col0 + j4 * 4;
                if (col < p.cout) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                  float4 v = make_float4(__uint_as_float(r[j4 * 4]) + b4.x,
                                         __uint_as_float(r[j4 * 4 + 1]) + b4.y,
                                         __uint_as_float(r[j4 * 4 + 2]) + b4.z,
                                         __uint_as_float(r[j4 * 4 + 3]) + b4.w);
                  if (p.relu) {
                    v.x = fmaxf(v.x, 0.f);
                    v.y = fmaxf(v.y, 0.f);
                    v.z = fmaxf(v.z, 0.f);
                    v.w = fmaxf(v.w, 0.f);
                  }
                  *reinterpret_cast<float4*>(dst + col) = v;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = col0 + j;
                if (valid && col < p.cout) {
                  float v = __uint_as_float(r[j]) + __ldg(p.bias + col);
                  if (p.relu) v = fmaxf(v, 0.f);
                  dst[col] = v;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);      // 128 arrivals hand the accumulator back
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (!OUT_F32 && issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, int TAPS, bool OUT_F32, bool RES = false, bool HALO = false>
int launch_one(const ConvIgemmParams& p, cudaStream_t stream) {
  using Cfg = IgemmCfg<BLOCK_N, RES, HALO>;
  auto kernel = conv_igemm_kernel<BLOCK_N, TAPS, OUT_F32, 0, RES, HALO>;
  static bool configured = false;
  if (!configured) {
    XV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes));
    configured = true;
  }
  const int total_tiles = p.N * p.tiles_y * p.tiles_x * p.n_blocks;
  const int grid = total_tiles < device_info().num_sms ? total_tiles : device_info().num_sms;
  kernel<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(p);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

template <int C1>
int launch_c1(const ConvIgemmParams& p, cudaStream_t stream) {
  using Cfg = IgemmCfg<64>;
  auto kernel = conv_igemm_kernel<64, 1, false, C1>;
  static bool configured = false;
  if (!configured) {
    XV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes));
    configured = true;
  }
  const int total_tiles = p.N * p.tiles_y * p.tiles_x * p.n_blocks;
  const int grid = total_tiles < device_info().num_sms ? total_tiles : device_info().num_sms;
  kernel<<<grid, kThreads + kPackThreads, Cfg::kSmemBytes, stream>>>(p);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

int launch_conv_igemm_c1(const ConvIgemmParams& p, int cin_raw, cudaStream_t stream) {
  XV_CHECK(p.th * p.tw == kBlockM, "conv_igemm_c1: tile must hold 128 pixels");
  XV_CHECK(p.cout == 64 && p.n_blocks == 1, "conv_igemm_c1: conv1_1 has 64 output channels");
  if (cin_raw == 1) return launch_c1<1>(p, stream);
  if (cin_raw == 2) return launch_c1<2>(p, stream);
  if (cin_raw == 3) return launch_c1<3>(p, stream);
  return fail("conv_igemm_c1: Cin must be 1, 2 or 3");
}

namespace {
}  // namespace

int conv_igemm_block_n(int cout) { return cout > 128 ? 256 : (cout > 64 ? 128 : 64); }

int launch_conv_igemm(const ConvIgemmParams& p, int block_n, int taps, bool out_f32,
                      cudaStream_t stream) {
  XV_CHECK(p.th * p.tw == kBlockM, "conv_igemm: tile must hold 128 pixels");
  XV_CHECK(p.cin % kBlockK == 0, "conv_igemm: Cin must be a multiple of 64");
  XV_CHECK(taps == 0 || taps == 1 || taps == 9,
           "conv_igemm: taps must be 1, 9 or 0 (geometry from the parameter block)");
  if (taps == 0)
    XV_CHECK(p.taps > 0 && p.kw > 0 && p.dil > 0, "conv_igemm: generic geometry not set");
  if (p.halo) {   // tmap_in box {64, tw, th + 2, 1}
    XV_CHECK(taps == 9 && !out_f32 && !p.has_residual && (p.tw == 8 || p.tw == 16) &&
                 (p.th + 2) * p.tw * 128 <= kACopyBytes,
             "conv_igemm: the halo variant needs a 3x3 bf16 layer and an 8x16 or 16x8 tile");
    if (block_n == 64) return launch_one<64, 9, false, false, true>(p, stream);
    if (block_n == 128) return launch_one<128, 9, false, false, true>(p, stream);
    if (block_n == 256) return launch_one<256, 9, false, false, true>(p, stream);
  }
  if (p.has_residual) {
    XV_CHECK(taps == 0 && block_n == 256 && !out_f32,
             "conv_igemm: the residual epilogue exists for generic bf16 layers with BLOCK_N = 256");
    return launch_one<256, 0, false, true>(p, stream);
  }
#define XV_IGEMM_CASE(BN)                                                              \
  if (block_n == BN) {                                                                 \
    if (taps == 0) {                                                                   \
      return out_f32 ? launch_one<BN, 0, true>(p, stream)                              \
                     : launch_one<BN, 0, false>(p, stream);                            \
    }                                                                                  \
    if (taps == 9) {                                                                   \
      return out_f32 ? launch_one<BN, 9, true>(p, stream)                              \
                     : launch_one<BN, 9, false>(p, stream);                            \
    }                                                                                  \
    return out_f32 ? launch_one<BN, 1, true>(p, stream) : launch_one<BN, 1, false>(p, stream); \
  }
  XV_IGEMM_CASE(64)
  XV_IGEMM_CASE(128)
  XV_IGEMM_CASE(256)
#undef XV_IGEMM_CASE
  return fail("conv_igemm: unsupported BLOCK_N");
}

}  // namespace xv

// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): 3x3 (any dilation, pad = dilation) and
// 1x1 convs with fused bias + ReLU, NHWC split-fp16 activations in and out.
//
// Replaces, for every Convolution layer of the deploy nets except conv1_1 and the 2/4/6/12-channel
// cls/bbox heads: caffe/src/caffe/layers/base_conv_layer.cpp:255-279 (im2col + sgemm + bias gemm),
// util/im2col.cpp:19-55, layers/cudnn_conv_layer.cu:11-46 and the in-place ReLU that follows
// (layers/relu_layer.cpp:9-19).  No column buffer is ever materialised: for tap (r,s) the A operand
// of the GEMM is the activation tensor itself, fetched by TMA at a shifted coordinate with the
// hardware's out-of-bounds zero fill standing in for Caffe's zero padding.
//
// GEMM view per image:  D[M = H*W pixels, N = Cout] = sum over taps, Cin of A[M, Cin] * B[Cout, Cin]^T
//   CTA tile: 128 output pixels (a TH x TW patch, TH*TW = 128) x BN output channels
//   K loop  : taps x (Cin / 64); one pipeline stage = A tile (128 x 64) + B tile (BN x 64), hi and lo planes
//   precision: x = hi + lo in fp16 for both operands; three kind::f16 MMAs per k-step
//              (hi*lo + lo*hi + hi*hi) accumulate in fp32 TMEM -> ~2^-22 relative operand error, which is
//              what the 1e-2 px box tolerance needs through 17-20 layers (DESIGN.md "precision").
//   roles   : warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warp 3 bias loader,
//             warps 4-7 epilogue (TMEM -> registers -> bias/ReLU/split -> swizzled smem -> TMA store)
#include "common.cuh"
#include "tma_host.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;                 // fp16 elements per 128-byte swizzled row
constexpr int kABytes = 2 * kTileM * 128;   // hi + lo planes of the A tile

template <int BN>
struct ConvCfg {
  static constexpr int kBBytes = 2 * BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 64) ? 4 : 3;
  static constexpr int kEpiBytes = 2 * kTileM * BN * 2;                 // hi + lo output tiles
  static constexpr int kPipeBytes = kStages * kStageBytes;
  static_assert(kEpiBytes <= kPipeBytes, "epilogue staging reuses the pipeline buffers");
  static constexpr int kSmemBytes = kPipeBytes + 1024 /*align*/ + 256 /*barriers*/ + BN * 4 /*bias*/;
  // Four fp32 accumulators per tile (see the MMA issuer): 3 take the hi*hi products round-robin over
  // k-steps, 1 takes the small hi*lo + lo*hi cross terms; the epilogue adds them with round-to-nearest.
  // The tensor core truncates when it aligns a product block to the running accumulator, so the error of
  // one accumulator grows linearly with the number of accumulations into it (measured 3.5e-9 * K relative
  // with a single accumulator); splitting cuts that by 9x at no extra MMA cost.
  static constexpr int kAccums = 4;
  static constexpr uint32_t kTmemCols = (kAccums * BN <= 256) ? 256 : 512;
  static_assert(kAccums * BN <= 512, "TMEM has 512 columns");
};

struct ConvParams {
  int H, W;                 // output (= input) spatial size
  int cin_chunks;           // Cin / 64
  int taps_h, taps_w;       // 3,3 or 1,1
  int dil, pad;
  int tw, th;               // tile patch, tw*th = 128
  int tiles_x;              // ceil(W / tw)
  int n_tiles;              // Cout / BN
  int cout_offset;          // channel offset inside the destination tensor (concat by construction)
  int relu;
  float out_scale;          // 2^-k undoing the weight pre-scale
  const float* bias;        // [Cout] or nullptr
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_o, const ConvParams p) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kPipeBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 1);
  float* bias_s = reinterpret_cast<float*>(smem + Cfg::kPipeBytes + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * Cfg::kStages);

  // tile coordinates: n-tile fastest so CTAs that share an activation patch are co-scheduled
  const int nt = blockIdx.x % p.n_tiles;
  const int tx = blockIdx.x / p.n_tiles;
  const int x0 = tx * p.tw;
  const int y0 = blockIdx.y * p.th;
  const int img = blockIdx.z;
  const int n0 = nt * BN;
  const int iters = p.taps_h * p.taps_w * p.cin_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_o);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  } else if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tmem_relinquish();
  } else if (warp == 3) {
    for (int c = lane; c < BN; c += 32) bias_s[c] = p.bias ? p.bias[n0 + c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int r = 0; r < p.taps_h; ++r) {
        for (int s = 0; s < p.taps_w; ++s) {
          const int ix = x0 - p.pad + s * p.dil;
          const int iy = y0 - p.pad + r * p.dil;
          const int tap = r * p.taps_w + s;
          for (int cc = 0; cc < p.cin_chunks; ++cc, ++it) {
            const int st = it % Cfg::kStages;
            const uint32_t ph = (it / Cfg::kStages) & 1;
            mbar_wait(empty_bar(st), ph ^ 1);
            mbar_arrive_expect_tx(full_bar(st), Cfg::kStageBytes);
            const uint32_t a_dst = smem_base + st * Cfg::kStageBytes;
            tma_load_5d(a_dst, &tmap_a, full_bar(st), cc * kChunkK, ix, iy, img, 0);
            tma_load_4d(a_dst + kABytes, &tmap_b, full_bar(st), cc * kChunkK, n0, tap, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    constexpr uint32_t idesc = umma_idesc_f16(kTileM, BN);
    uint32_t used = 0;                       // bit a set once accumulator a holds data
    for (int it = 0; it < iters; ++it) {
      const int st = it % Cfg::kStages;
      const uint32_t ph = (it / Cfg::kStages) & 1;
      mbar_wait(full_bar(st), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_base + st * Cfg::kStageBytes;
        const uint32_t a_lo = a_hi + kTileM * 128;
        const uint32_t b_hi = a_hi + kABytes;
        const uint32_t b_lo = b_hi + BN * 128;
#pragma unroll
        for (int k = 0; k < kChunkK / 16; ++k) {
          const uint32_t koff = k * 32;     // 16 fp16 = 32 bytes along the swizzled row
          const uint64_t da_hi = umma_desc_sw128(a_hi + koff), da_lo = umma_desc_sw128(a_lo + koff);
          const uint64_t db_hi = umma_desc_sw128(b_hi + koff), db_lo = umma_desc_sw128(b_lo + koff);
          const int main_acc = (it * (kChunkK / 16) + k) % 3;
          umma_f16(tmem_base + 3 * BN, da_hi, db_lo, idesc, (used >> 3) & 1u);      // cross terms -> accumulator 3
          umma_f16(tmem_base + 3 * BN, da_lo, db_hi, idesc, 1u);
          umma_f16(tmem_base + main_acc * BN, da_hi, db_hi, idesc, (used >> main_acc) & 1u);
          used |= 8u | (1u << main_acc);
        }
        umma_commit(empty_bar(st));          // frees the stage once these MMAs have read it
        if (it == iters - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int w = warp - 4;                  // TMEM lanes 32w .. 32w+31
    const int m = w * 32 + lane;             // accumulator row = pixel (y_local * tw + x_local)
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // all TMA loads have landed and all MMAs have drained: pipeline smem is free for staging
    uint8_t* stage_out = smem;
    const float scale = p.out_scale;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      {
        // sum of the four accumulators, smallest first (cross terms, then the three hi*hi partials)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)c0;
        const int n_main = iters * (kChunkK / 16) >= 3 ? 3 : iters * (kChunkK / 16);
        tmem_ld_32x32(lane_addr + 3 * BN, r);
        tmem_ld_wait();
        for (int a = 0; a < n_main; ++a) {
          uint32_t q[32];
          tmem_ld_32x32(lane_addr + a * BN, q);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__fadd_rn(__uint_as_float(r[e]), __uint_as_float(q[e])));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hi_pk[4], lo_pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c0 + j * 8 + e * 2;
          float v0 = __uint_as_float(r[j * 8 + e * 2]) * scale + bias_s[c];
          float v1 = __uint_as_float(r[j * 8 + e * 2 + 1]) * scale + bias_s[c + 1];
          if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
          __half h0, l0, h1, l1;
          split_h2(v0, h0, l0);
          split_h2(v1, h1, l1);
          hi_pk[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          lo_pk[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        const int c = c0 + j * 8;
        const int t = c >> 6;                          // 64-channel output tile
        const int chunk = (c & 63) >> 3;               // 16-byte chunk within the 128-byte row
        const uint32_t off = (uint32_t)t * (kTileM * 128) + (uint32_t)m * 128 + (uint32_t)((chunk ^ (m & 7)) << 4);
        *reinterpret_cast<uint4*>(stage_out + off) = make_uint4(hi_pk[0], hi_pk[1], hi_pk[2], hi_pk[3]);
        *reinterpret_cast<uint4*>(stage_out + (BN / 64) * (kTileM * 128) + off) =
            make_uint4(lo_pk[0], lo_pk[1], lo_pk[2], lo_pk[3]);
      }
    }
    fence_proxy_async_smem();                          // generic-proxy smem writes -> visible to TMA
    asm volatile("bar.sync 1, 128;" ::: "memory");     // the four epilogue warps only
    if (w == 0 && lane == 0) {
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
#pragma unroll
        for (int t = 0; t < BN / 64; ++t)
          tma_store_5d(&tmap_o, smem_base + (uint32_t)(pl * (BN / 64) + t) * (kTileM * 128),
                       p.cout_offset + n0 + t * 64, x0, y0, img, pl);
      tma_store_commit();
      tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// -------------------------------------------------------------------------------------------------
template <int BN>
int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const ConvParams& p, int batch,
                cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    SHF_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_set = true;
  }
  dim3 grid(p.tiles_x * p.n_tiles, (p.H + p.th - 1) / p.th, batch);
  conv_igemm_kernel<BN><<<grid, 256, Cfg::kSmemBytes, stream>>>(ta, tb, to, p);
  SHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// v1 (one TMA load per tap); the public shf_conv_igemm in conv_halo.cu dispatches here when asked to
int shf_conv_pertap_impl(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H,
                              int W, int cin, int cout, int ksize, int dilation, int out_channels_total,
                              int out_channel_offset, float out_scale, int relu, void* stream) {
  SHF_REQUIRE(ksize == 3 || ksize == 1, "shf_conv_igemm: kernel size %d (only 3x3 and 1x1 are on the hot path)", ksize);
  SHF_REQUIRE(cin % 64 == 0 && cin >= 64, "shf_conv_igemm: Cin=%d must be a multiple of 64", cin);
  SHF_REQUIRE(cout % 64 == 0 && cout >= 64, "shf_conv_igemm: Cout=%d must be a multiple of 64", cout);
  SHF_REQUIRE(out_channel_offset % 8 == 0 && out_channel_offset + cout <= out_channels_total &&
                  out_channels_total % 8 == 0,
              "shf_conv_igemm: bad destination channel window [%d,%d) of %d", out_channel_offset,
              out_channel_offset + cout, out_channels_total);
  SHF_REQUIRE(batch >= 1 && H >= 1 && W >= 1 && dilation >= 1, "shf_conv_igemm: bad geometry");
  const int bn = (cout % 128 == 0) ? 128 : 64;
  const int taps = ksize * ksize;
  ConvParams p;
  p.H = H; p.W = W;
  p.cin_chunks = cin / 64;
  p.taps_h = p.taps_w = ksize;
  p.dil = dilation;
  p.pad = (ksize == 3) ? dilation : 0;     // "same" convolution: every 3x3 on the path has pad == dilation
  p.tw = 16; p.th = 8;
  if (W <= 8) { p.tw = 8; p.th = 16; }
  p.tiles_x = (W + p.tw - 1) / p.tw;
  p.n_tiles = cout / bn;
  p.cout_offset = out_channel_offset;
  p.relu = relu;
  p.out_scale = out_scale;
  p.bias = bias;

  CUtensorMap ta, tb, to;
  {
    uint64_t d[5] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)batch, 2};
    uint32_t b[5] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1, 2};
    if (int e = shf_encode_f16_map(&ta, const_cast<void*>(in_h2), 5, d, b, "activations")) return e;
  }
  {
    uint64_t d[4] = {(uint64_t)cin, (uint64_t)cout, (uint64_t)taps, 2};
    uint32_t b[4] = {64, (uint32_t)bn, 1, 2};
    if (int e = shf_encode_f16_map(&tb, const_cast<void*>(w_h2), 4, d, b, "weights")) return e;
  }
  {
    uint64_t d[5] = {(uint64_t)out_channels_total, (uint64_t)W, (uint64_t)H, (uint64_t)batch, 2};
    uint32_t b[5] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1, 1};
    if (int e = shf_encode_f16_map(&to, out_h2, 5, d, b, "output")) return e;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return bn == 128 ? launch_conv<128>(ta, tb, to, p, batch, st) : launch_conv<64>(ta, tb, to, p, batch, st);
}

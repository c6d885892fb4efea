// Implicit-GEMM convolution v2 on tcgen05 tensor cores (sm_100a): the activation patch of a tile is loaded
// ONCE per 64-channel chunk as a halo tile and every 3x3 tap reads it through a shifted UMMA descriptor.
//
// Same math, layout, precision scheme and epilogue as conv_igemm.cu (see there for the reference citations:
// base_conv_layer.cpp:255-279, im2col.cpp:19-55, relu_layer.cpp:9-19).  What changes is the operand traffic,
// which profiles/r01_conv_igemm_ncu_full.md shows is what bounds v1 (L2->SM ~42 B/clk/SM, tensor pipe 19-57 %):
//
//   v1: per (tap, chunk) stage: A patch 32 KB + B slice 32 KB             -> 35 B per output element per chunk
//   v2: per chunk: one halo (TH+2d) x XW pixels, per tap only the B slice  -> 22 B (XW = 16) / 20 B (XW = TW+2d)
//
// Geometry: CTA tile = TH x TW = 16 x 8 output pixels (GEMM row m = y*8 + x), so one 8-row UMMA group is one
// image row of the tile.  The halo lives in smem as [y][x][64 ch] with 128-byte pixel rows, written by one
// rank-5 TMA box (64, XW, TH+2d, 1, 2 planes) with the 128B swizzle.  For tap (r, s) the A descriptor starts at
// pixel (r*d, s*d) of the halo: start += ((r*d)*XW + s*d)*128, stride between 8-row groups SBO = XW*128.
// The 128B swizzle is a function of the smem address bits [7,10) on both the TMA write and the UMMA read, so a
// start that is only 128-byte aligned reads consistent data; `base_offset` mirrors those bits in the descriptor.
//
// Two mbarrier rings: A halos (1-2 stages, one per chunk) and B weight slices (2-4 stages, one per tap).
#include "common.cuh"
#include "tma_host.cuh"
#include <stdlib.h>

int shf_conv_pertap_impl(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H, int W,
                         int cin, int cout, int ksize, int dilation, int out_channels_total, int out_channel_offset,
                         float out_scale, int relu, void* stream);

int shf_conv_stream_impl(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H, int W,
                         int cin, int cout, int ksize, int dilation, int out_channels_total, int out_channel_offset,
                         float out_scale, int relu, void* pool_out_h2, int pool_channels_total, int pool_channel_offset,
                         int ctas, int in_format, int out_format, void* stream);

namespace {

constexpr int kTileM = 128;
constexpr int kTH = 16, kTW = 8;
constexpr int kChunkK = 64;
constexpr int kAccums = 4;
constexpr int kMaxStages = 6;

struct HaloParams {
  int H, W;
  int cin_chunks, taps, dil, pad;
  int xw, xh;                // halo width / height in pixels
  int na, nb;                // ring depths
  int a_bytes, b_bytes;      // smem bytes per stage (both planes; A rounded up to 1024)
  int a_tx;                  // bytes one halo TMA load actually transfers
  int tiles_x, n_tiles;
  int cout_offset, relu;
  int bo_mode;               // 1: descriptor base_offset = (start >> 7) & 7
  int no_shift;              // timing probe only (wrong results): every tap reads the halo at offset 0
  float out_scale;
  const float* bias;
};

SHF_DEVICE uint64_t umma_desc_halo(uint32_t smem_addr, uint32_t sbo_bytes, int bo_mode) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  if (bo_mode) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_o, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  // all addresses below are derived with integer ops from the (warp-uniform) shared-window offset so that ptxas can
  // keep descriptors in uniform registers
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int pipe_bytes = p.na * p.a_bytes + p.nb * p.b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pipe_bytes);           // [fullA|emptyA|fullB|emptyB][kMaxStages] + tmem_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxStages + 1);
  float* bias_s = reinterpret_cast<float*>(smem + pipe_bytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_base + (uint32_t)pipe_bytes;
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto empty_a = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto full_b = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto empty_b = [&](int s) { return bar_base + 8u * (3 * kMaxStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (4 * kMaxStages);
  auto a_stage = [&](int s) { return smem_base + (uint32_t)(s * p.a_bytes); };
  auto b_stage = [&](int s) { return smem_base + (uint32_t)(p.na * p.a_bytes + s * p.b_bytes); };

  const int nt = blockIdx.x % p.n_tiles;
  const int tx = blockIdx.x / p.n_tiles;
  const int x0 = tx * kTW, y0 = blockIdx.y * kTH, img = blockIdx.z, n0 = nt * BN;
  // BN = 128: the whole tensor memory is allocated (one CTA per SM), so the allocation base is column 0 -- a
  // compile-time constant keeps every accumulator address uniform for the MMA issuers.  BN = 64 needs only half of
  // it, which lets two CTAs share an SM (one's epilogue overlaps the other's main loop); its base is read back.
  constexpr uint32_t kTmemCols = (kAccums * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_o);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(full_a(s), 1); mbar_init(empty_a(s), 2);       // 2 = the two MMA-issuing warps
      mbar_init(full_b(s), 1); mbar_init(empty_b(s), 2);
    }
    mbar_init(tmem_full_bar, 2);
    fence_mbar_init();
  } else if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
    for (int c = lane; c < BN; c += 32) bias_s[c] = p.bias ? p.bias[n0 + c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (kTmemCols == 512 && *tmem_slot != 0u) __trap();
  const uint32_t tmem_base = (kTmemCols == 512) ? 0u : *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: weight slices (one per tap) =====================
    if (lane == 0) {
      int itb = 0;
      for (int cc = 0; cc < p.cin_chunks; ++cc) {
        for (int tap = 0; tap < p.taps; ++tap, ++itb) {
          const int sb = itb % p.nb;
          mbar_wait(empty_b(sb), ((itb / p.nb) & 1) ^ 1);
          mbar_arrive_expect_tx(full_b(sb), p.b_bytes);
          tma_load_4d(b_stage(sb), &tmap_b, full_b(sb), cc * kChunkK, n0, tap, 0);
        }
      }
    }
  } else if (warp == 3) {
    // ===================== TMA producer: activation halos (one per 64-channel chunk) =====================
    // its own warp so the next halo is requested as soon as its stage drains, a whole chunk (9 taps) ahead
    if (lane == 0) {
      for (int cc = 0; cc < p.cin_chunks; ++cc) {
        const int sa = cc % p.na;
        mbar_wait(empty_a(sa), ((cc / p.na) & 1) ^ 1);
        mbar_arrive_expect_tx(full_a(sa), p.a_tx);
        tma_load_5d(a_stage(sa), &tmap_a, full_a(sa), cc * kChunkK, x0 - p.pad, y0 - p.pad, img, 0);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers (two warps, each converged with one elected lane issuing) =====================
    // The single issuing thread was the bottleneck of v1 (~23 SASS instructions per MMA vs a 64-cycle MMA), so the
    // three products of a k-step are split by accumulator: warp 1 issues hi*hi into accumulators 0..2 (round-robin),
    // warp 2 issues hi*lo and lo*hi into accumulator 3.  Disjoint accumulators -> no ordering between the two
    // threads is needed; both commit to the same empty/full barriers (arrival count 2).
    constexpr uint32_t idesc = umma_idesc_f16(kTileM, BN);
    const bool main_warp = (warp == 1);
    const uint32_t a_hi32 = (uint32_t)((p.xw * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B
    constexpr uint32_t b_hi32 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t lbo = 1u << 16;
    const uint32_t a_plane16 = ((uint32_t)(p.xh * p.xw) * 128u) >> 4;     // lo-plane offset in descriptor units
    const int ktaps = (p.taps == 9) ? 3 : 1;
    uint32_t used = 0;
    int itb = 0, ks = 0;
    for (int cc = 0; cc < p.cin_chunks; ++cc) {
      const int sa = cc % p.na;
      mbar_wait(full_a(sa), (cc / p.na) & 1);
      const uint32_t a_base = a_stage(sa);
      int r = 0, s = 0;
      for (int tap = 0; tap < p.taps; ++tap, ++itb) {
        const int sb = itb % p.nb;
        mbar_wait(full_b(sb), (itb / p.nb) & 1);
        tc_fence_after();
        const uint32_t a_off = p.no_shift ? 0u : (uint32_t)((r * p.dil) * p.xw + s * p.dil) * 128u;
        uint32_t a_lo32 = (((a_base + a_off) & 0x3FFFFu) >> 4) | lbo;     // descriptor low word of the hi plane
        uint32_t b_lo32 = ((b_stage(sb) & 0x3FFFFu) >> 4) | lbo;
        if (main_warp) {
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k, ++ks) {
            const int acc = ks % 3;
            umma_f16_elect_lohi(tmem_base + acc * BN, a_lo32, a_hi32, b_lo32, b_hi32, idesc, (used >> acc) & 1u);
            used |= 1u << acc;
            a_lo32 += 2; b_lo32 += 2;              // 16 fp16 = 32 bytes along the swizzled row (16-byte units)
          }
        } else {
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k) {
            umma_f16_elect_lohi(tmem_base + 3 * BN, a_lo32, a_hi32, b_lo32 + ((BN * 128) >> 4), b_hi32, idesc, used);
            umma_f16_elect_lohi(tmem_base + 3 * BN, a_lo32 + a_plane16, a_hi32, b_lo32, b_hi32, idesc, 1u);
            used = 1u;
            a_lo32 += 2; b_lo32 += 2;
          }
        }
        umma_commit_elect(empty_b(sb));
        if (tap == p.taps - 1) {
          umma_commit_elect(empty_a(sa));
          if (cc == p.cin_chunks - 1) umma_commit_elect(tmem_full_bar);
        }
        if (++s == ktaps) { s = 0; ++r; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (identical to v1) =====================
    const int w = warp - 4;
    const int m = w * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    uint8_t* stage_out = smem;
    const float scale = p.out_scale;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      {
        const uint32_t lane_addr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)c0;
        tmem_ld_32x32(lane_addr + 3 * BN, r);
        tmem_ld_wait();
#pragma unroll 1
        for (int a = 0; a < 3; ++a) {
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
        const int t = c >> 6, chunk = (c & 63) >> 3;
        const uint32_t off = (uint32_t)t * (kTileM * 128) + (uint32_t)m * 128 + (uint32_t)((chunk ^ (m & 7)) << 4);
        *reinterpret_cast<uint4*>(stage_out + off) = make_uint4(hi_pk[0], hi_pk[1], hi_pk[2], hi_pk[3]);
        *reinterpret_cast<uint4*>(stage_out + (BN / 64) * (kTileM * 128) + off) =
            make_uint4(lo_pk[0], lo_pk[1], lo_pk[2], lo_pk[3]);
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (w == 0 && lane == 0) {
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
#pragma unroll
        for (int t = 0; t < BN / 64; ++t)
          tma_store_5d(&tmap_o, smem_base + (uint32_t)(pl * (BN / 64) + t) * (kTileM * 128),
                       p.cout_offset + n0 + t * 64, x0, y0, img, pl);
      tma_store_commit();
      tma_store_wait_read();       // smem may be released once the bulk store has read it; the writes drain on their own
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// implementation selector (shf_set_conv_impl; tests / tuning):
//   0 = v1 per-tap loads (conv_igemm.cu)      1 = halo, XW = 16        3 = halo, XW = 8 + 2d (default)
//   5 = as 3 but 64-channel N tiles whenever Cout <= 128 (two CTAs per SM)
//   7 = v3 persistent streaming-drain kernel (conv_stream.cu), one CTA per tile
//   8 = the same kernel on CTA pairs (cta_group::2, weight stage split across the two SMs of a TPC) (default)
//   2 / 4 = as 1 / 3 with descriptor base_offset = (start >> 7) & 7 -- measured WRONG on B200: the UMMA swizzle is a
//           function of the absolute smem address, base_offset must stay 0 (kept only as a regression probe)
int g_conv_impl = 8;

template <int BN>
int launch_halo(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const HaloParams& p, int batch,
                int smem_bytes, cudaStream_t stream) {
  static int attr = 0;
  if (smem_bytes > attr) {
    SHF_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = 227 * 1024;
  }
  dim3 grid(p.tiles_x * p.n_tiles, (p.H + kTH - 1) / kTH, batch);
  conv_halo_kernel<BN><<<grid, 256, smem_bytes, stream>>>(ta, tb, to, p);
  SHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Convolution + ReLU + 2x2/2 max pooling in one launch (conv_stream.cu).  out_h2 may be NULL when only the pooled map
// is consumed downstream (conv1_2, conv2_2, conv3_3 of VGG16); conv4_3 needs both.
extern "C" int shf_conv_igemm_pool(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, void* pool_out_h2,
                                   int batch, int H, int W, int cin, int cout, int ksize, int dilation,
                                   int out_channels_total, int out_channel_offset, int pool_channels_total,
                                   int pool_channel_offset, float out_scale, int relu, int in_format, int out_format,
                                   void* stream) {
  SHF_REQUIRE(pool_out_h2 != nullptr, "shf_conv_igemm_pool: pool_out_h2 is NULL");
  return shf_conv_stream_impl(in_h2, w_h2, bias, out_h2, batch, H, W, cin, cout, ksize, dilation, out_channels_total,
                              out_channel_offset, out_scale, relu, pool_out_h2, pool_channels_total,
                              pool_channel_offset, g_conv_impl == 7 ? 1 : 2, in_format, out_format, stream);
}

extern "C" int shf_set_conv_impl(int impl) {
  SHF_REQUIRE(impl >= 0 && impl <= 8, "shf_set_conv_impl: %d", impl);
  g_conv_impl = impl;
  return 0;
}

// C ABI -- see include/shf_b200.h
extern "C" int shf_conv_igemm(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H,
                              int W, int cin, int cout, int ksize, int dilation, int out_channels_total,
                              int out_channel_offset, float out_scale, int relu, int in_format, int out_format,
                              void* stream) {
  if (g_conv_impl >= 7)
    return shf_conv_stream_impl(in_h2, w_h2, bias, out_h2, batch, H, W, cin, cout, ksize, dilation, out_channels_total,
                                out_channel_offset, out_scale, relu, nullptr, 0, 0, g_conv_impl == 7 ? 1 : 2, in_format,
                                out_format, stream);
  SHF_REQUIRE(in_format == SHF_FMT_H2 && out_format == SHF_FMT_H2,
              "shf_conv_igemm: the v1/v2 kernels (shf_set_conv_impl < 7) only know the h2 format");
  if (g_conv_impl == 0)
    return shf_conv_pertap_impl(in_h2, w_h2, bias, out_h2, batch, H, W, cin, cout, ksize, dilation, out_channels_total,
                                out_channel_offset, out_scale, relu, stream);
  SHF_REQUIRE(ksize == 3 || ksize == 1, "shf_conv_igemm: kernel size %d (only 3x3 and 1x1 are on the hot path)", ksize);
  SHF_REQUIRE(cin % 64 == 0 && cin >= 64, "shf_conv_igemm: Cin=%d must be a multiple of 64", cin);
  SHF_REQUIRE(cout % 64 == 0 && cout >= 64, "shf_conv_igemm: Cout=%d must be a multiple of 64", cout);
  SHF_REQUIRE(out_channel_offset % 8 == 0 && out_channel_offset + cout <= out_channels_total &&
                  out_channels_total % 8 == 0,
              "shf_conv_igemm: bad destination channel window [%d,%d) of %d", out_channel_offset,
              out_channel_offset + cout, out_channels_total);
  SHF_REQUIRE(batch >= 1 && H >= 1 && W >= 1 && dilation >= 1 && dilation <= 4, "shf_conv_igemm: bad geometry");
  const int bn = (cout % 128 == 0 && !(g_conv_impl == 5 && cout <= 128)) ? 128 : 64;
  HaloParams p;
  p.H = H; p.W = W;
  p.cin_chunks = cin / 64;
  p.taps = ksize * ksize;
  p.dil = (ksize == 3) ? dilation : 0;
  p.pad = p.dil;
  const bool narrow = g_conv_impl >= 3;
  p.xw = (ksize == 1) ? kTW : (narrow ? kTW + 2 * p.pad : 16);
  p.xh = kTH + 2 * p.pad;
  p.bo_mode = (g_conv_impl == 2 || g_conv_impl == 4) ? 1 : 0;
  p.no_shift = shf_probe_env("SHF_PROBE_NOSHIFT") ? 1 : 0;  // timing probe (wrong results): isolates the cost of unaligned groups
  p.a_tx = 2 * p.xh * p.xw * 128;
  p.a_bytes = (p.a_tx + 1023) & ~1023;               // keep every stage 1024-byte aligned
  p.b_bytes = 2 * bn * 128;
  const int budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - bn * 4 - 64;
  p.na = (2 * p.a_bytes + 3 * p.b_bytes <= budget && p.cin_chunks > 1) ? 2 : 1;
  p.nb = (budget - p.na * p.a_bytes) / p.b_bytes;
  if (p.nb > kMaxStages) p.nb = kMaxStages;
  if (const char* e = shf_probe_env("SHF_PROBE_NB")) { int v = atoi(e); if (v >= 2 && v < p.nb) p.nb = v; }
  if (const char* e = shf_probe_env("SHF_PROBE_NA")) { int v = atoi(e); if (v >= 1 && v <= 2 && v <= p.na) p.na = v; }
  if (bn == 64) {
    // 4 x 64 TMEM columns = half the SM's tensor memory: keep shared memory under half an SM as well so that two
    // CTAs are co-resident and one's epilogue overlaps the other's main loop
    const int half = (228 * 1024) / 2 - 1024 - 1024 - 256 - bn * 4 - 64;
    while (p.nb > 3 && p.na * p.a_bytes + p.nb * p.b_bytes > half) --p.nb;
    if (p.na == 2 && p.na * p.a_bytes + p.nb * p.b_bytes > half && p.a_bytes + 3 * p.b_bytes <= half) {
      p.na = 1;
      p.nb = (half - p.a_bytes) / p.b_bytes;
      if (p.nb > kMaxStages) p.nb = kMaxStages;
    }
  }
  SHF_REQUIRE(p.nb >= 2, "shf_conv_igemm: halo tile of %d bytes leaves no room for the weight ring", p.a_bytes);
  SHF_REQUIRE(p.na * p.a_bytes + p.nb * p.b_bytes >= 2 * kTileM * bn * 2, "shf_conv_igemm: staging does not fit");
  p.tiles_x = (W + kTW - 1) / kTW;
  p.n_tiles = cout / bn;
  p.cout_offset = out_channel_offset;
  p.relu = relu;
  p.out_scale = out_scale;
  p.bias = bias;
  const int smem_bytes = p.na * p.a_bytes + p.nb * p.b_bytes + 1024 + 256 + bn * 4 + 64;

  CUtensorMap ta, tb, to;
  {
    uint64_t d[5] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)batch, 2};
    uint32_t b[5] = {64, (uint32_t)p.xw, (uint32_t)p.xh, 1, 2};
    if (int e = shf_encode_f16_map(&ta, const_cast<void*>(in_h2), 5, d, b, "activations")) return e;
  }
  {
    uint64_t d[4] = {(uint64_t)cin, (uint64_t)cout, (uint64_t)p.taps, 2};
    uint32_t b[4] = {64, (uint32_t)bn, 1, 2};
    if (int e = shf_encode_f16_map(&tb, const_cast<void*>(w_h2), 4, d, b, "weights")) return e;
  }
  {
    uint64_t d[5] = {(uint64_t)out_channels_total, (uint64_t)W, (uint64_t)H, (uint64_t)batch, 2};
    uint32_t b[5] = {64, (uint32_t)kTW, (uint32_t)kTH, 1, 1};
    if (int e = shf_encode_f16_map(&to, out_h2, 5, d, b, "output")) return e;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return bn == 128 ? launch_halo<128>(ta, tb, to, p, batch, smem_bytes, st)
                   : launch_halo<64>(ta, tb, to, p, batch, smem_bytes, st);
}

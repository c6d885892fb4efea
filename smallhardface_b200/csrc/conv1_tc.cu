// conv1_1 on the tensor cores (sm_100a): 3 input channels, 3x3, pad 1, 64 output channels, + bias + ReLU.
// Replaces base_conv_layer.cpp:255-279 (im2col + sgemm) for the K = 27 first layer.
//
// The SIMT kernel (conv3x3_c3_kernel, simt_kernels.cu) needs 1728 fp32 FMAs per pixel and ran at ~30 % of the HBM
// roofline, fp32-issue bound.  Here the 27 taps of a pixel become ONE 128-byte K-major operand row
//     A[m] = [ x_hi(k = 0..31) | x_lo(k = 0..31) ]        k = c*9 + r*3 + s, zero for k >= 27, x = x_hi + x_lo (fp16)
// built by the pixel's own thread straight from the fp32 NCHW level blob (zero padding = skipped loads), and the layer
// is six M128 x N64 x K16 tcgen05 MMAs per 128-pixel tile:
//     D  = A[:, 0:64]  x [ w_hi | w_hi ]^T      (x_hi*w_hi + x_lo*w_hi,  4 k-steps)
//     D += A[:, 0:32]  x [ w_lo ]^T             (x_hi*w_lo,              2 k-steps)
// i.e. split-fp16 operands (2^-22) like the h2 path of conv_stream.cu; K is too short for the accumulator truncation
// to matter, so one accumulator takes all three products.  The epilogue is conv_stream's: bias, ReLU, conversion to
// h2 or hf8, rows staged through shared memory and written as full 128-byte lines.  What remains is the 256 B per
// pixel that must reach HBM.
#include "common.cuh"
#include "epilogue_store.cuh"

namespace {

struct C1Params {
  const float* in;          // (N, 3, H, W) fp32
  const __half* wpack;      // [2][64][64] fp16: [0] = rows [w_hi | w_hi], [1] = rows [w_lo | 0], of w * 2^k
  const float* bias;        // 64 or nullptr
  __half* out;              // activation tensor (N, H, W, 64), plane 0
  long long plane_elems;
  int N, H, W, relu, out_fmt;
  int tiles_x, tiles_y, total_tiles;
  float out_scale;          // 2^-k
  unsigned int* guard;      // range guard slot or nullptr (common.cuh)
};

constexpr int kC1Threads = 128;
constexpr int kTH = 16, kTW = 8;

__global__ void __launch_bounds__(kC1Threads, 4) conv1_tc_kernel(const C1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  // [A tile 16 KB][B1 8 KB][B2 8 KB][staging 4 x 4 KB][mbarrier 8][tmem slot 4][bias 256][halo 540 words]
  uint8_t* a_tile = smem;
  uint8_t* b1 = smem + 16384;
  uint8_t* b2 = smem + 24576;
  uint8_t* stage_s = smem + 32768;
  const uint32_t bar = smem_base + 49152;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 49152 + 8);
  float* bias_s = reinterpret_cast<float*>(smem + 49152 + 16);
  uint32_t* halo_s = reinterpret_cast<uint32_t*>(smem + 49152 + 16 + 256);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int m = threadIdx.x;                      // operand / accumulator row = pixel (y_local * 8 + x_local)

  // weights -> shared memory in the 128-byte-swizzled K-major layout the MMA descriptor expects (16-byte chunk j of
  // row n at position j ^ (n & 7))
  for (int i = threadIdx.x; i < 2 * 64 * 8; i += kC1Threads) {
    const int which = i >> 9, n = (i >> 3) & 63, j = i & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.wpack) + i);
    *reinterpret_cast<uint4*>((which ? b2 : b1) + n * 128 + ((j ^ (n & 7)) << 4)) = v;
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 64);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  constexpr uint32_t idesc = umma_idesc_f16(128, 64);
  const uint64_t a_desc = umma_desc_sw128(smem_base);
  const uint64_t b1_desc = umma_desc_sw128(smem_base + 16384);
  const uint64_t b2_desc = umma_desc_sw128(smem_base + 24576);
  const bool issuer = (warp == 0) && elect_one();
  uint32_t phase = 0;
  float gmax = 0.f;

  // ---- halo staging (r02b): the 18 x 10 x 3 input patch of a tile is fetched ONCE by the CTA (each thread owns up to
  //      five fixed patch positions), split to fp16 hi / lo there and parked in shared memory as (hi | lo << 16) words;
  //      a pixel's operand row is then 27 LDS + 28 byte permutes.  The per-pixel version issued 27 predicated global
  //      loads, 27 splits and their address arithmetic per thread: 630 of the loop's 1740 instructions in a kernel
  //      that ncu showed issue-bound (215 M warp instructions for 1 GB of output, 2.9 TB/s).
  constexpr int kHW = kTW + 2, kHH = kTH + 2, kHaloWords = 3 * kHH * kHW;      // 540
  constexpr int kPer = (kHaloWords + kC1Threads - 1) / kC1Threads;             // 5
  int h_rs[kPer];            // (channel << 16) | (row << 8) | column inside the patch, or -1 when the slot is unused
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int idx = (int)threadIdx.x + i * kC1Threads;
    const int c = idx / (kHH * kHW), rem = idx % (kHH * kHW), r = rem / kHW, s2 = rem % kHW;
    h_rs[i] = idx < kHaloWords ? ((c << 16) | (r << 8) | s2) : -1;
  }
  float nxt[kPer];           // the next tile's patch values, in flight across the current tile's MMA wait + epilogue
  int nx0 = 0, ny0 = 0, nimg = 0;                  // ... and its origin (decoded once per tile)
  auto prefetch = [&](int t) {
    const unsigned tx = (unsigned)p.tiles_x, ty = (unsigned)p.tiles_y;
    unsigned q = (unsigned)t;
    const int x0 = (int)(q % tx) * kTW;
    q /= tx;
    const int y0 = (int)(q % ty) * kTH, img = (int)(q / ty);
    nx0 = x0; ny0 = y0; nimg = img;
    const float* org = p.in + (size_t)img * 3 * p.H * p.W + ((long long)(y0 - 1) * p.W + (x0 - 1));
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int c = h_rs[i] >> 16, r = (h_rs[i] >> 8) & 255, s2 = h_rs[i] & 255;
      const int iy = y0 - 1 + r, ix = x0 - 1 + s2;
      const bool ok = h_rs[i] >= 0 && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
      nxt[i] = ok ? __ldg(org + ((c * p.H + r) * p.W + s2)) : 0.f;      // zero padding = skipped loads
    }
  };
  if ((int)blockIdx.x < p.total_tiles) prefetch(blockIdx.x);
  const uint32_t* my_halo = halo_s + (m >> 3) * kHW + (m & 7);

  for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
    const int x0 = nx0, y0 = ny0, img = nimg;
    const int y = y0 + (m >> 3), x = x0 + (m & 7);
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      __half hi, lo;
      split_h2(nxt[i], hi, lo);
      if (h_rs[i] >= 0)
        halo_s[threadIdx.x + i * kC1Threads] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
    }
    __syncthreads();
    // ---- this pixel's operand row: 27 taps as [hi(k = 0..31) | lo(k = 0..31)], k = c*9 + r*3 + s ----
    {
      uint32_t v[32];
#pragma unroll
      for (int k = 27; k < 32; ++k) v[k] = 0u;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) v[c * 9 + r * 3 + s2] = my_halo[(c * kHH + r) * kHW + s2];
      uint8_t* arow = a_tile + m * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k0 = 8 * (j & 3) + 2 * e;
          w4[e] = __byte_perm(v[k0], v[k0 + 1], j < 4 ? 0x5410 : 0x7632);     // the hi halves / the lo halves of two taps
        }
        *reinterpret_cast<uint4*>(arow + ((j ^ (m & 7)) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
    }
    fence_proxy_async_smem();                      // generic-proxy writes above -> visible to the tensor core
    __syncthreads();
    if (issuer) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_acc, a_desc + 2 * k, b1_desc + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 2; ++k) umma_f16(tmem_acc, a_desc + 2 * k, b2_desc + 2 * k, idesc, 1u);
      umma_commit(bar);
    }
    if (t + (int)gridDim.x < p.total_tiles) prefetch(t + (int)gridDim.x);
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    float v[64];
    {
      uint32_t r0[32], r1[32];
      const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16);
      tmem_ld_32x32(taddr, r0);
      tmem_ld_32x32(taddr + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        v[e] = fmaf(__uint_as_float(r0[e]), p.out_scale, bias_s[e]);
        v[32 + e] = fmaf(__uint_as_float(r1[e]), p.out_scale, bias_s[32 + e]);
      }
    }
    tc_fence_before();
    __syncthreads();                               // accumulator and operand tile may be overwritten by the next tile
    if (p.relu) {
#pragma unroll
      for (int e = 0; e < 64; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    if (p.guard && y < p.H && x < p.W) {
#pragma unroll
      for (int e = 0; e < 64; ++e) gmax = fmaxf(gmax, fabsf(v[e]));
    }
    uint8_t* stg = stage_s + warp * 4096;
    const int qy = y0 + warp * 4;
    __half* base = p.out + (((size_t)img * p.H + qy) * p.W + x0) * (size_t)64;
    const uint32_t pitch = (uint32_t)p.W * 64u;
    auto dst = [&](int row) -> __half* {
      const int dy = row >> 3, dx = row & 7;
      return (qy + dy < p.H && x0 + dx < p.W) ? base + ((uint32_t)dy * pitch + (uint32_t)dx * 64u) : nullptr;
    };
    store_plane<64>(stg, lane, v, 0, p.out_fmt, 0, (size_t)p.plane_elems, dst);
    store_plane<64>(stg, lane, v, 1, p.out_fmt, 0, (size_t)p.plane_elems, dst);
  }
  range_guard_commit(p.guard, gmax);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, 64);
  }
}

}  // namespace

// C ABI -- see include/shf_b200.h
extern "C" int shf_conv1_tc(const float* in_nchw, const void* w_packed, const float* bias, void* out_act, int batch,
                            int H, int W, int cout, float out_scale, int relu, int out_format,
                            unsigned int* range_guard, void* stream) {
  SHF_REQUIRE(cout == 64, "shf_conv1_tc: Cout=%d (the deploy nets' conv1_1 has 64)", cout);
  SHF_REQUIRE(out_format == SHF_FMT_H2 || out_format == SHF_FMT_HF8, "shf_conv1_tc: unknown activation format %d", out_format);
  SHF_REQUIRE(batch >= 1 && H >= 1 && W >= 1, "shf_conv1_tc: bad geometry");
  C1Params p;
  p.in = in_nchw;
  p.wpack = reinterpret_cast<const __half*>(w_packed);
  p.bias = bias;
  p.out = reinterpret_cast<__half*>(out_act);
  p.plane_elems = (long long)batch * H * W * 64;
  p.N = batch; p.H = H; p.W = W; p.relu = relu; p.out_fmt = out_format;
  p.tiles_x = (W + kTW - 1) / kTW;
  p.tiles_y = (H + kTH - 1) / kTH;
  p.total_tiles = p.tiles_x * p.tiles_y * batch;
  p.out_scale = out_scale;
  p.guard = range_guard;
  const int smem_bytes = 1024 + 49152 + 16 + 256 + 540 * 4;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  static bool attr[64] = {};                     // function attributes are per device
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    SHF_CUDA_CHECK(cudaFuncSetAttribute(conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int max_grid = sms * 4;                  // four resident CTAs per SM overlap gather, MMA and store phases
  const int grid = p.total_tiles < max_grid ? p.total_tiles : max_grid;
  conv1_tc_kernel<<<grid, kC1Threads, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  SHF_LAUNCH_CHECK();
  return 0;
}

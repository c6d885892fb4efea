// First convolutions (3 input channels -> 64) on the tensor cores (sm_100a), + bias + ReLU (BatchNorm / Scale folded in by
// the caller): the ResNet conv1 (7x7, stride 2) and the VGG16 conv1_1 (3x3, stride 1) as two instantiations of one kernel.
// conv_layer.cpp:8-28 / base_conv_layer.cpp:255-279 semantics; the fp32 SIMT twin is conv_first_kernel (resnet_kernels.cu),
// which stays the path for every other kernel size / stride.
//
// The K = 3*KS*KS taps of an output pixel (147 / 27) form one K-major operand row
//     A[m] = [ x_hi(k = 0..Kp-1) | x_lo(k = 0..Kp-1) ]     k = (c*KS + r)*KS + s, zero for k >= K, Kp = K rounded up to 16
// (160 / 32 halves per part: five / one 128-byte-swizzled column blocks of a 128-row tile), x = x_hi + x_lo (fp16), and the
// layer is 3 * Kp/16 (30 / 6) M128 x N64 x K16 tcgen05 MMAs per tile of 16 x 8 OUTPUT pixels:
//     D  = A[:,  0:Kp ] x W_hi^T   (x_hi * w_hi)
//     D += A[:, Kp:2Kp] x W_hi^T   (x_lo * w_hi)
//     D += A[:,  0:Kp ] x W_lo^T   (x_hi * w_lo)
// The fp32 input patch of a tile is fetched once per CTA (prefetched across the previous tile's epilogue), split to fp16
// hi / lo once and staged in shared memory as (hi | lo << 16) words; TWO threads share a pixel: each builds half of its
// operand row and, in the epilogue, converts and stores half of its 64 channels (32 accumulator registers per thread, so
// three 256-thread CTAs fit an SM for the 3x3 kernel: 24 warps against the 16 of conv1_tc.cu, its one-thread-per-pixel
// predecessor, which is kept as the regression twin).
#include "common.cuh"
#include "epilogue_store.cuh"

namespace {

struct FcParams {
  const float* in;          // (N, 3, H, W) fp32
  const __half* wpack;      // [2][64][64 * WB] fp16: [0] = w_hi(k), [1] = w_lo(k) of w * 2^e, k zero-padded to whole 64-half blocks
  const float* bias;        // 64 or nullptr
  __half* out;              // activation tensor (N, HO, WO, 64), plane 0
  long long plane_elems;
  int N, H, W, HO, WO, pad, relu, out_fmt;
  int tiles_x, tiles_y, total_tiles;
  float out_scale;          // 2^-e
  unsigned int* guard;
};

constexpr int kThreadsFc = 256;
constexpr int kTH = 16, kTW = 8;                       // output pixels per tile
constexpr int kABlock = 128 * 128, kBBlock = 64 * 128; // bytes of one 64-half column block of A / of a weight matrix

template <int KS, int ST>
struct Fc {
  static constexpr int kPH = (kTH - 1) * ST + KS;      // input rows / columns of a tile's patch (37 x 21, 18 x 10)
  static constexpr int kPW = (kTW - 1) * ST + KS;
  static constexpr int kHaloWords = 3 * kPH * kPW;
  static constexpr int kPer = (kHaloWords + kThreadsFc - 1) / kThreadsFc;
  static constexpr int kTaps = 3 * KS * KS;
  static constexpr int kKSteps = (kTaps + 15) / 16;    // k-steps (16 halves) per operand part
  static constexpr int kChunksPart = 2 * kKSteps;      // 16-byte chunks per part (hi or lo) of a row
  static constexpr int kABlocks = (2 * kChunksPart + 7) / 8;
  static constexpr int kWBlocks = (kChunksPart + 7) / 8;
  // shared memory: [A][W_hi][W_lo][staging 8 x 2 KB][barrier, tmem slot][bias][patch]
  static constexpr int kOffWhi = kABlocks * kABlock;
  static constexpr int kOffWlo = kOffWhi + kWBlocks * kBBlock;
  static constexpr int kOffStage = kOffWlo + kWBlocks * kBBlock;
  static constexpr int kOffBar = kOffStage + 8 * 2048;
  static constexpr int kOffBias = kOffBar + 16;
  static constexpr int kOffHalo = kOffBias + 256;
  static constexpr int kSmem = kOffHalo + kHaloWords * 4;
};

// Half of the chunks (8 taps each) of the hi part of an operand row and their lo twins, from the staged patch: every tap's
// patch offset is a compile-time constant.
template <int KS, int ST, int PART>
SHF_DEVICE void build_row_half(uint8_t* arow, const uint32_t* my_halo, int m) {
  using C = Fc<KS, ST>;
  constexpr int kMine = C::kChunksPart / 2;
#pragma unroll
  for (int qq = 0; qq < kMine; ++qq) {
    uint32_t v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = 8 * (kMine * PART + qq) + e;            // tap index (c * KS + r) * KS + s
      v[e] = k < C::kTaps ? my_halo[((k / (KS * KS)) * C::kPH + (k % (KS * KS)) / KS) * C::kPW + (k % KS)] : 0u;
    }
    uint32_t h4[4], l4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h4[e] = __byte_perm(v[2 * e], v[2 * e + 1], 0x5410);  // the hi halves of two taps
      l4[e] = __byte_perm(v[2 * e], v[2 * e + 1], 0x7632);  // ... and their lo halves
    }
    const int ch_hi = kMine * PART + qq, ch_lo = C::kChunksPart + ch_hi;   // chunk index inside the row
    *reinterpret_cast<uint4*>(arow + (ch_hi >> 3) * kABlock + (((ch_hi & 7) ^ (m & 7)) << 4)) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
    *reinterpret_cast<uint4*>(arow + (ch_lo >> 3) * kABlock + (((ch_lo & 7) ^ (m & 7)) << 4)) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
  }
}

template <int KS, int ST, int MINB>
__global__ void __launch_bounds__(kThreadsFc, MINB) conv_first_tc_kernel(const FcParams p) {
  using C = Fc<KS, ST>;
  constexpr int kPer = C::kPer, kPH = C::kPH, kPW = C::kPW, kHaloWords = C::kHaloWords, kKSteps = C::kKSteps;
  constexpr int kOffWhi = C::kOffWhi, kOffWlo = C::kOffWlo, kOffStage = C::kOffStage, kOffBar = C::kOffBar;
  constexpr int kOffBias = C::kOffBias, kOffHalo = C::kOffHalo, kST = ST;
  constexpr int kWChunks = C::kWBlocks * 8;               // 16-byte chunks per packed weight row

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* stage_s = smem + kOffStage;
  const uint32_t bar = smem_base + kOffBar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8);
  float* bias_s = reinterpret_cast<float*>(smem + kOffBias);
  uint32_t* halo_s = reinterpret_cast<uint32_t*>(smem + kOffHalo);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int quad = warp & 3, part = warp >> 2;   // warps w and w + 4 share a TMEM lane quadrant = 32 pixels
  const int m = quad * 32 + lane;                // operand / accumulator row = output pixel (y_local * 8 + x_local)

  // weights -> shared memory, 128-byte-swizzled K-major column blocks (16-byte chunk j of row n at j ^ (n & 7))
  for (int i = threadIdx.x; i < 2 * 64 * kWChunks; i += kThreadsFc) {
    const int which = i / (64 * kWChunks), n = (i / kWChunks) % 64, ch = i % kWChunks;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.wpack) + i);
    *reinterpret_cast<uint4*>(smem + (which ? kOffWlo : kOffWhi) + (ch >> 3) * kBBlock + n * 128 + (((ch & 7) ^ (n & 7)) << 4)) = v;
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
  const uint64_t whi_desc = umma_desc_sw128(smem_base + kOffWhi);
  const uint64_t wlo_desc = umma_desc_sw128(smem_base + kOffWlo);
  const bool issuer = (warp == 0) && elect_one();
  uint32_t phase = 0;
  float gmax = 0.f;

  int h_rs[kPer];            // (channel << 16) | (row << 8) | column of this thread's patch positions, -1 = unused slot
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int idx = (int)threadIdx.x + i * kThreadsFc;
    const int c = idx / (kPH * kPW), rem = idx % (kPH * kPW), r = rem / kPW, s2 = rem % kPW;
    h_rs[i] = idx < kHaloWords ? ((c << 16) | (r << 8) | s2) : -1;
  }
  float nxt[kPer];
  int nx0 = 0, ny0 = 0, nimg = 0;
  auto prefetch = [&](int t) {
    const unsigned tx = (unsigned)p.tiles_x, ty = (unsigned)p.tiles_y;
    unsigned q = (unsigned)t;
    const int x0 = (int)(q % tx) * kTW;
    q /= tx;
    const int y0 = (int)(q % ty) * kTH, img = (int)(q / ty);
    nx0 = x0; ny0 = y0; nimg = img;
    const int iy0 = y0 * kST - p.pad, ix0 = x0 * kST - p.pad;
    const float* org = p.in + (size_t)img * 3 * p.H * p.W;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int c = h_rs[i] >> 16, r = (h_rs[i] >> 8) & 255, s2 = h_rs[i] & 255;
      const int iy = iy0 + r, ix = ix0 + s2;
      const bool ok = h_rs[i] >= 0 && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
      nxt[i] = ok ? __ldg(org + ((size_t)(c * p.H + iy) * p.W + ix)) : 0.f;      // zero padding = skipped loads
    }
  };
  if ((int)blockIdx.x < p.total_tiles) prefetch(blockIdx.x);
  const uint32_t* my_halo = halo_s + ((m >> 3) * kST) * kPW + (m & 7) * kST;

  for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
    const int x0 = nx0, y0 = ny0, img = nimg;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      __half hi, lo;
      split_h2(nxt[i], hi, lo);
      if (h_rs[i] >= 0)
        halo_s[threadIdx.x + i * kThreadsFc] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
    }
    __syncthreads();
    // ---- this pixel's operand row: this thread builds half of the chunks of the hi part (8 taps each) and their twins in
    //      the lo part; part is warp-uniform ----
    if (part == 0) build_row_half<KS, ST, 0>(smem + m * 128, my_halo, m);
    else build_row_half<KS, ST, 1>(smem + m * 128, my_halo, m);
    fence_proxy_async_smem();                      // generic-proxy writes above -> visible to the tensor core
    __syncthreads();
    if (issuer) {
      tc_fence_after();
      // k-step j = halves [16 j, 16 j + 16) of a row: column block j / 4, 32-byte slice j % 4 (descriptor + 2 per slice)
      auto ka = [&](int j) { return a_desc + (uint64_t)((j >> 2) * (kABlock >> 4) + (j & 3) * 2); };
      auto kb = [&](uint64_t d, int j) { return d + (uint64_t)((j >> 2) * (kBBlock >> 4) + (j & 3) * 2); };
#pragma unroll
      for (int j = 0; j < kKSteps; ++j) umma_f16(tmem_acc, ka(j), kb(whi_desc, j), idesc, j ? 1u : 0u);
#pragma unroll
      for (int j = 0; j < kKSteps; ++j) umma_f16(tmem_acc, ka(kKSteps + j), kb(whi_desc, j), idesc, 1u);
#pragma unroll
      for (int j = 0; j < kKSteps; ++j) umma_f16(tmem_acc, ka(j), kb(wlo_desc, j), idesc, 1u);
      umma_commit(bar);
    }
    if (t + (int)gridDim.x < p.total_tiles) prefetch(t + (int)gridDim.x);
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    float v[32];
    {
      uint32_t r0[32];
      tmem_ld_32x32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(part * 32), r0);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = fmaf(__uint_as_float(r0[e]), p.out_scale, bias_s[part * 32 + e]);
    }
    tc_fence_before();
    __syncthreads();                               // accumulator, operand tile and halo may be overwritten by the next tile
    if (p.relu) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    const int y = y0 + (m >> 3), x = x0 + (m & 7);
    if (p.guard && y < p.HO && x < p.WO) {
#pragma unroll
      for (int e = 0; e < 32; ++e) gmax = fmaxf(gmax, fabsf(v[e]));
    }
    uint8_t* stg = stage_s + warp * 2048;
    const int qy = y0 + quad * 4;
    __half* base = p.out + (((size_t)img * p.HO + qy) * p.WO + x0) * (size_t)64;
    const uint32_t pitch = (uint32_t)p.WO * 64u;
    auto dst = [&](int row) -> __half* {
      const int dy = row >> 3, dx = row & 7;
      return (qy + dy < p.HO && x0 + dx < p.WO) ? base + ((uint32_t)dy * pitch + (uint32_t)dx * 64u) : nullptr;
    };
    store_plane<32>(stg, lane, v, 0, p.out_fmt, part * 32, (size_t)p.plane_elems, dst);
    store_plane<32>(stg, lane, v, 1, p.out_fmt, part * 32, (size_t)p.plane_elems, dst);
  }
  range_guard_commit(p.guard, gmax);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, 64);
  }
}

template <int KS, int ST, int MINB>
int launch_first_tc(FcParams& p, cudaStream_t stream) {
  using C = Fc<KS, ST>;
  p.HO = (p.H + 2 * p.pad - KS) / ST + 1;              // conv_layer.cpp:8-28
  p.WO = (p.W + 2 * p.pad - KS) / ST + 1;
  p.plane_elems = (long long)p.N * p.HO * p.WO * 64;
  p.tiles_x = (p.WO + kTW - 1) / kTW;
  p.tiles_y = (p.HO + kTH - 1) / kTH;
  p.total_tiles = p.tiles_x * p.tiles_y * p.N;
  const int smem_bytes = 1024 + C::kSmem;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  static bool attr[64] = {};                     // function attributes are per device
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    SHF_CUDA_CHECK(cudaFuncSetAttribute(conv_first_tc_kernel<KS, ST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int cap = sms * MINB;
  const int grid = p.total_tiles < cap ? p.total_tiles : cap;
  conv_first_tc_kernel<KS, ST, MINB><<<grid, kThreadsFc, smem_bytes, stream>>>(p);
  SHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// C ABI -- see include/shf_b200.h
extern "C" int shf_conv_first_tc(const float* in_nchw, const void* w_packed, const float* bias, void* out_act, int batch, int H,
                                 int W, int cout, int ksize, int stride, int pad, float out_scale, int relu, int out_format,
                                 unsigned int* range_guard, void* stream) {
  SHF_REQUIRE(cout == 64, "shf_conv_first_tc: Cout=%d (64 supported)", cout);
  SHF_REQUIRE((ksize == 7 && stride == 2) || (ksize == 3 && stride == 1),
              "shf_conv_first_tc: kernel %d stride %d (7x7/2 and 3x3/1 have tensor-core instantiations; use shf_conv_first)", ksize, stride);
  SHF_REQUIRE(out_format == SHF_FMT_H2 || out_format == SHF_FMT_HF8, "shf_conv_first_tc: unknown activation format %d", out_format);
  SHF_REQUIRE(batch >= 1 && pad >= 0 && pad < ksize && H + 2 * pad >= ksize && W + 2 * pad >= ksize, "shf_conv_first_tc: bad geometry");
  FcParams p;
  p.in = in_nchw;
  p.wpack = reinterpret_cast<const __half*>(w_packed);
  p.bias = bias;
  p.out = reinterpret_cast<__half*>(out_act);
  p.N = batch; p.H = H; p.W = W; p.pad = pad; p.relu = relu; p.out_fmt = out_format;
  p.out_scale = out_scale;
  p.guard = range_guard;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (ksize == 7) return launch_first_tc<7, 2, 1>(p, st);
  return launch_first_tc<3, 1, 3>(p, st);
}

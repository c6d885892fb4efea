// Layers a ResNet-style backbone adds to the VGG16 path (BASELINE north_star: "the ResNet/VGG backbone conv stack"; the
// reference ships no ResNet prototxt, SURVEY F1, so these follow Caffe's layer definitions, not a deployed net):
//   * EltwiseLayer SUM (eltwise_layer.cpp:37-77, coefficients included) + the in-place ReLU that follows a residual add,
//   * PoolingLayer MAX with any kernel / stride / pad (pooling_layer.cpp:79-123 output size, :140-187 clipped windows) --
//     the 3x3 stride-2 pool after conv1,
//   * the first convolution for 3-channel input with any square kernel / stride / pad (conv1 7x7 stride 2 pad 3 of
//     ResNet; base_conv_layer.cpp:255-279 semantics) as an fp32 SIMT kernel.
// Strided 1x1 convolutions (the projection shortcuts) need no kernel of their own: conv_stream.cu runs them over a
// strided TMA view of the input (shf_conv_igemm_strided).
#include "common.cuh"

namespace {

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 32;            // 148 SMs x resident CTAs; grid-stride loops cover the rest
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

constexpr int kMaxEltwise = 4;
struct EltwiseArgs {
  const __half* in[kMaxEltwise];
  float coeff[kMaxEltwise];
  int n_in;
};

// HBM-bound: n_in reads + 1 write of 4 bytes per element.  One thread = 8 consecutive channels of one pixel.
__global__ void __launch_bounds__(256) eltwise_sum_kernel(const EltwiseArgs a, __half* __restrict__ out, long long pixels,
                                                          int C, int CT, int c_off, int relu, int in_fmt, int out_fmt,
                                                          unsigned int* guard) {
  const int cv = C / 8;
  const long long total = pixels * cv;
  const size_t in_plane = (size_t)pixels * C, out_plane = (size_t)pixels * CT;
  float gmax = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    const long long pix = i / cv;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int t = 0; t < a.n_in; ++t) {
      float v[8];
      act_load8(a.in[t] + (size_t)pix * C, in_plane, c8 * 8, v, in_fmt);
      // eltwise_layer.cpp:52-57: top = coeff0 * bottom0, then axpy of the others, in bottom order
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = (t == 0) ? a.coeff[0] * v[j] : fmaf(a.coeff[t], v[j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (relu) acc[j] = fmaxf(acc[j], 0.f);
      gmax = fmaxf(gmax, fabsf(acc[j]));
    }
    act_store8(out + (size_t)pix * CT, out_plane, c_off + c8 * 8, acc, out_fmt);
  }
  range_guard_commit(guard, gmax);
}

__global__ void __launch_bounds__(256) maxpool_h2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int N,
                                                         int H, int W, int C, int KH, int KW, int SH, int SW, int PH,
                                                         int PW, int HO, int WO, int fmt) {
  const int cv = C / 8;
  const long long total = (long long)N * HO * WO * cv;
  const size_t in_plane = (size_t)N * H * W * C, out_plane = (size_t)N * HO * WO * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    const int ox = (int)((i / cv) % WO);
    const int oy = (int)((i / ((long long)cv * WO)) % HO);
    const int n = (int)(i / ((long long)cv * WO * HO));
    // pooling_layer.cpp:152-157: window start, end clipped to the image, start clipped to 0
    int hs = oy * SH - PH, ws = ox * SW - PW;
    const int he = min(hs + KH, H), we = min(ws + KW, W);
    hs = max(hs, 0);
    ws = max(ws, 0);
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -3.402823466e38f;      // -FLT_MAX, pooling_layer.cpp:147
    for (int iy = hs; iy < he; ++iy)
      for (int ix = ws; ix < we; ++ix) {
        float v[8];
        act_load8(in + (((size_t)n * H + iy) * W + ix) * C, in_plane, c8 * 8, v, fmt);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], v[j]);
      }
    act_store8(out + (((size_t)n * HO + oy) * WO + ox) * C, out_plane, c8 * 8, best, fmt);
  }
}

// First convolution, 3 input channels, square K x K kernel, stride S, pad P, COUT = 64 output channels: fp32 FMAs.
// A CTA owns a 16 x 8 tile of OUTPUT pixels: the fp32 input patch ((15 S + K) x (7 S + K) x 3) and the whole weight
// tensor (3 K K x 64 floats, transposed to [tap][cout]) sit in shared memory.
constexpr int kCfTH = 16, kCfTW = 8, kCfThreads = 128, kCfCout = 64;
__global__ void __launch_bounds__(kCfThreads) conv_first_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                const float* __restrict__ bias, __half* __restrict__ out,
                                                                int N, int H, int W, int HO, int WO, int K, int S, int P,
                                                                int relu, int out_fmt, int tiles_x, int tiles_y,
                                                                unsigned int* guard) {
  extern __shared__ float smem_f[];
  const int taps = 3 * K * K;
  const int ph = (kCfTH - 1) * S + K, pw = (kCfTW - 1) * S + K;
  float* w_s = smem_f;                           // [taps][64]
  float* patch = smem_f + taps * kCfCout;        // [3][ph][pw]
  for (int i = threadIdx.x; i < taps * kCfCout; i += kCfThreads) {
    const int o = i % kCfCout, t = i / kCfCout;  // w is OIHW: (o, c, r, s) -> tap index t = (c * K + r) * K + s
    w_s[i] = __ldg(w + (size_t)o * taps + t);
  }
  // thread = two horizontally adjacent output pixels x one half of the output channels (the half is warp-uniform, so the
  // weight loads stay broadcasts): per tap 2 + 8 shared-memory loads feed 64 FFMAs
  const int pair = threadIdx.x & 63, half = threadIdx.x >> 6;
  const int ly = pair >> 2, lx = (pair & 3) * 2;
  constexpr int kHalf = kCfCout / 2;
  const size_t out_plane = (size_t)N * HO * WO * kCfCout;
  float gmax = 0.f;
  const int total_tiles = tiles_x * tiles_y * N;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    int q = t;
    const int x0 = (q % tiles_x) * kCfTW;
    q /= tiles_x;
    const int y0 = (q % tiles_y) * kCfTH;
    const int img = q / tiles_y;
    __syncthreads();                             // previous tile's readers are done with the patch (and w_s is staged)
    const int iy0 = y0 * S - P, ix0 = x0 * S - P;
    for (int i = threadIdx.x; i < 3 * ph * pw; i += kCfThreads) {
      const int c = i / (ph * pw), rem = i % (ph * pw), r = rem / pw, s = rem % pw;
      const int iy = iy0 + r, ix = ix0 + s;
      patch[i] = ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
                     ? __ldg(in + (((size_t)img * 3 + c) * H + iy) * W + ix) : 0.f;      // zero padding
    }
    __syncthreads();
    float acc0[kHalf], acc1[kHalf];
#pragma unroll
    for (int o = 0; o < kHalf; ++o) { acc0[o] = 0.f; acc1[o] = 0.f; }
    const float* mine = patch + (ly * S) * pw + lx * S;
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < K; ++r)
        for (int s = 0; s < K; ++s) {
          const float xa = mine[(c * ph + r) * pw + s], xb = mine[(c * ph + r) * pw + s + S];
          const float4* wt = reinterpret_cast<const float4*>(w_s + ((c * K + r) * K + s) * kCfCout + half * kHalf);
#pragma unroll
          for (int o4 = 0; o4 < kHalf / 4; ++o4) {
            const float4 w4 = wt[o4];
            acc0[4 * o4 + 0] = fmaf(xa, w4.x, acc0[4 * o4 + 0]);
            acc0[4 * o4 + 1] = fmaf(xa, w4.y, acc0[4 * o4 + 1]);
            acc0[4 * o4 + 2] = fmaf(xa, w4.z, acc0[4 * o4 + 2]);
            acc0[4 * o4 + 3] = fmaf(xa, w4.w, acc0[4 * o4 + 3]);
            acc1[4 * o4 + 0] = fmaf(xb, w4.x, acc1[4 * o4 + 0]);
            acc1[4 * o4 + 1] = fmaf(xb, w4.y, acc1[4 * o4 + 1]);
            acc1[4 * o4 + 2] = fmaf(xb, w4.z, acc1[4 * o4 + 2]);
            acc1[4 * o4 + 3] = fmaf(xb, w4.w, acc1[4 * o4 + 3]);
          }
        }
    const int oy = y0 + ly;
#pragma unroll
    for (int pxi = 0; pxi < 2; ++pxi) {
      const int ox = x0 + lx + pxi;
      if (oy < HO && ox < WO) {
        __half* px = out + (((size_t)img * HO + oy) * WO + ox) * kCfCout;
#pragma unroll
        for (int o8 = 0; o8 < kHalf / 8; ++o8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float y = (pxi ? acc1[8 * o8 + j] : acc0[8 * o8 + j]) + (bias ? __ldg(bias + half * kHalf + 8 * o8 + j) : 0.f);
            if (relu) y = fmaxf(y, 0.f);
            gmax = fmaxf(gmax, fabsf(y));
            v[j] = y;
          }
          // hf8 rows are 64-channel blocks [64 x al8 | 64 x ah8]: 8-channel groups of either half land in the same block
          act_store8(px, out_plane, half * kHalf + 8 * o8, v, out_fmt);
        }
      }
    }
  }
  range_guard_commit(guard, gmax);
}

}  // namespace

// ---- C ABI (include/shf_b200.h) ---------------------------------------------------------------------------------
extern "C" int shf_eltwise_sum(const void* const* ins, const float* coeffs, int n_in, void* out, long long pixels, int C,
                               int out_channels_total, int out_channel_offset, int relu, int in_format, int out_format,
                               unsigned int* range_guard, void* stream) {
  SHF_REQUIRE(n_in >= 1 && n_in <= kMaxEltwise, "shf_eltwise_sum: %d bottoms (1..%d supported)", n_in, kMaxEltwise);
  SHF_REQUIRE(C % 8 == 0 && out_channel_offset % 8 == 0 && out_channels_total % 8 == 0 &&
                  out_channel_offset + C <= out_channels_total,
              "shf_eltwise_sum: channel counts must be multiples of 8");
  SHF_REQUIRE((in_format == SHF_FMT_H2 || (in_format == SHF_FMT_HF8 && C % 64 == 0)) &&
                  (out_format == SHF_FMT_H2 ||
                   (out_format == SHF_FMT_HF8 && out_channel_offset % 64 == 0 && out_channels_total % 64 == 0 && C % 64 == 0)),
              "shf_eltwise_sum: formats %d/%d need 64-aligned channel windows for hf8", in_format, out_format);
  SHF_REQUIRE(pixels >= 0, "shf_eltwise_sum: bad size");
  if (pixels == 0) return 0;
  EltwiseArgs a;
  a.n_in = n_in;
  for (int t = 0; t < kMaxEltwise; ++t) {
    a.in[t] = t < n_in ? reinterpret_cast<const __half*>(ins[t]) : nullptr;
    a.coeff[t] = (t < n_in && coeffs) ? coeffs[t] : 1.f;              // eltwise_layer.cpp:17-22: default coefficient 1
    SHF_REQUIRE(t >= n_in || a.in[t] != nullptr, "shf_eltwise_sum: bottom %d is NULL", t);
  }
  const long long total = pixels * (C / 8);
  eltwise_sum_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      a, reinterpret_cast<__half*>(out), pixels, C, out_channels_total, out_channel_offset, relu, in_format, out_format,
      range_guard);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_pool_out_size(int size, int k, int stride, int pad, int any_pad) {
  // pooling_layer.cpp:91-105: ceil((size + 2 pad - k) / stride) + 1; with padding (in either dimension) the last window
  // must start inside the image
  int o = (size + 2 * pad - k + stride - 1) / stride + 1;
  if (any_pad && (o - 1) * stride >= size + pad) --o;
  return o;
}

extern "C" int shf_maxpool(const void* in_h2, void* out_h2, int batch, int H, int W, int C, int kh, int kw, int sh, int sw,
                           int ph, int pw, int format, void* stream) {
  SHF_REQUIRE(C % 8 == 0, "shf_maxpool: C=%d must be a multiple of 8", C);
  SHF_REQUIRE(format == SHF_FMT_H2 || (format == SHF_FMT_HF8 && C % 64 == 0), "shf_maxpool: format %d with C=%d", format, C);
  SHF_REQUIRE(kh >= 1 && kw >= 1 && sh >= 1 && sw >= 1 && ph >= 0 && pw >= 0 && ph < kh && pw < kw,
              "shf_maxpool: kernel %dx%d stride %dx%d pad %dx%d (pad must be smaller than the kernel, pooling_layer.cpp:73-74)",
              kh, kw, sh, sw, ph, pw);
  SHF_REQUIRE(H + 2 * ph >= kh && W + 2 * pw >= kw, "shf_maxpool: %dx%d input smaller than the %dx%d window", H, W, kh, kw);
  const int any_pad = (ph || pw) ? 1 : 0;
  const int HO = shf_pool_out_size(H, kh, sh, ph, any_pad), WO = shf_pool_out_size(W, kw, sw, pw, any_pad);
  const long long total = (long long)batch * HO * WO * (C / 8);
  maxpool_h2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)in_h2, (__half*)out_h2, batch, H, W, C, kh, kw, sh, sw, ph, pw, HO, WO, format);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_conv_first(const float* in_nchw, const float* w_oihw, const float* bias, void* out_act, int batch, int H,
                              int W, int cout, int ksize, int stride, int pad, int relu, int out_format,
                              unsigned int* range_guard, void* stream) {
  SHF_REQUIRE(cout == kCfCout, "shf_conv_first: Cout=%d (64 supported)", cout);
  SHF_REQUIRE(ksize >= 1 && ksize <= 11 && stride >= 1 && stride <= 4 && pad >= 0 && pad < ksize,
              "shf_conv_first: kernel %d stride %d pad %d", ksize, stride, pad);
  SHF_REQUIRE(out_format == SHF_FMT_H2 || out_format == SHF_FMT_HF8, "shf_conv_first: unknown activation format %d", out_format);
  SHF_REQUIRE(batch >= 1 && H + 2 * pad >= ksize && W + 2 * pad >= ksize, "shf_conv_first: bad geometry");
  const int HO = (H + 2 * pad - ksize) / stride + 1, WO = (W + 2 * pad - ksize) / stride + 1;   // conv_layer.cpp:8-28
  const int ph = (kCfTH - 1) * stride + ksize, pw = (kCfTW - 1) * stride + ksize;
  const int smem = (3 * ksize * ksize * kCfCout + 3 * ph * pw) * (int)sizeof(float);
  SHF_REQUIRE(smem <= 200 * 1024, "shf_conv_first: %d bytes of shared memory", smem);
  static bool attr[64] = {};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    SHF_CUDA_CHECK(cudaFuncSetAttribute(conv_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_x = (WO + kCfTW - 1) / kCfTW, tiles_y = (HO + kCfTH - 1) / kCfTH;
  const long long total = (long long)tiles_x * tiles_y * batch;
  const int per_sm = smem > 0 ? (220 * 1024) / (smem + 1024) : 1;
  const long long cap = (long long)sms * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  conv_first_kernel<<<(int)(total < cap ? total : cap), kCfThreads, smem, (cudaStream_t)stream>>>(
      in_nchw, w_oihw, bias, reinterpret_cast<__half*>(out_act), batch, H, W, HO, WO, ksize, stride, pad, relu, out_format,
      tiles_x, tiles_y, range_guard);
  SHF_LAUNCH_CHECK();
  return 0;
}

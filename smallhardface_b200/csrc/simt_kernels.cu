// HBM-bound layers of the deploy nets as plain SIMT kernels (sm_100a): the 3-channel first conv,
// 2x2 max pooling, the depthwise 2x bilinear deconvolution, layout/precision converters, the image
// pyramid pre-processing, and a slow direct convolution used only to validate the tcgen05 kernel.
// Activations are NHWC in one of the two 4-byte formats of common.cuh (h2 = split fp16, hf8 = fp16 + 8-bit floats).
#include "common.cuh"

namespace {

SHF_DEVICE void unpack8(const uint4& v, __half (&h)[8]) {
  const __half2* p = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[2 * i] = __low2half(p[i]);
    h[2 * i + 1] = __high2half(p[i]);
  }
}
SHF_DEVICE uint4 pack8(const __half (&h)[8]) {
  uint4 v;
  __half2* p = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __halves2half2(h[2 * i], h[2 * i + 1]);
  return v;
}

// ---------------------------------------------------------------------------------------------------
// conv1_1: fp32 NCHW (N,3,H,W) -> h2 NHWC (N,H,W,64), 3x3 pad 1, + bias + ReLU.
// Replaces base_conv_layer.cpp:255-279 for the K=27 first layer (HBM-bound: AI 12.9 FLOP/B,
// SURVEY 8d) -- not worth a tensor-core tile; one thread per output pixel, 64 fp32 accumulators,
// weights broadcast from shared memory, 128-byte coalesced stores per plane.
// ---------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(128) conv3x3_c3_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                         const float* __restrict__ bias, __half* __restrict__ out,
                                                         int N, int H, int W, int relu, int out_fmt) {
  // Work split chosen for the STORE side (this layer writes 256 B per pixel and reads 12): four lanes share a run of
  // four horizontally adjacent pixels, each lane owning 16 of the 64 output channels.  A quad therefore writes each
  // pixel's 128-byte hi row (and lo row) as one full cache line, and one warp-wide store instruction covers 8 lines.
  // Register blocking 4 pixels x 16 channels: one LDS.128 of weights feeds 16 FMAs.
  static_assert(COUT == 64, "lane -> channel mapping assumes 64 output channels");
  __shared__ float4 ws[27][COUT / 4];     // [c*9 + r*3 + s][o/4]
  __shared__ float bs[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
    const int o = i / 27, k = i % 27;            // weights arrive OIHW: w[o][c][r][s]
    reinterpret_cast<float*>(&ws[k][0])[o] = w[i];
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int part = threadIdx.x & 3;              // channels [16*part, 16*part + 16)
  const int WG = (W + 3) / 4;                    // 4-pixel groups per row
  const long long ngroups = (long long)N * H * WG;
  const size_t plane = (size_t)N * H * W * COUT;
  for (long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2; g < ngroups;
       g += ((long long)gridDim.x * blockDim.x) >> 2) {
    const int xg = (int)(g % WG);
    const int y = (int)((g / WG) % H);
    const int n = (int)(g / ((long long)WG * H));
    const int x = xg * 4;
    float v[3][3][6];                            // [c][r][columns x-1 .. x+4]
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = y + r - 1;
        const float* row = in + (((size_t)n * 3 + c) * H + iy) * W;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int ix = x + q - 1;
          v[c][r][q] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(row + ix) : 0.f;
        }
      }
    float acc[4][16];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[px][j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 wv = ws[c * 9 + r * 3 + t][part * 4 + q];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              const float pv = v[c][r][t + px];
              acc[px][4 * q + 0] = fmaf(pv, wv.x, acc[px][4 * q + 0]);
              acc[px][4 * q + 1] = fmaf(pv, wv.y, acc[px][4 * q + 1]);
              acc[px][4 * q + 2] = fmaf(pv, wv.z, acc[px][4 * q + 2]);
              acc[px][4 * q + 3] = fmaf(pv, wv.w, acc[px][4 * q + 3]);
            }
          }
    const size_t pix0 = ((size_t)n * H + y) * W + x;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      if (x + px >= W) break;
      __half* opx = out + (pix0 + px) * COUT;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          t[j] = acc[px][h * 8 + j] + bs[part * 16 + h * 8 + j];
          if (relu) t[j] = fmaxf(t[j], 0.f);
        }
        act_store8(opx, plane, part * 16 + h * 8, t, out_fmt);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// 2x2 / stride-2 max pooling on h2 NHWC (pooling_layer.cpp:140-187; windows clipped to the image,
// ceil-mode output size).  One thread = one output pixel x 8 channels (16-byte vectors per plane).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2x2_h2_kernel(const __half* __restrict__ in, __half* __restrict__ out,
                                                            int N, int H, int W, int C, int HO, int WO, int fmt) {
  const int cv = C / 8;
  const long long total = (long long)N * HO * WO * cv;
  const size_t in_plane = (size_t)N * H * W * C, out_plane = (size_t)N * HO * WO * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    const int ox = (int)((i / cv) % WO);
    const int oy = (int)((i / ((long long)cv * WO)) % HO);
    const int n = (int)(i / ((long long)cv * WO * HO));
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -3.402823466e38f;
    for (int dy = 0; dy < 2; ++dy) {
      const int iy = oy * 2 + dy;
      if (iy >= H) continue;
      for (int dx = 0; dx < 2; ++dx) {
        const int ix = ox * 2 + dx;
        if (ix >= W) continue;
        float v[8];
        act_load8(in + (((size_t)n * H + iy) * W + ix) * C, in_plane, c8 * 8, v, fmt);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], v[j]);      // re-splitting the maximum reproduces its planes
      }
    }
    act_store8(out + (((size_t)n * HO + oy) * WO + ox) * C, out_plane, c8 * 8, best, fmt);
  }
}

// ---------------------------------------------------------------------------------------------------
// Depthwise transposed convolution (group == C, one filter per channel), e.g. conv5_256_up: k4 s2 p1.
// Replaces deconv_layer.cpp:30-46 -> base_conv_layer.cpp:281-297 (256 K=1 GEMMs) + col2im
// (im2col.cpp:163-197).  Writes into channels [c_off, c_off+C) of a wider NHWC tensor (the concat).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) deconv_dw_h2_kernel(const __half* __restrict__ in, const float* __restrict__ w,
                                                           __half* __restrict__ out, int N, int H, int W, int C,
                                                           int K, int S, int P, int HO, int WO, int CT, int c_off,
                                                           int in_fmt, int out_fmt, unsigned int* guard) {
  const int cv = C / 8;
  const long long total = (long long)N * HO * WO * cv;
  const size_t in_plane = (size_t)N * H * W * C, out_plane = (size_t)N * HO * WO * CT;
  float gmax = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    const int ox = (int)((i / cv) % WO);
    const int oy = (int)((i / ((long long)cv * WO)) % HO);
    const int n = (int)(i / ((long long)cv * WO * HO));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < K; ++ky) {                 // same (ky, kx) accumulation order as col2im
      const int ty = oy + P - ky;
      if (ty < 0 || ty % S) continue;
      const int iy = ty / S;
      if (iy >= H) continue;
      for (int kx = 0; kx < K; ++kx) {
        const int tx = ox + P - kx;
        if (tx < 0 || tx % S) continue;
        const int ix = tx / S;
        if (ix >= W) continue;
        float v[8];
        act_load8(in + (((size_t)n * H + iy) * W + ix) * C, in_plane, c8 * 8, v, in_fmt);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j] * __ldg(w + ((size_t)(c8 * 8 + j) * K + ky) * K + kx);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) gmax = fmaxf(gmax, fabsf(acc[j]));
    act_store8(out + (((size_t)n * HO + oy) * WO + ox) * CT, out_plane, c_off + c8 * 8, acc, out_fmt);
  }
  range_guard_commit(guard, gmax);
}

// Fast path for the net's own upsampler (k = 4, stride 2, pad 1): every output pixel has exactly 2 x 2 taps,
// ky = ((oy+1)&1) + {0,2} reading input rows (oy+1-ky)/2.  Per-channel 4x4 filters sit in shared memory; the tap
// loops are fully unrolled so the 8 operand loads of a thread are all in flight together.  Accumulation order
// (ky ascending, then kx ascending) is the same as the generic kernel / col2im.
__global__ void __launch_bounds__(256) deconv_k4s2p1_h2_kernel(const __half* __restrict__ in, const float* __restrict__ w,
                                                               __half* __restrict__ out, int N, int H, int W, int C,
                                                               int CT, int c_off, int in_fmt, int out_fmt,
                                                               unsigned int* guard) {
  extern __shared__ float wsm[];                 // [16 taps][C]: a warp's lanes (consecutive 8-channel groups) read
                                                 // consecutive 32-byte runs (the [C][16] layout was a 32-way bank conflict)
  for (int i = threadIdx.x; i < C * 16; i += blockDim.x) wsm[(i & 15) * C + (i >> 4)] = w[i];
  __syncthreads();
  const int HO = 2 * H, WO = 2 * W, cv = C / 8;
  const long long total = (long long)N * HO * WO * cv;
  const size_t in_plane = (size_t)N * H * W * C, out_plane = (size_t)N * HO * WO * CT;
  float gmax = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    const int ox = (int)((i / cv) % WO);
    const int oy = (int)((i / ((long long)cv * WO)) % HO);
    const int n = (int)(i / ((long long)cv * WO * HO));
    const int ky0 = (oy + 1) & 1, kx0 = (ox + 1) & 1;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ky = ky0 + 2 * a;
      const int iy = (oy + 1 - ky) >> 1;          // exact: oy+1-ky is even
      if (oy + 1 - ky < 0 || iy >= H) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int kx = kx0 + 2 * b;
        const int ix = (ox + 1 - kx) >> 1;
        if (ox + 1 - kx < 0 || ix >= W) continue;
        float v[8];
        act_load8(in + (((size_t)n * H + iy) * W + ix) * C, in_plane, c8 * 8, v, in_fmt);
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + (ky * 4 + kx) * C + c8 * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wsm + (ky * 4 + kx) * C + c8 * 8 + 4);
        acc[0] += v[0] * w0.x; acc[1] += v[1] * w0.y; acc[2] += v[2] * w0.z; acc[3] += v[3] * w0.w;
        acc[4] += v[4] * w1.x; acc[5] += v[5] * w1.y; acc[6] += v[6] * w1.z; acc[7] += v[7] * w1.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) gmax = fmaxf(gmax, fabsf(acc[j]));
    act_store8(out + (((size_t)n * HO + oy) * WO + ox) * CT, out_plane, c_off + c8 * 8, acc, out_fmt);
  }
  range_guard_commit(guard, gmax);
}

// ---------------------------------------------------------------------------------------------------
// Layout / precision converters (the Blob.data boundary: Caffe blobs are fp32 NCHW)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) h2_to_nchw_kernel(const __half* __restrict__ in, float* __restrict__ out, int N,
                                                         int H, int W, int CT, int c_off, int C, int fmt) {
  // one thread per (n, y, x, c): reads are channel-contiguous; a 32x32 smem transpose would make both
  // sides coalesced, but this only runs when a host reads an intermediate blob (debug / parity tests)
  const long long total = (long long)N * H * W * C;
  const size_t plane = (size_t)N * H * W * CT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pix = i / C;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    out[(((size_t)n * C + c) * H + y) * W + x] = act_load1(in + (size_t)pix * CT, plane, c_off + c, fmt);
  }
}

__global__ void __launch_bounds__(256) nchw_to_h2_kernel(const float* __restrict__ in, __half* __restrict__ out, int N,
                                                         int C, int H, int W, int CT, int c_off, int fmt) {
  const long long total = (long long)N * H * W * C;
  const size_t plane = (size_t)N * H * W * CT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pix = i / C;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    act_store1(out + (size_t)pix * CT, plane, c_off + c, in[(((size_t)n * C + c) * H + y) * W + x], fmt);
  }
}

// ---------------------------------------------------------------------------------------------------
// Pyramid level pre-processing: uint8 HWC BGR image -> mean-subtracted, bilinearly resized, optionally
// mirrored, zero-padded fp32 NCHW level blob.  Fuses lib/utils/test_utils.py:29-46 (mean subtract, then
// cv2.resize(fx, fy, INTER_LINEAR) on the float64 image), blob.py:16-32 (HWC->NCHW, float32),
// lib/test.py:147-155 (flip = mirror of the UNPADDED blob) and lib/test.py:30-38 (pad to x16).
// Arithmetic follows oracle/preprocess.py:resize_linear: double coefficients, lerp S0+(S1-S0)*a with
// fused multiply-add, horizontal then vertical (what opencv 4.13 computes for CV_64F).
// ---------------------------------------------------------------------------------------------------
SHF_DEVICE void preprocess_level_body(const uint8_t* __restrict__ img, int h, int w, float* __restrict__ out, int oh,
                                      int ow, int HP, int WP, double inv_scale, int identity, int flip, double m0,
                                      double m1, double m2) {
  const long long total = (long long)HP * WP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % WP), y = (int)(i / WP);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    if (x < ow && y < oh) {
      const int xs = flip ? (ow - 1 - x) : x;           // mirror the unpadded level
      const double mean[3] = {m0, m1, m2};
      double res[3];
      if (identity) {
        const uint8_t* px = img + ((size_t)y * w + xs) * 3;
        for (int c = 0; c < 3; ++c) res[c] = (double)(float)px[c] - mean[c];
      } else {
        double fx = __dadd_rn(__dmul_rn((double)xs + 0.5, inv_scale), -0.5);
        double fy = __dadd_rn(__dmul_rn((double)y + 0.5, inv_scale), -0.5);
        int sx = (int)floor(fx), sy = (int)floor(fy);
        double ax = fx - (double)sx, ay = fy - (double)sy;
        if (sx < 0) { sx = 0; ax = 0.0; }
        if (sx >= w - 1) { sx = w - 1; ax = 0.0; }
        if (sy < 0) { sy = 0; ay = 0.0; }
        if (sy >= h - 1) { sy = h - 1; ay = 0.0; }
        const int sx1 = min(sx + 1, w - 1), sy1 = min(sy + 1, h - 1);
        const uint8_t* p00 = img + ((size_t)sy * w + sx) * 3;
        const uint8_t* p01 = img + ((size_t)sy * w + sx1) * 3;
        const uint8_t* p10 = img + ((size_t)sy1 * w + sx) * 3;
        const uint8_t* p11 = img + ((size_t)sy1 * w + sx1) * 3;
        for (int c = 0; c < 3; ++c) {
          const double a = (double)(float)p00[c] - mean[c], b = (double)(float)p01[c] - mean[c];
          const double d = (double)(float)p10[c] - mean[c], e = (double)(float)p11[c] - mean[c];
          const double t0 = fma(b - a, ax, a);
          const double t1 = fma(e - d, ax, d);
          res[c] = fma(t1 - t0, ay, t0);
        }
      }
      r0 = (float)res[0]; r1 = (float)res[1]; r2 = (float)res[2];
    }
    out[(size_t)0 * HP * WP + i] = r0;
    out[(size_t)1 * HP * WP + i] = r1;
    out[(size_t)2 * HP * WP + i] = r2;
  }
}

__global__ void __launch_bounds__(256) preprocess_level_kernel(const uint8_t* __restrict__ img, int h, int w,
                                                               float* __restrict__ out, int oh, int ow, int HP, int WP,
                                                               double inv_scale, int identity, int flip, double m0,
                                                               double m1, double m2) {
  preprocess_level_body(img, h, w, out, oh, ow, HP, WP, inv_scale, identity, flip, m0, m1, m2);
}

// One launch for a whole level batch: blockIdx.y = slot j * passes + f of the (images x passes, 3, HP, WP) blob, image j
// of a stacked (N, h, w, 3) uint8 tensor, pass f = 0 plain / 1 mirrored (lib/test.py:147-155).
__global__ void __launch_bounds__(256) preprocess_level_batched_kernel(const uint8_t* __restrict__ imgs, int h, int w,
                                                                       float* __restrict__ out, int oh, int ow, int HP,
                                                                       int WP, double inv_scale, int identity, int passes,
                                                                       double m0, double m1, double m2) {
  const int slot = blockIdx.y, j = slot / passes, f = slot % passes;
  preprocess_level_body(imgs + (size_t)j * h * w * 3, h, w, out + (size_t)slot * 3 * HP * WP, oh, ow, HP, WP, inv_scale,
                        identity, f, m0, m1, m2);
}

// ---------------------------------------------------------------------------------------------------
// Validation-only direct convolution on h2 NHWC (fp32 FMA, one thread per output element).
// NOT on the product path: tests use it to check the tcgen05 kernel at sizes the CPU oracle is slow at.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_direct_h2_kernel(const __half* __restrict__ in, const float* __restrict__ w,
                                                             const float* __restrict__ bias, __half* __restrict__ out,
                                                             int N, int H, int W, int CI, int CO, int K, int dil,
                                                             int pad, int CT, int c_off, int relu, int in_fmt,
                                                             int out_fmt) {
  const long long total = (long long)N * H * W * CO;
  const size_t in_plane = (size_t)N * H * W * CI, out_plane = (size_t)N * H * W * CT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % CO);
    const long long pix = i / CO;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    float acc = 0.f;
    for (int r = 0; r < K; ++r) {
      const int iy = y - pad + r * dil;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < K; ++s) {
        const int ix = x - pad + s * dil;
        if (ix < 0 || ix >= W) continue;
        const __half* ph = in + (((size_t)n * H + iy) * W + ix) * CI;
        const float* pw = w + ((size_t)o * CI * K + r) * K + s;          // OIHW
        for (int c = 0; c < CI; ++c) acc = fmaf(act_load1(ph, in_plane, c, in_fmt), pw[(size_t)c * K * K], acc);
      }
    }
    acc += bias ? bias[o] : 0.f;
    if (relu) acc = fmaxf(acc, 0.f);
    act_store1(out + (size_t)pix * CT, out_plane, c_off + o, acc, out_fmt);
  }
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 32;            // 148 SMs x resident CTAs; grid-stride loops cover the rest
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int shf_conv1_c3(const float* in_nchw, const float* w_oihw, const float* bias, void* out_h2, int batch,
                            int H, int W, int cout, int relu, int out_format, void* stream) {
  SHF_REQUIRE(out_format == SHF_FMT_H2 || out_format == SHF_FMT_HF8, "shf_conv1_c3: unknown activation format %d", out_format);
  SHF_REQUIRE(cout == 64, "shf_conv1_c3: Cout=%d (the deploy nets' conv1_1 has 64)", cout);
  const long long npix = (long long)batch * H * ((W + 3) / 4) * 4;  // four threads per run of four pixels
  conv3x3_c3_kernel<64><<<grid_for(npix, 128), 128, 0, (cudaStream_t)stream>>>(in_nchw, w_oihw, bias, (__half*)out_h2,
                                                                               batch, H, W, relu, out_format);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_maxpool2x2(const void* in_h2, void* out_h2, int batch, int H, int W, int C, int format, void* stream) {
  SHF_REQUIRE(C % 8 == 0, "shf_maxpool2x2: C=%d must be a multiple of 8", C);
  SHF_REQUIRE(format == SHF_FMT_H2 || (format == SHF_FMT_HF8 && C % 64 == 0), "shf_maxpool2x2: format %d with C=%d", format, C);
  const int HO = (H + 1) / 2, WO = (W + 1) / 2;       // ceil mode, pooling_layer.cpp:91-94 with k=s=2, pad 0
  const long long total = (long long)batch * HO * WO * (C / 8);
  maxpool2x2_h2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)in_h2, (__half*)out_h2,
                                                                             batch, H, W, C, HO, WO, format);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_deconv_depthwise(const void* in_h2, const float* w, void* out_h2, int batch, int H, int W, int C,
                                    int ksize, int stride, int pad, int out_channels_total, int out_channel_offset,
                                    int in_format, int out_format, unsigned int* range_guard, void* stream) {
  SHF_REQUIRE(C % 8 == 0 && out_channel_offset % 8 == 0 && out_channels_total % 8 == 0,
              "shf_deconv_depthwise: channel counts must be multiples of 8");
  SHF_REQUIRE((in_format == SHF_FMT_H2 || (in_format == SHF_FMT_HF8 && C % 64 == 0)) &&
                  (out_format == SHF_FMT_H2 ||
                   (out_format == SHF_FMT_HF8 && out_channel_offset % 64 == 0 && out_channels_total % 64 == 0 && C % 64 == 0)),
              "shf_deconv_depthwise: formats %d/%d need 64-aligned channel windows for hf8", in_format, out_format);
  const int HO = stride * (H - 1) + ksize - 2 * pad, WO = stride * (W - 1) + ksize - 2 * pad;   // deconv_layer.cpp:8-28
  const long long total = (long long)batch * HO * WO * (C / 8);
  if (ksize == 4 && stride == 2 && pad == 1 && C * 16 * sizeof(float) <= 48 * 1024) {
    deconv_k4s2p1_h2_kernel<<<grid_for(total, 256), 256, C * 16 * sizeof(float), (cudaStream_t)stream>>>(
        (const __half*)in_h2, w, (__half*)out_h2, batch, H, W, C, out_channels_total, out_channel_offset, in_format,
        out_format, range_guard);
    SHF_LAUNCH_CHECK();
    return 0;
  }
  deconv_dw_h2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)in_h2, w, (__half*)out_h2, batch, H, W, C, ksize, stride, pad, HO, WO, out_channels_total,
      out_channel_offset, in_format, out_format, range_guard);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_h2_to_nchw(const void* in_h2, float* out_nchw, int batch, int H, int W, int channels_total,
                              int channel_offset, int channels, int format, void* stream) {
  SHF_REQUIRE(format == SHF_FMT_H2 || (format == SHF_FMT_HF8 && channels_total % 64 == 0), "shf_h2_to_nchw: format %d", format);
  const long long total = (long long)batch * H * W * channels;
  h2_to_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)in_h2, out_nchw, batch, H, W,
                                                                          channels_total, channel_offset, channels, format);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_nchw_to_h2(const float* in_nchw, void* out_h2, int batch, int channels, int H, int W,
                              int channels_total, int channel_offset, int format, void* stream) {
  SHF_REQUIRE(format == SHF_FMT_H2 || (format == SHF_FMT_HF8 && channels_total % 64 == 0), "shf_nchw_to_h2: format %d", format);
  const long long total = (long long)batch * H * W * channels;
  nchw_to_h2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in_nchw, (__half*)out_h2, batch, channels, H,
                                                                          W, channels_total, channel_offset, format);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_preprocess_level(const uint8_t* img_hwc, int h, int w, float* out_chw, int out_h, int out_w,
                                    int padded_h, int padded_w, double scale, int flip, const double* means,
                                    void* stream) {
  SHF_REQUIRE(out_h <= padded_h && out_w <= padded_w && h > 0 && w > 0, "shf_preprocess_level: bad geometry");
  const long long total = (long long)padded_h * padded_w;
  preprocess_level_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      img_hwc, h, w, out_chw, out_h, out_w, padded_h, padded_w, 1.0 / scale, scale == 1.0 ? 1 : 0, flip, means[0],
      means[1], means[2]);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_preprocess_level_batched(const uint8_t* imgs_nhwc, int num_images, int h, int w, float* out_nchw,
                                            int out_h, int out_w, int padded_h, int padded_w, double scale, int passes,
                                            const double* means, void* stream) {
  SHF_REQUIRE(out_h <= padded_h && out_w <= padded_w && h > 0 && w > 0 && num_images >= 1 && (passes == 1 || passes == 2),
              "shf_preprocess_level_batched: bad geometry");
  const long long total = (long long)padded_h * padded_w;
  long long gx = (total + 255) / 256;
  if (gx > 148 * 8) gx = 148 * 8;
  dim3 grid((unsigned)gx, (unsigned)(num_images * passes));
  preprocess_level_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(imgs_nhwc, h, w, out_nchw, out_h, out_w, padded_h,
                                                                         padded_w, 1.0 / scale, scale == 1.0 ? 1 : 0, passes,
                                                                         means[0], means[1], means[2]);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_debug_conv_direct(const void* in_h2, const float* w_oihw, const float* bias, void* out_h2, int batch,
                                     int H, int W, int cin, int cout, int ksize, int dilation, int pad,
                                     int out_channels_total, int out_channel_offset, int relu, int in_format,
                                     int out_format, void* stream) {
  const long long total = (long long)batch * H * W * cout;
  conv_direct_h2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)in_h2, w_oihw, bias, (__half*)out_h2, batch, H, W, cin, cout, ksize, dilation, pad,
      out_channels_total, out_channel_offset, relu, in_format, out_format);
  SHF_LAUNCH_CHECK();
  return 0;
}

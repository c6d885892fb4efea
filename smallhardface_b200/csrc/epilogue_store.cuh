// Coalesced NHWC row stores shared by the tcgen05 conv kernels (conv_stream.cu, conv1_tc.cu).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------------
// Epilogue store: a warp holds 32 pixels (lane = pixel) x KC channels (registers).  Written straight from registers
// every lane would store 16-byte (h2) or 8-byte (hf8) pieces of ITS OWN pixel row: 32 partial sectors per
// instruction, which made the LSU/L2 request rate the limiter of the short-K layers (tools/probe_epi.py).  Instead the
// warp stages its rows -- exactly the bytes of each pixel's slice of one plane -- in 4 KB of shared memory (16-byte
// chunks XOR-swizzled so both directions are conflict free) and copies them out with consecutive lanes on
// consecutive 16 bytes of a row: full 32-byte sectors, 64 or 128 contiguous bytes per pixel.
// `rows`: 32 (lane r = row r) or 8 (the fused-pool writers: row wr lives in lane ((wr >> 2) << 4) | ((wr & 3) << 1)).
// ---------------------------------------------------------------------------------------------------------------
template <int KC>
struct RowStore {
  static constexpr int kChunks = KC / 8;               // 16-byte chunks per row and plane (row = 2 * KC bytes)
  static constexpr int kRowBytes = KC * 2;
  SHF_DEVICE static int swz(int row, int k) { return KC == 64 ? (k ^ (row & 7)) : (k ^ ((row >> 1) & 3)); }
  SHF_DEVICE static void put(uint8_t* stg, int row, int k, const uint4& v) {
    *reinterpret_cast<uint4*>(stg + row * kRowBytes + swz(row, k) * 16) = v;
  }
  SHF_DEVICE static uint4 get(const uint8_t* stg, int row, int k) {
    return *reinterpret_cast<const uint4*>(stg + row * kRowBytes + swz(row, k) * 16);
  }
};

// Stage one plane of `v` (KC final values of this lane's pixel) and copy it out.  dst_px(row) must return the plane-0
// address of channel 0 of that row's pixel, or nullptr when the row is outside the image / not a writer.
template <int KC, typename DstFn>
SHF_DEVICE void store_plane(uint8_t* stg, int lane, bool lane_writes, int lane_row, int nrows, const float (&v)[KC], int plane,
                            int fmt, int c_first, size_t plane_elems, DstFn dst_px) {
  using RS = RowStore<KC>;
  if (lane_writes) {
    if (plane == 0 || fmt == SHF_FMT_H2) {
#pragma unroll
      for (int k = 0; k < KC / 8; ++k) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = v[8 * k + 2 * e], b = v[8 * k + 2 * e + 1];
          __half2 h = __floats2half2_rn(a, b);
          if (plane == 1) {                                  // h2 lo plane: rn16(x - hi)
            const float2 hf = __half22float2(h);
            h = __floats2half2_rn(a - hf.x, b - hf.y);
          }
          w[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        RS::put(stg, lane_row, k, make_uint4(w[0], w[1], w[2], w[3]));
      }
    } else {                                                 // hf8 plane 1: [KC x e4m3((x - hi) * 2^6) | KC x e4m3(hi * 2^-5)]
#pragma unroll
      for (int k = 0; k < KC / 16; ++k) {
        uint32_t wa[4], wb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint32_t a4 = 0u, b4 = 0u;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float a = v[16 * k + 4 * e + 2 * q], b = v[16 * k + 4 * e + 2 * q + 1];
            const float2 hf = __half22float2(__floats2half2_rn(a, b));
            const float2 lo = make_float2((a - hf.x) * kHf8AlScale, (b - hf.y) * kHf8AlScale);
            const float2 hs = make_float2(hf.x * kHf8AhScale, hf.y * kHf8AhScale);
            a4 |= (uint32_t)__nv_cvt_float2_to_fp8x2(lo, __NV_SATFINITE, __NV_E4M3) << (16 * q);
            b4 |= (uint32_t)__nv_cvt_float2_to_fp8x2(hs, __NV_SATFINITE, __NV_E4M3) << (16 * q);
          }
          wa[e] = a4;
          wb[e] = b4;
        }
        RS::put(stg, lane_row, k, make_uint4(wa[0], wa[1], wa[2], wa[3]));
        RS::put(stg, lane_row, KC / 16 + k, make_uint4(wb[0], wb[1], wb[2], wb[3]));
      }
    }
  }
  __syncwarp();
  const int items = nrows * RS::kChunks;
  for (int i = lane; i < items; i += 32) {
    const int row = i / RS::kChunks, k = i % RS::kChunks;
    __half* px = dst_px(row);
    if (px == nullptr) continue;
    const int srow = (nrows == 32) ? row : (((row >> 2) << 4) | ((row & 3) << 1));
    const uint4 val = RS::get(stg, srow, k);
    if (plane == 0) {
      *reinterpret_cast<uint4*>(px + c_first + 8 * k) = val;
    } else if (fmt == SHF_FMT_H2) {
      *reinterpret_cast<uint4*>(px + plane_elems + c_first + 8 * k) = val;
    } else {
      // chunks [0, KC/16) are al8 of channels c_first + 16k .., chunks [KC/16, KC/8) their ah8 twins 64 bytes further
      const int half = k / (KC / 16), kk = k % (KC / 16);
      uint8_t* p1 = reinterpret_cast<uint8_t*>(px + plane_elems) + hf8_off(c_first) + half * 64 + kk * 16;
      *reinterpret_cast<uint4*>(p1) = val;
    }
  }
  __syncwarp();
}

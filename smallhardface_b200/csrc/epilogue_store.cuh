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
//
// r02: the copy-out loop has a compile-time trip count (NR rows x KC/8 chunks is always a multiple of 32) and the
// callers hand in row pointers that are a per-tile base plus a 32-bit offset -- the 64-bit address arithmetic, integer
// divisions and loop branches of the r01 version were ~40 % of the epilogue's instructions, and the epilogue (not the
// tensor pipe) is what bounds the 64-channel layers (profiles/r02_ncu_full.md).
// ---------------------------------------------------------------------------------------------------------------
template <int KC>
struct RowStore {
  static constexpr int kChunks = KC / 8;               // 16-byte chunks per row and plane (row = 2 * KC bytes)
  static constexpr int kRowBytes = KC * 2;
  SHF_DEVICE static int swz(int row, int k) { return KC == 64 ? (k ^ (row & 7)) : (k ^ ((row >> 1) & 3)); }
  SHF_DEVICE static void put(uint8_t* stg, int row, int k, const uint4& v) {
    *reinterpret_cast<uint4*>(stg + row * kRowBytes + swz(row, k) * 16) = v;
  }
  SHF_DEVICE static void put8(uint8_t* stg, int row, int k, int half8, const uint2& v) {      // 8 bytes inside chunk k
    *reinterpret_cast<uint2*>(stg + row * kRowBytes + swz(row, k) * 16 + half8 * 8) = v;
  }
  SHF_DEVICE static uint4 get(const uint8_t* stg, int row, int k) {
    return *reinterpret_cast<const uint4*>(stg + row * kRowBytes + swz(row, k) * 16);
  }
};

// fp32 -> plane-0 halfs (hi) of 8 consecutive channels
SHF_DEVICE uint4 pack_hi8(const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    w[e] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// fp32 -> h2 plane-1 halfs: lo = rn16(x - hi)
SHF_DEVICE uint4 pack_lo8(const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = v[2 * e], b = v[2 * e + 1];
    const float2 hf = __half22float2(__floats2half2_rn(a, b));
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    w[e] = *reinterpret_cast<const uint32_t*>(&l);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// fp32 -> hf8 plane-1 bytes of 8 consecutive channels: al8 = e4m3((x - hi) * 2^6), ah8 = e4m3(hi * 2^-5)
SHF_DEVICE void pack_f8x8(const float* v, uint2& al, uint2& ah) {
  uint32_t a[2] = {0u, 0u}, b[2] = {0u, 0u};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float x0 = v[2 * e], x1 = v[2 * e + 1];
    const float2 hf = __half22float2(__floats2half2_rn(x0, x1));
    const float2 lo = make_float2((x0 - hf.x) * kHf8AlScale, (x1 - hf.y) * kHf8AlScale);
    const float2 hs = make_float2(hf.x * kHf8AhScale, hf.y * kHf8AhScale);
    a[e >> 1] |= (uint32_t)__nv_cvt_float2_to_fp8x2(lo, __NV_SATFINITE, __NV_E4M3) << (16 * (e & 1));
    b[e >> 1] |= (uint32_t)__nv_cvt_float2_to_fp8x2(hs, __NV_SATFINITE, __NV_E4M3) << (16 * (e & 1));
  }
  al = make_uint2(a[0], a[1]);
  ah = make_uint2(b[0], b[1]);
}

// Copy NR staged rows of one plane out to global memory.  dst_px(row) returns the plane-0 address of channel 0 of that
// row's pixel, or nullptr when the row is outside the image.  Lane l always serves chunk l % kChunks of rows
// l / kChunks + j * (32 / kChunks): NR * kChunks is a multiple of 32, so the loop has a constant trip count.
template <int KC, int NR, typename DstFn>
SHF_DEVICE void copy_rows_out(const uint8_t* stg, int lane, int plane, int fmt, int c_first, size_t plane_elems, DstFn dst_px) {
  using RS = RowStore<KC>;
  constexpr int kIters = NR * RS::kChunks / 32;
  static_assert(NR * RS::kChunks % 32 == 0 && kIters >= 1, "rows x chunks must fill whole warps");
  constexpr int kRowStep = 32 / RS::kChunks;
  const int k = lane % RS::kChunks, r0 = lane / RS::kChunks;
  // destination offset (in halfs from the pixel's plane-0 channel 0) of chunk k of this plane
  size_t off;
  if (plane == 0) {
    off = (size_t)(c_first + 8 * k);
  } else if (fmt == SHF_FMT_H2) {
    off = plane_elems + (size_t)(c_first + 8 * k);
  } else {
    // chunks [0, KC/16) are al8 of channels c_first + 16k .., chunks [KC/16, KC/8) their ah8 twins 64 bytes further
    const int half = k / (KC / 16), kk = k % (KC / 16);
    off = plane_elems + (size_t)((hf8_off(c_first) + half * 64 + kk * 16) >> 1);
  }
#pragma unroll
  for (int j = 0; j < kIters; ++j) {
    const int row = r0 + j * kRowStep;
    __half* px = dst_px(row);
    if (px != nullptr) *reinterpret_cast<uint4*>(px + off) = RS::get(stg, row, k);
  }
}

// Stage one plane of `v` (KC final values of this lane's pixel = staging row `lane`) and copy the warp's 32 rows out.
template <int KC, typename DstFn>
SHF_DEVICE void store_plane(uint8_t* stg, int lane, const float (&v)[KC], int plane, int fmt, int c_first, size_t plane_elems,
                            DstFn dst_px) {
  using RS = RowStore<KC>;
  if (plane == 0) {
#pragma unroll
    for (int k = 0; k < KC / 8; ++k) RS::put(stg, lane, k, pack_hi8(v + 8 * k));
  } else if (fmt == SHF_FMT_H2) {
#pragma unroll
    for (int k = 0; k < KC / 8; ++k) RS::put(stg, lane, k, pack_lo8(v + 8 * k));
  } else {                                                   // hf8 plane 1: [KC x al8 | KC x ah8]
#pragma unroll
    for (int k = 0; k < KC / 16; ++k) {
      uint2 al0, ah0, al1, ah1;
      pack_f8x8(v + 16 * k, al0, ah0);
      pack_f8x8(v + 16 * k + 8, al1, ah1);
      RS::put(stg, lane, k, make_uint4(al0.x, al0.y, al1.x, al1.y));
      RS::put(stg, lane, KC / 16 + k, make_uint4(ah0.x, ah0.y, ah1.x, ah1.y));
    }
  }
  __syncwarp();
  copy_rows_out<KC, 32>(stg, lane, plane, fmt, c_first, plane_elems, dst_px);
  __syncwarp();
}

// Mirror image of copy_rows_out: fetch NR rows of one plane into the staging buffer, consecutive lanes on consecutive
// 16 bytes of a row (rows outside the image are filled with zeros).
template <int KC, int NR, typename SrcFn>
SHF_DEVICE void copy_rows_in(uint8_t* stg, int lane, int plane, int fmt, int c_first, size_t plane_elems, SrcFn src_px) {
  using RS = RowStore<KC>;
  constexpr int kIters = NR * RS::kChunks / 32;
  constexpr int kRowStep = 32 / RS::kChunks;
  const int k = lane % RS::kChunks, r0 = lane / RS::kChunks;
  size_t off;
  if (plane == 0) {
    off = (size_t)(c_first + 8 * k);
  } else if (fmt == SHF_FMT_H2) {
    off = plane_elems + (size_t)(c_first + 8 * k);
  } else {
    const int half = k / (KC / 16), kk = k % (KC / 16);
    off = plane_elems + (size_t)((hf8_off(c_first) + half * 64 + kk * 16) >> 1);
  }
#pragma unroll
  for (int j = 0; j < kIters; ++j) {
    const int row = r0 + j * kRowStep;
    const __half* px = src_px(row);
    const uint4 v = px != nullptr ? __ldg(reinterpret_cast<const uint4*>(px + off)) : make_uint4(0u, 0u, 0u, 0u);
    RS::put(stg, row, k, v);
  }
}

// v[0, KC) += the values of this lane's pixel (staging row `lane`) of an activation tensor in format `fmt`: the residual
// add of a ResNet block inside the conv epilogue.  Rows travel through the staging buffer like the stores do, so the
// global loads are whole 128-byte (64-byte for KC = 32) rows per pixel instead of 16 bytes per lane and instruction.
template <int KC, typename SrcFn>
SHF_DEVICE void add_rows_in(uint8_t* stg, int lane, float (&v)[KC], int fmt, int c_first, size_t plane_elems, SrcFn src_px) {
  using RS = RowStore<KC>;
  copy_rows_in<KC, 32>(stg, lane, 0, fmt, c_first, plane_elems, src_px);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < KC / 8; ++k) {
    const uint4 q = RS::get(stg, lane, k);
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h[e]);
      v[8 * k + 2 * e] += f.x;
      v[8 * k + 2 * e + 1] += f.y;
    }
  }
  __syncwarp();
  copy_rows_in<KC, 32>(stg, lane, 1, fmt, c_first, plane_elems, src_px);
  __syncwarp();
  if (fmt == SHF_FMT_H2) {
#pragma unroll
    for (int k = 0; k < KC / 8; ++k) {
      const uint4 q = RS::get(stg, lane, k);
      const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        v[8 * k + 2 * e] += f.x;
        v[8 * k + 2 * e + 1] += f.y;
      }
    }
  } else {                                                   // chunks [0, KC/16): al8 of 16 channels each; the ah8 twins are not needed
#pragma unroll
    for (int k = 0; k < KC / 16; ++k) {
      const uint4 q = RS::get(stg, lane, k);
      const uint16_t* b2 = reinterpret_cast<const uint16_t*>(&q);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const __half2_raw r2 = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)b2[e], __NV_E4M3);
        const float2 f = __half22float2(__half2(r2));
        v[16 * k + 2 * e] = fmaf(f.x, kHf8AlInv, v[16 * k + 2 * e]);
        v[16 * k + 2 * e + 1] = fmaf(f.y, kHf8AlInv, v[16 * k + 2 * e + 1]);
      }
    }
  }
  __syncwarp();
}

// Fused 2x2 / stride-2 max pooling (pooling_layer.cpp:140-187) of a warp's 4 x 8 pixel patch (lane = y * 8 + x), then
// the same staged store for the 2 x 4 pooled pixels.  The four lanes of a window (l, l^1, l^8, l^9) SPLIT the channels
// while they reduce: the x-exchange leaves each lane with the pair maximum of one half of the channels, the y-exchange
// with the window maximum of one quarter -- 3/4 KC shuffles and maxima per lane instead of 2 KC, and the conversions
// and staging stores then run on KC / 4 values in all 32 lanes instead of KC values in the 8 window-origin lanes (the
// other 24 predicated off but still issued).  H and W are even and tiles are 16 x 8 aligned, so a window is never cut by
// the image border.  v is consumed (overwritten).
// post(value, channel) is applied to the KC / 4 window maxima this lane ends up owning (channel = index within the warp's
// KC): callers whose un-pooled tensor is not written pass the bias + ReLU there -- max pooling commutes exactly with a
// non-decreasing map, so it runs on a quarter of the values.
template <int KC, typename DstFn, typename PostFn>
SHF_DEVICE void store_pooled(uint8_t* stg, int lane, float (&v)[KC], int fmt, int c_first, size_t plane_elems, DstFn dst_px,
                             PostFn post) {
  using RS = RowStore<KC>;
  constexpr int H2c = KC / 2, Q = KC / 4;                    // channels kept after the x- / y-exchange
  const bool bx = lane & 1, by = lane & 8;
#pragma unroll
  for (int c = 0; c < H2c; ++c) {                            // in place: v[0, H2c) <- pair maxima of this lane's half
    const float mine = bx ? v[H2c + c] : v[c];
    const float send = bx ? v[c] : v[H2c + c];
    v[c] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 1));
  }
  float* q = v;                                             // v[0, Q) <- window maxima of this lane's quarter
#pragma unroll
  for (int c = 0; c < Q; ++c) {
    const float mine = by ? v[Q + c] : v[c];
    const float send = by ? v[c] : v[Q + c];
    v[c] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 8));
  }
  const int cs = (bx ? H2c : 0) + (by ? Q : 0);              // first channel (within the warp's KC) this lane now owns
#pragma unroll
  for (int c = 0; c < Q; ++c) v[c] = post(v[c], cs + c);
  const int prow = ((lane >> 4) << 2) | ((lane >> 1) & 3);   // pooled pixel: (y >> 1) * 4 + (x >> 1)
#pragma unroll
  for (int plane = 0; plane < 2; ++plane) {
    if (plane == 0 || fmt == SHF_FMT_H2) {
#pragma unroll
      for (int t = 0; t < Q / 8; ++t)
        RS::put(stg, prow, cs / 8 + t, plane == 0 ? pack_hi8(q + 8 * t) : pack_lo8(q + 8 * t));
    } else {
#pragma unroll
      for (int t = 0; t < Q / 8; ++t) {
        uint2 al, ah;
        pack_f8x8(q + 8 * t, al, ah);
        const int ch = cs + 8 * t;                           // al8 byte offset in the row; ah8 sits KC bytes further
        RS::put8(stg, prow, ch / 16, (ch >> 3) & 1, al);
        RS::put8(stg, prow, KC / 16 + ch / 16, (ch >> 3) & 1, ah);
      }
    }
    __syncwarp();
    copy_rows_out<KC, 8>(stg, lane, plane, fmt, c_first, plane_elems, dst_px);
    __syncwarp();
  }
}

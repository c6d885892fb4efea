// Detection tail on the device (sm_100a SIMT kernels; HBM/latency-bound integer + fp32 work):
//   head_decode   : the 1x1 cls/bbox convs + 2-way softmax + anchor decode + clip + candidate keys
//                   (replaces base_conv_layer.cpp:255-279 for the 2/4/6/12-channel heads, concat_layer.cpp,
//                    reshape_layer.cpp, softmax_layer.cpp:27-60 and lib/layers/proposal_layer.py:96-173)
//   sort          : (score desc, anchor index asc) segmented LSD radix sort, one CTA per segment, hand-written
//                   (proposal_layer.py:180-190, lib/test.py:182, cpu_nms.pyx:25)
//   gather        : top-K rows -> 'boxes' (R,5) / 'cls_prob' (R,2) blobs, plus the per-pass un-mirror,
//                   unscale and 0.05 threshold of lib/test.py:52-66,163-167 appended to the image's det list
//   nms / vote    : greedy NMS (lib/nms/cpu_nms.pyx:17-68, nms_kernel.cu:45-155 semantics selectable) and
//                   box voting (lib/test.py:181-217), batched over images: IoU bit masks computed by the whole GPU
//                   (64 x 64 tiles), a 64-rows-per-step sweep per image, one warp per cluster for the merged boxes;
//                   images with more rows than the mask workspace holds take the one-CTA serial sweep
//   bbox_overlaps : IoU / IoA / self-overlap matrices (lib/utils/bbox.pyx:14-142), float64
// All fp32 box arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) so that IoU
// values, and therefore keep/suppress decisions, are bit-identical to the NumPy/Cython reference.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kMaxAnchors = 8;

struct HeadParams {
  const __half* feat[kMaxAnchors];   // per-anchor head feature map, h2 NHWC [2][1][H][W][C]
  const float* wc;                   // [A][2][C]  (bg, fg) 1x1 weights
  const float* bc;                   // [A][2]
  const float* wb;                   // [A][4][C]  (dx, dy, dw, dh)
  const float* bb;                   // [A][4]
  float anchors[kMaxAnchors][4];     // base anchors (x1,y1,x2,y2), generate_anchors.py
  int A, H, W, C;
  long long plane_stride;            // elements between the hi and lo planes (= N*H*W*C of the batched tensor)
  long long image_stride;            // elements between consecutive images of the batch (H*W*C); batched launches only
  int batched;                       // 1: blockIdx.y = image, outputs strided per image, keys carry the image tag
  int feat_stride;
  float im_h, im_w;                  // unpadded level size (im_info[0:2])
  float min_size;                    // ANCHOR_MIN_SIZE * im_info[2]
  float score_thresh;                // SCORE_THRESH (0.002)
};

// Sort keys are ascending; the score field holds the bitwise complement of the usual order-preserving image of a float
// (sign bit flipped for x >= 0, all bits flipped for x < 0), so ascending keys = descending scores for ANY float --
// negative logits and NaNs included (NaNs first, like `argsort()[::-1]`) -- and ties fall back to the row index.
// A score field of 0xffffffff would be -NaN with a full mantissa: never produced by arithmetic, it marks sentinels.
SHF_DEVICE unsigned score_desc_bits(float s) {
  const unsigned b = __float_as_uint(s);
  return ~((b & 0x80000000u) ? ~b : (b | 0x80000000u));
}
SHF_DEVICE float score_from_desc_bits(unsigned k) {
  const unsigned asc = ~k;
  return __uint_as_float((asc & 0x80000000u) ? (asc & 0x7fffffffu) : ~asc);
}
SHF_DEVICE unsigned long long make_key(float score, unsigned idx) {
  return ((unsigned long long)score_desc_bits(score) << 32) | idx;
}
// Batched variant: [image : 5][~score bits : 32][row : 27] -- one device-wide sort leaves every image's rows
// contiguous (each image contributes exactly n keys, sentinels included) and ordered like make_key.
constexpr int kRowBits = 27;
SHF_DEVICE unsigned long long make_key_img(unsigned img, float score, unsigned idx) {
  return ((unsigned long long)img << 59) | ((unsigned long long)score_desc_bits(score) << kRowBits) | idx;
}
SHF_DEVICE unsigned long long sentinel_img(unsigned img) { return ((unsigned long long)img << 59) | ((1ull << 59) - 1); }
SHF_DEVICE float key_img_score(unsigned long long k) { return score_from_desc_bits((unsigned)((k >> kRowBits) & 0xffffffffu)); }
SHF_DEVICE unsigned key_img_row(unsigned long long k) { return (unsigned)(k & ((1u << kRowBits) - 1)); }

__global__ void __launch_bounds__(128) head_decode_kernel(const HeadParams p, float* __restrict__ prob,
                                                          float* __restrict__ delta, float* __restrict__ boxes,
                                                          unsigned long long* __restrict__ keys, int* __restrict__ count,
                                                          unsigned long long* __restrict__ best_key) {
  extern __shared__ float wsm[];                 // [A][6][C] then biases [A][6]
  const int A = p.A, C = p.C;
  for (int i = threadIdx.x; i < A * 6 * C; i += blockDim.x) {
    const int a = i / (6 * C), r = (i / C) % 6, c = i % C;
    wsm[i] = r < 2 ? p.wc[((size_t)a * 2 + r) * C + c] : p.wb[((size_t)a * 4 + (r - 2)) * C + c];
  }
  float* bsm = wsm + A * 6 * C;
  for (int i = threadIdx.x; i < A * 6; i += blockDim.x) {
    const int a = i / 6, r = i % 6;
    bsm[i] = r < 2 ? p.bc[a * 2 + r] : p.bb[a * 4 + (r - 2)];
  }
  __syncthreads();
  const int hw = p.H * p.W;
  const int n = hw * A;
  const size_t plane = (size_t)p.plane_stride;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // anchor row index, (h, w, a) order
  const unsigned img = p.batched ? blockIdx.y : 0u;
  if (p.batched) {                                          // per-image output slices
    prob += (size_t)img * 2 * A * hw;
    delta += (size_t)img * 4 * A * hw;
    boxes += (size_t)img * n * 4;
    keys += (size_t)img * n;
    count += img;
    best_key += img;
  }
  unsigned long long key = p.batched ? sentinel_img(img) : ~0ull, bkey = ~0ull;
  bool cand = false;
  if (i < n) {
    const int a = i % A, pix = i / A;
    const int x = pix % p.W, y = pix / p.W;
    const __half* fh = p.feat[a] + (size_t)img * p.image_stride + (size_t)pix * C;
    const __half* fl = fh + plane;
    const float* w = wsm + (size_t)a * 6 * C;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < C; c0 += 8) {
      const uint4 vh = __ldg(reinterpret_cast<const uint4*>(fh + c0));
      const uint4 vl = __ldg(reinterpret_cast<const uint4*>(fl + c0));
      const __half2* hh = reinterpret_cast<const __half2*>(&vh);
      const __half2* ll = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a2 = __half22float2(hh[j]), b2 = __half22float2(ll[j]);
        const float f0 = a2.x + b2.x, f1 = a2.y + b2.y;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          acc[r] = fmaf(f0, w[r * C + c0 + 2 * j], acc[r]);
          acc[r] = fmaf(f1, w[r * C + c0 + 2 * j + 1], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) acc[r] += bsm[a * 6 + r];
    // softmax over (bg, fg): softmax_layer.cpp:27-60 (subtract max, exp, divide by sum)
    const float mx = fmaxf(acc[0], acc[1]);
    const float e0 = expf(__fsub_rn(acc[0], mx)), e1 = expf(__fsub_rn(acc[1], mx));
    const float sum = __fadd_rn(e0, e1);
    const float pbg = __fdiv_rn(e0, sum), pfg = __fdiv_rn(e1, sum);
    prob[(size_t)a * hw + pix] = pbg;
    prob[(size_t)(A + a) * hw + pix] = pfg;
#pragma unroll
    for (int r = 0; r < 4; ++r) delta[(size_t)(a * 4 + r) * hw + pix] = acc[2 + r];
    // anchor decode: bbox_transform.py:33-77 in float32, mul and add NOT fused
    const float sx = (float)(x * p.feat_stride), sy = (float)(y * p.feat_stride);
    const float ax1 = p.anchors[a][0] + sx, ay1 = p.anchors[a][1] + sy;
    const float ax2 = p.anchors[a][2] + sx, ay2 = p.anchors[a][3] + sy;
    const float aw = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f), ah = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
    const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, aw)), cy = __fadd_rn(ay1, __fmul_rn(0.5f, ah));
    const float pcx = __fadd_rn(__fmul_rn(acc[2], aw), cx), pcy = __fadd_rn(__fmul_rn(acc[3], ah), cy);
    const float pw = __fmul_rn(expf(acc[4]), aw), ph = __fmul_rn(expf(acc[5]), ah);
    float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)), y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
    // clip_boxes, bbox_transform.py:80-93
    const float xm = __fsub_rn(p.im_w, 1.0f), ym = __fsub_rn(p.im_h, 1.0f);
    x1 = fmaxf(fminf(x1, xm), 0.f); y1 = fmaxf(fminf(y1, ym), 0.f);
    x2 = fmaxf(fminf(x2, xm), 0.f); y2 = fmaxf(fminf(y2, ym), 0.f);
    reinterpret_cast<float4*>(boxes)[i] = make_float4(x1, y1, x2, y2);
    // _filter_boxes (proposal_layer.py:231-236) then score threshold (:183-190)
    const float bw = __fadd_rn(__fsub_rn(x2, x1), 1.0f), bh = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
    if (bw >= p.min_size && bh >= p.min_size) {
      bkey = p.batched ? make_key_img(img, pfg, (unsigned)i) : make_key(pfg, (unsigned)i);
      if (pfg >= p.score_thresh) { key = bkey; cand = true; }
    }
    keys[i] = key;
  }
  // warp-aggregated candidate count and arg-best (the "nothing above threshold: keep the largest" rule)
  const unsigned ballot = __ballot_sync(0xffffffffu, cand);
  unsigned long long wbest = bkey;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, wbest, o);
    wbest = other < wbest ? other : wbest;
  }
  if ((threadIdx.x & 31) == 0) {
    if (ballot) atomicAdd(count, __popc(ballot));
    if (wbest != ~0ull) atomicMin(best_key, wbest);
  }
}

struct GatherParams {
  const unsigned long long* sorted;  // ascending keys (candidates first, ~0 sentinels last)
  const int* count;
  const unsigned long long* best_key;
  const float* prob;                 // [2A][hw]
  const float* boxes;                // [n][4]
  int A, hw, topn;
  float* out_boxes;                  // [topn][5]  (0, x1, y1, x2, y2)
  float* out_probs;                  // [topn][2]  (bg, fg)
  int* out_rows;                     // R
  // image-level accumulation (optional: dets == nullptr disables)
  float* dets;                       // [cap][5] (x1, y1, x2, y2, score) in raw-image coordinates
  int* pass_offsets;                 // [passes + 1]
  int pass, det_cap;
  int flip;
  float level_w;                     // unpadded level width (lib/test.py:52-54)
  float im_scale;
  float det_thresh;                  // 0.05, strict >
};

__global__ void __launch_bounds__(256) proposal_gather_kernel(const GatherParams g) {
  const int cnt = *g.count;
  const bool any = *g.best_key != ~0ull;
  const int R = cnt > 0 ? min(cnt, g.topn) : (any ? 1 : 0);
  __shared__ int s_keep, s_base;
  if (threadIdx.x == 0) {
    // rows are sorted by descending fg score: the rows with score > det_thresh are a prefix; find its length
    int lo = 0, hi = R;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const unsigned long long k = cnt > 0 ? g.sorted[mid] : *g.best_key;
      const float s = score_from_desc_bits((unsigned)(k >> 32));
      if (s > g.det_thresh) lo = mid + 1; else hi = mid;
    }
    s_keep = lo;
    s_base = g.dets ? g.pass_offsets[g.pass] : 0;
    if (blockIdx.x == 0) {
      *g.out_rows = R;
      if (g.dets) g.pass_offsets[g.pass + 1] = min(s_base + lo, g.det_cap);
    }
  }
  __syncthreads();
  const int keep = s_keep, base = s_base;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < R; j += gridDim.x * blockDim.x) {
    const unsigned long long k = cnt > 0 ? g.sorted[j] : *g.best_key;
    const unsigned i = (unsigned)(k & 0xffffffffu);
    const int a = i % g.A, pix = i / g.A;
    const float4 b = reinterpret_cast<const float4*>(g.boxes)[i];
    const float pbg = g.prob[(size_t)a * g.hw + pix], pfg = g.prob[(size_t)(g.A + a) * g.hw + pix];
    float* ob = g.out_boxes + (size_t)j * 5;
    ob[0] = 0.f; ob[1] = b.x; ob[2] = b.y; ob[3] = b.z; ob[4] = b.w;
    g.out_probs[(size_t)j * 2] = pbg;
    g.out_probs[(size_t)j * 2 + 1] = pfg;
    if (g.dets && j < keep && base + j < g.det_cap) {
      float x1 = b.x, x2 = b.z;
      if (g.flip) { x1 = __fsub_rn(g.level_w, b.z); x2 = __fsub_rn(g.level_w, b.x); }   // boxes[:, [1,3]] = w - boxes[:, [3,1]]
      float* d = g.dets + (size_t)(base + j) * 5;
      d[0] = __fdiv_rn(x1, g.im_scale); d[1] = __fdiv_rn(b.y, g.im_scale);
      d[2] = __fdiv_rn(x2, g.im_scale); d[3] = __fdiv_rn(b.w, g.im_scale);
      d[4] = pfg;
    }
  }
}

// Batched form used by the pyramid driver: one launch per level for all images of the batch and both mirror
// passes.  blockIdx.y = image; the two passes of an image (plain, mirrored) are appended back to back in the
// reference's order, which makes the pass-offset chain of lib/test.py:141-158 a per-image serial dependency that
// one block row resolves locally.
struct GatherBatchParams {
  const unsigned long long* sorted;  // [slots][n] image-tagged keys, each slot's rows contiguous and ordered
  const int* count;                  // [slots]
  const unsigned long long* best_key;// [slots]
  const float* prob;                 // [slots][2A][hw]
  const float* boxes;                // [slots][n][4]
  int A, hw, n, topn, nf;            // nf = passes per image at this level (1 or 2: plain, mirrored)
  float* dets;                       // [images][det_cap][5]
  int* pass_offsets;                 // [images][passes_total + 1]
  int image_base, passes_total, pass_base, det_cap;
  float level_w, im_scale, det_thresh;
};

__global__ void __launch_bounds__(256) gather_dets_batched_kernel(const GatherBatchParams g) {
  const int j = blockIdx.y;                              // image within the level batch
  const int image = g.image_base + j;
  int* offs = g.pass_offsets + (size_t)image * (g.passes_total + 1);
  float* dets = g.dets + (size_t)image * g.det_cap * 5;
  __shared__ int s_R[2], s_keep[2], s_base[2];
  if (threadIdx.x < g.nf) {
    const int f = threadIdx.x, slot = j * g.nf + f;
    const int cnt = g.count[slot];
    const bool any = g.best_key[slot] != ~0ull;
    const int R = cnt > 0 ? min(cnt, g.topn) : (any ? 1 : 0);
    const unsigned long long* keys = g.sorted + (size_t)slot * g.n;
    int lo = 0, hi = R;                                  // rows with score > det_thresh are a prefix
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const unsigned long long k = cnt > 0 ? keys[mid] : g.best_key[slot];
      if (key_img_score(k) > g.det_thresh) lo = mid + 1; else hi = mid;
    }
    s_R[f] = R;
    s_keep[f] = lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int base = offs[g.pass_base];
    for (int f = 0; f < g.nf; ++f) {
      s_base[f] = base;
      base = min(base + s_keep[f], g.det_cap);
      if (blockIdx.x == 0) offs[g.pass_base + f + 1] = base;
    }
  }
  __syncthreads();
  for (int f = 0; f < g.nf; ++f) {
    const int slot = j * g.nf + f;
    const int keep = s_keep[f], base = s_base[f];
    const int cnt = g.count[slot];
    const unsigned long long* keys = g.sorted + (size_t)slot * g.n;
    const float* prob = g.prob + (size_t)slot * 2 * g.A * g.hw;
    const float4* boxes = reinterpret_cast<const float4*>(g.boxes) + (size_t)slot * g.n;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < keep; r += gridDim.x * blockDim.x) {
      if (base + r >= g.det_cap) break;
      const unsigned long long k = cnt > 0 ? keys[r] : g.best_key[slot];
      const unsigned i = key_img_row(k);
      const int a = i % g.A, pix = i / g.A;
      const float4 b = boxes[i];
      const float pfg = prob[(size_t)(g.A + a) * g.hw + pix];
      float x1 = b.x, x2 = b.z;
      if (f == 1) { x1 = __fsub_rn(g.level_w, b.z); x2 = __fsub_rn(g.level_w, b.x); }      // un-mirror, lib/test.py:52-54
      float* d = dets + (size_t)(base + r) * 5;
      d[0] = __fdiv_rn(x1, g.im_scale); d[1] = __fdiv_rn(b.y, g.im_scale);
      d[2] = __fdiv_rn(x2, g.im_scale); d[3] = __fdiv_rn(b.w, g.im_scale);
      d[4] = pfg;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Segmented radix sort: `max_score.argsort()[::-1]` (proposal_layer.py:181), `det[:, 4].argsort()[::-1]`
// (lib/test.py:182) and `scores.argsort()[::-1]` (cpu_nms.pyx:25) as ONE launch.
//
// One CTA (1024 threads) per segment.  Keys arrive in ascending row order with sentinels (score field all ones) for
// rows that are not candidates.  Pass 0 streams the segment once: a stable compaction drops the sentinels (usually
// ~90 % of an anchor grid) and the four 8-bit digit histograms of the 32-bit score field are counted on the way.
// Four stable LSD passes then order the survivors by score; rows with equal scores keep their arrival order = lower
// row first, which is the order the oracle defines for ties.  The segment's working set (<= 1.5 MB) stays in L2.
//   stable ranking inside a tile of 4096 keys: warp w owns 4 rounds of 32 consecutive keys; __match_any_sync groups
//   the lanes of a round by digit, the group's first lane bumps the warp's private histogram, and a 32-step scan
//   per digit across the warps (threads 0..255) turns the private histograms into scatter offsets.
// ---------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 1024;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 4;                               // keys per thread and tile
constexpr int kSortTile = kSortThreads * kSortRounds;

__global__ void __launch_bounds__(kSortThreads) segmented_sort_kernel(
    const unsigned long long* __restrict__ keys_in, unsigned long long* __restrict__ keys_out,
    unsigned long long* __restrict__ tmp, const int* __restrict__ seg_begin, const int* __restrict__ seg_end,
    int fixed_len, long long stride, int begin_bit, int* __restrict__ out_count) {
  __shared__ unsigned s_wh[kSortWarps][256];                 // per-warp digit histograms -> scatter offsets
  __shared__ unsigned s_off[4][256];                         // per pass: global digit histogram -> running bin offset
  __shared__ unsigned s_wcnt[kSortWarps];
  __shared__ unsigned s_total;
  const int seg = blockIdx.x;
  const int n = seg_end ? max(0, min(seg_end[seg] - (seg_begin ? seg_begin[seg] : 0), fixed_len)) : fixed_len;
  const unsigned long long* in = keys_in + (size_t)seg * stride;
  unsigned long long* a = keys_out + (size_t)seg * stride;
  unsigned long long* b = tmp + (size_t)seg * stride;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (int i = tid; i < 4 * 256; i += kSortThreads) (&s_off[0][0])[i] = 0u;
  if (tid == 0) s_total = 0u;
  __syncthreads();
  // ---- pass 0: stable compaction of the non-sentinel keys into `a`, digit histograms ----
  for (int base = 0; base < n; base += kSortTile) {
    unsigned long long k[kSortRounds];
    unsigned bal[kSortRounds];
    unsigned cnt = 0;
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const int idx = base + (w * kSortRounds + r) * 32 + lane;
      k[r] = idx < n ? in[idx] : ~0ull;
      const bool valid = (unsigned)(k[r] >> begin_bit) != 0xffffffffu;
      bal[r] = __ballot_sync(0xffffffffu, valid);
      cnt += __popc(bal[r]);
    }
    if (lane == 0) s_wcnt[w] = cnt;
    __syncthreads();
    unsigned v = s_wcnt[lane];                               // every warp scans the 32 warp counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    const unsigned tile_total = __shfl_sync(0xffffffffu, v, 31);
    unsigned dst = s_total + __shfl_sync(0xffffffffu, v, w) - s_wcnt[w];
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      if (bal[r] & (1u << lane)) {
        a[dst + __popc(bal[r] & lt)] = k[r];
        const unsigned f = (unsigned)(k[r] >> begin_bit);
        atomicAdd(&s_off[0][f & 255u], 1u);
        atomicAdd(&s_off[1][(f >> 8) & 255u], 1u);
        atomicAdd(&s_off[2][(f >> 16) & 255u], 1u);
        atomicAdd(&s_off[3][f >> 24], 1u);
      }
      dst += __popc(bal[r]);
    }
    __syncthreads();
    if (tid == 0) s_total += tile_total;
    __syncthreads();
  }
  const int m = (int)s_total;
  if (tid == 0 && out_count) out_count[seg] = m;
  // histograms -> exclusive offsets (warp p scans pass p's 256 bins, 8 per lane)
  if (w < 4) {
    unsigned c[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = s_off[w][lane * 8 + j]; sum += c[j]; }
    unsigned v = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    unsigned run = v - sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_off[w][lane * 8 + j] = run; run += c[j]; }
  }
  __syncthreads();
  // ---- four stable LSD passes over the m survivors: a -> b -> a -> b -> a ----
  unsigned long long* src = a;
  unsigned long long* dstp = b;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = begin_bit + 8 * pass;
    for (int base = 0; base < m; base += kSortTile) {
      for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_wh[0][0])[i] = 0u;
      __syncthreads();
      unsigned long long k[kSortRounds];
      unsigned d[kSortRounds], lr[kSortRounds];
#pragma unroll
      for (int r = 0; r < kSortRounds; ++r) {
        const int idx = base + (w * kSortRounds + r) * 32 + lane;
        const bool valid = idx < m;
        k[r] = valid ? src[idx] : 0ull;
        d[r] = valid ? (unsigned)(k[r] >> shift) & 255u : 0xffffffffu;      // invalid lanes form their own group
        const unsigned peers = __match_any_sync(0xffffffffu, d[r]);
        const int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (valid && lane == leader) {
          old = s_wh[w][d[r]];
          s_wh[w][d[r]] = old + __popc(peers);
        }
        __syncwarp();
        old = __shfl_sync(0xffffffffu, old, leader);
        lr[r] = old + __popc(peers & lt);
      }
      __syncthreads();
      if (tid < 256) {
        unsigned run = s_off[pass][tid];
#pragma unroll 8
        for (int ww = 0; ww < kSortWarps; ++ww) {
          const unsigned t = s_wh[ww][tid];
          s_wh[ww][tid] = run;
          run += t;
        }
        s_off[pass][tid] = run;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kSortRounds; ++r)
        if (d[r] != 0xffffffffu) dstp[s_wh[w][d[r]] + lr[r]] = k[r];
      __syncthreads();
    }
    unsigned long long* t = src; src = dstp; dstp = t;
  }
  // after an even number of passes the sorted keys are back in `a` = keys_out; pad the tail with sentinels so that
  // readers that scan past `count` (none today) see the old convention
  for (int i = m + tid; i < n; i += kSortThreads) a[i] = ~0ull;
}

// ---------------------------------------------------------------------------------------------------
// image-level greedy sweeps
// ---------------------------------------------------------------------------------------------------
__global__ void det_keys_kernel(const float* __restrict__ dets, const int* __restrict__ seg_begin,
                                const int* __restrict__ seg_end, int cap_per_image, unsigned long long* __restrict__ keys) {
  // keys[img][0..n): (score desc, row index asc) for the image's n rows -- nothing is written (or sorted) beyond n
  const int img = blockIdx.y;
  const int n = min(seg_end[img] - seg_begin[img], cap_per_image);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    keys[(size_t)img * cap_per_image + j] = make_key(dets[(size_t)(seg_begin[img] + j) * 5 + 4], (unsigned)j);
}

SHF_DEVICE float box_area(const float4& b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}
SHF_DEVICE float box_iou(const float4& a, float aa, const float4& b, float ab) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
  const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
}
// mode 0: cpu_nms.pyx:65  (double)ovr >= thresh ; mode 1: nms_kernel.cu:82  ovr > (float)thresh ;
// mode 2: lib/test.py:198 (bbox_vote) / py_cpu_nms  float32 compare against (float)thresh, >=
SHF_DEVICE bool overlaps(float ovr, double thr, int mode) {
  if (mode == 0) return (double)ovr >= thr;
  if (mode == 1) return ovr > (float)thr;
  return ovr >= (float)thr;
}

constexpr int kSweepThreads = 1024;

// One CTA per image.  sorted_keys gives the descending-score order; alive[] lives in global scratch.
// NMS : emits kept ORIGINAL row indices (into the image's dets segment) in kept order.
// VOTE: emits merged boxes (x1,y1,x2,y2,score) in emission order.
template <bool VOTE>
__global__ void __launch_bounds__(kSweepThreads) greedy_sweep_kernel(
    const float* __restrict__ dets, const int* __restrict__ seg_begin, const int* __restrict__ seg_end,
    const unsigned long long* __restrict__ sorted_keys, int cap_per_image, unsigned char* __restrict__ alive_all,
    float4* __restrict__ sbox_all, double thr, int mode, int* __restrict__ out_idx, float* __restrict__ out_dets,
    int* __restrict__ out_count, int out_cap, int mask_rows) {
  const int img = blockIdx.x;
  const int n = min(seg_end[img] - seg_begin[img], cap_per_image);
  if (n <= mask_rows) return;                  // the bit-mask kernels own this image
  const float* d = dets + (size_t)seg_begin[img] * 5;
  const unsigned long long* keys = sorted_keys + (size_t)img * cap_per_image;
  unsigned char* alive = alive_all + (size_t)img * cap_per_image;
  float4* sbox = sbox_all + (size_t)img * cap_per_image;       // boxes in sorted order
  int* oidx = out_idx ? out_idx + (size_t)img * out_cap : nullptr;
  float* odet = out_dets ? out_dets + (size_t)img * out_cap * 5 : nullptr;
  __shared__ double red[5][kSweepThreads / 32];
  __shared__ float redmax[kSweepThreads / 32];
  __shared__ int redcnt[kSweepThreads / 32];
  __shared__ int s_next, s_emit;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int j = tid; j < n; j += kSweepThreads) alive[j] = 1;      // sbox[] was filled by sorted_boxes_kernel
  if (tid == 0) { s_next = 0; s_emit = 0; }
  __syncthreads();
  if (n == 0) {
    if (tid == 0) {
      if (VOTE) {       // lib/test.py:184-186: empty input -> [[10, 10, 20, 20, 0.0001]]
        odet[0] = 10.f; odet[1] = 10.f; odet[2] = 20.f; odet[3] = 20.f; odet[4] = 0.0001f;
        out_count[img] = 1;
      } else {
        out_count[img] = 0;
      }
    }
    return;
  }
  int i = 0;
  while (true) {
    // i = first alive row (uniform across the CTA)
    i = s_next;
    if (i >= n) break;
    const float4 bi = sbox[i];
    const float ai = box_area(bi);
    const float si = d[(size_t)(keys[i] & 0xffffffffu) * 5 + 4];
    // pass 1: mark members (IoU test against the current top), accumulate cluster statistics
    double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0, ss = 0;
    float smax = -1.f;
    int members = 0, first_other = n;
    for (int j = i + tid; j < n; j += kSweepThreads) {
      if (!alive[j]) continue;
      const float4 bj = sbox[j];
      const bool hit = (j == i) || overlaps(box_iou(bi, ai, bj, box_area(bj)), thr, mode);
      if (hit) {
        if (VOTE) {
          const float sj = d[(size_t)(keys[j] & 0xffffffffu) * 5 + 4];
          // det_accu[:, 0:4] * score in float32 (lib/test.py:207), then summed
          sx1 += (double)__fmul_rn(bj.x, sj); sy1 += (double)__fmul_rn(bj.y, sj);
          sx2 += (double)__fmul_rn(bj.z, sj); sy2 += (double)__fmul_rn(bj.w, sj);
          ss += (double)sj;
          smax = fmaxf(smax, sj);
        }
        ++members;
        alive[j] = 0;
      } else {
        first_other = min(first_other, j);
      }
    }
    // block reductions: members (sum), first surviving row (min), cluster sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      members += __shfl_xor_sync(0xffffffffu, members, o);
      first_other = min(first_other, __shfl_xor_sync(0xffffffffu, first_other, o));
      if (VOTE) {
        sx1 += __shfl_xor_sync(0xffffffffu, sx1, o); sy1 += __shfl_xor_sync(0xffffffffu, sy1, o);
        sx2 += __shfl_xor_sync(0xffffffffu, sx2, o); sy2 += __shfl_xor_sync(0xffffffffu, sy2, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
      }
    }
    __syncthreads();            // previous iteration's readers of s_next / red are done
    if (lane == 0) {
      redcnt[wid] = members;
      redmax[wid] = smax;
      red[0][wid] = sx1; red[1][wid] = sy1; red[2][wid] = sx2; red[3][wid] = sy2; red[4][wid] = ss;
      // reuse red slot for first_other via integer min in smem below
    }
    __shared__ int redmin[kSweepThreads / 32];
    if (lane == 0) redmin[wid] = first_other;
    __syncthreads();
    if (tid == 0) {
      int tot = 0, nxt = n;
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
      float mx = -1.f;
      for (int k = 0; k < kSweepThreads / 32; ++k) {
        tot += redcnt[k];
        nxt = min(nxt, redmin[k]);
        a0 += red[0][k]; a1 += red[1][k]; a2 += red[2][k]; a3 += red[3][k]; a4 += red[4][k];
        mx = fmaxf(mx, redmax[k]);
      }
      const int e = s_emit;
      if (!VOTE) {
        if (e < out_cap) oidx[e] = (int)(keys[i] & 0xffffffffu);
        s_emit = e + 1;
      } else {
        if (tot <= 1) {
          // singleton: emitted only when nothing remains (lib/test.py:200-206)
          if (nxt >= n) {
            if (e < out_cap) {
              odet[(size_t)e * 5] = bi.x; odet[(size_t)e * 5 + 1] = bi.y; odet[(size_t)e * 5 + 2] = bi.z;
              odet[(size_t)e * 5 + 3] = bi.w; odet[(size_t)e * 5 + 4] = si;
            }
            s_emit = e + 1;
          }
        } else {
          if (e < out_cap) {
            const float fs = (float)a4;
            odet[(size_t)e * 5] = __fdiv_rn((float)a0, fs); odet[(size_t)e * 5 + 1] = __fdiv_rn((float)a1, fs);
            odet[(size_t)e * 5 + 2] = __fdiv_rn((float)a2, fs); odet[(size_t)e * 5 + 3] = __fdiv_rn((float)a3, fs);
            odet[(size_t)e * 5 + 4] = mx;
          }
          s_emit = e + 1;
        }
      }
      s_next = nxt;
    }
    __syncthreads();
  }
  if (tid == 0) out_count[img] = s_emit;       // the TRUE count: the caller must treat count > out_cap as an error
}

// ---------------------------------------------------------------------------------------------------
// Bit-mask NMS / box voting (images with n <= mask_rows rows; shape of lib/nms/nms_kernel.cu:45-89, semantics of
// cpu_nms.pyx / lib/test.py:181-217 selectable through `mode`):
//   sorted_boxes : boxes gathered into descending-score order (float4 per row)
//   iou_mask     : M[i][w] bit c = row j = 64 w + c overlaps row i (j > i only), 64 x 64 tiles, persistent grid
//   mask_sweep   : one CTA per image walks 64 rows per step.  The first still-alive row i is a cluster head / a kept
//                  box; every alive j with M[i][j] set joins it and dies.  Inside a step the 64 x 64 diagonal tile is
//                  resolved serially on register bit masks, then thread w applies the step's heads IN ORDER to its own
//                  64-bit word of the alive set.  For voting the row of a head is overwritten with its MEMBER mask.
//   vote_reduce  : one warp per head: score-weighted box sums over the member mask (float32 products accumulated in
//                  double, as greedy_sweep_kernel), singleton clusters dropped unless they are the last head
//                  (lib/test.py:200-206), emission order = head order.
// ---------------------------------------------------------------------------------------------------
__global__ void sorted_boxes_kernel(const float* __restrict__ dets, const int* __restrict__ seg_begin,
                                    const int* __restrict__ seg_end, const unsigned long long* __restrict__ sorted_keys,
                                    int cap_per_image, float4* __restrict__ sbox_all) {
  const int img = blockIdx.y;
  const int n = min(seg_end[img] - seg_begin[img], cap_per_image);
  const float* d = dets + (size_t)seg_begin[img] * 5;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const unsigned r = (unsigned)(sorted_keys[(size_t)img * cap_per_image + j] & 0xffffffffu);
    sbox_all[(size_t)img * cap_per_image + j] =
        make_float4(d[(size_t)r * 5], d[(size_t)r * 5 + 1], d[(size_t)r * 5 + 2], d[(size_t)r * 5 + 3]);
  }
}

__global__ void __launch_bounds__(64) iou_mask_kernel(const float4* __restrict__ sbox_all, const int* __restrict__ seg_begin,
                                                      const int* __restrict__ seg_end, int cap_per_image, int mask_rows,
                                                      int mask_words, unsigned long long* __restrict__ mask_all,
                                                      double thr, int mode, int num_images) {
  __shared__ float4 cb[64];
  __shared__ float ca[64];
  const int tid = threadIdx.x;
  // persistent walk: the tiles (row block rb, column block cbk >= rb) of an image are numbered row block by row block;
  // CTA c takes tiles c, c + gridDim.x, ... of the concatenation over images (all loop bounds are CTA-uniform)
  long long g = blockIdx.x;
  for (int img = 0; img < num_images; ++img) {
    const int n = min(seg_end[img] - seg_begin[img], cap_per_image);
    const int nb = (n <= mask_rows) ? (n + 63) >> 6 : 0;
    const long long tiles = (long long)nb * (nb + 1) / 2;
    const float4* sb = sbox_all + (size_t)img * cap_per_image;
    for (; g < tiles; g += gridDim.x) {
      // row block rb starts at offset rb * nb - rb (rb - 1) / 2
      int rb = (int)(((2.0f * nb + 1.0f) - sqrtf((2.0f * nb + 1.0f) * (2.0f * nb + 1.0f) - 8.0f * (float)g)) * 0.5f);
      rb = max(0, min(rb, nb - 1));
      while (rb > 0 && (long long)rb * nb - (long long)rb * (rb - 1) / 2 > g) --rb;
      while ((long long)(rb + 1) * nb - (long long)(rb + 1) * rb / 2 <= g) ++rb;
      const int cbk = rb + (int)(g - ((long long)rb * nb - (long long)rb * (rb - 1) / 2));
      const int j0 = cbk * 64;
      __syncthreads();
      if (j0 + tid < n) { cb[tid] = sb[j0 + tid]; ca[tid] = box_area(cb[tid]); }
      __syncthreads();
      const int i = rb * 64 + tid;
      if (i < n) {
        const float4 bi = sb[i];
        const float ai = box_area(bi);
        unsigned long long bits = 0ull;
        const int cmax = min(64, n - j0);
        const int c0 = (cbk == rb) ? tid + 1 : 0;            // strictly later rows only
        for (int c = c0; c < cmax; ++c)
          if (overlaps(box_iou(bi, ai, cb[c], ca[c]), thr, mode)) bits |= 1ull << c;
        mask_all[((size_t)img * mask_rows + i) * mask_words + cbk] = bits;
      }
    }
    g -= tiles;
  }
}

constexpr int kMaskSweepThreads = 256;       // one thread per 64-bit word of the alive set: mask_rows <= 64 * 256

template <bool VOTE>
__global__ void __launch_bounds__(kMaskSweepThreads) mask_sweep_kernel(
    const int* __restrict__ seg_begin, const int* __restrict__ seg_end, int cap_per_image, int mask_rows, int mask_words,
    unsigned long long* __restrict__ mask_all, int* __restrict__ heads_all, int* __restrict__ head_count) {
  const int img = blockIdx.x;
  const int n = min(seg_end[img] - seg_begin[img], cap_per_image);
  if (n > mask_rows) return;                                  // greedy_sweep_kernel owns this image
  unsigned long long* M = mask_all + (size_t)img * mask_rows * mask_words;
  int* heads = heads_all + (size_t)img * mask_rows;
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_alive;
  __shared__ int s_heads;
  const int tid = threadIdx.x;
  const int nb = (n + 63) >> 6;
  // thread w owns the alive bits of rows 64 w .. 64 w + 63
  unsigned long long alive = 0ull;
  if (tid < nb) alive = (tid == nb - 1 && (n & 63)) ? ((1ull << (n & 63)) - 1ull) : ~0ull;
  if (tid == 0) s_heads = 0;
  for (int b = 0; b < nb; ++b) {
    __syncthreads();
    if (tid == b) s_alive = alive;
    if (tid < 64 && b * 64 + tid < n) s_diag[tid] = M[(size_t)(b * 64 + tid) * mask_words + b];
    __syncthreads();
    // every thread resolves the diagonal tile redundantly (uniform control flow, register bit masks)
    unsigned long long cur = s_alive, hbits = 0ull;
    const int hbase = s_heads;
    while (cur) {
      const int r = __ffsll((long long)cur) - 1;
      hbits |= 1ull << r;
      const unsigned long long mem = s_diag[r] & cur;        // later rows of this block that join head r
      if (VOTE && tid == 0) M[(size_t)(b * 64 + r) * mask_words + b] = mem;
      cur &= ~(mem | (1ull << r));
    }
    if (tid == b) alive = 0ull;                               // every row of the block is now a head or a member
    // heads of this block, in order, applied to this thread's word (w > b)
    const int nh = __popcll(hbits);
    if (tid > b && tid < nb) {
      unsigned long long hb = hbits;
      while (hb) {
        // up to 32 independent loads in flight (the step is L2-latency bound: one round trip per batch), then the
        // order-dependent bit logic
        constexpr int kBatch = 32;
        unsigned long long rows[kBatch];
        unsigned long long todo = hb;
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
          if (todo) {
            const int r = __ffsll((long long)todo) - 1;
            todo &= todo - 1;
            rows[q] = M[(size_t)(b * 64 + r) * mask_words + tid];
          }
        }
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
          if (hb) {
            const int r = __ffsll((long long)hb) - 1;
            hb &= hb - 1;
            const unsigned long long mem = rows[q] & alive;
            if (VOTE && mem != rows[q]) M[(size_t)(b * 64 + r) * mask_words + tid] = mem;
            alive &= ~mem;
          }
        }
      }
    }
    if (tid < 64 && (hbits >> tid) & 1ull)
      heads[hbase + __popcll(hbits & ((1ull << tid) - 1ull))] = b * 64 + tid;
    __syncthreads();
    if (tid == 0) s_heads = hbase + nh;
  }
  __syncthreads();
  if (tid == 0) head_count[img] = s_heads;
}

// NMS: kept rows = heads; emit their ORIGINAL row indices in kept order
__global__ void nms_emit_kernel(const int* __restrict__ seg_begin, const int* __restrict__ seg_end, int cap_per_image,
                                int mask_rows, const unsigned long long* __restrict__ sorted_keys,
                                const int* __restrict__ heads_all, const int* __restrict__ head_count,
                                int* __restrict__ out_idx, int* __restrict__ out_count, int out_cap) {
  const int img = blockIdx.y;
  const int n = min(seg_end[img] - seg_begin[img], cap_per_image);
  if (n > mask_rows) return;
  const int nh = head_count[img];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nh && k < out_cap; k += gridDim.x * blockDim.x)
    out_idx[(size_t)img * out_cap + k] =
        (int)(sorted_keys[(size_t)img * cap_per_image + heads_all[(size_t)img * mask_rows + k]] & 0xffffffffu);
  if (blockIdx.x == 0 && threadIdx.x == 0) out_count[img] = nh;
}

constexpr int kVoteThreads = 1024;

__global__ void __launch_bounds__(kVoteThreads) vote_reduce_kernel(
    const float* __restrict__ dets, const int* __restrict__ seg_begin, const int* __restrict__ seg_end, int cap_per_image,
    int mask_rows, int mask_words, const unsigned long long* __restrict__ sorted_keys, const float4* __restrict__ sbox_all,
    const unsigned long long* __restrict__ mask_all, const int* __restrict__ heads_all, const int* __restrict__ head_count,
    float* __restrict__ stat_all, float* __restrict__ out_dets, int* __restrict__ out_count, int out_cap) {
  const int img = blockIdx.x;
  const int n = min(seg_end[img] - seg_begin[img], cap_per_image);
  if (n > mask_rows) return;
  float* odet = out_dets + (size_t)img * out_cap * 5;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (n == 0) {                 // lib/test.py:184-186: empty input -> [[10, 10, 20, 20, 0.0001]]
    if (tid == 0) {
      odet[0] = 10.f; odet[1] = 10.f; odet[2] = 20.f; odet[3] = 20.f; odet[4] = 0.0001f;
      out_count[img] = 1;
    }
    return;
  }
  const float* d = dets + (size_t)seg_begin[img] * 5;
  const unsigned long long* keys = sorted_keys + (size_t)img * cap_per_image;
  const float4* sbox = sbox_all + (size_t)img * cap_per_image;
  const unsigned long long* M = mask_all + (size_t)img * mask_rows * mask_words;
  const int* heads = heads_all + (size_t)img * mask_rows;
  float* stat = stat_all + (size_t)img * mask_rows * 6;      // per head: x1, y1, x2, y2, score, emit flag
  const int nh = head_count[img];
  const int nb = (n + 63) >> 6;
  // ---- phase A: one warp per head ----
  for (int k = wid; k < nh; k += kVoteThreads / 32) {
    const int i = heads[k];
    const float4 bi = sbox[i];
    const float si = d[(size_t)(keys[i] & 0xffffffffu) * 5 + 4];
    double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0, ss = 0;
    float smax = -1.f;
    int members = 0;
    if (lane == 0) {             // the head itself
      sx1 = (double)__fmul_rn(bi.x, si); sy1 = (double)__fmul_rn(bi.y, si);
      sx2 = (double)__fmul_rn(bi.z, si); sy2 = (double)__fmul_rn(bi.w, si);
      ss = (double)si; smax = si; members = 1;
    }
    for (int w = (i >> 6) + lane; w < nb; w += 32) {
      unsigned long long bits = M[(size_t)i * mask_words + w];
      while (bits) {
        const int c = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        const int j = w * 64 + c;
        const float4 bj = sbox[j];
        const float sj = d[(size_t)(keys[j] & 0xffffffffu) * 5 + 4];
        // det_accu[:, 0:4] * score in float32 (lib/test.py:207), then summed
        sx1 += (double)__fmul_rn(bj.x, sj); sy1 += (double)__fmul_rn(bj.y, sj);
        sx2 += (double)__fmul_rn(bj.z, sj); sy2 += (double)__fmul_rn(bj.w, sj);
        ss += (double)sj;
        smax = fmaxf(smax, sj);
        ++members;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      members += __shfl_xor_sync(0xffffffffu, members, o);
      sx1 += __shfl_xor_sync(0xffffffffu, sx1, o); sy1 += __shfl_xor_sync(0xffffffffu, sy1, o);
      sx2 += __shfl_xor_sync(0xffffffffu, sx2, o); sy2 += __shfl_xor_sync(0xffffffffu, sy2, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    }
    if (lane == 0) {
      float* st = stat + (size_t)k * 6;
      if (members <= 1) {         // singleton: emitted only when it is the last head (nothing remains after it)
        st[0] = bi.x; st[1] = bi.y; st[2] = bi.z; st[3] = bi.w; st[4] = si;
        st[5] = (k == nh - 1) ? 1.f : 0.f;
      } else {
        const float fs = (float)ss;
        st[0] = __fdiv_rn((float)sx1, fs); st[1] = __fdiv_rn((float)sy1, fs);
        st[2] = __fdiv_rn((float)sx2, fs); st[3] = __fdiv_rn((float)sy2, fs);
        st[4] = smax;
        st[5] = 1.f;
      }
    }
  }
  __syncthreads();
  // ---- phase B: emission slots = exclusive scan of the emit flags, in head order ----
  __shared__ int s_warp[kVoteThreads / 32];
  __shared__ int s_base;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int k0 = 0; k0 < nh; k0 += kVoteThreads) {
    const int k = k0 + tid;
    const int flag = (k < nh && stat[(size_t)k * 6 + 5] != 0.f) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int v = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    const int chunk_total = __shfl_sync(0xffffffffu, v, 31);
    const int e = s_base + __shfl_sync(0xffffffffu, v, wid) - s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
    if (flag && e < out_cap) {
#pragma unroll
      for (int q = 0; q < 5; ++q) odet[(size_t)e * 5 + q] = stat[(size_t)k * 6 + q];
    }
    __syncthreads();
    if (tid == 0) s_base += chunk_total;
    __syncthreads();
  }
  if (tid == 0) out_count[img] = s_base;
}

// ---------------------------------------------------------------------------------------------------
// lib/utils/bbox.pyx: kind 0 = IoU (:14-54), 1 = IoA with zeroed diagonal (:56-102), 2 = itself (:106-142)
// ---------------------------------------------------------------------------------------------------
__global__ void bbox_overlaps_kernel(const double* __restrict__ boxes, const double* __restrict__ query, int N, int K,
                                     int kind, double* __restrict__ out) {
  const long long total = (long long)N * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K), n = (int)(i / K);
    const double* b = boxes + (size_t)n * 4;
    const double* q = query + (size_t)k * 4;
    double r = 0.0;
    const double iw = __dadd_rn(__dsub_rn(fmin(b[2], q[2]), fmax(b[0], q[0])), 1.0);
    if (iw > 0) {
      const double ih = __dadd_rn(__dsub_rn(fmin(b[3], q[3]), fmax(b[1], q[1])), 1.0);
      if (ih > 0) {
        const double barea = __dmul_rn(__dadd_rn(__dsub_rn(b[2], b[0]), 1.0), __dadd_rn(__dsub_rn(b[3], b[1]), 1.0));
        const double qarea = __dmul_rn(__dadd_rn(__dsub_rn(q[2], q[0]), 1.0), __dadd_rn(__dsub_rn(q[3], q[1]), 1.0));
        const double inter = __dmul_rn(iw, ih);
        if (kind == 0) r = __ddiv_rn(inter, __dsub_rn(__dadd_rn(barea, qarea), inter));
        else r = __ddiv_rn(inter, barea);
        if (kind == 1 && n == k) r = 0.0;
      }
    }
    out[i] = r;
  }
}

}  // namespace

// =================================================================================================
// C ABI (declared in include/shf_b200.h)
// =================================================================================================
extern "C" int shf_head_decode(const void* const* feat_h2, long long feat_plane_stride, int num_anchors,
                               const float* w_cls, const float* b_cls, const float* w_box, const float* b_box,
                               const float* base_anchors, int H, int W, int C, int feat_stride, float im_h, float im_w, float min_size, float score_thresh, float* prob,
                               float* delta, float* boxes, unsigned long long* keys, int* count,
                               unsigned long long* best_key, void* stream) {
  SHF_REQUIRE(num_anchors >= 1 && num_anchors <= kMaxAnchors, "shf_head_decode: %d anchors (max %d)", num_anchors,
              kMaxAnchors);
  SHF_REQUIRE(C % 8 == 0, "shf_head_decode: C=%d must be a multiple of 8", C);
  HeadParams p;
  for (int a = 0; a < num_anchors; ++a) {
    p.feat[a] = (const __half*)feat_h2[a];
    for (int k = 0; k < 4; ++k) p.anchors[a][k] = base_anchors[a * 4 + k];
  }
  p.wc = w_cls; p.bc = b_cls; p.wb = w_box; p.bb = b_box;
  p.A = num_anchors; p.H = H; p.W = W; p.C = C; p.feat_stride = feat_stride;
  p.plane_stride = feat_plane_stride > 0 ? feat_plane_stride : (long long)H * W * C;
  p.image_stride = 0; p.batched = 0;
  p.im_h = im_h; p.im_w = im_w; p.min_size = min_size; p.score_thresh = score_thresh;
  cudaStream_t st = (cudaStream_t)stream;
  SHF_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
  SHF_CUDA_CHECK(cudaMemsetAsync(best_key, 0xff, sizeof(unsigned long long), st));
  const int n = H * W * num_anchors;
  const size_t smem = (size_t)num_anchors * 6 * (C + 1) * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    SHF_CUDA_CHECK(cudaFuncSetAttribute(head_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  head_decode_kernel<<<(n + 127) / 128, 128, smem, st>>>(p, prob, delta, boxes, keys, count, best_key);
  SHF_LAUNCH_CHECK();
  return 0;
}

// Batched head decode for the pyramid driver: `num_images` images of one level batch in a single launch.
// Outputs are per image ([img][...]) and the sort keys carry the image index in their top bits (see make_key_img),
// so a single shf_sort_keys over num_images * n keys orders every image's rows.
extern "C" int shf_head_decode_batched(const void* const* feat_h2, long long feat_plane_stride, long long feat_image_stride,
                                       int num_images, int num_anchors, const float* w_cls, const float* b_cls,
                                       const float* w_box, const float* b_box, const float* base_anchors, int H, int W,
                                       int C, int feat_stride, float im_h, float im_w, float min_size, float score_thresh,
                                       float* prob, float* delta, float* boxes, unsigned long long* keys, int* count,
                                       unsigned long long* best_key, void* stream) {
  SHF_REQUIRE(num_anchors >= 1 && num_anchors <= kMaxAnchors, "shf_head_decode_batched: %d anchors (max %d)", num_anchors,
              kMaxAnchors);
  SHF_REQUIRE(num_images >= 1 && num_images <= 32, "shf_head_decode_batched: %d images (1..32 per launch)", num_images);
  SHF_REQUIRE(C % 8 == 0 && (long long)H * W * num_anchors < (1ll << kRowBits), "shf_head_decode_batched: bad geometry");
  HeadParams p;
  for (int a = 0; a < num_anchors; ++a) {
    p.feat[a] = (const __half*)feat_h2[a];
    for (int k = 0; k < 4; ++k) p.anchors[a][k] = base_anchors[a * 4 + k];
  }
  p.wc = w_cls; p.bc = b_cls; p.wb = w_box; p.bb = b_box;
  p.A = num_anchors; p.H = H; p.W = W; p.C = C; p.feat_stride = feat_stride;
  p.plane_stride = feat_plane_stride; p.image_stride = feat_image_stride; p.batched = 1;
  p.im_h = im_h; p.im_w = im_w; p.min_size = min_size; p.score_thresh = score_thresh;
  cudaStream_t st = (cudaStream_t)stream;
  SHF_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int) * num_images, st));
  SHF_CUDA_CHECK(cudaMemsetAsync(best_key, 0xff, sizeof(unsigned long long) * num_images, st));
  const int n = H * W * num_anchors;
  const size_t smem = (size_t)num_anchors * 6 * (C + 1) * sizeof(float);
  SHF_REQUIRE(smem <= 48 * 1024, "shf_head_decode_batched: head weights do not fit the default shared memory");
  dim3 grid((n + 127) / 128, num_images);
  head_decode_kernel<<<grid, 128, smem, st>>>(p, prob, delta, boxes, keys, count, best_key);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_gather_dets_batched(const unsigned long long* sorted_keys, const int* count,
                                       const unsigned long long* best_key, const float* prob, const float* boxes,
                                       int num_anchors, int hw, int topn, int num_images, int passes_per_image,
                                       float* dets, int* pass_offsets, int image_base, int passes_total, int pass_base,
                                       int det_cap, float level_w, float im_scale, float det_thresh, void* stream) {
  SHF_REQUIRE(passes_per_image == 1 || passes_per_image == 2, "shf_gather_dets_batched: %d passes per image", passes_per_image);
  GatherBatchParams g;
  g.sorted = sorted_keys; g.count = count; g.best_key = best_key; g.prob = prob; g.boxes = boxes;
  g.A = num_anchors; g.hw = hw; g.n = hw * num_anchors; g.topn = topn; g.nf = passes_per_image;
  g.dets = dets; g.pass_offsets = pass_offsets; g.image_base = image_base; g.passes_total = passes_total;
  g.pass_base = pass_base; g.det_cap = det_cap; g.level_w = level_w; g.im_scale = im_scale; g.det_thresh = det_thresh;
  dim3 grid(8, num_images);
  gather_dets_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" long long shf_sort_keys_workspace(int total_keys) {
  return (long long)sizeof(unsigned long long) * (total_keys > 0 ? total_keys : 1);      // the ping-pong buffer
}

extern "C" int shf_sort_keys(const unsigned long long* keys_in, unsigned long long* keys_out, int num_segments,
                             int segment_stride, const int* segment_len, int fixed_len, int begin_bit, void* workspace,
                             long long workspace_bytes, void* stream) {
  SHF_REQUIRE(begin_bit >= 0 && begin_bit <= 32, "shf_sort_keys: begin_bit %d (the 32 bits above it are sorted)", begin_bit);
  SHF_REQUIRE(num_segments >= 1 && fixed_len >= 0 && fixed_len <= segment_stride, "shf_sort_keys: bad segment geometry");
  SHF_REQUIRE(keys_in != keys_out, "shf_sort_keys: in-place sorting is not supported");
  SHF_REQUIRE(workspace_bytes >= shf_sort_keys_workspace(num_segments * segment_stride), "shf_sort_keys: workspace too small");
  segmented_sort_kernel<<<num_segments, kSortThreads, 0, (cudaStream_t)stream>>>(
      keys_in, keys_out, (unsigned long long*)workspace, nullptr, segment_len, fixed_len, (long long)segment_stride,
      begin_bit, nullptr);
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_proposal_gather(const unsigned long long* sorted_keys, const int* count,
                                   const unsigned long long* best_key, const float* prob, const float* boxes,
                                   int num_anchors, int hw, int topn, float* out_boxes, float* out_probs, int* out_rows,
                                   float* dets, int* pass_offsets, int pass, int det_cap, int flip, float level_w,
                                   float im_scale, float det_thresh, void* stream) {
  GatherParams g;
  g.sorted = sorted_keys; g.count = count; g.best_key = best_key; g.prob = prob; g.boxes = boxes;
  g.A = num_anchors; g.hw = hw; g.topn = topn;
  g.out_boxes = out_boxes; g.out_probs = out_probs; g.out_rows = out_rows;
  g.dets = dets; g.pass_offsets = pass_offsets; g.pass = pass; g.det_cap = det_cap;
  g.flip = flip; g.level_w = level_w; g.im_scale = im_scale; g.det_thresh = det_thresh;
  const int blocks = (topn + 255) / 256;
  proposal_gather_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, (cudaStream_t)stream>>>(g);
  SHF_LAUNCH_CHECK();
  return 0;
}

// Workspace layout of shf_postprocess (per image unless noted)
constexpr int kMaskRowsMax = 16384;            // images with more rows take the serial sweep
struct PostWs {
  unsigned long long *keys_in, *keys_out, *sort_tmp, *mask;
  float4* sbox;
  unsigned char* alive;
  int *heads, *head_count;
  float* stat;
  int mask_rows, mask_words;
  size_t bytes;
};
static PostWs post_layout(void* base, int num_images, int cap) {
  PostWs w;
  w.mask_rows = cap < kMaskRowsMax ? ((cap + 63) & ~63) : kMaskRowsMax;
  w.mask_words = w.mask_rows / 64;
  const size_t n = (size_t)num_images * cap;
  uint8_t* p = (uint8_t*)base;
  auto take = [&](size_t bytes) { uint8_t* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  w.keys_in = (unsigned long long*)take(n * 8);
  w.keys_out = (unsigned long long*)take(n * 8);
  w.sort_tmp = (unsigned long long*)take(n * 8);
  w.sbox = (float4*)take(n * 16);
  w.alive = (unsigned char*)take(n);
  w.mask = (unsigned long long*)take((size_t)num_images * w.mask_rows * w.mask_words * 8);
  w.heads = (int*)take((size_t)num_images * w.mask_rows * 4);
  w.head_count = (int*)take((size_t)num_images * 4);
  w.stat = (float*)take((size_t)num_images * w.mask_rows * 6 * 4);
  w.bytes = (size_t)(p - (uint8_t*)base);
  return w;
}

extern "C" long long shf_postprocess_workspace(int num_images, int cap_per_image) {
  return (long long)post_layout(nullptr, num_images, cap_per_image).bytes + 256;
}

// method 0 = NMS (out_idx), 1 = box voting (out_dets).  mode: see overlaps().
extern "C" int shf_postprocess(const float* dets, const int* seg_begin, const int* seg_end, int num_images,
                               int cap_per_image, double thresh, int method, int mode, int* out_idx, float* out_dets,
                               int* out_count, int out_cap, void* workspace, long long workspace_bytes, void* stream) {
  SHF_REQUIRE(num_images >= 1 && cap_per_image >= 1, "shf_postprocess: bad sizes");
  SHF_REQUIRE(workspace_bytes >= shf_postprocess_workspace(num_images, cap_per_image),
              "shf_postprocess: workspace too small");
  SHF_REQUIRE(method == 0 || method == 1, "shf_postprocess: method %d", method);
  if (method == 0) SHF_REQUIRE(out_idx != nullptr, "shf_postprocess: NMS needs out_idx");
  else SHF_REQUIRE(out_dets != nullptr, "shf_postprocess: voting needs out_dets");
  cudaStream_t st = (cudaStream_t)stream;
  const PostWs w = post_layout((void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255), num_images, cap_per_image);
  const int cmp = method == 1 ? 2 : mode;                    // bbox_vote compares in float32 with >= (lib/test.py:198)
  dim3 kg(std::min((cap_per_image + 255) / 256, 64), num_images);
  det_keys_kernel<<<kg, 256, 0, st>>>(dets, seg_begin, seg_end, cap_per_image, w.keys_in);
  SHF_LAUNCH_CHECK();
  // (keys are written in row order: the stable sort of the 32 score bits leaves ties in row order)
  segmented_sort_kernel<<<num_images, kSortThreads, 0, st>>>(w.keys_in, w.keys_out, w.sort_tmp, seg_begin, seg_end,
                                                            cap_per_image, (long long)cap_per_image, 32, nullptr);
  SHF_LAUNCH_CHECK();
  sorted_boxes_kernel<<<kg, 256, 0, st>>>(dets, seg_begin, seg_end, w.keys_out, cap_per_image, w.sbox);
  SHF_LAUNCH_CHECK();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  iou_mask_kernel<<<sms * 16, 64, 0, st>>>(w.sbox, seg_begin, seg_end, cap_per_image, w.mask_rows, w.mask_words, w.mask,
                                           thresh, cmp, num_images);
  SHF_LAUNCH_CHECK();
  if (method == 0) {
    mask_sweep_kernel<false><<<num_images, kMaskSweepThreads, 0, st>>>(seg_begin, seg_end, cap_per_image, w.mask_rows,
                                                                      w.mask_words, w.mask, w.heads, w.head_count);
    SHF_LAUNCH_CHECK();
    dim3 eg(std::min((w.mask_rows + 255) / 256, 16), num_images);
    nms_emit_kernel<<<eg, 256, 0, st>>>(seg_begin, seg_end, cap_per_image, w.mask_rows, w.keys_out, w.heads, w.head_count,
                                        out_idx, out_count, out_cap);
    SHF_LAUNCH_CHECK();
    if (cap_per_image > w.mask_rows)
      greedy_sweep_kernel<false><<<num_images, kSweepThreads, 0, st>>>(dets, seg_begin, seg_end, w.keys_out, cap_per_image,
                                                                      w.alive, w.sbox, thresh, cmp, out_idx, nullptr,
                                                                      out_count, out_cap, w.mask_rows);
  } else {
    mask_sweep_kernel<true><<<num_images, kMaskSweepThreads, 0, st>>>(seg_begin, seg_end, cap_per_image, w.mask_rows,
                                                                     w.mask_words, w.mask, w.heads, w.head_count);
    SHF_LAUNCH_CHECK();
    vote_reduce_kernel<<<num_images, kVoteThreads, 0, st>>>(dets, seg_begin, seg_end, cap_per_image, w.mask_rows,
                                                            w.mask_words, w.keys_out, w.sbox, w.mask, w.heads,
                                                            w.head_count, w.stat, out_dets, out_count, out_cap);
    SHF_LAUNCH_CHECK();
    if (cap_per_image > w.mask_rows)
      greedy_sweep_kernel<true><<<num_images, kSweepThreads, 0, st>>>(dets, seg_begin, seg_end, w.keys_out, cap_per_image,
                                                                     w.alive, w.sbox, thresh, cmp, nullptr, out_dets,
                                                                     out_count, out_cap, w.mask_rows);
  }
  SHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int shf_bbox_overlaps(const double* boxes, const double* query, int n, int k, int kind, double* out,
                                 void* stream) {
  if (n == 0 || k == 0) return 0;
  const long long total = (long long)n * k;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  bbox_overlaps_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(boxes, query, n, k, kind, out);
  SHF_LAUNCH_CHECK();
  return 0;
}

// Detection tail on the device (sm_100a SIMT kernels; HBM/latency-bound integer + fp32 work):
//   head_decode   : the 1x1 cls/bbox convs + 2-way softmax + anchor decode + clip + candidate keys
//                   (replaces base_conv_layer.cpp:255-279 for the 2/4/6/12-channel heads, concat_layer.cpp,
//                    reshape_layer.cpp, softmax_layer.cpp:27-60 and lib/layers/proposal_layer.py:96-173)
//   sort          : (score desc, anchor index asc) 64-bit radix sort  (proposal_layer.py:180-190)
//   gather        : top-K rows -> 'boxes' (R,5) / 'cls_prob' (R,2) blobs, plus the per-pass un-mirror,
//                   unscale and 0.05 threshold of lib/test.py:52-66,163-167 appended to the image's det list
//   nms / vote    : greedy NMS (lib/nms/cpu_nms.pyx:17-68, nms_kernel.cu:45-155 semantics selectable) and
//                   box voting (lib/test.py:181-217) -- one CTA per image, batched
//   bbox_overlaps : IoU / IoA / self-overlap matrices (lib/utils/bbox.pyx:14-142), float64
// All fp32 box arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) so that IoU
// values, and therefore keep/suppress decisions, are bit-identical to the NumPy/Cython reference.
#include <cub/device/device_radix_sort.cuh>

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

SHF_DEVICE unsigned long long make_key(float score, unsigned idx) {
  return ((unsigned long long)(~__float_as_uint(score)) << 32) | idx;     // ascending key = descending score, then index
}
// Batched variant: [image : 5][~score bits : 32][row : 27] -- one device-wide sort leaves every image's rows
// contiguous (each image contributes exactly n keys, sentinels included) and ordered like make_key.
constexpr int kRowBits = 27;
SHF_DEVICE unsigned long long make_key_img(unsigned img, float score, unsigned idx) {
  return ((unsigned long long)img << 59) | ((unsigned long long)(~__float_as_uint(score)) << kRowBits) | idx;
}
SHF_DEVICE unsigned long long sentinel_img(unsigned img) { return ((unsigned long long)img << 59) | ((1ull << 59) - 1); }
SHF_DEVICE float key_img_score(unsigned long long k) { return __uint_as_float(~(unsigned)((k >> kRowBits) & 0xffffffffu)); }
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
      const float s = __uint_as_float(~(unsigned)(k >> 32));
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
// image-level greedy sweeps
// ---------------------------------------------------------------------------------------------------
__global__ void det_keys_kernel(const float* __restrict__ dets, const int* __restrict__ seg_begin,
                                const int* __restrict__ seg_end, int cap_per_image, unsigned long long* __restrict__ keys) {
  // keys[img][cap]: (score desc, row index asc); rows beyond the image's count get the ~0 sentinel
  const int img = blockIdx.y;
  const int n = seg_end[img] - seg_begin[img];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cap_per_image; j += gridDim.x * blockDim.x)
    keys[(size_t)img * cap_per_image + j] =
        j < n ? make_key(dets[(size_t)(seg_begin[img] + j) * 5 + 4], (unsigned)j) : ~0ull;
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
    int* __restrict__ out_count, int out_cap) {
  const int img = blockIdx.x;
  const int n = min(seg_end[img] - seg_begin[img], cap_per_image);
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
  for (int j = tid; j < n; j += kSweepThreads) {
    const unsigned r = (unsigned)(keys[j] & 0xffffffffu);
    sbox[j] = make_float4(d[(size_t)r * 5], d[(size_t)r * 5 + 1], d[(size_t)r * 5 + 2], d[(size_t)r * 5 + 3]);
    alive[j] = 1;
  }
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
  if (tid == 0) out_count[img] = min(s_emit, out_cap);
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

extern "C" long long shf_sort_keys_workspace(int n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, n);
  return (long long)bytes;
}

extern "C" int shf_sort_keys(const unsigned long long* keys_in, unsigned long long* keys_out, int n, int begin_bit,
                             void* workspace, long long workspace_bytes, void* stream) {
  SHF_REQUIRE(begin_bit >= 0 && begin_bit < 64, "shf_sort_keys: begin_bit %d", begin_bit);
  size_t bytes = (size_t)workspace_bytes;
  // LSD radix sort is stable: keys that arrive in ascending order of their low `begin_bit` bits (the row index) need
  // only the bits above them sorted -- 4-5 passes instead of 8
  SHF_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(workspace, bytes, keys_in, keys_out, n, begin_bit, 64, (cudaStream_t)stream));
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

extern "C" long long shf_postprocess_workspace(int num_images, int cap_per_image) {
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                 cap_per_image);
  const size_t n = (size_t)num_images * cap_per_image;
  // keys_in + keys_out + sorted boxes + alive flags + cub temp (per image, reused)
  return (long long)(n * 8 * 2 + n * 16 + ((n + 255) & ~(size_t)255) + ((sort_bytes + 255) & ~(size_t)255) + 1024);
}

// method 0 = NMS (out_idx), 1 = box voting (out_dets).  mode: see overlaps().
extern "C" int shf_postprocess(const float* dets, const int* seg_begin, const int* seg_end, int num_images,
                               int cap_per_image, double thresh, int method, int mode, int* out_idx, float* out_dets,
                               int* out_count, int out_cap, void* workspace, long long workspace_bytes, void* stream) {
  SHF_REQUIRE(num_images >= 1 && cap_per_image >= 1, "shf_postprocess: bad sizes");
  SHF_REQUIRE(workspace_bytes >= shf_postprocess_workspace(num_images, cap_per_image),
              "shf_postprocess: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)num_images * cap_per_image;
  uint8_t* ws = (uint8_t*)workspace;
  unsigned long long* keys_in = (unsigned long long*)ws;
  unsigned long long* keys_out = keys_in + n;
  float4* sbox = (float4*)(keys_out + n);
  unsigned char* alive = (unsigned char*)(sbox + n);
  uint8_t* cub_tmp = (uint8_t*)(((uintptr_t)(alive + n) + 255) & ~(uintptr_t)255);
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, keys_in, keys_out, cap_per_image);
  dim3 kg((cap_per_image + 255) / 256, num_images);
  det_keys_kernel<<<kg, 256, 0, st>>>(dets, seg_begin, seg_end, cap_per_image, keys_in);
  SHF_LAUNCH_CHECK();
  for (int i = 0; i < num_images; ++i) {
    size_t b = sort_bytes;
    SHF_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(cub_tmp, b, keys_in + (size_t)i * cap_per_image,
                                                  keys_out + (size_t)i * cap_per_image, cap_per_image, 32, 64, st));
    // (keys are written in row order, so the stable sort only needs the 32 score bits)
  }
  if (method == 0) {
    SHF_REQUIRE(out_idx != nullptr, "shf_postprocess: NMS needs out_idx");
    greedy_sweep_kernel<false><<<num_images, kSweepThreads, 0, st>>>(dets, seg_begin, seg_end, keys_out, cap_per_image,
                                                                    alive, sbox, thresh, mode, out_idx, nullptr,
                                                                    out_count, out_cap);
  } else {
    SHF_REQUIRE(out_dets != nullptr, "shf_postprocess: voting needs out_dets");
    greedy_sweep_kernel<true><<<num_images, kSweepThreads, 0, st>>>(dets, seg_begin, seg_end, keys_out, cap_per_image,
                                                                   alive, sbox, thresh, 2, nullptr, out_dets, out_count,
                                                                   out_cap);
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

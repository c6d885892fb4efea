// C-ABI glue: error reporting, library info and the reference's own `_nms` entry point
// (lib/nms/gpu_nms.hpp:1-2), kept symbol-for-symbol so the reference's Cython wrapper
// (lib/nms/gpu_nms.pyx:11-30) could bind this library unchanged.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void shf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* shf_last_error(void) { return g_err; }

extern "C" int shf_abi_version(void) { return 8; }

extern "C" int shf_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, long long* total_mem) {
  cudaDeviceProp prop;
  SHF_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  *sm_count = prop.multiProcessorCount;
  *cc_major = prop.major;
  *cc_minor = prop.minor;
  *total_mem = (long long)prop.totalGlobalMem;
  return 0;
}

extern "C" long long shf_postprocess_workspace(int num_images, int cap_per_image);
extern "C" int shf_postprocess(const float* dets, const int* seg_begin, const int* seg_end, int num_images,
                               int cap_per_image, double thresh, int method, int mode, int* out_idx, float* out_dets,
                               int* out_count, int out_cap, void* workspace, long long workspace_bytes, void* stream);

// Host-pointer greedy NMS with the reference GPU module's exact contract:
//   boxes_host (boxes_num x boxes_dim floats, [x1,y1,x2,y2,score,...]) ALREADY sorted by descending score,
//   keep_out (caller-allocated, boxes_num ints) receives kept row indices, *num_out their count.
//   Suppression rule is the reference kernel's `IoU > thresh` in float32 (nms_kernel.cu:82).
// Differences from the reference: CUDA errors are reported (shf_last_error / *num_out = -1) instead of
// being printed and ignored (nms_kernel.cu:12-19), the greedy sweep runs on the device, and the device / pinned
// staging buffers are kept between calls (the reference cudaMallocs and cudaFrees three buffers per call,
// nms_kernel.cu:101-143) -- grow-only, one set per device, so the entry points are not re-entrant across threads.
struct HostPostBuffers {
  float* d_dets = nullptr;     // [n][5]
  int* d_meta = nullptr;       // seg_begin, seg_end, count, pad, then out_idx[n]
  float* d_out = nullptr;      // [n/2+1][5] voted boxes
  void* d_ws = nullptr;
  float* h_pin = nullptr;      // pinned staging: max(n * 5 floats in, results out)
  int cap = 0;
  long long ws_bytes = 0;
};
static HostPostBuffers g_host_post[64];

static int host_post_reserve(HostPostBuffers& b, int n) {
  if (n <= b.cap) return 0;
  cudaFree(b.d_dets); cudaFree(b.d_meta); cudaFree(b.d_out); cudaFree(b.d_ws); cudaFreeHost(b.h_pin);
  b = HostPostBuffers();
  const int cap = n + n / 2 + 64;
  b.ws_bytes = shf_postprocess_workspace(1, cap);
  SHF_CUDA_CHECK(cudaMalloc(&b.d_dets, (size_t)cap * 5 * sizeof(float)));
  SHF_CUDA_CHECK(cudaMalloc(&b.d_meta, (size_t)(cap + 4) * sizeof(int)));
  SHF_CUDA_CHECK(cudaMalloc(&b.d_out, (size_t)(cap / 2 + 1) * 5 * sizeof(float)));
  SHF_CUDA_CHECK(cudaMalloc(&b.d_ws, (size_t)b.ws_bytes));
  SHF_CUDA_CHECK(cudaMallocHost(&b.h_pin, (size_t)cap * 5 * sizeof(float)));
  b.cap = cap;
  return 0;
}

// method 0: NMS -> keep_out / *num_out;  method 1: box voting -> out_dets (n/2+1 rows of 5 floats max) / *num_out
static int post_host_impl(int method, int* keep_out, float* out_dets, int* num_out, const float* boxes_host, int boxes_num,
                          int boxes_dim, double thresh, int mode, int device_id) {
  *num_out = 0;
  if (boxes_num <= 0 && method == 0) return 0;
  SHF_REQUIRE(boxes_dim >= 5, "_nms: boxes_dim=%d, need at least [x1,y1,x2,y2,score]", boxes_dim);
  SHF_REQUIRE(device_id >= 0 && device_id < 64, "_nms: device %d", device_id);
  int cur = -1;
  SHF_CUDA_CHECK(cudaGetDevice(&cur));
  if (cur != device_id) SHF_CUDA_CHECK(cudaSetDevice(device_id));
  HostPostBuffers& b = g_host_post[device_id];
  const int n = boxes_num > 0 ? boxes_num : 0;
  if (int rc = host_post_reserve(b, n > 0 ? n : 1)) return rc;
  for (int i = 0; i < n; ++i) memcpy(b.h_pin + (size_t)i * 5, boxes_host + (size_t)i * boxes_dim, 5 * sizeof(float));
  const int meta[4] = {0, n, 0, 0};
  SHF_CUDA_CHECK(cudaMemcpyAsync(b.d_dets, b.h_pin, (size_t)n * 5 * sizeof(float), cudaMemcpyHostToDevice, nullptr));
  SHF_CUDA_CHECK(cudaMemcpyAsync(b.d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice, nullptr));
  const int out_cap = method == 0 ? b.cap : b.cap / 2 + 1;
  int rc = shf_postprocess(b.d_dets, b.d_meta, b.d_meta + 1, 1, b.cap, thresh, method, mode, b.d_meta + 4, b.d_out,
                           b.d_meta + 2, out_cap, b.d_ws, b.ws_bytes, nullptr);
  if (rc) { *num_out = -1; return rc; }
  int cnt = 0;
  SHF_CUDA_CHECK(cudaMemcpy(&cnt, b.d_meta + 2, sizeof(int), cudaMemcpyDeviceToHost));
  SHF_REQUIRE(cnt >= 0 && cnt <= out_cap, "_nms: %d result rows for %d inputs", cnt, n);
  if (method == 0) SHF_CUDA_CHECK(cudaMemcpy(keep_out, b.d_meta + 4, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost));
  else SHF_CUDA_CHECK(cudaMemcpy(out_dets, b.d_out, (size_t)cnt * 5 * sizeof(float), cudaMemcpyDeviceToHost));
  *num_out = cnt;
  return 0;
}

extern "C" void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                     float nms_overlap_thresh, int device_id) {
  if (post_host_impl(0, keep_out, nullptr, num_out, boxes_host, boxes_num, boxes_dim, (double)nms_overlap_thresh, 1, device_id))
    *num_out = -1;
}

// Same contract, selectable comparison (0: cpu_nms `>=` in double, 1: gpu `>` float, 2: `>=` float) and an int status.
extern "C" int shf_nms_host(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                            double thresh, int mode, int device_id) {
  const int rc = post_host_impl(0, keep_out, nullptr, num_out, boxes_host, boxes_num, boxes_dim, thresh, mode, device_id);
  if (rc) *num_out = -1;
  return rc;
}

// lib/test.py:181-217 `bbox_vote(det)` with host buffers: dets_host (n x 5 float32, any order) -> out_dets (caller-allocated,
// at least n / 2 + 1 rows of 5 floats; an empty input yields the reference's single [10,10,20,20,1e-4] row), *num_out rows.
extern "C" int shf_bbox_vote_host(float* out_dets, int* num_out, const float* dets_host, int num_dets, double thresh,
                                  int device_id) {
  const int rc = post_host_impl(1, nullptr, out_dets, num_out, dets_host, num_dets, 5, thresh, 2, device_id);
  if (rc) *num_out = -1;
  return rc;
}

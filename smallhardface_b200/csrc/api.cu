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

extern "C" int shf_abi_version(void) { return 3; }

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
// being printed and ignored (nms_kernel.cu:12-19), and the greedy sweep runs on the device.
static int nms_host_impl(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                         double thresh, int mode, int device_id) {
  *num_out = 0;
  if (boxes_num <= 0) return 0;
  SHF_REQUIRE(boxes_dim >= 5, "_nms: boxes_dim=%d, need at least [x1,y1,x2,y2,score]", boxes_dim);
  int cur = -1;
  SHF_CUDA_CHECK(cudaGetDevice(&cur));
  if (cur != device_id) SHF_CUDA_CHECK(cudaSetDevice(device_id));
  float* packed = (float*)malloc((size_t)boxes_num * 5 * sizeof(float));
  SHF_REQUIRE(packed != nullptr, "_nms: out of host memory");
  for (int i = 0; i < boxes_num; ++i) memcpy(packed + (size_t)i * 5, boxes_host + (size_t)i * boxes_dim, 5 * sizeof(float));
  const long long ws_bytes = shf_postprocess_workspace(1, boxes_num);
  float* d_dets = nullptr;
  int* d_meta = nullptr;     // seg_begin, seg_end, count, then keep[boxes_num]
  void* d_ws = nullptr;
  int rc = 0;
  cudaError_t e;
  if ((e = cudaMalloc(&d_dets, (size_t)boxes_num * 5 * sizeof(float))) != cudaSuccess ||
      (e = cudaMalloc(&d_meta, (size_t)(boxes_num + 3) * sizeof(int))) != cudaSuccess ||
      (e = cudaMalloc(&d_ws, (size_t)ws_bytes)) != cudaSuccess) {
    shf_set_error("_nms: cudaMalloc failed: %s", cudaGetErrorString(e));
    rc = -2;
  }
  if (!rc) {
    const int meta[3] = {0, boxes_num, 0};
    if ((e = cudaMemcpy(d_dets, packed, (size_t)boxes_num * 5 * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice)) != cudaSuccess) {
      shf_set_error("_nms: H2D copy failed: %s", cudaGetErrorString(e));
      rc = -2;
    }
  }
  if (!rc)
    rc = shf_postprocess(d_dets, d_meta, d_meta + 1, 1, boxes_num, thresh, 0, mode, d_meta + 3, nullptr, d_meta + 2,
                         boxes_num, d_ws, ws_bytes, nullptr);
  if (!rc) {
    int n = 0;
    if ((e = cudaMemcpy(&n, d_meta + 2, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess ||
        (e = cudaMemcpy(keep_out, d_meta + 3, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) {
      shf_set_error("_nms: D2H copy failed: %s", cudaGetErrorString(e));
      rc = -2;
    } else {
      *num_out = n;
    }
  }
  free(packed);
  cudaFree(d_dets);
  cudaFree(d_meta);
  cudaFree(d_ws);
  if (rc) *num_out = -1;
  return rc;
}

extern "C" void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                     float nms_overlap_thresh, int device_id) {
  nms_host_impl(keep_out, num_out, boxes_host, boxes_num, boxes_dim, (double)nms_overlap_thresh, 1, device_id);
}

// Same contract, selectable comparison (0: cpu_nms `>=` in double, 1: gpu `>` float, 2: `>=` float) and an int status.
extern "C" int shf_nms_host(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                            double thresh, int mode, int device_id) {
  return nms_host_impl(keep_out, num_out, boxes_host, boxes_num, boxes_dim, thresh, mode, device_id);
}

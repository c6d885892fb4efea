// Shared device/host helpers for the sm_100a kernels: PTX wrappers (mbarrier, TMA, tcgen05, TMEM),
// the split-fp16 ("h2") activation format, and error plumbing for the C ABI.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SHF_DEVICE __device__ __forceinline__

// ----------------------------------------------------------------------------------------------
// error plumbing (C ABI returns 0 / negative code; message via shf_last_error())
// ----------------------------------------------------------------------------------------------
extern "C" const char* shf_last_error(void);
void shf_set_error(const char* fmt, ...);

#define SHF_CUDA_CHECK(expr)                                                                    \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      shf_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
      return -2;                                                                                \
    }                                                                                           \
  } while (0)

#define SHF_REQUIRE(cond, ...)                                                                  \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      shf_set_error(__VA_ARGS__);                                                               \
      return -1;                                                                                \
    }                                                                                           \
  } while (0)

#define SHF_LAUNCH_CHECK() SHF_CUDA_CHECK(cudaGetLastError())

// Tuning / timing probes (SHF_PROBE_* environment variables; some give WRONG results by design) are honoured only by a
// library built with -DSHF_PROBES (`make clean && make PROBES=1`): a stray variable cannot perturb a product build.
#include <stdlib.h>
inline const char* shf_probe_env(const char* name) {
#ifdef SHF_PROBES
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

// ----------------------------------------------------------------------------------------------
// Activation formats.  Both are NHWC, 4 bytes per element in two equally sized planes, C a multiple of 8:
//
//  SHF_FMT_H2  ("h2", precise): x = hi + lo, hi = rn_f16(x), lo = rn_f16(x - hi)   (|err| <= 2^-22 |x|).
//      planes [2][N][H][W][C] of __half; the conv issues hi*hi + hi*lo + lo*hi as three kind::f16 MMAs.
//
//  SHF_FMT_HF8 ("hf8", fast): plane 0 = hi = rn_f16(x) as above; plane 1 holds, per pixel and per block of 64
//      channels, 64 bytes  al8 = e4m3((x - hi) * 2^6)  followed by 64 bytes  ah8 = e4m3(hi * 2^-5)  (saturating) -- i.e.
//      one 128-byte K = 128 row of 8-bit floats.  The conv issues hi*hi as ONE kind::f16 MMA and the first-order
//      correction al*wh + ah*wl as ONE kind::f8f6f4 MMA (K = 32 per instruction, same issue time) against a weight
//      plane laid out the same way ([wh8 = e4m3(wh * 2^-6) | wl8 = e4m3(wl * 2^5)]): 2 MMAs per 16 channels, not 3.
//      x is recovered as hi + al8 * 2^-6.  The fixed exponents give the residual its full 4 significant bits for
//      2 <~ |x| < 16384 and hi its 4 bits for 0.5 <= |x| < 14336 (smaller values keep fewer bits -- their absolute
//      error is negligible next to the large activations'; larger ones saturate towards plain-fp16 accuracy);
//      C must be a multiple of 64.  tools/precision_model.py quantifies the accuracy per pyramid level (row h2f8v2; the
//      range-free e5m2 variant h2f8 has 1.5x the error).
// ----------------------------------------------------------------------------------------------
enum { SHF_FMT_H2 = 0, SHF_FMT_HF8 = 1 };

SHF_DEVICE void split_h2(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
SHF_DEVICE float join_h2(__half hi, __half lo) { return __half2float(hi) + __half2float(lo); }

constexpr float kHf8AlScale = 64.f, kHf8AlInv = 0.015625f;       // al8 = e4m3(lo * 2^6)
constexpr float kHf8AhScale = 0.03125f;                           // ah8 = e4m3(hi * 2^-5)
SHF_DEVICE uint8_t f32_to_f8(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }
SHF_DEVICE float f8_to_f32(uint8_t b) { return __half2float(__half(__nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3))); }
SHF_DEVICE void split_hf8(float v, __half& hi, uint8_t& al8, uint8_t& ah8) {
  hi = __float2half_rn(v);
  const float h = __half2float(hi);
  al8 = f32_to_f8((v - h) * kHf8AlScale);
  ah8 = f32_to_f8(h * kHf8AhScale);
}
SHF_DEVICE float join_hf8(__half hi, uint8_t al8) { return fmaf(f8_to_f32(al8), kHf8AlInv, __half2float(hi)); }
// byte offset of channel c's al8 inside a pixel's plane-1 row (its ah8 twin sits 64 bytes further)
SHF_DEVICE int hf8_off(int c) { return ((c >> 6) << 7) + (c & 63); }

// One element: px0 = plane-0 address of channel 0 of the pixel, plane_elems = elements per plane.
SHF_DEVICE float act_load1(const __half* px0, size_t plane_elems, int c, int fmt) {
  if (fmt == SHF_FMT_H2) return join_h2(px0[c], px0[plane_elems + c]);
  return join_hf8(px0[c], reinterpret_cast<const uint8_t*>(px0 + plane_elems)[hf8_off(c)]);
}
SHF_DEVICE void act_store1(__half* px0, size_t plane_elems, int c, float v, int fmt) {
  if (fmt == SHF_FMT_H2) {
    __half hi, lo;
    split_h2(v, hi, lo);
    px0[c] = hi;
    px0[plane_elems + c] = lo;
  } else {
    __half hi;
    uint8_t a, b;
    split_hf8(v, hi, a, b);
    px0[c] = hi;
    uint8_t* p1 = reinterpret_cast<uint8_t*>(px0 + plane_elems) + hf8_off(c);
    p1[0] = a;
    p1[64] = b;
  }
}
// Eight consecutive channels c .. c+7 (c % 8 == 0): 16-byte vectors on plane 0 (and on plane 1 for h2), two 8-byte
// vectors on plane 1 for hf8.
SHF_DEVICE void act_load8(const __half* px0, size_t plane_elems, int c, float (&v)[8], int fmt) {
  const uint4 vh = __ldg(reinterpret_cast<const uint4*>(px0 + c));
  const __half2* hh = reinterpret_cast<const __half2*>(&vh);
  if (fmt == SHF_FMT_H2) {
    const uint4 vl = __ldg(reinterpret_cast<const uint4*>(px0 + plane_elems + c));
    const __half2* ll = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(hh[j]), b = __half22float2(ll[j]);
      v[2 * j] = a.x + b.x;
      v[2 * j + 1] = a.y + b.y;
    }
  } else {
    const uint2 va = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(px0 + plane_elems) + hf8_off(c)));
    const uint8_t* ab = reinterpret_cast<const uint8_t*>(&va);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(hh[j]);
      const __half2_raw r2 = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)((unsigned short)ab[2 * j] | ((unsigned short)ab[2 * j + 1] << 8)), __NV_E4M3);
      const float2 l2 = __half22float2(__half2(r2));
      v[2 * j] = fmaf(l2.x, kHf8AlInv, a.x);
      v[2 * j + 1] = fmaf(l2.y, kHf8AlInv, a.y);
    }
  }
}
SHF_DEVICE void act_store8(__half* px0, size_t plane_elems, int c, const float (&v)[8], int fmt) {
  uint32_t hp[4];
  if (fmt == SHF_FMT_H2) {
    uint32_t lp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __half h0, l0, h1, l1;
      split_h2(v[2 * e], h0, l0);
      split_h2(v[2 * e + 1], h1, l1);
      hp[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      lp[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4*>(px0 + c) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(px0 + plane_elems + c) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
  } else {
    uint32_t ap[2] = {0u, 0u}, bp[2] = {0u, 0u};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // pairs: one F2FP packs two fp16 / two e4m3 values
      const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      const float2 hf = __half22float2(h);
      const float2 lo = make_float2((v[2 * e] - hf.x) * kHf8AlScale, (v[2 * e + 1] - hf.y) * kHf8AlScale);
      const float2 hs = make_float2(hf.x * kHf8AhScale, hf.y * kHf8AhScale);
      const uint32_t a2 = (uint32_t)__nv_cvt_float2_to_fp8x2(lo, __NV_SATFINITE, __NV_E4M3);
      const uint32_t b2 = (uint32_t)__nv_cvt_float2_to_fp8x2(hs, __NV_SATFINITE, __NV_E4M3);
      hp[e] = *reinterpret_cast<const uint32_t*>(&h);
      ap[e >> 1] |= a2 << (16 * (e & 1));
      bp[e >> 1] |= b2 << (16 * (e & 1));
    }
    *reinterpret_cast<uint4*>(px0 + c) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    uint8_t* p1 = reinterpret_cast<uint8_t*>(px0 + plane_elems) + hf8_off(c);
    *reinterpret_cast<uint2*>(p1) = make_uint2(ap[0], ap[1]);
    *reinterpret_cast<uint2*>(p1 + 64) = make_uint2(bp[0], bp[1]);
  }
}

// ----------------------------------------------------------------------------------------------
// Range guard: every kernel that WRITES an activation tensor can publish max |x| of what it wrote (float bits of a
// non-negative value order like unsigned integers) into one 32-bit slot.  The host reads the slots at its next
// synchronisation point: a value >= 65504 means the fp16 hi plane overflowed (either format); for hf8 tensors a value
// >= 14336 means the ah8 = e4m3(hi * 2^-5) bytes saturated and a small maximum means the residual bytes
// e4m3((x - hi) * 2^6) have underflowed their 4 significant bits -- the caller then repeats the level on split fp16
// (smallhardface_b200/engine.py: GpuNet.check_ranges).  One warp-level reduction + at most one atomic per warp per
// launch: call once, warp-converged, after the kernel's persistent / grid-stride loop.
// ----------------------------------------------------------------------------------------------
SHF_DEVICE void range_guard_commit(unsigned int* guard, float thread_abs_max) {
  if (guard == nullptr) return;
  const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(thread_abs_max));
  if ((threadIdx.x & 31) == 0 && m > *reinterpret_cast<volatile unsigned int*>(guard)) atomicMax(guard, m);
}

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
SHF_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

SHF_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

SHF_DEVICE void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
SHF_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SHF_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

SHF_DEVICE void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SHF_DEVICE void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
SHF_DEVICE bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU box.
SHF_DEVICE void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("shf: mbarrier timeout block (%d,%d,%d) thread %d bar %u parity %u\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---- TMA ---------------------------------------------------------------------------------------
SHF_DEVICE void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
SHF_DEVICE void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
SHF_DEVICE void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
SHF_DEVICE void tma_store_5d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
SHF_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SHF_DEVICE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
SHF_DEVICE void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- tcgen05 / TMEM ------------------------------------------------------------------------------
SHF_DEVICE void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
SHF_DEVICE void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
SHF_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
SHF_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SHF_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 operands, fp32 accumulate)
SHF_DEVICE void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-uniform variants for the MMA-issuing warp: every lane executes them with identical (uniform) operands and
// one elected lane issues.  Keeping the control flow convergent lets ptxas hold the descriptors in uniform
// registers; a divergent `if (lane == 0)` region costs ~23 SASS instructions per MMA (R2UR/ELECT/VOTE loops),
// which made the issuing thread -- not the tensor pipe -- the bottleneck (profiles/r01_*).
SHF_DEVICE void umma_f16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the 64-bit descriptors assembled inside the asm from a per-kernel-constant high word and a low word
// (start address >> 4) that is the only part that changes between MMAs: no 64-bit arithmetic in the issue loop.
SHF_DEVICE void umma_f16_elect_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                    uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4 twin (8-bit float operands, K = 32 per instruction, fp32 accumulate): same descriptors, same 32-byte
// K advance per instruction as kind::f16
SHF_DEVICE void umma_f8_elect_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                   uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
SHF_DEVICE void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
      ::"r"(bar) : "memory");
}
// ---- single-thread variants (call inside ONE `if (issuer)` region per group of MMAs; issuer = elect_one() once) ----
// Electing per instruction wraps every MMA in its own ELECT / branch / reconvergence pair (~50 cycles per MMA measured:
// slower than a 32-cycle N = 64 MMA, so the tensor pipe of the 64-channel layers idled 40 % of the time).
#define SHF_DEFINE_UMMA(NAME, GROUP, KIND)                                                                          \
  SHF_DEVICE void NAME(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, \
                       uint32_t accumulate) {                                                                       \
    asm volatile(                                                                                                   \
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"                                                             \
        "mov.b64 da, {%1, %2};\n\t"                                                                                 \
        "mov.b64 db, {%3, %4};\n\t"                                                                                 \
        "setp.ne.b32 p, %6, 0;\n\t"                                                                                 \
        "tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%0], da, db, %5, p;\n\t}\n" ::"r"(d_tmem),                \
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)                                     \
        : "memory");                                                                                                \
  }
SHF_DEFINE_UMMA(umma1_f16, "1", "f16")
SHF_DEFINE_UMMA(umma1_f8, "1", "f8f6f4")
SHF_DEFINE_UMMA(umma2_f16, "2", "f16")
SHF_DEFINE_UMMA(umma2_f8, "2", "f8f6f4")
#undef SHF_DEFINE_UMMA
SHF_DEVICE void umma2_commit(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n"
      ::"r"(bar) : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
SHF_DEVICE void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pairs (cluster of 2, tcgen05 cta_group::2) -----------------------------------------------------------
// The two CTAs of a pair run on the two SMs of one TPC.  One tcgen05.mma.cta_group::2 issued by the leader (cluster
// rank 0) drives both tensor cores: each CTA supplies its own 128 A rows and HALF of the B rows from its own shared
// memory, so the weight tile crosses L2->SM once per pair instead of once per CTA.
SHF_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
SHF_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
SHF_DEVICE uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
SHF_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: the only thing handed over is tensor memory, ordered by tcgen05.wait::ld + tcgen05.fence (a release here
  // costs a MEMBAR per drain)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion bytes are credited to an mbarrier that may live in the peer CTA (cluster address)
SHF_DEVICE void tma_load_5d_pair(uint32_t dst, const void* tmap, uint32_t bar_cluster, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
SHF_DEVICE void tma_load_4d_pair(uint32_t dst, const void* tmap, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
SHF_DEVICE void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
SHF_DEVICE void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
SHF_DEVICE void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
SHF_DEVICE void umma_f16_pair_elect_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
SHF_DEVICE void umma_f8_pair_elect_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %5, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all MMAs issued so far by this thread have completed) on the barrier at this offset in BOTH CTAs
SHF_DEVICE void umma_commit_pair_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n"
      ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane)
SHF_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
SHF_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 B apart
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
//  version=1 [46,48), layout SWIZZLE_128B=2 [61,64))
SHF_DEVICE uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO (unused for swizzled K-major; canonical value 1)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=f16, both K-major, MxN
// instruction descriptor, kind::f8f6f4: D=f32, A = e5m2 (format 1) or e4m3 (0), B likewise, both K-major
constexpr uint32_t umma_idesc_f8(int M, int N, uint32_t a_fmt, uint32_t b_fmt) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Implicit-GEMM convolution v3 on tcgen05 tensor cores (sm_100a): persistent CTAs, halo-tile operand reuse, and
// accumulators that are DRAINED into registers every few dozen MMAs while the tensor pipe keeps running.
//
// Replaces ConvolutionLayer<float>::Forward (conv_layer.cpp:30-46 -> base_conv_layer.cpp:255-279: im2col + sgemm) and
// the in-place ReLU / 2x2 max pooling that follow it.  The round-1 predecessors (one TMA load per (tap, chunk); a
// non-persistent halo-tile kernel) are gone from the tree; what this kernel does differently from them, and why
// (numbers from profiles/r01_*):
//  * persistent: one CTA per SM walks tiles t = blockIdx.x + i*gridDim.x.  v2 paid ~15k cycles per tile for CTA
//    launch, barrier init, TMEM alloc, first-load latency and a serial epilogue; here the TMA rings simply keep
//    running into the next tile and the epilogue of tile i overlaps the main loop of tile i+1.
//  * N-stacked MMAs: the weight stage is [B_hi ; B_lo] (2*BN rows), so  A_hi x [B_hi;B_lo]  is ONE N = 2*BN MMA
//    that yields hi*hi in columns [0,BN) and hi*lo in [BN,2BN); a second N = BN MMA adds A_lo x B_hi to the
//    cross columns.  8 MMAs per stage instead of 12 (the single issuing thread was the bottleneck), A_hi is read
//    from shared memory once instead of twice.
//  * streaming drain: tensor memory holds two accumulator sets of 2*BN columns.  The MMA issuer alternates sets
//    every "phase" (= the 9 taps of one 64-channel chunk: 36 k-steps); the four epilogue warps add the finished
//    set into fp32 registers with round-to-nearest and hand it back.  The tensor core truncates when it aligns
//    products to a running accumulator, so its error grows with the accumulation count: a phase is 36
//    accumulations against 96 (v2) or 864 (a single accumulator over K = 4608).
//  * epilogue writes NHWC hi/lo rows straight from registers (each thread owns one pixel: 2*BN contiguous bytes
//    per plane), so no staging buffer competes with the operand rings for shared memory.
//  * CTA pairs (CTAS = 2, the default): a cluster of two CTAs works on two horizontally adjacent pixel tiles with ONE
//    tcgen05.mma.cta_group::2 stream issued by the leader.  Each CTA loads only HALF of every weight stage (its
//    half of the output channels, hi and lo planes) -- the tensor cores read the other half from the peer's shared
//    memory -- so L2->SM weight traffic per SM halves.  r01 profiles showed that traffic (not DRAM, not the MMA
//    pipe) bounding the single-CTA kernel at ~9 TB/s.  Barriers: "full" lives in the leader and is credited by both
//    CTAs' TMA loads; "empty"/"accumulator full" are multicast commits to both CTAs; "accumulator drained" collects
//    the eight epilogue warps of the pair in the leader.
#include <stdlib.h>

#include "common.cuh"
#include "epilogue_store.cuh"
#include "tma_host.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kTH = 16, kTW = 8;
constexpr int kChunkK = 64;
constexpr int kMaxA = 4, kMaxB = 9;
constexpr int kThreads = 384;        // warps 0-3: TMA B / MMA / TMEM alloc / TMA A; warps 4-11: drain + epilogue

struct StreamParams {
  int H, W, batch;
  int cin_chunks, taps, dil, pad;
  int xw, xh;                 // halo width / height (pixels)
  int na, nb;                 // ring depths
  int a_bytes, a_tx, b_bytes;
  int tiles_x, tiles_y, n_tiles, total_tiles;
  int chunks_per_phase;       // G
  int ctot, cout_offset, relu;
  int in_fmt, out_fmt;        // SHF_FMT_* of the input (and weight) planes / of the tensors written
  int group_rows;             // streaming path: one MMA-issue region per kernel row (3 weight stages) instead of per tap
  int b_resident;             // the whole weight tensor fits the B ring (one stage per (chunk, tap)): load it once per CTA
  int probe;                  // SHF_PROBE_EPI timing probes (wrong results): 1 = no global stores, 2 = drain only
  float out_scale;
  const float* bias;
  __half* out;                // h2 destination tensor base (plane 0); plane 1 at + plane_elems (nullptr: skip)
  long long plane_elems;
  // fused 2x2/2 max pooling (pooling_layer.cpp:140-187) of the post-ReLU result: every warp of the epilogue holds
  // a 4x8 pixel patch, so the 2x2 window partners are lanes ^1 (x) and ^8 (y)
  __half* pool_out;           // nullptr: no pooling
  long long pool_plane_elems;
  int pool_ctot, pool_coffset;
  unsigned int* guard;        // range guard slot (max |x| written, float bits) or nullptr
  // fused residual add (EltwiseLayer SUM with unit coefficients, eltwise_layer.cpp:52-57) before the ReLU: the shortcut
  // tensor of a ResNet block, same N x H x W, read at a channel offset like the output is written
  const __half* res;          // nullptr: none
  long long res_plane_elems;
  int res_ctot, res_coffset, res_fmt;
  // 3x3 stride-2 mode (template S3): cin_chunks counts (tap, 64-channel chunk) pairs, cin_real the chunks per tap
  int cin_real;
};

// Activation tensor maps of a launch.  Every mode but S3 uses m[0]; the 3x3 stride-2 mode reads the input through its four
// parity views (rows / columns of one parity each: pixel and row pitch doubled, base shifted by one row / pixel).
struct TmapA4 {
  CUtensorMap m[4];
};

SHF_DEVICE void mbar_arrive_cnt(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

#ifdef SHF_PROBES
constexpr bool kProbes = true;
#else
constexpr bool kProbes = false;               // product builds: the timing probes are compiled out of the hot loops
#endif

// FMT = operand format of the activations read AND of the packed weights (SHF_FMT_*): a template parameter so that the
// MMA-issuing thread's loop carries no format / probe / residency tests (r02: ncu showed that thread, at ~110 SASS
// instructions and 550 cycles per 8-MMA weight stage, pacing every layer -- the N = 64 ones at half the tensor rate).
// RES = the launch adds a residual tensor in its epilogue (ResNet blocks): an instantiation of its own, so that the staged
// row loads do not cost the VGG path registers (the 128-wide variants sit at the 168-register cap).
// S3 = 3x3 convolution with stride 2 (pad 1): output pixel (y, x) reads input (2y + r - 1, 2x + s - 1).  Tap (r, s) lives in
// the parity view ((r + 1) & 1, (s + 1) & 1) of the input at view coordinates (y - [r == 0], x - [s == 0]), so the layer
// runs as a 1x1-style K loop over 9 * Cin / 64 (tap, chunk) pairs, each with ONE un-haloed TMA load from its tap's view
// (TMA zero fill outside a view = the zero padding) and the tap's weight stage -- no gather pass, no wasted MMAs.
template <int BN, int CTAS, int FMT, bool RES = false, bool S3 = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_stream_kernel(const __grid_constant__ TmapA4 tmaps_a, const __grid_constant__ CUtensorMap tmap_b,
                   const StreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int pipe_bytes = p.na * p.a_bytes + p.nb * p.b_bytes;
  // barriers: fullA[4] emptyA[4] fullB[6] emptyB[6] accFull[2] accEmpty[2]
  const uint32_t bar_base = smem_base + (uint32_t)pipe_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pipe_bytes + 8 * (2 * kMaxA + 2 * kMaxB + 4));
  float* bias_s = reinterpret_cast<float*>(smem + pipe_bytes + 512);      // [2][BN]: this tile's bias slice, double-buffered
  uint8_t* stage_s = smem + pipe_bytes + 1536;                            // 8 x 4 KB: row staging of the epilogue warps
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto empty_a = [&](int s) { return bar_base + 8u * (kMaxA + s); };
  auto full_b = [&](int s) { return bar_base + 8u * (2 * kMaxA + s); };
  auto empty_b = [&](int s) { return bar_base + 8u * (2 * kMaxA + kMaxB + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (2 * kMaxA + 2 * kMaxB + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (2 * kMaxA + 2 * kMaxB + 2 + s); };
  auto a_stage = [&](int s) { return smem_base + (uint32_t)(s * p.a_bytes); };
  auto b_stage = [&](int s) { return smem_base + (uint32_t)(p.na * p.a_bytes + s * p.b_bytes); };

  // warp index through a shuffle: ptxas then KNOWS it is warp-uniform and keeps the role branches, loop counters and
  // MMA descriptors in uniform registers (otherwise every tcgen05.mma operand costs an R2UR in the issue loop)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = (CTAS == 2) ? cluster_ctarank() : 0u;      // 0 = leader (issues the MMAs of the pair)
  const int cta_lin = (int)(blockIdx.x / CTAS), cta_cnt = (int)(gridDim.x / CTAS);   // persistent walk over (pair) tiles
  constexpr uint32_t kTmemCols = 512;          // whole tensor memory: two sets of 2*BN columns at fixed addresses
  constexpr uint32_t kSetCols = 2 * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmaps_a.m[0]);
    if (S3) { tma_prefetch_desc(&tmaps_a.m[1]); tma_prefetch_desc(&tmaps_a.m[2]); tma_prefetch_desc(&tmaps_a.m[3]); }
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kMaxA; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
    for (int s = 0; s < kMaxB; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 8 * CTAS); }
    fence_mbar_init();
  } else if (warp == 2) {
    if (CTAS == 2) { tmem_alloc_pair(smem_u32(tmem_slot), kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(tmem_slot), kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();              // the full allocation starts at column 0
  const int phases_per_tile = (p.cin_chunks + p.chunks_per_phase - 1) / p.chunks_per_phase;

  // tile -> (n tile, x tile, y tile, image); n fastest so neighbouring CTAs share the activation patch in L2
  auto decode_tile = [&](int t, int& nt, int& x0, int& y0, int& img) {
    nt = t % p.n_tiles;
    int q = t / p.n_tiles;
    x0 = ((q % p.tiles_x) * CTAS + (int)rank) * kTW;      // p.tiles_x counts tile PAIRS when CTAS == 2
    q /= p.tiles_x;
    y0 = (q % p.tiles_y) * kTH;
    img = q / p.tiles_y;
  };

  if (warp == 0) {
    // ===================== TMA producer: weight stages [B_hi ; B_lo], one per (chunk, tap) =====================
    if (lane == 0) {
      // ring position kept as (stage, phase bit): a runtime `it % nb`, `it / nb` costs an integer division per use,
      // which in the MMA-issuing warp was ~30 % of its cycles (profiles/r01 source view)
      int sb = 0;
      uint32_t phb = 0;
      bool first = true;
      for (int t = cta_lin; t < p.total_tiles; t += cta_cnt) {
        int nt, x0, y0, img;
        decode_tile(t, nt, x0, y0, img);
        if (p.b_resident && !first) break;     // weights stay in shared memory for every tile of this CTA
        first = false;
        for (int cc = 0; cc < p.cin_chunks; ++cc)
          for (int tap0 = 0; tap0 < p.taps; ++tap0) {
            // S3: the K loop walks (tap, chunk) pairs; the weight stage of pair cc is chunk cc % cin_real of tap cc / cin_real
            const int tap = S3 ? cc / p.cin_real : tap0, wc = S3 ? cc % p.cin_real : cc;
            mbar_wait(empty_b(sb), phb ^ 1);
            if (CTAS == 2) {
              // both CTAs credit the LEADER's barrier; each loads its half of the output channels (hi and lo plane)
              if (rank == 0) mbar_arrive_expect_tx(full_b(sb), 2 * p.b_bytes);
              tma_load_4d_pair(b_stage(sb), &tmap_b, mapa_shared(full_b(sb), 0), wc * kChunkK,
                               nt * BN + (int)rank * (BN / 2), tap, 0);
            } else {
              mbar_arrive_expect_tx(full_b(sb), p.b_bytes);
              tma_load_4d(b_stage(sb), &tmap_b, full_b(sb), wc * kChunkK, nt * BN, tap, 0);
            }
            if (++sb == p.nb) { sb = 0; phb ^= 1; }
          }
      }
      if (CTAS == 2 && !p.b_resident)      // tail: the leader's multicast commits must have landed here before this CTA may exit
        for (int k = 0; k < p.nb; ++k) {
          mbar_wait(empty_b(sb), phb ^ 1);
          if (++sb == p.nb) { sb = 0; phb ^= 1; }
        }
    }
  } else if (warp == 3) {
    // ===================== TMA producer: activation halos, one per 64-channel chunk =====================
    if (lane == 0) {
      int sa = 0;
      uint32_t pha = 0;
      for (int t = cta_lin; t < p.total_tiles; t += cta_cnt) {
        int nt, x0, y0, img;
        decode_tile(t, nt, x0, y0, img);
        for (int cc = 0; cc < p.cin_chunks; ++cc) {
          mbar_wait(empty_a(sa), pha ^ 1);
          if (kProbes && (p.probe & 4)) {       // timing probe: no activation loads at all
            if (rank == 0) mbar_arrive(full_a(sa));
          } else {
            // S3: tap (r, s) of pair cc -> parity view and view-coordinate shift (see the template comment)
            const int tap = S3 ? cc / p.cin_real : 0, ac = S3 ? cc % p.cin_real : cc;
            const int r = tap / 3, s3 = tap - 3 * r;
            const CUtensorMap* ma = S3 ? &tmaps_a.m[(((r + 1) & 1) << 1) | ((s3 + 1) & 1)] : &tmaps_a.m[0];
            const int ax = S3 ? x0 - (s3 == 0 ? 1 : 0) : x0 - p.pad, ay = S3 ? y0 - (r == 0 ? 1 : 0) : y0 - p.pad;
            if (CTAS == 2) {
              if (rank == 0) mbar_arrive_expect_tx(full_a(sa), 2 * p.a_tx);
              tma_load_5d_pair(a_stage(sa), ma, mapa_shared(full_a(sa), 0), ac * kChunkK, ax, ay, img, 0);
            } else {
              mbar_arrive_expect_tx(full_a(sa), p.a_tx);
              tma_load_5d(a_stage(sa), ma, full_a(sa), ac * kChunkK, ax, ay, img, 0);
            }
          }
          if (++sa == p.na) { sa = 0; pha ^= 1; }
        }
      }
      if (CTAS == 2)
        for (int k = 0; k < p.na; ++k) {
          mbar_wait(empty_a(sa), pha ^ 1);
          if (++sa == p.na) { sa = 0; pha ^= 1; }
        }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (warp converged, one elected lane issues; leader CTA only) =====================
    constexpr uint32_t idesc_wide = umma_idesc_f16(kTileM * CTAS, 2 * BN);     // A_hi x [B_hi ; B_lo]      (CTAS == 1)
    constexpr uint32_t idesc_half = umma_idesc_f16(kTileM * CTAS, BN);         // one operand-plane pair
    constexpr uint32_t idesc_f8 = umma_idesc_f8(kTileM * CTAS, BN, 0u, 0u);    // A and B = e4m3 (format code 0)
    const uint32_t a_hi32 = (uint32_t)((p.xw * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO | version | SWIZZLE_128B
    constexpr uint32_t b_hi32 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t lbo = 1u << 16;
    constexpr uint32_t b_plane16 = (uint32_t)((BN / CTAS) * 128) >> 4;           // hi rows -> lo rows of a weight stage
    const uint32_t a_plane16 = ((uint32_t)(p.xh * p.xw) * 128u) >> 4;
    const int ktaps = (p.taps == 9) ? 3 : 1;
    const uint32_t tap_dx = (uint32_t)p.dil * 128u;                              // bytes to the next tap column
    const uint32_t tap_dy = (uint32_t)(p.dil * p.xw) * 128u - (uint32_t)ktaps * tap_dx;   // ... and on to the next tap row
    const bool issuer = elect_one();          // the one lane that talks to the tensor core
    // loop-invariant parameters in registers (the asm volatile MMAs would otherwise force constant-bank reloads per stage)
    const int taps = p.taps, nb = p.nb, na = p.na, cin_chunks = p.cin_chunks, cpp = p.chunks_per_phase;
    const uint32_t a_bytes = (uint32_t)p.a_bytes, b_bytes = (uint32_t)p.b_bytes;
    const uint32_t b_ring = smem_base + (uint32_t)na * a_bytes;
    const bool skip_mma = kProbes && (p.probe & 8);          // timing probe: no MMAs, only the barrier traffic
    // the 8 (hf8) or 12 (h2) MMAs of one weight stage = one tap of one 64-channel chunk
    auto issue_stage = [&](uint32_t d_main, uint32_t a_lo32, uint32_t b_lo32, uint32_t opened) {
      if (skip_mma) return;
      if (FMT == SHF_FMT_HF8) {
        // hi*hi as one f16 MMA, the first-order correction [al8 | ah8] x [wh8 | wl8] as one f8 MMA over the
        // same 32 bytes of K per operand row (plane 1 of either stage), both into the same accumulator
#pragma unroll
        for (int k = 0; k < kChunkK / 16; ++k) {
          if (CTAS == 2) {
            umma2_f16(d_main, a_lo32, a_hi32, b_lo32, b_hi32, idesc_half, opened);
            umma2_f8(d_main, a_lo32 + a_plane16, a_hi32, b_lo32 + b_plane16, b_hi32, idesc_f8, 1u);
          } else {
            umma1_f16(d_main, a_lo32, a_hi32, b_lo32, b_hi32, idesc_half, opened);
            umma1_f8(d_main, a_lo32 + a_plane16, a_hi32, b_lo32 + b_plane16, b_hi32, idesc_f8, 1u);
          }
          opened = 1u;
          a_lo32 += 2; b_lo32 += 2;
        }
      } else if (CTAS == 2) {
        // the pair's weight stage is split by output channel, so hi and lo rows are separate N = BN operands:
        //   main += A_hi x B_hi ;  cross += A_hi x B_lo ;  cross += A_lo x B_hi
#pragma unroll
        for (int k = 0; k < kChunkK / 16; ++k) {
          umma2_f16(d_main, a_lo32, a_hi32, b_lo32, b_hi32, idesc_half, opened);
          umma2_f16(d_main + BN, a_lo32, a_hi32, b_lo32 + b_plane16, b_hi32, idesc_half, opened);
          umma2_f16(d_main + BN, a_lo32 + a_plane16, a_hi32, b_lo32, b_hi32, idesc_half, 1u);
          opened = 1u;
          a_lo32 += 2; b_lo32 += 2;
        }
      } else {
#pragma unroll
        for (int k = 0; k < kChunkK / 16; ++k) {
          // [main | cross] += A_hi x [B_hi ; B_lo]   (first MMA of a phase overwrites both halves)
          umma1_f16(d_main, a_lo32, a_hi32, b_lo32, b_hi32, idesc_wide, opened);
          // cross += A_lo x B_hi
          umma1_f16(d_main + BN, a_lo32 + a_plane16, a_hi32, b_lo32, b_hi32, idesc_half, 1u);
          opened = 1u;
          a_lo32 += 2; b_lo32 += 2;
        }
      }
    };
    auto commit = [&](uint32_t bar) { if (CTAS == 2) umma2_commit(bar); else umma_commit(bar); };
    int sa = 0, gp = 0;
    uint32_t pha = 0;
    const bool grouped = (taps == 9) && (nb >= 4) && p.group_rows;   // streaming path: three weight stages per single-thread region
    const uint32_t row_pitch = (uint32_t)(p.dil * p.xw) * 128u;   // bytes from one kernel row of the halo to the next
    if (p.b_resident) {
      // ---- resident weights (conv1_2: one N tile, all taps in the ring): wait for them ONCE, then every halo is one
      //      single-thread region that issues all of its taps back to back -- no per-tap barrier wait, election or
      //      ring bookkeeping between the MMAs
      for (int s2 = 0; s2 < nb; ++s2) mbar_wait(full_b(s2), 0u);
      tc_fence_after();
      for (int t = cta_lin; t < p.total_tiles; t += cta_cnt) {
        for (int ph = 0; ph < phases_per_tile; ++ph, ++gp) {
          const int set = gp & 1;
          const uint32_t d_main = (uint32_t)set * kSetCols;
          mbar_wait(acc_empty(set), ((gp >> 1) & 1) ^ 1);
          const int c_begin = ph * cpp;
          const int c_end = min(c_begin + cpp, cin_chunks);
          uint32_t opened = 0;
          for (int cc = c_begin; cc < c_end; ++cc) {
            mbar_wait(full_a(sa), pha);
            tc_fence_after();
            if (issuer) {
              const uint32_t a_base = smem_base + (uint32_t)sa * a_bytes;
              uint32_t a_off = 0, b_addr = b_ring + (uint32_t)(cc * taps) * b_bytes;
              int s = 0;
              for (int tap = 0; tap < taps; ++tap) {
                issue_stage(d_main, (((a_base + a_off) & 0x3FFFFu) >> 4) | lbo, ((b_addr & 0x3FFFFu) >> 4) | lbo, opened);
                opened = 1u;
                b_addr += b_bytes;
                a_off += tap_dx;
                if (++s == ktaps) { s = 0; a_off += tap_dy; }
              }
              commit(empty_a(sa));
              if (cc == c_end - 1) commit(acc_full(set));
            }
            __syncwarp();
            opened = 1u;
            if (++sa == na) { sa = 0; pha ^= 1; }
          }
        }
      }
    } else {
      int sb = 0;
      uint32_t phb = 0;
      for (int t = cta_lin; t < p.total_tiles; t += cta_cnt) {
        for (int ph = 0; ph < phases_per_tile; ++ph, ++gp) {
          const int set = gp & 1;
          const uint32_t d_main = (uint32_t)set * kSetCols;
          mbar_wait(acc_empty(set), ((gp >> 1) & 1) ^ 1);           // drained by the epilogue warps (of both CTAs)
          tc_fence_after();
          const int c_begin = ph * cpp;
          const int c_end = min(c_begin + cpp, cin_chunks);
          uint32_t opened = 0;                                      // 0 until the set has been overwritten once
          for (int cc = c_begin; cc < c_end; ++cc) {
            mbar_wait(full_a(sa), pha);
            const uint32_t a_base = smem_base + (uint32_t)sa * a_bytes;
            uint32_t a_off = 0;
            if (grouped) {
              // 3x3 kernels, ring of >= 4 weight stages: one single-thread region per KERNEL ROW (three weight stages,
              // 24 / 36 MMAs) -- the barrier waits, the election and the ring bookkeeping are paid once per three stages;
              // every stage is still released by its own commit as soon as its MMAs retire
              for (int row = 0; row < 3; ++row) {
                int s3[3];
                uint32_t p3[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  s3[j] = sb; p3[j] = phb;
                  if (++sb == nb) { sb = 0; phb ^= 1; }
                }
#pragma unroll
                for (int j = 0; j < 3; ++j) mbar_wait(full_b(s3[j]), p3[j]);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                  for (int j = 0; j < 3; ++j) {
                    issue_stage(d_main, (((a_base + a_off + (uint32_t)j * tap_dx) & 0x3FFFFu) >> 4) | lbo,
                                (((b_ring + (uint32_t)s3[j] * b_bytes) & 0x3FFFFu) >> 4) | lbo, j == 0 ? opened : 1u);
                    commit(empty_b(s3[j]));
                  }
                  if (row == 2) {                 // last row of this halo: hand the A stage back as well
                    commit(empty_a(sa));
                    if (cc == c_end - 1) commit(acc_full(set));     // ... and, after the phase's last chunk, the accumulators
                  }
                }
                __syncwarp();
                opened = 1u;
                a_off += row_pitch;
              }
            } else {
            int s = 0;
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(full_b(sb), phb);
              tc_fence_after();
              if (issuer) {                       // ONE single-thread region per weight stage: 8 or 12 MMAs + the stage release
                issue_stage(d_main, (((a_base + a_off) & 0x3FFFFu) >> 4) | lbo,
                            (((b_ring + (uint32_t)sb * b_bytes) & 0x3FFFFu) >> 4) | lbo, opened);
                commit(empty_b(sb));
                if (tap == taps - 1) {            // last tap of this halo: hand the A stage back as well
                  commit(empty_a(sa));
                  if (cc == c_end - 1) commit(acc_full(set));       // ... and, after the phase's last chunk, the accumulators
                }
              }
              __syncwarp();
              opened = 1u;
              if (++sb == nb) { sb = 0; phb ^= 1; }
              a_off += tap_dx;
              if (++s == ktaps) { s = 0; a_off += tap_dy; }
            }
            }
            if (++sa == na) { sa = 0; pha ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== drain + epilogue warps =====================
    // Eight warps: warp w may only touch TMEM lanes 32*(w%4)..+31, so warps w and w+4 share a pixel quadrant and split
    // its channels (kCols each).  Two epilogue warps per scheduler hide each other's shuffle / conversion / store
    // latencies; with four, the epilogue of the 64-channel layers was slower than their main loop.
    constexpr int kCols = BN / 2;
    const int w = warp - 4;
    const int quad = w & 3, half = w >> 2;
    const int m = quad * 32 + lane;              // accumulator row = pixel (y_local * 8 + x_local)
    const int col0 = half * kCols;               // first channel of the n-tile this warp owns
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float scale = p.out_scale;
    const uint32_t drained0 = (CTAS == 2) ? mapa_shared(acc_empty(0), 0) : acc_empty(0);   // in the leader
    const uint32_t drained1 = (CTAS == 2) ? mapa_shared(acc_empty(1), 0) : acc_empty(1);
    int gp = 0, tile_it = 0, last_nt = -1;
    float gmax = 0.f;                            // range guard: max |x| this thread has written
    for (int t = cta_lin; t < p.total_tiles; t += cta_cnt) {
      int nt, x0, y0, img;
      decode_tile(t, nt, x0, y0, img);
      // stage this tile's bias slice in shared memory now, so that its global-load latency hides behind the main loop
      // (read per 8-channel group straight from global memory it serialised ~16 L2 round trips into every epilogue)
      // Layers with ONE channel tile (conv1_2) keep the same slice for every tile: it is staged once, and the eight warps
      // are not forced back into lockstep by a barrier per tile (their drains / shuffles / stores then overlap freely).
      if (nt != last_nt) {
        ++tile_it;
        float* bias_w = bias_s + (tile_it & 1) * BN;
        if (w * 32 + lane < BN) bias_w[w * 32 + lane] = p.bias ? __ldg(p.bias + nt * BN + w * 32 + lane) : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");     // the eight epilogue warps only
        last_nt = nt;
      }
      float* bias_t = bias_s + (tile_it & 1) * BN;
      float acc[kCols];
#pragma unroll
      for (int c = 0; c < kCols; ++c) acc[c] = 0.f;
      for (int ph = 0; ph < phases_per_tile; ++ph, ++gp) {
        const int set = gp & 1;
        mbar_wait(acc_full(set), (gp >> 1) & 1);
        tc_fence_after();
        const uint32_t base = lane_addr + (uint32_t)set * kSetCols;
        if (FMT == SHF_FMT_HF8) {
#pragma unroll
          for (int c0 = 0; c0 < kCols; c0 += 32) {
            uint32_t mq[32];
            tmem_ld_32x32(base + col0 + c0, mq);     // hi*hi + correction partial
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[c0 + e] = __fadd_rn(acc[c0 + e], __uint_as_float(mq[e]));
          }
        } else {
#pragma unroll
          for (int c0 = 0; c0 < kCols; c0 += 32) {
            uint32_t mq[32], cq[32];
            tmem_ld_32x32(base + col0 + c0, mq);            // hi*hi partial
            tmem_ld_32x32(base + BN + col0 + c0, cq);       // cross partial
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e)
              acc[c0 + e] = __fadd_rn(__fadd_rn(acc[c0 + e], __uint_as_float(cq[e])), __uint_as_float(mq[e]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                          // 8 (x2) warps -> the set is free again
          if (CTAS == 2) mbar_arrive_cluster(set ? drained1 : drained0);
          else mbar_arrive_cnt(acc_empty(set));
        }
      }
      // ---- epilogue for this tile: bias, ReLU, (2x2 max pool), split to hi/lo, NHWC rows straight to global memory ----
      const int y = y0 + (m >> 3), x = x0 + (m & 7);
      const bool inside = (y < p.H && x < p.W);
      const int n0 = nt * BN;
      if (kProbes && (p.probe & 3) == 2) continue;
      // When only the pooled map is written (conv1_2, conv2_2, conv3_3) bias + ReLU + guard run AFTER the pooling, on the
      // quarter of the values each lane keeps: max commutes exactly with x -> relu(x * scale + bias), scale = 2^-k > 0.
      const bool post_pool = p.pool_out && !p.out;
      if (!post_pool) {
#pragma unroll
        for (int c = 0; c < kCols; ++c) acc[c] = fmaf(acc[c], scale, bias_t[col0 + c]);
        if (RES && p.res != nullptr) {                        // warp-uniform: all lanes take part in the staged row loads
          const int ry = y0 + quad * 4;
          const __half* rbase = p.res + (((size_t)img * p.H + ry) * p.W + x0) * (size_t)p.res_ctot;
          const uint32_t rpitch = (uint32_t)p.W * (uint32_t)p.res_ctot, rct = (uint32_t)p.res_ctot;
          auto src = [&](int row) -> const __half* {
            const int dy = row >> 3, dx = row & 7;
            return (ry + dy < p.H && x0 + dx < p.W) ? rbase + ((uint32_t)dy * rpitch + (uint32_t)dx * rct) : nullptr;
          };
          add_rows_in<kCols>(stage_s + w * 4096, lane, acc, p.res_fmt, p.res_coffset + nt * BN + col0,
                             (size_t)p.res_plane_elems, src);
        }
        if (p.relu) {
#pragma unroll
          for (int c = 0; c < kCols; ++c) acc[c] = fmaxf(acc[c], 0.f);
        }
        if (p.guard && inside) {
#pragma unroll
          for (int c = 0; c < kCols; ++c) gmax = fmaxf(gmax, fabsf(acc[c]));
        }
      }
      uint8_t* stg = stage_s + w * 4096;
      const int c_first = n0 + col0;
      const int qy = y0 + quad * 4;                         // this warp's 4 x 8 pixel patch starts at (qy, x0)
      if (p.out && !(kProbes && (p.probe & 3) == 1)) {
        // row pointers = one 64-bit base per tile + 32-bit offsets (W * ctot < 2^31 elements is checked on the host)
        __half* base = p.out + (((size_t)img * p.H + qy) * p.W + x0) * (size_t)p.ctot;
        const uint32_t pitch = (uint32_t)p.W * (uint32_t)p.ctot, ct = (uint32_t)p.ctot;
        auto dst = [&](int row) -> __half* {
          const int dy = row >> 3, dx = row & 7;
          return (qy + dy < p.H && x0 + dx < p.W) ? base + ((uint32_t)dy * pitch + (uint32_t)dx * ct) : nullptr;
        };
        store_plane<kCols>(stg, lane, acc, 0, p.out_fmt, p.cout_offset + c_first, (size_t)p.plane_elems, dst);
        store_plane<kCols>(stg, lane, acc, 1, p.out_fmt, p.cout_offset + c_first, (size_t)p.plane_elems, dst);
      }
      if (p.pool_out && !(kProbes && (p.probe & 3) == 1)) {  // warp-uniform branch: all lanes take part in the shuffles
        const int ph2 = p.H >> 1, pw2 = p.W >> 1;
        __half* base = p.pool_out + (((size_t)img * ph2 + (qy >> 1)) * pw2 + (x0 >> 1)) * (size_t)p.pool_ctot;
        const uint32_t pitch = (uint32_t)pw2 * (uint32_t)p.pool_ctot, ct = (uint32_t)p.pool_ctot;
        auto dst = [&](int row) -> __half* {                // row = pooled pixel (2 rows x 4 columns per warp)
          const int dy = row >> 2, dx = row & 3;
          return (qy + 2 * dy < p.H && x0 + 2 * dx < p.W) ? base + ((uint32_t)dy * pitch + (uint32_t)dx * ct) : nullptr;
        };
        if (post_pool) {
          const float* bias_w = bias_t + col0;
          const bool relu = p.relu != 0, guard = p.guard != nullptr && inside;
          store_pooled<kCols>(stg, lane, acc, p.out_fmt, p.pool_coffset + c_first, (size_t)p.pool_plane_elems, dst,
                              [&](float x, int ch) {
                                float y = fmaf(x, scale, bias_w[ch]);
                                if (relu) y = fmaxf(y, 0.f);
                                if (guard) gmax = fmaxf(gmax, fabsf(y));
                                return y;
                              });
        } else {
          store_pooled<kCols>(stg, lane, acc, p.out_fmt, p.pool_coffset + c_first, (size_t)p.pool_plane_elems, dst,
                              [](float x, int) { return x; });
        }
      }
    }
    range_guard_commit(p.guard, gmax);
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(0u, kTmemCols); else tmem_dealloc(0u, kTmemCols);
  }
}

template <int BN, int CTAS, int FMT, bool RES = false, bool S3 = false>
int launch_stream(const TmapA4& ta, const CUtensorMap& tb, const StreamParams& p, int smem_bytes, int grid,
                  cudaStream_t stream) {
  static bool attr[64] = {};                   // function attributes are per device
  int dev = 0;
  SHF_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    SHF_CUDA_CHECK(cudaFuncSetAttribute(conv_stream_kernel<BN, CTAS, FMT, RES, S3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  if (CTAS == 1) {
    conv_stream_kernel<BN, CTAS, FMT, RES, S3><<<grid, kThreads, smem_bytes, stream>>>(ta, tb, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CTAS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SHF_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_stream_kernel<BN, CTAS, FMT, RES, S3>, ta, tb, p));
  }
  SHF_LAUNCH_CHECK();
  return 0;
}

int sm_count() {
  static int cached[64] = {};                  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  int n = (dev >= 0 && dev < 64) ? cached[dev] : 0;
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
  }
  return n;
}

}  // namespace

static int shf_conv_stream_impl(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H, int W,
                         int cin, int cout, int ksize, int dilation, int out_channels_total, int out_channel_offset,
                         float out_scale, int relu, void* pool_out_h2, int pool_channels_total, int pool_channel_offset,
                         int ctas, int in_format, int out_format, unsigned int* range_guard, void* stream,
                         int in_stride = 1, int in_H = 0, int in_W = 0, const void* residual = nullptr,
                         int res_channels_total = 0, int res_channel_offset = 0, int res_format = 0, int s3 = 0) {
  // in_stride > 1 (1x1 kernels only): H x W are the OUTPUT dims, the activations are read through a strided TMA view
  // of the in_H x in_W input (every in_stride-th pixel) -- conv_layer.cpp:8-28 with kernel 1, pad 0
  SHF_REQUIRE(ctas == 1 || ctas == 2, "shf_conv_igemm: %d CTAs per tile group", ctas);
  // s3: 3x3 kernel, stride 2, pad 1 (H x W = OUTPUT dims, in_H x in_W the input): runs as a 1x1-style K loop over (tap, chunk)
  // pairs on the four parity views of the input (template S3 of the kernel)
  SHF_REQUIRE(in_stride >= 1 && (in_stride == 1 || ((ksize == 1 || s3) && pool_out_h2 == nullptr)),
              "shf_conv_igemm: stride %d needs a 1x1 kernel (or the 3x3 stride-2 mode) without fused pooling", in_stride);
  SHF_REQUIRE(!s3 || (ksize == 3 && in_stride == 2 && dilation == 1 && ctas == 2 && residual == nullptr),
              "shf_conv3x3_s2: needs a 3x3 kernel, stride 2, no dilation, the CTA-pair kernel, no residual");
  SHF_REQUIRE((in_format == SHF_FMT_H2 || in_format == SHF_FMT_HF8) && (out_format == SHF_FMT_H2 || out_format == SHF_FMT_HF8),
              "shf_conv_igemm: unknown activation format %d / %d", in_format, out_format);
  if (out_format == SHF_FMT_HF8)
    SHF_REQUIRE(out_channel_offset % 64 == 0 && out_channels_total % 64 == 0 &&
                    (!pool_out_h2 || (pool_channel_offset % 64 == 0 && pool_channels_total % 64 == 0)),
                "shf_conv_igemm: hf8 tensors need channel windows aligned to 64");
  SHF_REQUIRE(out_h2 != nullptr || pool_out_h2 != nullptr, "shf_conv_igemm: no destination");
  if (residual != nullptr) {
    SHF_REQUIRE(pool_out_h2 == nullptr && out_h2 != nullptr, "shf_conv_igemm_res: no fused pooling with a residual");
    SHF_REQUIRE(res_format == SHF_FMT_H2 || res_format == SHF_FMT_HF8, "shf_conv_igemm_res: unknown residual format %d", res_format);
    SHF_REQUIRE(res_channel_offset % 8 == 0 && res_channels_total % 8 == 0 && res_channel_offset + cout <= res_channels_total &&
                    (res_format == SHF_FMT_H2 || (res_channel_offset % 64 == 0 && res_channels_total % 64 == 0)),
                "shf_conv_igemm_res: bad residual channel window [%d,%d) of %d", res_channel_offset,
                res_channel_offset + cout, res_channels_total);
  }
  if (pool_out_h2)
    SHF_REQUIRE(H % 2 == 0 && W % 2 == 0 && pool_channel_offset % 8 == 0 && pool_channels_total % 8 == 0 &&
                    pool_channel_offset + cout <= pool_channels_total,
                "shf_conv_igemm_pool: fused pooling needs even H, W (got %dx%d) and an 8-aligned channel window", H, W);
  SHF_REQUIRE(ksize == 3 || ksize == 1, "shf_conv_igemm: kernel size %d (only 3x3 and 1x1 are on the hot path)", ksize);
  SHF_REQUIRE(cin % 64 == 0 && cin >= 64, "shf_conv_igemm: Cin=%d must be a multiple of 64", cin);
  SHF_REQUIRE(cout % 64 == 0 && cout >= 64, "shf_conv_igemm: Cout=%d must be a multiple of 64", cout);
  SHF_REQUIRE(out_h2 == nullptr || (out_channel_offset % 8 == 0 && out_channel_offset + cout <= out_channels_total &&
                                     out_channels_total % 8 == 0),
              "shf_conv_igemm: bad destination channel window [%d,%d) of %d", out_channel_offset,
              out_channel_offset + cout, out_channels_total);
  SHF_REQUIRE(batch >= 1 && H >= 1 && W >= 1 && dilation >= 1 && dilation <= 4, "shf_conv_igemm: bad geometry");
  SHF_REQUIRE((long long)W * res_channels_total < (1ll << 29), "shf_conv_igemm_res: residual row overflows 32-bit offsets");
  SHF_REQUIRE((long long)W * out_channels_total < (1ll << 29) && (long long)W * pool_channels_total < (1ll << 29),
              "shf_conv_igemm: a row of %d pixels x %d channels overflows the epilogue's 32-bit in-tile offsets", W,
              out_channels_total);
  // N tile: 128 output channels, or 64 when the layer has 64-channel granularity -- or when 128-wide tiles would leave
  // more than half of the CTA pairs idle (the deep layers of the small pyramid levels, batch 1 on the plugin path:
  // conv4_x of a 304-px level is 9 pixel tile pairs x 4 channel tiles on 74 SM pairs, each walking K = 4608 alone).
  // Twice the tiles at half the MMA time each; the weight packing does not depend on the tile width.
  int bn = (cout % 128 == 0) ? 128 : 64;
  static int small_bn64 = -1;                 // A/B switch: SHF_CONV_SMALL_BN64=0 keeps 128-wide tiles everywhere
  if (small_bn64 < 0) { const char* e = getenv("SHF_CONV_SMALL_BN64"); small_bn64 = (e && e[0] == '0') ? 0 : 1; }
  if (bn == 128 && small_bn64) {
    const long long px_tiles = (long long)(((W + kTW - 1) / kTW + ctas - 1) / ctas) * ((H + kTH - 1) / kTH) * batch;
    if (px_tiles * (cout / 128) * 2 <= sm_count() / ctas) bn = 64;
  }
  StreamParams p;
  p.H = H; p.W = W; p.batch = batch;
  p.cin_real = cin / 64;
  p.cin_chunks = s3 ? 9 * (cin / 64) : cin / 64;
  p.taps = s3 ? 1 : ksize * ksize;
  p.dil = (ksize == 3 && !s3) ? dilation : 0;
  p.pad = p.dil;
  p.xw = kTW + 2 * p.pad;
  p.xh = kTH + 2 * p.pad;
  p.a_tx = 2 * p.xh * p.xw * 128;
  p.a_bytes = (p.a_tx + 1023) & ~1023;
  p.b_bytes = 2 * (bn / ctas) * 128;          // per CTA: its share of the output channels, hi + lo plane
  const int budget = 227 * 1024 - 1024 - 1536 - 8 * 4096;      // alignment slack; barriers (512) + staged bias (1024); epilogue row staging
  // A-ring depth: a halo stage lasts taps x 4 k-steps of MMAs; when that is short (one 64-channel chunk per tile at
  // BN = 64: ~2-3k cycles, or 1x1 convs) two stages do not cover the ~4k cycles a 46 KB halo takes to arrive
  p.na = (p.taps == 1) ? kMaxA : ((bn == 64 || p.cin_chunks == 1) ? 3 : 2);
  if (const char* e = shf_probe_env("SHF_PROBE_NA")) { int v = atoi(e); if (v >= 1 && v <= kMaxA) p.na = v; }
  while (p.na > 1 && p.na * p.a_bytes + 4 * p.b_bytes > budget) --p.na;
  p.nb = (budget - p.na * p.a_bytes) / p.b_bytes;
  if (p.nb > kMaxB) p.nb = kMaxB;
  // conv1_2-class layers (one N tile, 9 stages of weights in all): keep the weights in shared memory for the whole
  // kernel -- streaming them again for every 128-pixel tile was most of that layer's L2->SM traffic
  p.b_resident = (!s3 && cout == bn && p.taps * p.cin_chunks <= p.nb && p.na >= 2 && !shf_probe_env("SHF_PROBE_NO_RESIDENT")) ? 1 : 0;
  if (p.b_resident) p.nb = p.taps * p.cin_chunks;
  if (const char* e = shf_probe_env("SHF_PROBE_NB")) { int v = atoi(e); if (v >= 2 && v < p.nb) p.nb = v; }
  // (a resident weight tensor may be a single stage: the 64 -> 64 1x1 convolutions of a ResNet bottleneck)
  SHF_REQUIRE(p.nb >= 2 || p.b_resident, "shf_conv_igemm: halo tile of %d bytes leaves no room for the weight ring", p.a_bytes);
  p.tiles_x = ((W + kTW - 1) / kTW + ctas - 1) / ctas;          // tile pairs along x when ctas == 2
  p.tiles_y = (H + kTH - 1) / kTH;
  p.n_tiles = cout / bn;
  p.total_tiles = p.tiles_x * p.tiles_y * p.n_tiles * batch;
  // Accumulator drain period in 64-channel chunks: one (36 k-steps) for 3x3 kernels.  Four chunks on the fast format
  // measured 1-2 % faster in a same-box probe (profiles/r02f_probe_g.txt) but add ~1e-6 of accumulator truncation, which
  // the hf8 operand-model test (3e-6) does not allow: not taken.
  p.chunks_per_phase = (p.taps == 9) ? 1 : 9;
  if (const char* e = shf_probe_env("SHF_PROBE_G")) { int v = atoi(e); if (v >= 1) p.chunks_per_phase = v; }
  p.ctot = out_channels_total;
  p.cout_offset = out_channel_offset;
  p.relu = relu;
  p.in_fmt = in_format;
  {
    // A/B switch for profiling runs.  Same-box alternating runs (profiles/r02_ab_group_rows.txt) put the per-tap regions
    // ~1 % ahead of the per-kernel-row grouping (2048 level 691 vs 672 TFLOP/s, bench 106.5 vs 105.3 images/s), so
    // per-tap is the default and SHF_CONV_GROUP_ROWS=1 selects the grouping.
    static int group_rows = -1;
    if (group_rows < 0) { const char* e = getenv("SHF_CONV_GROUP_ROWS"); group_rows = (e && e[0] == '1') ? 1 : 0; }
    p.group_rows = group_rows;
  }
  p.probe = 0;
  if (const char* e = shf_probe_env("SHF_PROBE_EPI")) p.probe = atoi(e);
  p.out_fmt = out_format;
  p.out_scale = out_scale;
  p.bias = bias;
  p.out = reinterpret_cast<__half*>(out_h2);
  p.plane_elems = (long long)batch * H * W * out_channels_total;
  p.pool_out = reinterpret_cast<__half*>(pool_out_h2);
  p.pool_plane_elems = (long long)batch * (H / 2) * (W / 2) * pool_channels_total;
  p.pool_ctot = pool_channels_total;
  p.pool_coffset = pool_channel_offset;
  p.guard = range_guard;
  p.res = reinterpret_cast<const __half*>(residual);
  p.res_plane_elems = (long long)batch * H * W * res_channels_total;
  p.res_ctot = res_channels_total;
  p.res_coffset = res_channel_offset;
  p.res_fmt = res_format;
  const int smem_bytes = p.na * p.a_bytes + p.nb * p.b_bytes + 1024 + 1536 + 8 * 4096;

  TmapA4 ta4;
  CUtensorMap tb;
  CUtensorMap& ta = ta4.m[0];
  {
    uint64_t d[5] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)batch, 2};
    uint32_t b[5] = {64, (uint32_t)p.xw, (uint32_t)p.xh, 1, 2};
    if (s3) {
      const uint64_t px = (uint64_t)cin * 2, row = px * (uint64_t)in_W, img = row * (uint64_t)in_H;
      const uint64_t bs[4] = {px * 2, row * 2, img, img * (uint64_t)batch};
      for (int py = 0; py < 2; ++py)
        for (int pxl = 0; pxl < 2; ++pxl) {
          // rows py, py + 2, ... and columns pxl, pxl + 2, ... of the input; an empty view (1-pixel input) keeps one
          // zero-filled... cannot be encoded with a zero dim, so H, W >= 2 is required below
          uint64_t dv[5] = {(uint64_t)cin, (uint64_t)((in_W - pxl + 1) / 2), (uint64_t)((in_H - py + 1) / 2), (uint64_t)batch, 2};
          uint8_t* basev = reinterpret_cast<uint8_t*>(const_cast<void*>(in_h2)) + (uint64_t)py * row + (uint64_t)pxl * px;
          if (int e = shf_encode_f16_map(&ta4.m[py * 2 + pxl], basev, 5, dv, b, "parity view", bs)) return e;
        }
    } else if (in_stride == 1) {
      if (int e = shf_encode_f16_map(&ta, const_cast<void*>(in_h2), 5, d, b, "activations")) return e;
      ta4.m[1] = ta4.m[2] = ta4.m[3] = ta;
    } else {
      const uint64_t px = (uint64_t)cin * 2, row = px * (uint64_t)in_W, img = row * (uint64_t)in_H;
      const uint64_t bs[4] = {px * (uint64_t)in_stride, row * (uint64_t)in_stride, img, img * (uint64_t)batch};
      if (int e = shf_encode_f16_map(&ta, const_cast<void*>(in_h2), 5, d, b, "strided activations", bs)) return e;
      ta4.m[1] = ta4.m[2] = ta4.m[3] = ta;
    }
  }
  {
    uint64_t d[4] = {(uint64_t)cin, (uint64_t)cout, (uint64_t)(s3 ? 9 : p.taps), 2};
    uint32_t b[4] = {64, (uint32_t)(bn / ctas), 1, 2};
    if (int e = shf_encode_f16_map(&tb, const_cast<void*>(w_h2), 4, d, b, "weights")) return e;
  }
  const int groups = sm_count() / ctas;
  const int grid = (p.total_tiles < groups ? p.total_tiles : groups) * ctas;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool f8 = in_format == SHF_FMT_HF8;
  if (s3) {
    if (bn == 128) return f8 ? launch_stream<128, 2, SHF_FMT_HF8, false, true>(ta4, tb, p, smem_bytes, grid, st)
                             : launch_stream<128, 2, SHF_FMT_H2, false, true>(ta4, tb, p, smem_bytes, grid, st);
    return f8 ? launch_stream<64, 2, SHF_FMT_HF8, false, true>(ta4, tb, p, smem_bytes, grid, st)
              : launch_stream<64, 2, SHF_FMT_H2, false, true>(ta4, tb, p, smem_bytes, grid, st);
  }
  if (p.res != nullptr) {
    SHF_REQUIRE(ctas == 2, "shf_conv_igemm_res: the residual epilogue exists for the CTA-pair kernel only");
    if (bn == 128) return f8 ? launch_stream<128, 2, SHF_FMT_HF8, true>(ta4, tb, p, smem_bytes, grid, st)
                             : launch_stream<128, 2, SHF_FMT_H2, true>(ta4, tb, p, smem_bytes, grid, st);
    return f8 ? launch_stream<64, 2, SHF_FMT_HF8, true>(ta4, tb, p, smem_bytes, grid, st)
              : launch_stream<64, 2, SHF_FMT_H2, true>(ta4, tb, p, smem_bytes, grid, st);
  }
  if (ctas == 2) {
    if (bn == 128) return f8 ? launch_stream<128, 2, SHF_FMT_HF8>(ta4, tb, p, smem_bytes, grid, st)
                             : launch_stream<128, 2, SHF_FMT_H2>(ta4, tb, p, smem_bytes, grid, st);
    return f8 ? launch_stream<64, 2, SHF_FMT_HF8>(ta4, tb, p, smem_bytes, grid, st)
              : launch_stream<64, 2, SHF_FMT_H2>(ta4, tb, p, smem_bytes, grid, st);
  }
  if (bn == 128) return f8 ? launch_stream<128, 1, SHF_FMT_HF8>(ta4, tb, p, smem_bytes, grid, st)
                           : launch_stream<128, 1, SHF_FMT_H2>(ta4, tb, p, smem_bytes, grid, st);
  return f8 ? launch_stream<64, 1, SHF_FMT_HF8>(ta4, tb, p, smem_bytes, grid, st)
            : launch_stream<64, 1, SHF_FMT_H2>(ta4, tb, p, smem_bytes, grid, st);
}

// ---- C ABI (include/shf_b200.h) ---------------------------------------------------------------------------------
static int g_conv_ctas = 2;

// 8 = CTA pairs (tcgen05 cta_group::2, the product path), 7 = the same kernel with one CTA per tile (kept as the
// regression twin of the pair protocol: same math, no cluster).
extern "C" int shf_set_conv_impl(int impl) {
  SHF_REQUIRE(impl == 7 || impl == 8, "shf_set_conv_impl: %d (7 = single CTA, 8 = CTA pairs)", impl);
  g_conv_ctas = impl == 7 ? 1 : 2;
  return 0;
}

extern "C" int shf_conv_igemm(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H,
                              int W, int cin, int cout, int ksize, int dilation, int out_channels_total,
                              int out_channel_offset, float out_scale, int relu, int in_format, int out_format,
                              unsigned int* range_guard, void* stream) {
  return shf_conv_stream_impl(in_h2, w_h2, bias, out_h2, batch, H, W, cin, cout, ksize, dilation, out_channels_total,
                              out_channel_offset, out_scale, relu, nullptr, 0, 0, g_conv_ctas, in_format, out_format,
                              range_guard, stream);
}

// Convolution + ReLU + 2x2/2 max pooling in one launch.  out_h2 may be NULL when only the pooled map is consumed
// downstream (conv1_2, conv2_2, conv3_3 of VGG16); conv4_3 needs both.
extern "C" int shf_conv_igemm_pool(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, void* pool_out_h2,
                                   int batch, int H, int W, int cin, int cout, int ksize, int dilation,
                                   int out_channels_total, int out_channel_offset, int pool_channels_total,
                                   int pool_channel_offset, float out_scale, int relu, int in_format, int out_format,
                                   unsigned int* range_guard, void* stream) {
  SHF_REQUIRE(pool_out_h2 != nullptr, "shf_conv_igemm_pool: pool_out_h2 is NULL");
  return shf_conv_stream_impl(in_h2, w_h2, bias, out_h2, batch, H, W, cin, cout, ksize, dilation, out_channels_total,
                              out_channel_offset, out_scale, relu, pool_out_h2, pool_channels_total,
                              pool_channel_offset, g_conv_ctas, in_format, out_format, range_guard, stream);
}

// 1x1 convolution with stride s (pad 0): the projection shortcuts / downsampling 1x1 convs of a ResNet bottleneck.  H x W
// are the INPUT dims; the output is ((H - 1) / s + 1) x ((W - 1) / s + 1).  Same kernel, strided TMA view of the input.
extern "C" int shf_conv_igemm_strided(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch,
                                      int H, int W, int stride, int cin, int cout, int out_channels_total,
                                      int out_channel_offset, float out_scale, int relu, int in_format, int out_format,
                                      unsigned int* range_guard, void* stream) {
  SHF_REQUIRE(stride >= 1 && stride <= 8 && H >= 1 && W >= 1, "shf_conv_igemm_strided: stride %d on %dx%d", stride, H, W);
  const int HO = (H - 1) / stride + 1, WO = (W - 1) / stride + 1;
  return shf_conv_stream_impl(in_h2, w_h2, bias, out_h2, batch, HO, WO, cin, cout, 1, 1, out_channels_total,
                              out_channel_offset, out_scale, relu, nullptr, 0, 0, g_conv_ctas, in_format, out_format,
                              range_guard, stream, stride, H, W);
}

// shf_conv_igemm + the residual add of a ResNet block in the epilogue: out = [relu](conv(in) * scale + bias + residual).
// Replaces Convolution (+ BatchNorm + Scale) -> Eltwise SUM (-> ReLU) without writing and re-reading the branch output.
extern "C" int shf_conv_igemm_res(const void* in_h2, const void* w_h2, const float* bias, const void* residual, void* out_h2,
                                  int batch, int H, int W, int cin, int cout, int ksize, int dilation, int out_channels_total,
                                  int out_channel_offset, int res_channels_total, int res_channel_offset, float out_scale,
                                  int relu, int in_format, int res_format, int out_format, unsigned int* range_guard,
                                  void* stream) {
  SHF_REQUIRE(residual != nullptr, "shf_conv_igemm_res: residual is NULL");
  return shf_conv_stream_impl(in_h2, w_h2, bias, out_h2, batch, H, W, cin, cout, ksize, dilation, out_channels_total,
                              out_channel_offset, out_scale, relu, nullptr, 0, 0, g_conv_ctas, in_format, out_format,
                              range_guard, stream, 1, 0, 0, residual, res_channels_total, res_channel_offset, res_format);
}

// 3x3 convolution with stride 2 and pad 1 (stage transitions of torchvision-style ResNets and similar backbones): H x W are
// the INPUT dims (both >= 2), the output is ((H - 1) / 2 + 1) x ((W - 1) / 2 + 1) (conv_layer.cpp:8-28).  w_h2: the 3x3 packing
// of shf_conv_igemm.  The tcgen05 kernel walks (tap, 64-channel chunk) pairs and reads each tap from the parity view of the
// input it lives in (strided TMA maps): no gather pass, no MMAs on pixels that are thrown away.
extern "C" int shf_conv3x3_s2(const void* in_h2, const void* w_h2, const float* bias, void* out_h2, int batch, int H, int W,
                              int cin, int cout, int out_channels_total, int out_channel_offset, float out_scale, int relu,
                              int in_format, int out_format, unsigned int* range_guard, void* stream) {
  SHF_REQUIRE(H >= 2 && W >= 2, "shf_conv3x3_s2: input %dx%d (at least 2 x 2)", H, W);
  const int HO = (H - 1) / 2 + 1, WO = (W - 1) / 2 + 1;
  return shf_conv_stream_impl(in_h2, w_h2, bias, out_h2, batch, HO, WO, cin, cout, 3, 1, out_channels_total,
                              out_channel_offset, out_scale, relu, nullptr, 0, 0, 2, in_format, out_format, range_guard,
                              stream, 2, H, W, nullptr, 0, 0, 0, 1);
}

// az_net_tc.cu — the residual tower on 5th-generation tensor cores (sm_100a only).
//
// One kernel = one 3x3 convolution layer as an implicit GEMM over the padded-row activation matrix
// (layout: az_net.cu header):   D[m, co] = sum_{tap, ci} X[m + off(tap), ci] * W[tap][co][ci]
//   M tile 128 rows (UMMA_M = 128, cta_group::1), N = all output channels (64 / 128 / 256),
//   K = 9 taps x Cin, consumed 64 channels (= one 128-byte swizzle atom of bf16) per pipeline stage.
// Warp roles (192 threads, persistent over M tiles; the leaf count is read from device memory so the
// launch never waits for the host):
//   warp 0   TMA producer: per stage one 2-D box of activations (128 rows x 64 ch, rows shifted by the
//            tap offset) and one of weights (N rows x 64 ch), SWIZZLE_128B, completing on an mbarrier
//   warp 1   allocates TMEM, then one elected lane issues tcgen05.mma.kind::f16 (bf16 x bf16 -> f32)
//            into one of two TMEM accumulator stages and commits stage-free / accumulator-full barriers
//   warps 2-5 epilogue: tcgen05.ld the accumulator (each warp owns its 32-lane TMEM quarter), add the
//            folded-BN bias, the residual, ReLU, force the padding rows to zero, store bf16 rows
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <thread>

#include "az_net_impl.h"
#include "az_net_kernels.cuh"

#ifndef AZ_TC_MODE_DEFAULT
#define AZ_TC_MODE_DEFAULT 5
#endif
#define TC_STAGES 4
#define TC_BM 128
#define TC_BK 64
#define TC_THREADS 192

struct AzNetTc {
  CUtensorMap map_in, map_x, map_mid;
  CUtensorMap hmap_in[2], hmap_x[2], hmap_mid[2];  // halo kernel: [0] 256-row box, [1] tail box (AR-256 rows)
  int pdl = 1;             // AZ_PDL (default 1): conv layers 1.. are launched with programmatic stream serialization
  int split = 0;           // AZ_NET_BF16X3: activation rows [hi | lo | hi] (3 C channels), weights [W_hi | W_hi | W_lo]
  int mode = 0;            // AZ_TC_MODE: 0 = one TMA box per tap, 1/2 = halo tile + row-shifted descriptors (base-offset variants)
  int halo = 0, AR = 0;
  int res_l2 = 0;
  size_t halo_smem = 0;
  // dense-x kernel (mode 5): 3-D maps [board row][x][channel] of the three activation buffers, tile geometry
  CUtensorMap xmap_in, xmap_x, xmap_mid;
  int x_TH = 0, x_sub_bytes = 0, x_sub_stride = 0;
  size_t x_smem = 0;
  std::vector<CUtensorMap> map_w;
  std::vector<CUtensorMap> map_w_half;  // pair kernel: box of cout/2 weight rows
  std::vector<__nv_bfloat16*> w_dev;
  int cin0 = 64;
  size_t smem_bytes = 0;
  int num_sms = 148;
};

struct TcLayer {
  int cin, cout;       // channels (cin multiple of 64)
  int relu, has_res;
  int Wr, Hc, RP, guard;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: rows 128 B apart, 8-row atoms 1024 B apart (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// issue only: the registers may be read after tc_wait_ld()
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const float* __restrict__ bias,
          const __nv_bfloat16* res, __nv_bfloat16* out, const int32_t* __restrict__ n_rows, TcLayer L) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: stages of {A 16 KB, B cout*128 B}, then barriers
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TC_BM * TC_BK * 2, b_bytes = (uint32_t)L.cout * TC_BK * 2, stage_bytes = a_bytes + b_bytes;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)TC_STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* tfull_bar = empty_bar + TC_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long M = (long long)(*n_rows) * L.RP;
  const int num_tiles = (int)((M + TC_BM - 1) / TC_BM);
  const int kc = L.cin / TC_BK, KS = 9 * kc;
  uint32_t tmem_cols = 2 * (uint32_t)L.cout;  // two accumulator stages
  tmem_cols = tmem_cols <= 32 ? 32 : tmem_cols <= 64 ? 64 : tmem_cols <= 128 ? 128 : tmem_cols <= 256 ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int row0 = L.guard + t * TC_BM;
        for (int ks = 0; ks < KS; ++ks) {
          const int tap = ks / kc, kk = ks - tap * kc;
          const int toff = (tap / 3 - 1) * L.Wr + (tap % 3 - 1);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          unsigned char* sa = smem + (size_t)stage * stage_bytes;
          tma_load_2d(sa, &map_a, &full_bar[stage], kk * TC_BK, row0 + toff);
          tma_load_2d(sa + a_bytes, &map_b, &full_bar[stage], kk * TC_BK, tap * L.cout);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(L.cout >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * L.cout);
        for (int ks = 0; ks < KS; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = tc_smem_desc(sa), bdesc = tc_smem_desc(sa + a_bytes);
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
            // advance 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (>>4) start-address field
            tc_mma(d_tmem, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), idesc, (ks | k4) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      const long long m = (long long)t * TC_BM + q * 32 + lane;
      const int r = (int)(m % L.RP);
      const int yy = r / L.Wr, xx = r - yy * L.Wr;
      const bool valid = (m < M) && yy < L.Hc && xx < L.Hc;
      const size_t grow = ((size_t)L.guard + (size_t)m) * L.cout;
      for (int c0 = 0; c0 < L.cout; c0 += 32) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * L.cout + c0), v);
        if (m < M) {
          __align__(16) __nv_bfloat16 o[32];
          if (valid) {
            __align__(16) __nv_bfloat16 rr[32];
            if (L.has_res) {
              const uint4* rp = reinterpret_cast<const uint4*>(res + grow + c0);
#pragma unroll
              for (int k = 0; k < 4; ++k) reinterpret_cast<uint4*>(rr)[k] = rp[k];
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float f = __uint_as_float(v[j]) + bias[c0 + j];
              if (L.has_res) f += __bfloat162float(rr[j]);
              if (L.relu) f = fmaxf(f, 0.f);
              o[j] = __float2bfloat16(f);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float2bfloat16(0.f);
          }
          uint4* op = reinterpret_cast<uint4*>(out + grow + c0);
#pragma unroll
          for (int k = 0; k < 4; ++k) op[k] = reinterpret_cast<const uint4*>(o)[k];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------
// Halo variant (default for <= 128 filters): the activation tile (256 output rows + one board row + 1 of
// halo on each side, all input channels) is loaded ONCE per tile; the nine taps are row-shifted views of
// it, addressed by moving the UMMA shared-memory descriptor's start address by whole 128-byte rows (the
// hardware applies the 128B swizzle on absolute shared-memory address bits, so an unaligned start needs
// no base offset — measured on B200: base_offset = 0 is exact, (addr>>7)&7 is wrong).  Weights stream
// through a 3-stage ring, each 16 KB chunk feeding 8 MMAs (two 128-row halves x four K=16 steps).
// L2->smem operand traffic per 128 output rows: ~180 KB instead of 576 KB for the per-tap kernel.
// Epilogue: 8 warps (two per TMEM lane quarter, splitting the columns); residual rows are fetched and
// results written back through a per-warp shared-memory staging tile so that every global access is a
// full 128-byte line (the naive one-row-per-lane pattern issues 32 partial sectors per instruction).
#define H_BSTAGES 4
#define H_EPI_WARPS 8
#define H_MMA_WARPS 2
#define H_THREADS (32 * (1 + H_MMA_WARPS + H_EPI_WARPS))

struct HaloLayer {
  int cin, cout, relu, has_res;
  int Wr, Hc, RP, guard;
  int halo, AR;   // halo rows on each side; rows of the staged tile (multiple of 8)
  int bo_mode;    // experiment switch: 1 puts (start >> 7) & 7 into the descriptor's base-offset field (wrong on B200)
  int res_l2;     // 1: the producer asks TMA to pull the residual rows of the NEXT tile into L2 (AZ_TC_RESPF)
};


// ---- 2-CTA (cta_group::2) helpers ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (same shared-memory offset, CTA-rank bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(x), "r"(y)
               : "memory");
}
// L2 prefetch of a tensor-map box (no shared-memory destination, no barrier): a hint, results never depend on it
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit that arrives on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <bool PAIR>
__global__ void __launch_bounds__(H_THREADS, 1)
k_conv_tc_halo(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_r, const float* __restrict__ bias, const __nv_bfloat16* res, __nv_bfloat16* out, const int32_t* __restrict__ n_rows, HaloLayer L) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kc = L.cin / TC_BK;
  const uint32_t a_chunk = (uint32_t)L.AR * 128u, a_buf = (uint32_t)kc * a_chunk;
  const uint32_t a_buf_max = 2u * a_chunk;  // buffers are carved for cin = 128
  const uint32_t b_bytes = (uint32_t)(PAIR ? L.cout / 2 : L.cout) * TC_BK * 2;  // a pair splits every weight chunk along N
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  constexpr int NB = PAIR ? 2 * H_BSTAGES : H_BSTAGES;                               // same bytes, twice the depth
  const int cw = L.cout / 2;                 // columns per epilogue warp, processed 32 at a time
  const uint32_t srow = 32 * 2 + 16;         // staging row pitch (bytes): 32 bf16 + 16 keeps 16-byte accesses conflict-free
  unsigned char* smA = smem;
  unsigned char* smB = smem + 2 * (size_t)a_buf_max;
  unsigned char* smS = smB + (size_t)NB * b_bytes;
  float* s_bias = (float*)(smS + (size_t)H_EPI_WARPS * 32 * srow);
  uint64_t* a_full = (uint64_t*)(s_bias + L.cout);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + NB;
  uint64_t* tfull_bar = b_empty + NB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long M = (long long)(*n_rows) * L.RP;
  const int num_tiles256 = (int)((M + 255) / 256);
  // work unit: one 256-row tile per CTA; a pair takes two consecutive tiles (its CTAs differ in `crank`)
  const int num_units = PAIR ? (num_tiles256 + 1) / 2 : num_tiles256;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ustep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  uint32_t tmem_cols = 4 * (uint32_t)L.cout;  // 2 halves x 2 accumulator stages
  tmem_cols = tmem_cols <= 32 ? 32 : tmem_cols <= 64 ? 64 : tmem_cols <= 128 ? 128 : tmem_cols <= 256 ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], H_MMA_WARPS);
      mbar_init(&tfull_bar[s], H_MMA_WARPS);
      mbar_init(&tempty_bar[s], PAIR ? 2 * H_EPI_WARPS : H_EPI_WARPS);
    }
    for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], H_MMA_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = threadIdx.x; c < L.cout; c += blockDim.x) s_bias[c] = bias[c];
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers must exist before remote arrives / 2-SM TMA completions target them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer ======================================================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      int bstage = 0, it = 0;
      uint32_t bphase = 0;
      // the activation tile of iteration `j` (tile index tt) goes to buffer j&1; it is requested half a tile ahead
      // u = work unit; this CTA's 256-row tile is u (single) or 2u + crank (pair).  In a pair only the leader arms the
      // "full" barriers (with the bytes of BOTH CTAs); each CTA issues its own loads, credited to the leader's barrier.
      auto load_a = [&](int j, int u) {
        const int buf = j & 1;
        const int tt = PAIR ? 2 * u + (int)crank : u;
        const int g0 = L.guard + tt * 256 - L.halo;
        mbar_wait(&a_empty[buf], ((j >> 1) & 1) ^ 1);
        if (!PAIR) mbar_expect_tx(&a_full[buf], a_buf);
        else if (crank == 0) mbar_expect_tx(&a_full[buf], 2 * a_buf);
        for (int kk = 0; kk < kc; ++kk) {
          unsigned char* dst = smA + (size_t)buf * a_buf_max + (size_t)kk * a_chunk;
          if (PAIR) {
            tma_load_2d_pair(dst, &map_a0, &a_full[buf], kk * TC_BK, g0);
            tma_load_2d_pair(dst + 256 * 128, &map_a1, &a_full[buf], kk * TC_BK, g0 + 256);
          } else {
            tma_load_2d(dst, &map_a0, &a_full[buf], kk * TC_BK, g0);
            tma_load_2d(dst + 256 * 128, &map_a1, &a_full[buf], kk * TC_BK, g0 + 256);
          }
        }
      };
      // residual rows (256 x cout of the block input) of a tile, pulled into L2 one tile before the epilogue's register
      // prefetch asks for them: the epilogue's dependent loads then see L2 latency instead of HBM latency
      auto res_to_l2 = [&](int u) {
        const int tt = PAIR ? 2 * u + (int)crank : u;
        for (int c = 0; c < L.cout; c += TC_BK) tma_prefetch_l2_2d(&map_r, c, L.guard + tt * 256);
      };
      if (unit0 < num_units) load_a(0, unit0);
      for (int u = unit0; u < num_units; u += ustep, ++it) {
        if (L.res_l2 && u + ustep < num_units) res_to_l2(u + ustep);
        for (int tap = 0; tap < 9; ++tap) {
          if (tap == 3 && u + ustep < num_units) load_a(it + 1, u + ustep);
          for (int kk = 0; kk < kc; ++kk) {
            mbar_wait(&b_empty[bstage], bphase ^ 1);
            if (PAIR) {
              if (crank == 0) mbar_expect_tx(&b_full[bstage], 2 * b_bytes);
              tma_load_2d_pair(smB + (size_t)bstage * b_bytes, &map_b, &b_full[bstage], kk * TC_BK, tap * L.cout + (int)crank * (L.cout / 2));
            } else {
              mbar_expect_tx(&b_full[bstage], b_bytes);
              tma_load_2d(smB + (size_t)bstage * b_bytes, &map_b, &b_full[bstage], kk * TC_BK, tap * L.cout);
            }
            if (++bstage == NB) { bstage = 0; bphase ^= 1; }
          }
        }
      }
    }
  } else if (warp <= H_MMA_WARPS) {
   if (!PAIR || crank == 0) {
    // ===== MMA issuers: warp 1 owns the upper 128 rows of the tile, warp 2 the lower 128 ==========
    // The whole warp runs the loop (uniform control flow, descriptors precomputed); one elected lane issues.
    const int h = warp - 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(L.cout >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO | version 1 | SWIZZLE_128B  (bits 32..63)
    uint32_t b_lo[NB];
#pragma unroll
    for (int s2 = 0; s2 < NB; ++s2) b_lo[s2] = ((smem_u32(smB + (size_t)s2 * b_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
    const bool leader = elect_one();
    int bstage = 0, it = 0;
    uint32_t bphase = 0;
    for (int u = unit0; u < num_units; u += ustep, ++it) {
      const int buf = it & 1, acc = it & 1;
      mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
      mbar_wait(&a_full[buf], (it >> 1) & 1);
      tc_fence_after();
      // start-address field (>>4) of output row 0 of this half, tap (0,0), channel chunk 0
      const uint32_t a_lo0 = (((smem_u32(smA + (size_t)buf * a_buf_max) + (uint32_t)(h * 128 + L.halo) * 128u) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t d_tmem = tmem_base + (uint32_t)((acc * 2 + h) * L.cout);
      for (int tap = 0; tap < 9; ++tap) {
        const int toff8 = ((tap / 3 - 1) * L.Wr + (tap % 3 - 1)) * 8;  // rows -> 16-byte units
        for (int kk = 0; kk < kc; ++kk) {
          mbar_wait(&b_full[bstage], bphase);
          tc_fence_after();
          if (leader) {
            const uint32_t alo = (uint32_t)((int)a_lo0 + toff8) + (uint32_t)kk * (a_chunk >> 4);
            const uint32_t blo = b_lo[bstage];
#pragma unroll
            for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(alo + (uint32_t)k4 * 2u);
              const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(blo + (uint32_t)k4 * 2u);
              if (PAIR) tc_mma_pair(d_tmem, ad, bd, idesc, (tap | kk | k4) != 0 ? 1u : 0u);
              else tc_mma(d_tmem, ad, bd, idesc, (tap | kk | k4) != 0 ? 1u : 0u);
            }
            if (PAIR) tc_commit_pair(&b_empty[bstage]);
            else tc_commit(&b_empty[bstage]);
          }
          __syncwarp();
          if (++bstage == NB) { bstage = 0; bphase ^= 1; }
        }
      }
      if (leader) {
        if (PAIR) { tc_commit_pair(&tfull_bar[acc]); tc_commit_pair(&a_empty[buf]); }
        else { tc_commit(&tfull_bar[acc]); tc_commit(&a_empty[buf]); }
      }
      __syncwarp();
    }
   }
  } else {
    // ===== epilogue: 8 warps, two per TMEM lane quarter (column halves) ==========================
    // A pass = 32 rows x 32 columns: residual (prefetched into registers one pass ahead) -> staging,
    // tcgen05.ld, bias/residual/ReLU, bf16 -> staging, staging -> global.  Global accesses are 64-byte
    // row segments (two full sectors), 8 rows per instruction.
    const int ew = warp - (1 + H_MMA_WARPS);  // 0..7
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int ch = ew >> 2;                   // column half
    const int col0 = ch * cw;
    const int npass_c = cw / 32;              // column passes per half-tile (2 for 128 filters, 1 for 64)
    unsigned char* stage = smS + (size_t)ew * 32 * srow;
    const int crow = lane >> 2, cchunk = lane & 3;  // coalesced phases: 4 lanes x 16 B per row, 8 rows per instruction
    uint4 pre[4];
    auto prefetch = [&](long long m_base, int cc) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long mr = m_base + i * 8 + crow;
        pre[i] = make_uint4(0, 0, 0, 0);
        if (mr < M) pre[i] = *reinterpret_cast<const uint4*>(res + ((size_t)L.guard + (size_t)mr) * L.cout + cc + cchunk * 8);
      }
    };
    int it = 0;
    const int tstep = PAIR ? 2 * ustep : ustep;  // distance between this CTA's consecutive 256-row tiles
    const int t_first = PAIR ? 2 * unit0 + (int)crank : unit0;
    if (L.has_res && L.bo_mode != 3 && unit0 < num_units) prefetch((long long)t_first * 256 + q * 32, col0);
    for (int u = unit0; u < num_units; u += ustep, ++it) {
      const int t = PAIR ? 2 * u + (int)crank : u;
      const int acc = it & 1;
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ps = 0; ps < 2 * npass_c; ++ps) {
        const int h = ps / npass_c, pc = ps - h * npass_c;
        const int cc = col0 + pc * 32;
        const long long m_base = (long long)t * 256 + h * 128 + q * 32;
        if (L.bo_mode == 3) {
          // experiment: no shared-memory staging, every lane reads / writes its own 64-byte row segment directly
          const long long m = m_base + lane;
          const int r = (int)(m % L.RP);
          const int yy = r / L.Wr, xx = r - yy * L.Wr;
          const bool valid = (m < M) && yy < L.Hc && xx < L.Hc;
          const size_t grow = ((size_t)L.guard + (size_t)m) * L.cout + cc;
          __align__(16) __nv_bfloat16 rr[32];
          if (L.has_res && valid) {
#pragma unroll
            for (int k = 0; k < 4; ++k) reinterpret_cast<uint4*>(rr)[k] = reinterpret_cast<const uint4*>(res + grow)[k];
          }
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * 2 + h) * L.cout + cc), v);
          if (m < M) {
            __align__(16) __nv_bfloat16 o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float f = __uint_as_float(v[j]) + s_bias[cc + j];
              if (L.has_res) f += __bfloat162float(rr[j]);
              if (L.relu) f = fmaxf(f, 0.f);
              o[j] = __float2bfloat16(valid ? f : 0.f);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) reinterpret_cast<uint4*>(out + grow)[k] = reinterpret_cast<const uint4*>(o)[k];
          }
          continue;
        }
        uint32_t v[32];  // accumulator row segment of this lane: requested now, consumed after the residual staging below
        tc_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * 2 + h) * L.cout + cc), v);
        if (L.has_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stage + (size_t)(i * 8 + crow) * srow + cchunk * 16) = pre[i];
          // next pass of this tile, or the first pass of this CTA's next tile
          const int nps = ps + 1;
          if (nps < 2 * npass_c) {
            const int nh = nps / npass_c, npc = nps - nh * npass_c;
            prefetch((long long)t * 256 + nh * 128 + q * 32, col0 + npc * 32);
          } else if (u + ustep < num_units) {
            prefetch((long long)(t + tstep) * 256 + q * 32, col0);
          }
        }
        __syncwarp();
        const long long m = m_base + lane;
        const int r = (int)(m % L.RP);
        const int yy = r / L.Wr, xx = r - yy * L.Wr;
        const bool valid = (m < M) && yy < L.Hc && xx < L.Hc;
        unsigned char* myrow = stage + (size_t)lane * srow;
        {
          tc_wait_ld();
          __align__(16) __nv_bfloat16 o[32];
          if (valid) {
            __align__(16) __nv_bfloat16 rr[32];
            if (L.has_res) {
#pragma unroll
              for (int k = 0; k < 4; ++k) reinterpret_cast<uint4*>(rr)[k] = *reinterpret_cast<const uint4*>(myrow + k * 16);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float f = __uint_as_float(v[j]) + s_bias[cc + j];
              if (L.has_res) f += __bfloat162float(rr[j]);
              if (L.relu) f = fmaxf(f, 0.f);
              o[j] = __float2bfloat16(f);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float2bfloat16(0.f);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(myrow + k * 16) = reinterpret_cast<const uint4*>(o)[k];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long long mr = m_base + i * 8 + crow;
          if (mr < M)
            *reinterpret_cast<uint4*>(out + ((size_t)L.guard + (size_t)mr) * L.cout + cc + cchunk * 8) =
                *reinterpret_cast<const uint4*>(stage + (size_t)(i * 8 + crow) * srow + cchunk * 16);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_remote(&tempty_bar[acc], 0);  // the leader's MMA warps wait for both CTAs' drains
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------
// Dense-x variant (AZ_TC_MODE=5, the default for <= 128 filters): the activation rows of a leaf are y*Hc + x with ONE zero board row per leaf
// and NO separator column (90 rows per 81 positions at 9x9 instead of 100): 10 % fewer MMA rows than the halo kernel.
// Without a separator column the dx = -1 / +1 taps cannot be row shifts of the same tile (they would wrap around the board
// edge), so TMA supplies them: the activation matrix is described as a 3-D tensor [board row][x][channel] and the tile is
// loaded three times with the x coordinate starting at -1, 0, +1 — the out-of-bounds column is filled with zeros by the
// hardware, which is exactly the padding a 3x3 convolution needs.  The dy taps stay row-shifted UMMA descriptors (+-Hc rows)
// over each copy.  Sub-tiles (one dx copy of one 64-channel chunk, NBR board rows = TH + 2 Hc rows and change) stream through
// a 4-slot ring; each feeds 3 taps x 4 K-steps x 2 halves = 24 MMAs.  A tile is TH = floor(256 / Hc) * Hc output rows, so
// every tile starts on a board-row boundary and both CTAs of a pair use the same descriptor offsets; the 256 - TH surplus
// rows of the M = 256 MMA are computed and dropped.
#define X_ASLOTS 4
#define X_BSTAGES 6

struct XLayer {
  int cin, cout, relu, has_res;
  int Hc, RP, guard;
  int TH;          // output rows per tile (multiple of Hc, <= 256)
  int sub_bytes;   // bytes of one TMA box = NBR * Hc * 128
  int sub_stride;  // the same rounded up to 1024 (swizzle atom alignment of the next slot)
  int bstages;     // weight ring depth: X_BSTAGES for a pair (half chunks), half of it for a single CTA (same bytes)
  int ostride;     // elements per activation row of `out` / `res`: cout, or 3 * cout for the split-bf16 rows [hi | lo | hi]
  int ksteps;      // K = 16 steps per 64-channel chunk that carry data: 4, or 2 for the input layer (17 planes in channels 0..31)
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// SPLIT (AZ_NET_BF16X3): every fp32 value travels as two bf16 terms, value = hi + lo (16 significand bits).  Activation rows
// are [hi(C) | lo(C) | hi(C)] and the weights of a layer are packed [W_hi | W_hi | W_lo] along K, so the unchanged main loop
// accumulates a_hi*w_hi + a_lo*w_hi + a_hi*w_lo in fp32 (the dropped a_lo*w_lo term is 2^-16 relative): 3x the MMAs of the
// bf16 tower, ~2^-16 instead of 2^-8 relative error per layer.  Only the epilogue differs: it splits its fp32 result again.
template <bool PAIR, bool SPLIT>
__global__ void __launch_bounds__(H_THREADS, 1)
k_conv_tc_x(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const float* __restrict__ bias,
            const __nv_bfloat16* res, __nv_bfloat16* out, const int32_t* __restrict__ n_rows, XLayer L) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kc = L.cin / TC_BK;
  const uint32_t b_bytes = (uint32_t)(PAIR ? L.cout / 2 : L.cout) * TC_BK * 2;  // a pair splits every weight chunk along N
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const int cw = L.cout / 2;                 // columns per epilogue warp, processed 32 at a time
  const uint32_t srow = 32 * 2 + 16;         // staging row pitch (bytes)
  unsigned char* smA = smem;
  unsigned char* smB = smem + (size_t)X_ASLOTS * L.sub_stride;
  unsigned char* smS = smB + (size_t)L.bstages * b_bytes;
  float* s_bias = (float*)(smS + (size_t)H_EPI_WARPS * 32 * srow);
  uint64_t* a_full = (uint64_t*)(s_bias + L.cout);
  uint64_t* a_empty = a_full + X_ASLOTS;
  uint64_t* b_full = a_empty + X_ASLOTS;
  uint64_t* b_empty = b_full + X_BSTAGES;
  uint64_t* tfull_bar = b_empty + X_BSTAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long M = (long long)(*n_rows) * L.RP;
  const int num_tiles = (int)((M + L.TH - 1) / L.TH);
  const int num_units = PAIR ? (num_tiles + 1) / 2 : num_tiles;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ustep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  uint32_t tmem_cols = 4 * (uint32_t)L.cout;  // 2 halves x 2 accumulator stages
  tmem_cols = tmem_cols <= 32 ? 32 : tmem_cols <= 64 ? 64 : tmem_cols <= 128 ? 128 : tmem_cols <= 256 ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int s = 0; s < X_ASLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], H_MMA_WARPS); }
    for (int s = 0; s < L.bstages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], H_MMA_WARPS); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], H_MMA_WARPS); mbar_init(&tempty_bar[s], PAIR ? 2 * H_EPI_WARPS : H_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = threadIdx.x; c < L.cout; c += blockDim.x) s_bias[c] = bias[c];
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch (AZ_PDL, default on): the next layer's CTAs may be placed as soon as this grid's CTAs leave their SMs, and
  // this grid's set-up above (barriers, TMEM, cluster handshake, bias) overlapped the previous layer's tail; nothing written by the
  // previous layer is touched before this point.  Both instructions are no-ops for a launch without the attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer: sub-tiles (kk, dx) and weight chunks (kk, dx, dy) in the order the MMA warps consume them =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      int bstage = 0;
      uint32_t bphase = 0, ai = 0;
      for (int u = unit0; u < num_units; u += ustep) {
        const int tt = PAIR ? 2 * u + (int)crank : u;
        const int br0 = (L.guard + tt * L.TH) / L.Hc - 1;  // first board row of the box: one above the tile
        for (int kk = 0; kk < kc; ++kk) {
          for (int dxi = 0; dxi < 3; ++dxi, ++ai) {
            const uint32_t slot = ai % X_ASLOTS, par = (ai / X_ASLOTS) & 1u;
            mbar_wait(&a_empty[slot], par ^ 1u);
            unsigned char* dst = smA + (size_t)slot * L.sub_stride;
            if (PAIR) {
              if (crank == 0) mbar_expect_tx(&a_full[slot], 2u * (uint32_t)L.sub_bytes);
              tma_load_3d_pair(dst, &map_a, &a_full[slot], kk * TC_BK, dxi - 1, br0);
            } else {
              mbar_expect_tx(&a_full[slot], (uint32_t)L.sub_bytes);
              tma_load_3d(dst, &map_a, &a_full[slot], kk * TC_BK, dxi - 1, br0);
            }
            for (int dyi = 0; dyi < 3; ++dyi) {
              const int tap = dyi * 3 + dxi;
              mbar_wait(&b_empty[bstage], bphase ^ 1);
              if (PAIR) {
                if (crank == 0) mbar_expect_tx(&b_full[bstage], 2 * b_bytes);
                tma_load_2d_pair(smB + (size_t)bstage * b_bytes, &map_b, &b_full[bstage], kk * TC_BK, tap * L.cout + (int)crank * (L.cout / 2));
              } else {
                mbar_expect_tx(&b_full[bstage], b_bytes);
                tma_load_2d(smB + (size_t)bstage * b_bytes, &map_b, &b_full[bstage], kk * TC_BK, tap * L.cout);
              }
              if (++bstage == L.bstages) { bstage = 0; bphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp <= H_MMA_WARPS) {
   if (!PAIR || crank == 0) {
    // ===== MMA issuers: warp 1 owns rows 0..127 of the tile, warp 2 rows 128..255 =================================
    const int h = warp - 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(L.cout >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO | version 1 | SWIZZLE_128B  (bits 32..63)
    uint32_t b_lo[X_BSTAGES];
#pragma unroll
    for (int s2 = 0; s2 < X_BSTAGES; ++s2) b_lo[s2] = ((smem_u32(smB + (size_t)s2 * b_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
    // output row 0 of the tile is row Hc of every box (the box starts one board row above); (>>4) units: 8 per 128-byte row
    const uint32_t row0_16 = (uint32_t)(L.Hc + h * 128) * 8u;
    const bool leader = elect_one();
    int bstage = 0, it = 0;
    uint32_t bphase = 0, ai = 0;
    for (int u = unit0; u < num_units; u += ustep, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)((acc * 2 + h) * L.cout);
      for (int kk = 0; kk < kc; ++kk) {
        for (int dxi = 0; dxi < 3; ++dxi, ++ai) {
          const uint32_t slot = ai % X_ASLOTS, par = (ai / X_ASLOTS) & 1u;
          mbar_wait(&a_full[slot], par);
          tc_fence_after();
          const uint32_t a_lo0 = (((smem_u32(smA + (size_t)slot * L.sub_stride) & 0x3FFFFu) >> 4) + row0_16) | (1u << 16);
          for (int dyi = 0; dyi < 3; ++dyi) {
            mbar_wait(&b_full[bstage], bphase);
            tc_fence_after();
            if (leader) {
              const uint32_t alo = (uint32_t)((int)a_lo0 + (dyi - 1) * L.Hc * 8);
              const uint32_t blo = b_lo[bstage];
#pragma unroll
              for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                if (k4 < L.ksteps) {  // uniform: the input layer's channels 32..63 are zero, its chunk needs two steps only
                  const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(alo + (uint32_t)k4 * 2u);
                  const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(blo + (uint32_t)k4 * 2u);
                  const uint32_t accum = (kk | dxi | dyi | k4) != 0 ? 1u : 0u;
                  if (PAIR) tc_mma_pair(d_tmem, ad, bd, idesc, accum);
                  else tc_mma(d_tmem, ad, bd, idesc, accum);
                }
              }
              if (PAIR) tc_commit_pair(&b_empty[bstage]);
              else tc_commit(&b_empty[bstage]);
            }
            __syncwarp();
            if (++bstage == L.bstages) { bstage = 0; bphase ^= 1; }
          }
          if (leader) {
            if (PAIR) tc_commit_pair(&a_empty[slot]);
            else tc_commit(&a_empty[slot]);
          }
          __syncwarp();
        }
      }
      if (leader) {
        if (PAIR) tc_commit_pair(&tfull_bar[acc]);
        else tc_commit(&tfull_bar[acc]);
      }
      __syncwarp();
    }
   }
  } else {
    // ===== epilogue: as in k_conv_tc_halo; rows >= TH of the M = 256 tile belong to the next tile and are dropped =====
    const int ew = warp - (1 + H_MMA_WARPS);  // 0..7
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int ch = ew >> 2;                   // column half
    const int col0 = ch * cw;
    const int npass_c = cw / 32;
    unsigned char* stage = smS + (size_t)ew * 32 * srow;
    const int crow = lane >> 2, cchunk = lane & 3;
    uint4 pre[4];
    // rows of a pass: tile row l0 + i*8 + crow (coalesced phases) / l0 + lane (TMEM phase)
    auto prefetch = [&](long long tile_row0, int l0, int cc) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int l = l0 + i * 8 + crow;
        const long long mr = tile_row0 + l;
        pre[i] = make_uint4(0, 0, 0, 0);
        if (l < L.TH && mr < M) pre[i] = *reinterpret_cast<const uint4*>(res + ((size_t)L.guard + (size_t)mr) * L.cout + cc + cchunk * 8);
      }
    };
    int it = 0;
    const int tstep = PAIR ? 2 * ustep : ustep;
    const int t_first = PAIR ? 2 * unit0 + (int)crank : unit0;
    if (!SPLIT && L.has_res && unit0 < num_units) prefetch((long long)t_first * L.TH, q * 32, col0);
    for (int u = unit0; u < num_units; u += ustep, ++it) {
      const int t = PAIR ? 2 * u + (int)crank : u;
      const long long tile_row0 = (long long)t * L.TH;
      const int acc = it & 1;
      // which of this lane's two rows of the tile (h = 0, 1) hold an output: once per tile, in 32-bit arithmetic (the row index
      // of the whole batch fits: checked at create), instead of a 64-bit modulo and a division in every pass
      uint32_t vmask = 0;  // bit h: row h * 128 + q * 32 + lane of the tile holds an output
      {
        const uint32_t lrow = (uint32_t)(q * 32 + lane);
        const uint32_t mrow = (uint32_t)tile_row0 + lrow;
        const uint32_t r0 = mrow % (uint32_t)L.RP;
        const uint32_t r1 = (r0 + 128u) % (uint32_t)L.RP;
        const uint32_t hw = (uint32_t)(L.Hc * L.Hc);  // rows >= Hc*Hc of a leaf: its zero board row
        if (lrow < (uint32_t)L.TH && (long long)mrow < M && r0 < hw) vmask |= 1u;
        if (lrow + 128u < (uint32_t)L.TH && (long long)mrow + 128 < M && r1 < hw) vmask |= 2u;
      }
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      bool arrived = false;
#pragma unroll 1
      for (int ps = 0; ps < 2 * npass_c; ++ps) {
        const int h = ps >= npass_c ? 1 : 0, pc = ps - h * npass_c;
        const int cc = col0 + pc * 32;
        const int l0 = h * 128 + q * 32;
        if constexpr (SPLIT) {
          // every lane owns one row: 64-byte segments of the hi / lo / hi column groups, read and written directly
          const int l = l0 + lane;
          const long long m = tile_row0 + l;
          const bool inb = (l < L.TH) && (m < M);
          const bool valid = ((vmask >> h) & 1u) != 0;
          const size_t grow = ((size_t)L.guard + (size_t)(inb ? m : 0)) * L.ostride + cc;
          __align__(16) __nv_bfloat16 rh[32], rl[32];
          if (L.has_res && valid) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              reinterpret_cast<uint4*>(rh)[k] = reinterpret_cast<const uint4*>(res + grow)[k];
              reinterpret_cast<uint4*>(rl)[k] = reinterpret_cast<const uint4*>(res + grow + L.cout)[k];
            }
          }
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * 2 + h) * L.cout + cc), v);
          if (inb) {
            __align__(16) __nv_bfloat16 oh[32], ol[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float f = __uint_as_float(v[j]) + s_bias[cc + j];
              if (L.has_res && valid) f += __bfloat162float(rh[j]) + __bfloat162float(rl[j]);
              if (L.relu) f = fmaxf(f, 0.f);
              if (!valid) f = 0.f;
              const __nv_bfloat16 hb = __float2bfloat16(f);
              oh[j] = hb;
              ol[j] = __float2bfloat16(f - __bfloat162float(hb));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              reinterpret_cast<uint4*>(out + grow)[k] = reinterpret_cast<const uint4*>(oh)[k];
              reinterpret_cast<uint4*>(out + grow + L.cout)[k] = reinterpret_cast<const uint4*>(ol)[k];
              reinterpret_cast<uint4*>(out + grow + 2 * L.cout)[k] = reinterpret_cast<const uint4*>(oh)[k];
            }
          }
          continue;
        }
        uint32_t v[32];
        tc_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * 2 + h) * L.cout + cc), v);
        if (L.has_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stage + (size_t)(i * 8 + crow) * srow + cchunk * 16) = pre[i];
          const int nps = ps + 1;
          if (nps < 2 * npass_c) {
            const int nh = nps >= npass_c ? 1 : 0, npc = nps - nh * npass_c;
            prefetch(tile_row0, nh * 128 + q * 32, col0 + npc * 32);
          } else if (u + ustep < num_units) {
            prefetch((long long)(t + tstep) * L.TH, q * 32, col0);
          }
        }
        __syncwarp();
        const bool valid = ((vmask >> h) & 1u) != 0;
        unsigned char* myrow = stage + (size_t)lane * srow;
        {
          tc_wait_ld();
          if (ps == 2 * npass_c - 1) {
            // the accumulator stage is free as soon as its last TMEM read has landed in registers: hand it back to the MMA warps
            // BEFORE this pass's arithmetic and global stores (the release of a remote arrive otherwise waits for those stores)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR && crank != 0) mbar_arrive_remote(&tempty_bar[acc], 0);
              else mbar_arrive(&tempty_bar[acc]);
            }
            arrived = true;
          }
          __align__(16) __nv_bfloat16 o[32];
          if (valid) {
            __align__(16) __nv_bfloat16 rr[32];
            if (L.has_res) {
#pragma unroll
              for (int k = 0; k < 4; ++k) reinterpret_cast<uint4*>(rr)[k] = *reinterpret_cast<const uint4*>(myrow + k * 16);
            }
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + cc);  // 16-byte aligned: s_bias follows 16-byte multiples, cc % 32 == 0
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bq = b4[j4];
              const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = j4 * 4 + jj;
                float f = __uint_as_float(v[j]) + bb[jj];
                if (L.has_res) f += __bfloat162float(rr[j]);
                if (L.relu) f = fmaxf(f, 0.f);
                o[j] = __float2bfloat16(f);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float2bfloat16(0.f);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(myrow + k * 16) = reinterpret_cast<const uint4*>(o)[k];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ls = l0 + i * 8 + crow;
          const long long mr = tile_row0 + ls;
          if (ls < L.TH && mr < M)
            *reinterpret_cast<uint4*>(out + ((size_t)L.guard + (size_t)mr) * L.cout + cc + cchunk * 8) =
                *reinterpret_cast<const uint4*>(stage + (size_t)(i * 8 + crow) * srow + cchunk * 16);
        }
        __syncwarp();
      }
      if (!arrived) {  // split path: every pass ends with its own tcgen05.wait::ld
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && crank != 0) mbar_arrive_remote(&tempty_bar[acc], 0);
          else mbar_arrive(&tempty_bar[acc]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

static int make_map(CUtensorMap* m, void* base, uint64_t inner, uint64_t rows, uint32_t box_rows, std::string& err) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { err = "cuTensorMapEncodeTiled not available"; return AZ_ERR_CUDA; }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r); return AZ_ERR_CUDA; }
  return 0;
}

static int make_map_3d(CUtensorMap* m, void* base, uint64_t inner, uint64_t width, uint64_t board_rows, uint32_t box_rows, std::string& err) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { err = "cuTensorMapEncodeTiled not available"; return AZ_ERR_CUDA; }
  cuuint64_t dims[3] = {inner, width, board_rows};
  cuuint64_t strides[2] = {inner * 2, inner * 2 * width};
  cuuint32_t box[3] = {TC_BK, (cuuint32_t)width, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled (3-D) failed with code " + std::to_string((int)r); return AZ_ERR_CUDA; }
  return 0;
}

int aznet_tc_create(AzNet* n, AzRt& rt, std::string& err) {
  AzNetTc* tc = new AzNetTc();
  n->tc = tc;
  // the bf16 tower reads 64-channel input rows (one swizzle atom); re-allocate the input buffer accordingly
  n->g.cin_pad = 64;
  rt_free(n->act_in);
  n->act_in = rt_alloc(n->rows_total * 64 * 2);
  if (!n->act_in) { err = "out of device memory"; return AZ_ERR_CUDA; }
  if (n->C != 64 && n->C != 128 && n->C != 256) { err = "tensor-core towers support num_filters <= 128 (padded to 64 / 128) or 193..256 (padded to 256)"; return AZ_ERR_BAD_ARG; }
  if ((unsigned long long)n->max_leaves * (unsigned long long)n->g.RP + (unsigned long long)2 * n->g.guard >= (1ull << 31)) {
    err = "tensor-core tower: leaf batch too large (row index must fit in 31 bits)";
    return AZ_ERR_BAD_ARG;
  }
  tc->split = n->precision == AZ_NET_BF16X3;
  const uint64_t CW = (uint64_t)(tc->split ? 3 * n->C : n->C);  // channels of a stored activation row
  int rc = make_map(&tc->map_in, n->act_in, 64, n->rows_total, TC_BM, err);
  if (!rc) rc = make_map(&tc->map_x, n->act_x, CW, n->rows_total, TC_BM, err);
  if (!rc) rc = make_map(&tc->map_mid, n->act_mid, CW, n->rows_total, TC_BM, err);
  if (rc) return rc;
  tc->smem_bytes = (size_t)TC_STAGES * (TC_BM * TC_BK * 2 + (size_t)n->C * TC_BK * 2) + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc->smem_bytes);
  if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return AZ_ERR_CUDA; }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&tc->num_sms, cudaDevAttrMultiProcessorCount, dev);
  const char* rp = getenv("AZ_TC_RESPF");
  tc->res_l2 = rp ? atoi(rp) : 0;
  const char* pd = getenv("AZ_PDL");
  tc->pdl = pd ? atoi(pd) : 1;
  const char* md = getenv("AZ_TC_MODE");
  // 5 = dense-x layout + 2-CTA pairs (default), 6 = dense-x single CTA, 4 = halo tile + 2-CTA pairs, 2 = halo tile single CTA,
  // 0 = one TMA box per tap
  tc->mode = md ? atoi(md) : AZ_TC_MODE_DEFAULT;
  if (n->C > 128) tc->mode = 0;  // the halo tile of a 256-channel layer does not fit next to the weight ring
  if (tc->split && tc->mode != 5 && tc->mode != 6) { err = "bf16x3 tower runs on the dense-x kernel only (num_filters <= 128, AZ_TC_MODE 5 or 6)"; return AZ_ERR_BAD_ARG; }
  if (tc->mode == 5 || tc->mode == 6) {
    // dense-x layout: rows y*Hc + x, one zero board row per leaf, no separator column.  The input and head kernels follow
    // NetGeom, so only the geometry changes for them.  6 = the single-CTA build of the same kernel.  A canvas too wide for the
    // shared-memory ring (Hc > 19 with 128 filters) keeps the halo kernel.
    const int Hc = n->g.Hc;
    const int nbr = (256 + 2 * Hc + Hc - 1) / Hc;
    const int sub_bytes = nbr * Hc * 128, sub_stride = (sub_bytes + 1023) / 1024 * 1024;
    // weight ring: X_BSTAGES half chunks per CTA of a pair == X_BSTAGES / 2 whole chunks for a single CTA
    const size_t x_smem = (size_t)X_ASLOTS * sub_stride + (size_t)X_BSTAGES * (n->C / 2) * TC_BK * 2 + (size_t)H_EPI_WARPS * 32 * 80 + (size_t)n->C * 4 +
                          (size_t)(2 * X_ASLOTS + 2 * X_BSTAGES + 4) * 8 + 16 + 1024;
    if (x_smem > 227 * 1024) {
      if (tc->split) { err = "bf16x3 tower: the dense-x ring does not fit this canvas"; return AZ_ERR_BAD_ARG; }
      tc->mode = 4;
    } else {
      NetGeom& g = n->g;
      g.Wr = Hc;
      g.RP = Hc * (Hc + 1);
      g.guard = Hc * 16;  // a whole number of board rows, so that board rows of the 3-D view start at matrix row 0
      tc->x_TH = 256 / Hc * Hc;
      tc->x_sub_bytes = sub_bytes;
      tc->x_sub_stride = sub_stride;
      tc->x_smem = x_smem;
      const uint64_t board_rows = n->rows_total / (uint64_t)Hc;
      rc = make_map_3d(&tc->xmap_in, n->act_in, 64, Hc, board_rows, nbr, err);
      if (!rc) rc = make_map_3d(&tc->xmap_x, n->act_x, CW, Hc, board_rows, nbr, err);
      if (!rc) rc = make_map_3d(&tc->xmap_mid, n->act_mid, CW, Hc, board_rows, nbr, err);
      if (rc) return rc;
      g.in_dup = tc->split;  // layer 0 of the split tower reads the planes twice: [x | x] against [W_hi | W_lo]
      e = cudaFuncSetAttribute(k_conv_tc_x<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc->x_smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_tc_x<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc->x_smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_tc_x<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc->x_smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_tc_x<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc->x_smem);
      if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(dense-x): ") + cudaGetErrorString(e); return AZ_ERR_CUDA; }
      return 0;
    }
  }
  if (tc->mode) {
    tc->halo = n->g.Wr + 1;
    tc->AR = (256 + 2 * tc->halo + 7) / 8 * 8;
    const uint32_t tail = (uint32_t)(tc->AR - 256);
    rc = make_map(&tc->hmap_in[0], n->act_in, 64, n->rows_total, 256, err);
    if (!rc) rc = make_map(&tc->hmap_in[1], n->act_in, 64, n->rows_total, tail, err);
    if (!rc) rc = make_map(&tc->hmap_x[0], n->act_x, n->C, n->rows_total, 256, err);
    if (!rc) rc = make_map(&tc->hmap_x[1], n->act_x, n->C, n->rows_total, tail, err);
    if (!rc) rc = make_map(&tc->hmap_mid[0], n->act_mid, n->C, n->rows_total, 256, err);
    if (!rc) rc = make_map(&tc->hmap_mid[1], n->act_mid, n->C, n->rows_total, tail, err);
    if (rc) return rc;
    tc->halo_smem = (size_t)2 * 2 * tc->AR * 128 + (size_t)H_BSTAGES * n->C * TC_BK * 2 + (size_t)H_EPI_WARPS * 32 * 80 + (size_t)n->C * 4 + 1024 + 256;
    if (tc->halo_smem > 227 * 1024) { tc->mode = 0; return 0; }
    e = cudaFuncSetAttribute(k_conv_tc_halo<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc->halo_smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_tc_halo<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc->halo_smem);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(halo): ") + cudaGetErrorString(e); return AZ_ERR_CUDA; }
  }
  return 0;
}

void aznet_tc_destroy(AzNet* n) {
  if (!n || !n->tc) return;
  for (auto* p : n->tc->w_dev) rt_free(p);
  delete n->tc;
  n->tc = nullptr;
}

// K extent of a layer's packed weights / of the activation rows it reads
static inline int tc_layer_cin(const AzNetTc* tc, int C, int li) { return li == 0 ? 64 : (tc->split ? 3 * C : C); }

// Fold BatchNorm and pack straight into bf16 [9][cout][cin] (K-major B operand), one host thread per group of layers;
// device buffers and TMA maps are created on the first call and refreshed in place afterwards.
int aznet_tc_set_weights(AzNet* n, AzRt& rt, std::string& err) {
  AzNetTc* tc = n->tc;
  const int C = n->C;
  const int n_conv = (int)n->layer_src.size();
  std::vector<std::vector<__nv_bfloat16>> pk(n_conv);
  auto work = [&](int t0, int step) {
    for (int li = t0; li < n_conv; li += step) {
      const float* const* L = n->layer_src[li];
      const int cin_src = n->layer_cin[li], cin = tc_layer_cin(tc, C, li);
      pk[li].assign((size_t)9 * C * cin, __float2bfloat16(0.f));
      for (int co = 0; co < C; ++co) {
        const float sc = L[1][co] / sqrtf(L[4][co] + 1e-5f);
        for (int ci = 0; ci < cin_src; ++ci) {
          const float* w9 = L[0] + ((size_t)co * cin_src + ci) * 9;
          for (int t = 0; t < 9; ++t) {
            const float wf = w9[t] * sc;
            const __nv_bfloat16 hi = __float2bfloat16(wf);
            __nv_bfloat16* row = &pk[li][((size_t)t * C + co) * cin];
            row[ci] = hi;
            if (tc->split) {
              const __nv_bfloat16 lo = __float2bfloat16(wf - __bfloat162float(hi));
              if (li == 0) row[32 + ci] = lo;                    // [W_hi(32) | W_lo(32)] against the duplicated planes
              else { row[C + ci] = hi; row[2 * C + ci] = lo; }  // [W_hi | W_hi | W_lo] against [a_hi | a_lo | a_hi]
            }
          }
        }
      }
    }
  };
  {
    const int nthreads = std::min(8, n_conv);
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t, nthreads);
    work(0, nthreads);
    for (auto& th : pool) th.join();
  }
  const bool first = tc->w_dev.empty();
  for (int li = 0; li < n_conv; ++li) {
    const int cin = tc_layer_cin(tc, C, li);
    if (first) {
      __nv_bfloat16* d = (__nv_bfloat16*)rt_alloc(pk[li].size() * 2);
      if (!d) { err = "out of device memory"; return AZ_ERR_CUDA; }
      tc->w_dev.push_back(d);
      CUtensorMap m, mh;
      int rc = make_map(&m, d, (uint64_t)cin, (uint64_t)9 * C, (uint32_t)C, err);
      if (!rc) rc = make_map(&mh, d, (uint64_t)cin, (uint64_t)9 * C, (uint32_t)(C / 2), err);
      if (rc) return rc;
      tc->map_w.push_back(m);
      tc->map_w_half.push_back(mh);
    }
    rt_h2d(rt, tc->w_dev[li], pk[li].data(), pk[li].size() * 2);
  }
  return 0;
}

// One launch of the tower: conv layer `li` (0 = input layer act_in -> act_x; odd = first conv of a block act_x -> act_mid; even >= 2 =
// second conv of a block act_mid -> act_x, with the block input act_x as residual when `with_res`).  The kernel follows tc->mode.
int aznet_tc_layer(AzNet* n, AzRt& rt, int li, bool with_res, const int32_t* n_rows_dev, int max_rows) {
  AzNetTc* tc = n->tc;
  const NetGeom& g = n->g;
  const long long Mmax = (long long)max_rows * g.RP;
  __nv_bfloat16* X = (__nv_bfloat16*)n->act_x;
  __nv_bfloat16* MID = (__nv_bfloat16*)n->act_mid;
  const int src = li == 0 ? 0 : ((li & 1) ? 1 : 2);  // which buffer the layer reads: act_in, act_x, act_mid
  __nv_bfloat16* outp = (li & 1) ? MID : X;
  const __nv_bfloat16* resp = (with_res && li >= 2 && !(li & 1)) ? X : nullptr;
  const float* bias = n->conv_b[li];
  const int cin = tc_layer_cin(tc, n->C, li);
  if (tc->mode == 5 || tc->mode == 6) {
    const bool pair = tc->mode == 5;
    const long long tiles = (Mmax + tc->x_TH - 1) / tc->x_TH;
    const int xgrid = pair ? (int)std::max<long long>(2, std::min<long long>((tiles + 1) / 2 * 2, tc->num_sms & ~1))
                           : (int)std::max<long long>(1, std::min<long long>(tiles, tc->num_sms));
    XLayer X5;
    X5.Hc = g.Hc; X5.RP = g.RP; X5.guard = g.guard; X5.TH = tc->x_TH; X5.sub_bytes = tc->x_sub_bytes; X5.sub_stride = tc->x_sub_stride;
    X5.bstages = pair ? X_BSTAGES : X_BSTAGES / 2;
    X5.cin = cin; X5.cout = n->C; X5.relu = 1; X5.has_res = resp ? 1 : 0; X5.ostride = tc->split ? 3 * n->C : n->C;
    X5.ksteps = (li == 0 && !tc->split && g.planes <= 32) ? 2 : 4;
    const CUtensorMap& ma = src == 0 ? tc->xmap_in : (src == 1 ? tc->xmap_x : tc->xmap_mid);
    if (!pair) {
      if (tc->split) k_conv_tc_x<false, true><<<xgrid, H_THREADS, tc->x_smem, rt.stream>>>(ma, tc->map_w[li], bias, resp, outp, n_rows_dev, X5);
      else k_conv_tc_x<false, false><<<xgrid, H_THREADS, tc->x_smem, rt.stream>>>(ma, tc->map_w[li], bias, resp, outp, n_rows_dev, X5);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)xgrid);
      cfg.blockDim = dim3(H_THREADS);
      cfg.dynamicSmemBytes = tc->x_smem;
      cfg.stream = rt.stream;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = (tc->pdl && li > 0) ? 2 : 1;  // layer 0 follows k_net_input, an ordinary launch
      if (tc->split) cudaLaunchKernelEx(&cfg, k_conv_tc_x<true, true>, ma, tc->map_w_half[li], bias, resp, outp, n_rows_dev, X5);
      else cudaLaunchKernelEx(&cfg, k_conv_tc_x<true, false>, ma, tc->map_w_half[li], bias, resp, outp, n_rows_dev, X5);
    }
  } else if (tc->mode) {
    const int hgrid = (int)std::max<long long>(2, std::min<long long>((Mmax + 255) / 256, tc->num_sms));
    HaloLayer H;
    H.Wr = g.Wr; H.Hc = g.Hc; H.RP = g.RP; H.guard = g.guard; H.halo = tc->halo; H.AR = tc->AR; H.bo_mode = tc->mode == 1 ? 1 : (tc->mode == 3 ? 3 : 0);
    H.cin = cin; H.cout = n->C; H.relu = 1; H.has_res = resp ? 1 : 0; H.res_l2 = resp ? tc->res_l2 : 0;  // the residual is always act_x (hmap_x)
    const CUtensorMap* ma = src == 0 ? tc->hmap_in : (src == 1 ? tc->hmap_x : tc->hmap_mid);
    if (tc->mode != 4) {
      k_conv_tc_halo<false><<<hgrid, H_THREADS, tc->halo_smem, rt.stream>>>(ma[0], ma[1], tc->map_w[li], tc->hmap_x[0], bias, resp, outp, n_rows_dev, H);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(hgrid & ~1));
      cfg.blockDim = dim3(H_THREADS);
      cfg.dynamicSmemBytes = tc->halo_smem;
      cfg.stream = rt.stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k_conv_tc_halo<true>, ma[0], ma[1], tc->map_w_half[li], tc->hmap_x[0], bias, resp, outp, n_rows_dev, H);
    }
  } else {
    const int grid = (int)std::max<long long>(1, std::min<long long>((Mmax + TC_BM - 1) / TC_BM, tc->num_sms));
    TcLayer L;
    L.Wr = g.Wr; L.Hc = g.Hc; L.RP = g.RP; L.guard = g.guard;
    L.cin = cin; L.cout = n->C; L.relu = 1; L.has_res = resp ? 1 : 0;
    const CUtensorMap& ma = src == 0 ? tc->map_in : (src == 1 ? tc->map_x : tc->map_mid);
    k_conv_tc<<<grid, TC_THREADS, tc->smem_bytes, rt.stream>>>(ma, tc->map_w[li], bias, resp, outp, n_rows_dev, L);
  }
  rt.launches++;
  return AZ_OK;
}

int aznet_tc_forward(AzNet* n, AzRt& rt, const int8_t* obs_base, const int32_t* row_list, const int32_t* n_rows_dev, int max_rows,
                     float* priors_base, float* values_base, int pri_stride) {
  AzNetTc* tc = n->tc;
  const NetGeom& g = n->g;
  {
    long long work = (long long)max_rows * g.nc;
    int blocks = (int)std::min<long long>((work + 255) / 256, 148 * 8);
    k_net_input<__nv_bfloat16><<<blocks, 256, 0, rt.stream>>>(obs_base, row_list, n_rows_dev, (__nv_bfloat16*)n->act_in, g);
    rt.launches++;
  }
  const int n_conv = 1 + 2 * n->blocks;
  cudaEvent_t* evp = n->ev_tower[n->n_forwards % AzNet::kTowerRing];
  cudaEventRecord(evp[0], rt.stream);
  for (int li = 0; li < n_conv; ++li) aznet_tc_layer(n, rt, li, true, n_rows_dev, max_rows);
  cudaEventRecord(evp[1], rt.stream);
  n->n_forwards++;
  launch_heads<__nv_bfloat16>(rt.stream, (const __nv_bfloat16*)n->act_x, row_list, n_rows_dev, n->hp, g, n->C, tc->split ? 3 * n->C : n->C, tc->split, n->A, n->fc,
                              priors_base, values_base, pri_stride, max_rows);
  rt.launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_az_error = std::string("tensor-core tower launch (AZ_TC_MODE ") + std::to_string(tc->mode) + "): " + cudaGetErrorString(e); return AZ_ERR_CUDA; }
  return AZ_OK;
}

int aznet_tc_mode(const AzNet* n) { return n && n->tc ? n->tc->mode : -1; }

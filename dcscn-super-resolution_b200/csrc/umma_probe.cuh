// Measurement kernel: kind::f16 tcgen05.mma throughput in isolation (no TMA, no epilogue, operands resident in shared
// memory).  It answers two questions the roofline of the conv kernels rests on:
//   * what the tensor pipe delivers for kind::f16 on THIS box (the denominator bench.py reports next to the driver's
//     bf16 matmul figure) - N = 256, M = 256 per CTA pair, back-to-back UMMAs on every SM;
//   * what one K = 16 slice of the three-product scheme (a_lo*w_hi, a_hi*w_lo, a_hi*w_hi) costs as a function of the
//     layer width N - the thin layers of DCSCN (N = 48..112) are far from the N = 256 rate - and whether the
//     "stacked" form (one UMMA a_hi x [w_hi ; w_lo] of width 2N plus one a_lo x w_hi of width N) would be cheaper.
// Not on any product path; exported as dcscn_umma_probe for bench.py and scripts/umma_probe.py.
#pragma once
#include "conv_tc_pair.cuh"

namespace dcscn {

struct UmmaProbeParams {
  int n;             // accumulator width of one product (multiple of 16, 16..256)
  int mode;          // 0: three UMMAs of width n per slice; 1: stacked (2n then n; cta_group::1 only); 2: one UMMA of width n
  int iters;         // outer iterations; each issues 4 slices (one 64-channel stage)
  unsigned long long* cycles;   // [clusters] clock64 span of the issuing thread
};

// smem: A hi/lo planes (128 rows x 128 B each) then B planes (rows x 128 B); all SWIZZLE_128B K-major tiles.
template <int GROUP>
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const UmmaProbeParams p) {
  extern __shared__ __align__(1024) uint8_t probe_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(probe_smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_PLANE = 128 * 128;
  const int b_rows = (GROUP == 2) ? p.n / 2 : p.n;          // rows of one weight plane held by this CTA
  const int B_PLANE = b_rows * 128;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + 2 * A_PLANE;
  // a second stage so that consecutive iterations touch different shared-memory lines, as the real ring does
  const int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  // plausible operand values: fp16 numbers in (-2, 2) from a hash (zeros would lower the power draw and flatter the clock)
  {
    uint16_t* w = reinterpret_cast<uint16_t*>(smem);
    const int total = STAGE;   // halves: 2 * STAGE bytes / 2
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      const uint16_t sign = (h & 1u) ? 0x8000u : 0u;
      const uint16_t expo = (uint16_t)(12 + ((h >> 1) % 4)) << 10;      // 2^-3 .. 2^0
      w[i] = sign | expo | (uint16_t)((h >> 8) & 0x3FFu);
    }
  }
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = (GROUP == 2) ? ptx::cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar[0], 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) {
    if (GROUP == 2) { ptx::tmem_alloc_2sm(tmem_slot, 512); ptx::tmem_relinquish_2sm(); }
    else { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (GROUP == 2) ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && rank == 0) {
    const uint32_t m = (GROUP == 2) ? 256u : 128u;
    const uint32_t idesc_n = (1u << 4) | ((uint32_t)(p.n >> 3) << 17) | ((m >> 4) << 24);
    const uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * p.n) >> 3) << 17) | ((m >> 4) << 24);
    constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);      // SBO 1024, SW128
    const uint32_t tmem_d = tmem_base, tmem_c = tmem_base + (uint32_t)((p.mode == 1) ? p.n : 256);
    const uint32_t a_u32 = ptx::smem_u32(smem_a), b_u32 = ptx::smem_u32(smem_b);
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      const uint32_t so = (it & 1) ? (uint32_t)STAGE : 0u;
      const uint32_t ah0 = desc_lo_t<64>(a_u32 + so), al0 = desc_lo_t<64>(a_u32 + so + A_PLANE);
      const uint32_t bh0 = desc_lo_t<64>(b_u32 + so), bl0 = desc_lo_t<64>(b_u32 + so + B_PLANE);
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t da_hi = ((uint64_t)kDescHi << 32) | (ah0 + 2u * ks), da_lo = ((uint64_t)kDescHi << 32) | (al0 + 2u * ks);
          const uint64_t db_hi = ((uint64_t)kDescHi << 32) | (bh0 + 2u * ks), db_lo = ((uint64_t)kDescHi << 32) | (bl0 + 2u * ks);
          if (GROUP == 2) {
            if (p.mode == 0) {
              ptx::mma_f16_ss_2sm_acc(tmem_c, da_lo, db_hi, idesc_n);
              ptx::mma_f16_ss_2sm_acc(tmem_c, da_hi, db_lo, idesc_n);
            }
            ptx::mma_f16_ss_2sm_acc(tmem_d, da_hi, db_hi, idesc_n);
          } else {
            if (p.mode == 0) {
              ptx::mma_f16_ss(tmem_c, da_lo, db_hi, idesc_n, 1);
              ptx::mma_f16_ss(tmem_c, da_hi, db_lo, idesc_n, 1);
              ptx::mma_f16_ss(tmem_d, da_hi, db_hi, idesc_n, 1);
            } else if (p.mode == 1) {
              ptx::mma_f16_ss(tmem_d, da_hi, db_hi, idesc_2n, 1);     // [w_hi ; w_lo] are contiguous rows: D = [dominant | correction]
              ptx::mma_f16_ss(tmem_c, da_lo, db_hi, idesc_n, 1);
            } else {
              ptx::mma_f16_ss(tmem_d, da_hi, db_hi, idesc_n, 1);
            }
          }
        }
      }
      __syncwarp();
    }
    if (ptx::elect_one()) {
      if (GROUP == 2) ptx::mma_commit_2sm(&bar[0], 1);
      else ptx::mma_commit(&bar[0]);
    }
    __syncwarp();
    ptx::mbar_wait(&bar[0], 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0 && p.cycles != nullptr) p.cycles[blockIdx.x / GROUP] = (unsigned long long)(t1 - t0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (GROUP == 2) ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (GROUP == 2) ptx::tmem_dealloc_2sm(tmem_base, 512);
    else ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dcscn

// Single-box halo variant (validated on B200: bit-level agreement with the three-box variant up to fp32 summation order): ONE {64, 10, 18, 1} TMA box per channel chunk serves all
// nine taps.  The smem rows are (hy * 10 + hx); tap (dy, dx) starts (dy * 10 + dx) rows into the box and the 16 eight-row
// groups of the UMMA A operand are 10 rows (1280 B) apart, so operand rows are NOT aligned to the 1024-byte swizzle
// atom.  This relies on the tensor core applying the 128B-swizzle XOR on absolute shared-memory address bits [7,10)
// (like TMA does when writing) - measured: correct with the descriptor base-offset field left 0, wrong with it set
// (`base_mode` = 1 keeps that experiment reproducible).
#pragma once
#include "conv_tc_halo.cuh"

namespace dcscn {

constexpr int kHalo1W = kHaloTW + 2;                              // 10 columns
constexpr int kHalo1Rows = (kHaloTH + 2) * kHalo1W;               // 180 rows
constexpr int kHalo1PlaneBytes = ((kHalo1Rows * 128 + 1023) / 1024) * 1024;  // 23552, keeps planes 1024-aligned

__host__ __device__ inline size_t tc_halo1_a_slot_bytes(int nplanes) { return (size_t)nplanes * kHalo1PlaneBytes; }

__device__ __forceinline__ uint64_t make_desc64_halo1(uint32_t saddr, int base_mode) {
  constexpr uint32_t hi = (uint32_t)((kHalo1W * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO = 1280, SW128
  uint32_t h = hi;
  if (base_mode == 1) h |= ((saddr >> 7) & 7u) << 17;             // base_offset, descriptor bits [49,52)
  return ((uint64_t)h << 32) | (uint64_t)(((saddr & 0x3FFFFu) >> 4) | (1u << 16));
}

template <int NPLANES>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_halo1_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                    const __grid_constant__ CUtensorMap tm_w, const ConvTCParams p, const int num_a, const int num_b, const int base_mode) {
  constexpr int KC = 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int AH_BYTES = kHalo1PlaneBytes;            // one plane of the 18 x 10 box
  constexpr int A_SLOT = NPLANES * AH_BYTES;
  const int half_rows = p.n_pad >> 1;                 // weight-tile rows staged by each CTA of the pair
  const int BH_BYTES = half_rows * KC * 2;            // one plane of this CTA's weight half
  const int B_STAGE = NPLANES * BH_BYTES;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)num_a * A_SLOT;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + (size_t)num_b * B_STAGE);
  uint64_t* a_empty = a_full + kMaxStages;
  uint64_t* b_full = a_empty + kMaxStages;
  uint64_t* b_empty = b_full + kMaxStages;
  uint64_t* acc_full = b_empty + kMaxStages;
  uint64_t* acc_empty = acc_full + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);
  float* s_rdot = reinterpret_cast<float*>(tmem_slot + 4);   // 16-byte aligned (barriers start 1024-aligned)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_a; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < num_b; ++s) {
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&acc_full[s], 1);                     // leader's tcgen05.commit (multicast)
      ptx::mbar_init(&acc_empty[s], 2 * kEpiWarps);        // epilogue warps of both CTAs (used on the leader only)
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_slot, kAccStages * kAccStride);
    ptx::tmem_relinquish_2sm();
  }
  if (p.epi.mode == EPI_D2S_RDOT)
    for (int i = threadIdx.x; i < p.epi.rdot_taps * p.epi.d2s_cout; i += blockDim.x) s_rdot[i] = p.epi.rdot_w[i];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const ConvGeom& g = p.g;
  const int tiles_per_img = g.tiles_x * g.tiles_y;
  const int num_tiles = g.n_img * tiles_per_img;
  const int groups = (num_tiles + 1) >> 1;
  const int num_items = groups * p.n_tiles;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int total_a = p.chunks * 3;                   // segment units per tile: (channel chunk, dx); one A slot per chunk

  if (warp < kEpiWarp0) {
    ptx::setmaxnreg_dec<kRegsIssue>();
    if (warp == 0) {
      // ============================== TMA producer: A boxes (both CTAs) ==============================
      if (lane == 0) {
        ptx::prefetch_tensormap(&tm_hi);
        if (NPLANES == 2) ptx::prefetch_tensormap(&tm_lo);
        int a = 0;
        uint32_t pha = 0;
        for (int item = cluster_id; item < num_items; item += num_clusters) {
          int tile = (item / p.n_tiles) * 2 + (int)rank;
          if (tile >= num_tiles) tile = num_tiles - 1;       // lockstep filler (stores are masked)
          const int img = tile / tiles_per_img;
          const int t2 = tile - img * tiles_per_img;
          const int ty = t2 / g.tiles_x, tx = t2 - ty * g.tiles_x;
          for (int ch = 0; ch < p.chunks; ++ch) {
            ptx::mbar_wait(&a_empty[a], pha ^ 1);
            uint8_t* slot = smem_a + (size_t)a * A_SLOT;
            const uint32_t lead = ptx::mapa_shared(ptx::smem_u32(&a_full[a]), 0);
            if (leader) ptx::mbar_arrive_expect_tx(&a_full[a], (uint32_t)(2 * NPLANES * kHalo1Rows * 128));
            ptx::tma_load_4d_2sm(slot, &tm_hi, lead, ch * KC, tx * kHaloTW - 1, ty * kHaloTH - 1, img);
            if (NPLANES == 2)
              ptx::tma_load_4d_2sm(slot + AH_BYTES, &tm_lo, lead, ch * KC, tx * kHaloTW - 1, ty * kHaloTH - 1, img);
            if (++a == num_a) { a = 0; pha ^= 1; }
          }
        }
      }
    } else if (warp == 2) {
      // ============================== TMA producer: weight halves (both CTAs) ==============================
      if (lane == 0) {
        ptx::prefetch_tensormap(&tm_w);
        const int wrows = NPLANES * half_rows;               // rows of one (tile, rank) block in the packed weights
        int b = 0;
        uint32_t phb = 0;
        for (int item = cluster_id; item < num_items; item += num_clusters) {
          const int n_tile = item % p.n_tiles;
          for (int ai = 0; ai < total_a; ++ai) {
            const int ch = ai / 3, dx = ai - ch * 3;
            for (int dy = 0; dy < 3; ++dy) {
              ptx::mbar_wait(&b_empty[b], phb ^ 1);
              const uint32_t lead = ptx::mapa_shared(ptx::smem_u32(&b_full[b]), 0);
              if (leader) ptx::mbar_arrive_expect_tx(&b_full[b], (uint32_t)(2 * B_STAGE));
              const int wblock = ((n_tile * 9 + dy * 3 + dx) * p.chunks + ch) * 2 + (int)rank;
              ptx::tma_load_2d_2sm(smem_b + (size_t)b * B_STAGE, &tm_w, lead, 0, wblock * wrows);
              if (++b == num_b) { b = 0; phb ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1 && leader) {
      // ============================== MMA issuer (leader CTA only) ================================
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t sa_u32 = ptx::smem_u32(smem_a), sb_u32 = ptx::smem_u32(smem_b);
      int a = 0, b = 0;
      uint32_t pha = 0, phb = 0;
      uint32_t seg_count = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        for (int a0 = 0; a0 < total_a; a0 += p.seg_chunks) {
          const int acc = seg_count & 1;
          ptx::mbar_wait(&acc_empty[acc], ((seg_count >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kAccStride);
          uint32_t accumulate = 0;
          const int a1 = (a0 + p.seg_chunks < total_a) ? a0 + p.seg_chunks : total_a;
          // Pass A: correction products as the slots / stages of this segment land.
          int sa = a, sb = b;
          uint32_t spa = pha, spb = phb;
          for (int ai = a0; ai < a1; ++ai) {
            const int ch = ai / 3, dx = ai - ch * 3;
            int ksteps = (p.cin_pad - ch * KC);
            ksteps = (ksteps > KC ? KC : ksteps) >> 4;
            if (dx == 0) ptx::mbar_wait(&a_full[sa], spa);
            for (int dy = 0; dy < 3; ++dy) {
              ptx::mbar_wait(&b_full[sb], spb);
              ptx::tc_fence_after();
              if (NPLANES == 2) {
                const uint32_t a_addr = sa_u32 + (uint32_t)sa * (uint32_t)A_SLOT + (uint32_t)(dy * kHalo1W + dx) * 128u;
                const uint32_t b_addr = sb_u32 + (uint32_t)sb * (uint32_t)B_STAGE;
                const uint32_t lb_hi = desc_lo_t<KC>(b_addr), lb_lo = desc_lo_t<KC>(b_addr + BH_BYTES);
                if (ptx::elect_one()) {
#pragma unroll 1
                  for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t kadd = (uint32_t)ks * 2u;
                    ptx::mma_f16_ss_2sm(tmem_d, make_desc64_halo1(a_addr + AH_BYTES + ks * 32, base_mode), make_desc64_t<KC>(lb_hi + kadd), idesc, accumulate);
                    ptx::mma_f16_ss_2sm(tmem_d, make_desc64_halo1(a_addr + ks * 32, base_mode), make_desc64_t<KC>(lb_lo + kadd), idesc, 1);
                    accumulate = 1;
                  }
                }
                accumulate = 1;
                __syncwarp();
              }
              if (++sb == num_b) { sb = 0; spb ^= 1; }
            }
            if (dx == 2 && ++sa == num_a) { sa = 0; spa ^= 1; }
          }
          // Pass B: dominant a_hi*w_hi products; release weight stages and A slots.
          sa = a;
          sb = b;
          for (int ai = a0; ai < a1; ++ai) {
            const int ch = ai / 3, dx = ai - ch * 3;
            int ksteps = (p.cin_pad - ch * KC);
            ksteps = (ksteps > KC ? KC : ksteps) >> 4;
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t a_addr = sa_u32 + (uint32_t)sa * (uint32_t)A_SLOT + (uint32_t)(dy * kHalo1W + dx) * 128u;
              const uint32_t lb_hi = desc_lo_t<KC>(sb_u32 + (uint32_t)sb * (uint32_t)B_STAGE);
              if (ptx::elect_one()) {
#pragma unroll 1
                for (int ks = 0; ks < ksteps; ++ks) {
                  const uint32_t kadd = (uint32_t)ks * 2u;
                  ptx::mma_f16_ss_2sm(tmem_d, make_desc64_halo1(a_addr + ks * 32, base_mode), make_desc64_t<KC>(lb_hi + kadd), idesc, accumulate);
                  accumulate = 1;
                }
                ptx::mma_commit_2sm(&b_empty[sb], 3);
                if (dy == 2 && dx == 2) ptx::mma_commit_2sm(&a_empty[sa], 3);
              }
              accumulate = 1;
              __syncwarp();
              if (++sb == num_b) sb = 0;
            }
            if (dx == 2 && ++sa == num_a) sa = 0;
          }
          a = sa; pha = spa; b = sb; phb = spb;
          if (ptx::elect_one()) ptx::mma_commit_2sm(&acc_full[acc], 3);
          __syncwarp();
          ++seg_count;
        }
      }
    }
  } else {
    ptx::setmaxnreg_inc<kRegsEpilogue>();
    // ============================== epilogue (both CTAs, own 128 TMEM lanes) ==================================
    const int ew = warp - kEpiWarp0;
    const int quad = warp & 3;
    const int grp = ew >> 2;
    const int row = quad * 32 + lane;
    const int py = row / kHaloTW, px = row - py * kHaloTW;
    const int n_total = p.n_tiles * p.n_pad;
    const int nch = p.n_pad >> 4;
    const int per = (nch + kColSplit - 1) / kColSplit;
    const int first_chunk = grp * per;
    const int my_chunks = (nch - first_chunk) < per ? ((nch - first_chunk) > 0 ? nch - first_chunk : 0) : per;
    const int col_base = first_chunk * 16;
    const int nseg = (total_a + p.seg_chunks - 1) / p.seg_chunks;
    const uint32_t lead_acc_empty0 = ptx::mapa_shared(ptx::smem_u32(&acc_empty[0]), 0);
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col_base;
    uint32_t seg_count = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      const int n_tile = item % p.n_tiles;
      const int tile = (item / p.n_tiles) * 2 + (int)rank;
      const bool real = tile < num_tiles;
      const int img = tile / tiles_per_img;
      const int t2 = tile - img * tiles_per_img;
      const int ty = t2 / g.tiles_x, tx = t2 - ty * g.tiles_x;
      const int y = ty * kHaloTH + py, x = tx * kHaloTW + px;
      const bool valid = real && (y < g.H) && (x < g.W);

      float sum[kMaxColChunks][16];
#pragma unroll
      for (int j = 0; j < kMaxColChunks; ++j)
#pragma unroll
        for (int i = 0; i < 16; ++i) sum[j][i] = 0.f;
      for (int s = 0; s < nseg; ++s) {
        const int acc = seg_count & 1;
        ptx::mbar_wait(&acc_full[acc], (seg_count >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr = taddr0 + (uint32_t)(acc * kAccStride);
        // fp32 round-to-nearest promotion of the segment: wide TMEM loads (64 / 32 columns per instruction); columns
        // past this thread's share may be read (they stay inside the accumulator stage) but are never stored.
#pragma unroll
        for (int j = 0; j < kMaxColChunks; j += 4) {
          if (j < my_chunks) {
            if (my_chunks - j > 2) {
              float v[64];
              ptx::tmem_ld64(taddr + j * 16, v);
#pragma unroll
              for (int i = 0; i < 64; ++i) sum[j + (i >> 4)][i & 15] += v[i];
            } else {
              float v[32];
              ptx::tmem_ld32(taddr + j * 16, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[j + (i >> 4)][i & 15] += v[i];
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(lead_acc_empty0 + (uint32_t)(acc * sizeof(uint64_t)));
        ++seg_count;
      }
      if (p.epi.mode == EPI_D2S_RDOT) {
        float v[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) v[t] = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxColChunks; ++j) {
          if (j < my_chunks) {
            const int cg = n_tile * p.n_pad + col_base + j * 16;
            if (cg < p.epi.n_valid) {
              const int ij = cg / p.epi.d2s_cout, c = cg - ij * p.epi.d2s_cout;
              rdot_accumulate16(p.epi, s_rdot, cg, c, sum[j], v);
              if (c + 16 == p.epi.d2s_cout && valid) rdot_flush(p.epi, g, img, y, x, ij, v);
            }
          }
        }
      } else if (valid) {
#pragma unroll
        for (int j = 0; j < kMaxColChunks; ++j)
          if (j < my_chunks) epilogue_store16(p.epi, g, n_total, img, y, x, n_tile * p.n_pad + col_base + j * 16, sum[j]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, kAccStages * kAccStride);
  }
}

}  // namespace dcscn

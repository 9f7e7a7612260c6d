// tcgen05 single-head attention core for the NCSN++ AttnBlockpp (reference layerspp.py:82-86):
//   w = softmax_keys( einsum(q, k) * C^-0.5 ) ;  h = einsum(w, v)
// over HW = 64/128/256 tokens of C = 64..256 channels per sample (16x16 and 8x8 maps).
//
// One CTA = one sample x 128 queries.  All HW keys fit in one accumulator, so there is no
// online-softmax rescaling:
//   S[128 x HW]  = Q K^T      tcgen05.mma, A = Q chunk, B = K chunk (both K-major, 64-channel
//                             chunks streamed by TMA through a 3-slot ring), accumulator in TMEM
//   P            = exp2(S*c - max*c) -> bf16, written by the 128 softmax threads (one row each)
//                  straight into shared memory in the 128B-swizzled K-major UMMA layout
//   O[128 x C]   = P V        per 64-channel group g: A = P, B = V_g as an MN-major operand
//                             (V is [keys, channels] in memory = contiguous along N)
//   out          = O / rowsum  -> bf16 NHWC
// TMEM: HW columns for S + C columns for O (<= 512).  q|k|v arrive packed along the channel axis
// of one [N, HW, 3C] tensor (the fused NIN_0/1/2 GEMM), addressed with ONE 3-D tensor map.

#include <cuda.h>

#include <new>

#include "common.cuh"
#include "tc_common.cuh"
#include "conv_tc_common.cuh"

namespace psld {

constexpr int AT_P_TILE = 128 * 256 * 2;       // P (or normalised O) as 4 K-major tiles of [128 x 64] bf16
constexpr int AT_THREADS = 192;

// kX3 = split-bf16 operands (the fp32-tolerance tier): q|k|v rows are [3C hi | 3C lo], every GEMM
// (S = Q K^T, O = P V, Y = O W3^T) is three MMA groups hi*hi + hi*lo + lo*hi, P and the normalised
// O are written to shared memory as hi and lo tiles, the output is split bf16.  Operand tiles are
// twice as large, so the ring has ONE slot (96 KB) next to the 128 KB of P; the projection
// epilogue's staging tiles alias the slot (every MMA has retired by then).
template <bool kX3>
struct AtCfg {
  static constexpr int kQBytes = 16384 * (kX3 ? 2 : 1);      // Q chunk [128 x 64] (hi | lo)
  static constexpr int kKVBytes = 32768 * (kX3 ? 2 : 1);     // K chunk / V group / W3 chunk [<=256 x 64]
  static constexpr int kSlotBytes = kQBytes + kKVBytes;
  static constexpr int kSlots = kX3 ? 1 : 3;
  static constexpr int kPBytes = AT_P_TILE * (kX3 ? 2 : 1);
  static constexpr int kSmem = kSlots * kSlotBytes + kPBytes + 1024 + 256;
  // fused output projection: 4 KB staging and 256 B additive vector per softmax/epilogue warp
  static constexpr int kSmemProj = kSmem + (kX3 ? 0 : 4 * 4096 + 4 * 256);
};

struct AttnTcParams {
  __nv_bfloat16* out;
  int HW, C, N;
  int q_rows;        // rows of the Q box: min(128, HW)
  float scale_log2;  // C^-0.5 * log2(e)
  ConvTcParams ep;   // kProj: epilogue of the fused NIN_3 projection (bias, residual, scale, stats)
};

struct AttnTcState {
  CUtensorMap tq, tkv, tw;
  bool proj, x3;
  AttnTcParams p;
  dim3 grid;
};

// kProj: the block's output projection rides along (AttnBlockpp, layerspp.py:87-91):
//   h = (NIN_3(O / rowsum) + x) * scale.  The normalised O goes to shared memory as bf16 in the
//   P buffer (K-major over channels), W3 streams through the ring as 64-channel chunks, the
//   product accumulates in the TMEM columns S occupied, and the conv epilogue (bias, residual,
//   scale, bf16 store, GroupNorm statistics) finishes the tile: no O round trip through HBM and
//   no separate 1x1 convolution launch.
template <bool kProj, bool kX3>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
               const __grid_constant__ CUtensorMap tmW, const AttnTcParams p) {
  using Cfg = AtCfg<kX3>;
  constexpr int AT_SLOTS = Cfg::kSlots;
  constexpr int AT_SLOT_BYTES = Cfg::kSlotBytes;
  constexpr uint32_t QB = Cfg::kQBytes;              // offset of the K / V / W3 tile inside a slot
  constexpr uint32_t A_LO = 16384u;                  // lo half of the Q chunk (kX3)
  constexpr uint32_t B_LO = 32768u;                  // lo half of the K / V / W3 tile (kX3)
  constexpr uint32_t P_LO = AT_P_TILE;               // lo tiles of P / normalised O (kX3)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t p_base = base + AT_SLOTS * AT_SLOT_BYTES;
  const uint32_t bar_base = p_base + Cfg::kPBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (AT_SLOTS + s); };
  const uint32_t s_full = bar_base + 8u * (2 * AT_SLOTS);
  const uint32_t p_ready = s_full + 8u;
  const uint32_t o_full = s_full + 16u;
  const uint32_t o_ready = s_full + 24u;
  const uint32_t y_full = s_full + 32u;
  const uint32_t tmem_slot = s_full + 40u;
  // kProj only; kX3: aliases the (by then idle) operand slot
  const uint32_t stg_base = kX3 ? base : bar_base + 256u;
  const uint32_t addv_base = stg_base + 4u * 4096u;
  const int lo_c = 3 * p.C;                              // kX3: lo half of a q|k|v row
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y, q0 = blockIdx.x * 128;
  const int nck = p.C / 64;          // channel chunks (S GEMM K loop) = channel groups (PV GEMM)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
    for (int s = 0; s < AT_SLOTS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    mbar_init(o_ready, 128);
    mbar_init(y_full, 1);
    if (kProj) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                  // q|k|v come from the previous kernel (prologue above overlaps its tail)
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_S = tmem_base;                 // columns [0, HW)
  const uint32_t tmem_O = tmem_base + 256u;          // columns [256, 256 + C)

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      // Q/K channel chunks
      for (int c = 0; c < nck; ++c) {
        mbar_wait(empty_bar(slot), phase ^ 1);
        mbar_arrive_expect_tx(full_bar(slot), (uint32_t)(p.q_rows + p.HW) * 128u * (kX3 ? 2u : 1u));
        const uint32_t sq = base + slot * AT_SLOT_BYTES;
        tma_load_3d(sq, &tmQ, full_bar(slot), c * 64, q0, n);
        tma_load_3d(sq + QB, &tmKV, full_bar(slot), p.C + c * 64, 0, n);
        if (kX3) {
          tma_load_3d(sq + A_LO, &tmQ, full_bar(slot), lo_c + c * 64, q0, n);
          tma_load_3d(sq + QB + B_LO, &tmKV, full_bar(slot), lo_c + p.C + c * 64, 0, n);
        }
        if (++slot == AT_SLOTS) { slot = 0; phase ^= 1; }
      }
      // V channel groups
      for (int g = 0; g < nck; ++g) {
        mbar_wait(empty_bar(slot), phase ^ 1);
        mbar_arrive_expect_tx(full_bar(slot), (uint32_t)p.HW * 128u * (kX3 ? 2u : 1u));
        const uint32_t sv = base + slot * AT_SLOT_BYTES + QB;
        tma_load_3d(sv, &tmKV, full_bar(slot), 2 * p.C + g * 64, 0, n);
        if (kX3) tma_load_3d(sv + B_LO, &tmKV, full_bar(slot), lo_c + 2 * p.C + g * 64, 0, n);
        if (++slot == AT_SLOTS) { slot = 0; phase ^= 1; }
      }
      if (kProj) {       // W3 [C out rows x 64 input channels] per chunk
        for (int c = 0; c < nck; ++c) {
          mbar_wait(empty_bar(slot), phase ^ 1);
          mbar_arrive_expect_tx(full_bar(slot), (uint32_t)p.C * 128u * (kX3 ? 2u : 1u));
          tma_load_2d(base + slot * AT_SLOT_BYTES + QB, &tmW, full_bar(slot), c * 64, 0);
          if (kX3) tma_load_2d(base + slot * AT_SLOT_BYTES + QB + B_LO, &tmW, full_bar(slot), c * 64, p.C);
          if (++slot == AT_SLOTS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      // S = Q K^T : M = 128, N = HW, both operands K-major
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.HW >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
      for (int c = 0; c < nck; ++c) {
        mbar_wait(full_bar(slot), phase);
        tc_fence_after();
        const uint32_t sq = base + slot * AT_SLOT_BYTES;
        const uint64_t adesc = make_sw128_desc(sq);
        const uint64_t bdesc = make_sw128_desc(sq + QB);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(tmem_S, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_s,
                      (c > 0 || k > 0) ? 1u : 0u);
        if (kX3) {
          const uint64_t adesc_lo = make_sw128_desc(sq + A_LO);
          const uint64_t bdesc_lo = make_sw128_desc(sq + QB + B_LO);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_S, adesc + (uint64_t)(2 * k), bdesc_lo + (uint64_t)(2 * k), idesc_s, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_S, adesc_lo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_s, 1u);
        }
        tc_commit(empty_bar(slot));
        if (++slot == AT_SLOTS) { slot = 0; phase ^= 1; }
      }
      tc_commit(s_full);
      // O_g = P V_g : M = 128, N = 64, A = P (K-major over keys), B = V_g (MN-major)
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) |
                               ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(p_ready, 0);
      tc_fence_after();
      for (int g = 0; g < nck; ++g) {
        mbar_wait(full_bar(slot), phase);
        tc_fence_after();
        const uint32_t sv = base + slot * AT_SLOT_BYTES + QB;
        for (int ks = 0; ks < p.HW / 16; ++ks) {
          const uint64_t adesc = make_sw128_desc(p_base + (uint32_t)(ks >> 2) * 16384u) +
                                 (uint64_t)(2 * (ks & 3));
          const uint64_t bdesc = make_sw128_desc(sv + (uint32_t)ks * 2048u);   // 16 keys x 128 B
          tc_mma_bf16(tmem_O + (uint32_t)g * 64u, adesc, bdesc, idesc_o, ks > 0 ? 1u : 0u);
          if (kX3) {
            const uint64_t adesc_lo = make_sw128_desc(p_base + P_LO + (uint32_t)(ks >> 2) * 16384u) +
                                      (uint64_t)(2 * (ks & 3));
            const uint64_t bdesc_lo = make_sw128_desc(sv + B_LO + (uint32_t)ks * 2048u);
            tc_mma_bf16(tmem_O + (uint32_t)g * 64u, adesc, bdesc_lo, idesc_o, 1u);
            tc_mma_bf16(tmem_O + (uint32_t)g * 64u, adesc_lo, bdesc, idesc_o, 1u);
          }
        }
        tc_commit(empty_bar(slot));
        if (++slot == AT_SLOTS) { slot = 0; phase ^= 1; }
      }
      tc_commit(o_full);
      if (kProj) {
        // Y[128 x C] = Onorm W3^T : A = Onorm (bf16, K-major over channels, in the P buffer),
        // B = W3 chunk (K-major), accumulator in the columns S used
        const uint32_t idesc_y = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.C >> 3) << 17) |
                                 ((uint32_t)(128 >> 4) << 24);
        mbar_wait(o_ready, 0);
        tc_fence_after();
        for (int c = 0; c < nck; ++c) {
          mbar_wait(full_bar(slot), phase);
          tc_fence_after();
          const uint64_t adesc = make_sw128_desc(p_base + (uint32_t)c * 16384u);
          const uint64_t bdesc = make_sw128_desc(base + slot * AT_SLOT_BYTES + QB);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_S, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_y,
                        (c > 0 || k > 0) ? 1u : 0u);
          if (kX3) {
            const uint64_t adesc_lo = make_sw128_desc(p_base + P_LO + (uint32_t)c * 16384u);
            const uint64_t bdesc_lo = make_sw128_desc(base + slot * AT_SLOT_BYTES + QB + B_LO);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(tmem_S, adesc + (uint64_t)(2 * k), bdesc_lo + (uint64_t)(2 * k), idesc_y, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(tmem_S, adesc_lo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_y, 1u);
          }
          tc_commit(empty_bar(slot));
          if (++slot == AT_SLOTS) { slot = 0; phase ^= 1; }
        }
        tc_commit(y_full);
      }
    }
  } else {
    // ===== softmax + epilogue: 128 threads, one query row each =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    mbar_wait(s_full, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int ch = 0; ch < p.HW; ch += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_S + lane_addr + (uint32_t)ch, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
    const float mxs = mx * p.scale_log2;
    float sum = 0.f;
    for (int ch = 0; ch < p.HW; ch += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_S + lane_addr + (uint32_t)ch, r);
      tmem_ld_wait();
      // P tile (ch / 64), 16-byte chunks (ch % 64) / 8 .. +3, swizzled by (row % 8)
      const uint32_t tile = p_base + (uint32_t)(ch >> 6) * 16384u + (uint32_t)row * 128u;
      const int cbase = (ch & 63) >> 3;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t w[4], wl[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float e0 = exp2f(fmaf(__uint_as_float(r[q * 8 + 2 * t]), p.scale_log2, -mxs));
          const float e1 = exp2f(fmaf(__uint_as_float(r[q * 8 + 2 * t + 1]), p.scale_log2, -mxs));
          if (kX3) {
            split_bf2(e0, e1, w[t], wl[t]);
            sum += e0 + e1;
          } else {
            __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
            // accumulate the sum from the ROUNDED values so that P / sum is a true softmax of P
            const float2 f = __bfloat1622float2(h);
            sum += f.x + f.y;
            w[t] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        const uint32_t dst = tile + (uint32_t)(((cbase + q) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        if (kX3) sts128(dst + P_LO, wl[0], wl[1], wl[2], wl[3]);
      }
    }
    // make the generic-proxy smem writes visible to the tensor-core (async) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    mbar_arrive(p_ready);
    const float inv = 1.0f / sum;
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int q = q0 + row;
    const bool valid = q < p.HW;
    if (kProj) {
      // normalised O -> bf16 in the P buffer, same swizzled K-major tiles as P (tile = ch / 64)
      for (int ch = 0; ch < p.C; ch += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_O + lane_addr + (uint32_t)ch, r);
        tmem_ld_wait();
        const uint32_t tile = p_base + (uint32_t)(ch >> 6) * 16384u + (uint32_t)row * 128u;
        const int cbase = (ch & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w[4], wl[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float o0 = __uint_as_float(r[j * 8 + 2 * t]) * inv;
            const float o1 = __uint_as_float(r[j * 8 + 2 * t + 1]) * inv;
            if (kX3) split_bf2(o0, o1, w[t], wl[t]);
            else w[t] = f2_to_bf2(o0, o1);
          }
          sts128(tile + (uint32_t)(((cbase + j) ^ (row & 7)) << 4), w[0], w[1], w[2], w[3]);
          if (kX3)
            sts128(tile + P_LO + (uint32_t)(((cbase + j) ^ (row & 7)) << 4), wl[0], wl[1], wl[2], wl[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(o_ready);
      mbar_wait(y_full, 0);
      tc_fence_after();
      const int m_tile = n * (p.HW >> 7) + (int)blockIdx.x;
      const int ew = warp - 2;                     // softmax / epilogue warps are warps 2..5
#pragma unroll 1
      for (int half = 0; half < 2; ++half)
        tc_epilogue_tile<true, kX3 ? 32 : 64, !kX3, kX3>(p.ep, tmem_S, 0, m_tile, 0, quarter, half, lane,
                                                         stg_base + (uint32_t)ew * 4096u,
                                                         addv_base + (uint32_t)ew * 256u, []() {}, []() {});
    } else {
    __nv_bfloat16* orow = p.out + ((int64_t)n * p.HW + q) * (p.C * (kX3 ? 2 : 1));
    for (int ch = 0; ch < p.C; ch += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_O + lane_addr + (uint32_t)ch, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t w[4], wl[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float o0 = __uint_as_float(r[j + 2 * t]) * inv;
            const float o1 = __uint_as_float(r[j + 2 * t + 1]) * inv;
            if (kX3) split_bf2(o0, o1, w[t], wl[t]);
            else w[t] = f2_to_bf2(o0, o1);
          }
          *reinterpret_cast<uint4*>(orow + ch + j) = make_uint4(w[0], w[1], w[2], w[3]);
          if (kX3) *reinterpret_cast<uint4*>(orow + p.C + ch + j) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
      }
      __syncwarp();
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ---------------------------------------------------------------- host side
static int encode_qkv_map(CUtensorMap* tm, const void* ptr, int N, int HW, int C3, int rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return PSLD_ECUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)C3, (cuuint64_t)HW, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)C3 * 2, (cuuint64_t)HW * C3 * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(qkv) failed: %d", (int)r); return PSLD_ECUDA; }
  return PSLD_OK;
}

int prepare_attn_tc(psld_op& op) {
  const int N = op.i[PSLD_ATTN_N], HW = op.i[PSLD_ATTN_HW], C = op.i[PSLD_ATTN_C];
  auto unsupported = [&](const char* why) {
    set_error("attn_tc: not eligible (%s): N=%d HW=%d C=%d", why, N, HW, C);
    return PSLD_EUNSUPPORTED;
  };
  const int adt = op.i[PSLD_ATTN_DTYPE];
  if (adt != PSLD_BF16 && adt != PSLD_BF16S) return unsupported("dtype must be bf16 or split bf16");
  const bool x3 = adt == PSLD_BF16S;
  const int cm = x3 ? 2 : 1;
  if (C % 64 || C < 64 || C > 256) return unsupported("C must be 64..256, multiple of 64");
  if (HW != 64 && HW != 128 && HW != 256) return unsupported("HW must be 64, 128 or 256");
  if (!op.in[0] || !op.out[0]) { set_error("attn_tc: null pointer"); return PSLD_EINVAL; }
  AttnTcState* st = new (std::nothrow) AttnTcState();
  if (!st) { set_error("attn_tc: out of host memory"); return PSLD_ECUDA; }
  const int q_rows = HW < 128 ? HW : 128;
  st->x3 = x3;
  int rc = encode_qkv_map(&st->tq, op.in[0], N, HW, cm * 3 * C, q_rows);
  if (rc == PSLD_OK) rc = encode_qkv_map(&st->tkv, op.in[0], N, HW, cm * 3 * C, HW);
  if (rc != PSLD_OK) { delete st; return rc; }
  st->p.out = (__nv_bfloat16*)op.out[0];
  st->p.HW = HW; st->p.C = C; st->p.N = N; st->p.q_rows = q_rows;
  st->p.scale_log2 = op.f[0] * 1.4426950408889634f;
  st->grid = dim3((unsigned)((HW + 127) / 128), (unsigned)N);
  // optional fused output projection (in[1] = W3 bf16 [C out, C in], in[2] = bias, in[3] = residual)
  st->proj = op.i[PSLD_ATTN_PROJ] != 0;
  if (st->proj) {
    if (HW % 128 || !op.in[1] || !op.in[3]) { delete st; return unsupported("fused projection needs HW %% 128 == 0, weight and residual"); }
    EncodeTiledFn enc = get_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)(cm * C)};      // split: planes [2][C out, C in]
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)C};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc ? enc(&st->tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(op.in[1]), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                     : CUDA_ERROR_UNKNOWN;
    if (r != CUDA_SUCCESS) { delete st; set_error("cuTensorMapEncodeTiled(proj weight) failed: %d", (int)r); return PSLD_ECUDA; }
    ConvTcParams& e = st->p.ep;
    e = ConvTcParams{};
    e.bias = (const float*)op.in[2];
    e.temb = nullptr;
    e.res = (const __nv_bfloat16*)op.in[3];
    e.y = (__nv_bfloat16*)op.out[0];
    e.y_nchw = nullptr;
    e.mg_stats = (double*)op.out[1];
    e.cout_valid = C;
    e.scale = op.f[1];
    e.HW = HW; e.Cout = C; e.H = 0; e.W = 0;
    e.block_n = C; e.n_tiles_n = 1;
    e.M = (int64_t)N * HW;
  } else {
    st->tw = st->tq;       // unused
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<false, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg<false>::kSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               AtCfg<false>::kSmemProj);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               AtCfg<true>::kSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               AtCfg<true>::kSmemProj);
    if (e != cudaSuccess) {
      set_error("attn_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      delete st;
      return PSLD_ECUDA;
    }
    attr_set = true;
  }
  op.aux = st;
  return PSLD_OK;
}

int release_attn_tc(psld_op& op) {
  if (op.aux) {
    delete (AttnTcState*)op.aux;
    op.aux = nullptr;
  }
  return PSLD_OK;
}

int run_attn_tc(const psld_op& op, cudaStream_t s) {
  const AttnTcState* st = (const AttnTcState*)op.aux;
  PSLD_CHECK_ARG(st != nullptr, "attn_tc: op not prepared (call psld_op_prepare)");
#define ATTN_LAUNCH(PROJ, X3)                                                                  \
  PSLD_CHECK_CUDA(launch_pdl(attn_tc_kernel<PROJ, X3>, st->grid, dim3(AT_THREADS),                \
                             PROJ ? AtCfg<X3>::kSmemProj : AtCfg<X3>::kSmem, s, 1, st->tq, st->tkv, \
                             st->tw, st->p))
  if (st->proj) { if (st->x3) ATTN_LAUNCH(true, true); else ATTN_LAUNCH(true, false); }
  else { if (st->x3) ATTN_LAUNCH(false, true); else ATTN_LAUNCH(false, false); }
#undef ATTN_LAUNCH
  PSLD_CHECK_LAUNCH();
  return PSLD_OK;
}

}  // namespace psld

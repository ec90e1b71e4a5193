// tcgen05 flash attention for the causal-LM path (head_dim 64, GQA): forward.
//
// One CTA per (128-query tile, head, batch); 6 warps, warp-specialised like the GEMM:
//   warp 0      TMA producer: Q tile once; K and V tiles (128 keys) in single buffers with separate full/empty barriers
//               (K_{j+1} streams in as soon as S_j's MMAs retire, V_{j+1} as soon as the PV_j MMAs retire)
//   warp 1      MMA issuer  : S = Q K^T (128x128x64, fp32 in TMEM) and O += P V (128x64x128), tcgen05.mma kind::f16
//   warps 2..5  softmax     : ONE THREAD PER QUERY ROW (TMEM lane == row, so row max / sum need no shuffles):
//                             tcgen05.ld S -> scale, causal mask, online softmax (exp2) -> P (bf16) written straight
//                             into the 128B-swizzled K-major smem layout the PV MMA reads as its A operand; when the
//                             running max moved, O is rescaled in TMEM (tcgen05.ld / tcgen05.st).
// TMEM: 256 columns per CTA (S: 128, O: 64) so two CTAs are resident per SM and one CTA's softmax overlaps the other's
// MMAs.  q/k/v are column slices of the fused projection [B*T, ld]; one tensor map serves all three.
#include "kernels.h"
#include <cudaTypedefs.h>
#include <stdlib.h>

namespace {

constexpr int AT_BR = 128;   // query rows per CTA
constexpr int AT_BC = 128;   // keys per step
constexpr int AT_THREADS = 192;
constexpr uint32_t SQ_BYTES = AT_BR * 128;        // 16 KB: [128 rows][64 dims] bf16
constexpr uint32_t SKV_BYTES = AT_BC * 128;       // 16 KB each for K and V
constexpr uint32_t SP_BYTES = 2 * AT_BR * 128;    // 32 KB: two 64-key K-blocks of [128 rows][64 keys]
constexpr uint32_t AT_SMEM = SQ_BYTES + 3 * SKV_BYTES + SP_BYTES + 256 + 1024;   // 97.25 KB -> 2 CTAs / SM

SK_DEVINL void tmem_ld_32(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }
SK_DEVINL void tmem_st_32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
SK_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ o, float* __restrict__ lse, int T, int ldo,
                   int H, int KVH, float scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + SQ_BYTES;           // two K stages
  const uint32_t sV = sK + 2 * SKV_BYTES;
  const uint32_t sP = sV + SKV_BYTES;
  const uint32_t bar = sP + SP_BYTES;
  const uint32_t q_full = bar, k_full = bar + 8 /*[2]*/, k_empty = bar + 24 /*[2]*/, v_full = bar + 40, v_empty = bar + 48,
                 s_full = bar + 56, s_empty = bar + 64, p_full = bar + 72, o_done = bar + 80, tmem_slot = bar + 88;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (T + AT_BR - 1) / AT_BR;
  const int qt = n_qt - 1 - (int)blockIdx.x;  // heaviest query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int g = h / (H / KVH);
  const int q0 = qt * AT_BR;
  const int row_base = b * T;                 // row of this sequence in the [B*T, ld] activation
  const int n_kv = CAUSAL ? (min(T - 1, q0 + AT_BR - 1) / AT_BC + 1) : (T + AT_BC - 1) / AT_BC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_full + 8, 1);
    mbar_init(k_empty, 1);
    mbar_init(k_empty + 8, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, SQ_BYTES);
      tma_load_2d(sQ, &tmQKV, q_full, h * 64, row_base + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait_sleep(k_empty + 8 * st, ((j >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(k_full + 8 * st, SKV_BYTES);
        tma_load_2d(sK + st * SKV_BYTES, &tmQKV, k_full + 8 * st, (H + g) * 64, row_base + j * AT_BC);
        if (j >= 1) {   // V_{j-1} is issued one step behind K_j so that K_0 and K_1 go out back to back
          mbar_wait_sleep(v_empty, ((j - 1) & 1) ^ 1u);
          mbar_arrive_expect_tx(v_full, SKV_BYTES);
          tma_load_2d(sV, &tmQKV, v_full, (H + KVH + g) * 64, row_base + (j - 1) * AT_BC);
        }
      }
      mbar_wait_sleep(v_empty, ((n_kv - 1) & 1) ^ 1u);
      mbar_arrive_expect_tx(v_full, SKV_BYTES);
      tma_load_2d(sV, &tmQKV, v_full, (H + KVH + g) * 64, row_base + (n_kv - 1) * AT_BC);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(1u, 0u, 0u, 128, 128);   // S = Q K^T: A, B K-major
      constexpr uint32_t idesc_o = umma_idesc(1u, 0u, 1u, 128, 64);    // O += P V: A K-major, B (V) MN-major
      auto issue_s = [&](int j) {
        const uint32_t sKj = sK + (j & 1) * SKV_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tS, umma_desc_sw128(sQ + k * 32, 16, 1024), umma_desc_sw128(sKj + k * 32, 16, 1024), idesc_s, k > 0);
        tc_commit(s_full);
        tc_commit(k_empty + 8 * (j & 1));      // K stage free once these MMAs retire
      };
      mbar_wait_sleep(q_full, 0);
      mbar_wait_sleep(k_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          // S_{j+1} is issued BEFORE P_j is needed: the softmax warps copied S_j to registers (s_empty), so the tensor
          // pipe computes the next scores while they do exp / pack / correction for this tile
          mbar_wait_sleep(k_full + 8 * ((j + 1) & 1), ((j + 1) >> 1) & 1);
          mbar_wait_sleep(s_empty, j & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait_sleep(p_full, j & 1);          // P_j in smem, O rescaled
        mbar_wait_sleep(v_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t pa = sP + (k >> 2) * (AT_BR * 128) + (k & 3) * 32;
          tc_mma_f16(tO, umma_desc_sw128(pa, 16, 1024), umma_desc_sw128(sV + k * 2048, 8192, 1024), idesc_o,
                     (j > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(v_empty);
        tc_commit(o_done);
      }
    }
  } else {
    // ===== softmax / correction / epilogue: thread == query row =====
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // row inside the tile == TMEM lane
    const int qrow = q0 + r;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float sl2 = scale * 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int k0 = j * AT_BC;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const bool need_mask = (CAUSAL && k0 + AT_BC - 1 > q0) || (k0 + AT_BC > T);
      // the whole 128-key row of S lives in registers: ONE pass over TMEM (all four loads in flight, one wait)
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32(tS + lane_off + c * 32, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);       // S consumed: the next S MMA may overwrite it
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int key = k0 + c * 32 + i;
            if (key >= T || (CAUSAL && key > qrow)) v[c][i] = 0xff800000u;  // -inf
          }
      }
      float mx = m_run;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[c][i]));
      const float m_new = mx;
      const float mb = (m_new == -INFINITY) ? 0.f : m_new * sl2;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run * sl2 - mb);
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[c][2 * i]), sl2, -mb));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[c][2 * i + 1]), sl2, -mb));
          rs0 += p0;
          rs1 += p1;
          pk[i] = pack_bf16(p0, p1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int chunk = c * 4 + u;           // 16-byte chunk (8 keys) index 0..15 within the 128-key row
          const uint32_t dst = sP + (chunk >> 3) * (AT_BR * 128) + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * u]), "r"(pk[4 * u + 1]),
                       "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3])
                       : "memory");
        }
      }
      const float rs = rs0 + rs1;

      l_run = l_run * alpha + rs;
      m_run = m_new;
      // correction: rescale O when this row's max moved (skipped warp-wide when no lane needs it)
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);          // previous PV MMA retired
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32(tO + lane_off + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32(tO + lane_off + c * 32, v);
          }
          tmem_st_wait();
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue
    mbar_wait(o_done, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
    bf16* op = o + ((size_t)(row_base + qrow)) * ldo + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32(tO + lane_off + c * 32, v);
      tmem_ld_wait();
      if (qrow < T) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(v[8 * u]) * inv, __uint_as_float(v[8 * u + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(v[8 * u + 2]) * inv, __uint_as_float(v[8 * u + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(v[8 * u + 4]) * inv, __uint_as_float(v[8 * u + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(v[8 * u + 6]) * inv, __uint_as_float(v[8 * u + 7]) * inv);
          stg128(op + c * 32 + u * 8, w);
        }
      }
    }
    if (lse && qrow < T) lse[((size_t)b * H + h) * T + qrow] = m_run * scale + logf(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// q/k/v: column slices of one [B*T, ld] bf16 buffer starting at `qkv` (q heads first, then KVH k heads, then KVH v heads)
int sk_attn_tc_fwd_launch(const bf16* qkv, bf16* o, float* lse, int B, int T, int H, int KVH, int ld, int ldo, int causal,
                          float scale, cudaStream_t s) {
  SK_REQUIRE(H % KVH == 0, "attention: H must be a multiple of KVH");
  SK_REQUIRE(ld % 8 == 0 && ldo % 8 == 0, "attention: leading dims must be multiples of 8");
  CUtensorMap tm;
  int rc = sk_make_tmap_2d(&tm, qkv, 2, (uint64_t)(H + 2 * KVH) * 64, (uint64_t)B * T, (uint64_t)ld, 64, 128);
  if (rc) return rc;
  static bool init = false;
  if (!init) {
    SK_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    SK_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    init = true;
  }
  dim3 grid((T + AT_BR - 1) / AT_BR, H, B);
  sk_prof_begin(1, s);
  if (causal) attn_tc_fwd_kernel<true><<<grid, AT_THREADS, AT_SMEM, s>>>(tm, o, lse, T, ldo, H, KVH, scale);
  else attn_tc_fwd_kernel<false><<<grid, AT_THREADS, AT_SMEM, s>>>(tm, o, lse, T, ldo, H, KVH, scale);
  sk_prof_end(s);
  SK_LAUNCH_CHECK();
  return 0;
}

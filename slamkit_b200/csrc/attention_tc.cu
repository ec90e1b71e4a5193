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
#include <type_traits>
#define SK_TRY_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

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
SK_DEVINL void tmem_ld_16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// inverse rotary embedding (the backward of apply_rotary_pos_emb) on 8 (x1, x2) pairs held as packed bf16, with the
// rounding points of the bf16 autograd chain (the same as rope_kernel's inverse mode): o1 = bf16(x1 c) + bf16(x2 s),
// o2 = bf16(x2 c) + bf16(-x1 s)
SK_DEVINL void rope_inv8(const uint32_t (&x1)[4], const uint32_t (&x2)[4], const uint4 cv, const uint4 sv, uint4& o1, uint4& o2) {
  const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w}, sw[4] = {sv.x, sv.y, sv.z, sv.w};
  uint32_t a[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 p = unpack_bf16(x1[k]), q = unpack_bf16(x2[k]), c = unpack_bf16(cw[k]), sn = unpack_bf16(sw[k]);
    a[k] = pack_bf16(bf16_round(p.x * c.x) + bf16_round(q.x * sn.x), bf16_round(p.y * c.y) + bf16_round(q.y * sn.y));
    b[k] = pack_bf16(bf16_round(q.x * c.x) + bf16_round(-p.x * sn.x), bf16_round(q.y * c.y) + bf16_round(-p.y * sn.y));
  }
  o1 = make_uint4(a[0], a[1], a[2], a[3]);
  o2 = make_uint4(b[0], b[1], b[2], b[3]);
}

template <bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ o, float* __restrict__ lse, int T, int ldo,
                   int H, int KVH, float scale, const int* __restrict__ seg_start) {
  griddep_launch();
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
  // grid = (head, batch, tile): CTAs are dispatched x-fastest, so every (head, batch) pair's heaviest causal tile is
  // handed out before any lighter one (longest-processing-time order: ~95 % slot efficiency instead of ~74 % when the
  // tile index varies fastest)
  const int qt = n_qt - 1 - (int)blockIdx.z;  // heaviest query tiles first
  const int h = blockIdx.x, b = blockIdx.y;
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
  griddep_wait();   // prologue (barriers, TMEM) may overlap the previous kernel's tail; its outputs are visible from here
  const uint32_t tS = tmem_base, tO = tmem_base + 128;
  // Packed batches (several documents in one row, position_ids restarting at 0): seg_start[token] is the in-row index of
  // the first token of its document.  Keys before it are masked (block-diagonal causal attention, what HF's varlen
  // flash-attention path computes from the same position_ids) and whole key tiles before the tile's first document
  // are skipped (seg_start is non-decreasing along a row, so the first query row has the smallest bound).
  const int j_begin = seg_start ? seg_start[row_base + q0] / AT_BC : 0;
  const int n_it = n_kv - j_begin;             // loop counters below start at 0 (barrier parities), tile index = j_begin + jj

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, SQ_BYTES);
      tma_load_2d(sQ, &tmQKV, q_full, h * 64, row_base + q0);
      for (int j = 0; j < n_it; ++j) {
        const int st = j & 1;
        mbar_wait_sleep(k_empty + 8 * st, ((j >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(k_full + 8 * st, SKV_BYTES);
        tma_load_2d(sK + st * SKV_BYTES, &tmQKV, k_full + 8 * st, (H + g) * 64, row_base + (j_begin + j) * AT_BC);
        if (j >= 1) {   // V_{j-1} is issued one step behind K_j so that K_0 and K_1 go out back to back
          mbar_wait_sleep(v_empty, ((j - 1) & 1) ^ 1u);
          mbar_arrive_expect_tx(v_full, SKV_BYTES);
          tma_load_2d(sV, &tmQKV, v_full, (H + KVH + g) * 64, row_base + (j_begin + j - 1) * AT_BC);
        }
      }
      mbar_wait_sleep(v_empty, ((n_it - 1) & 1) ^ 1u);
      mbar_arrive_expect_tx(v_full, SKV_BYTES);
      tma_load_2d(sV, &tmQKV, v_full, (H + KVH + g) * 64, row_base + (j_begin + n_it - 1) * AT_BC);
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
      for (int j = 0; j < n_it; ++j) {
        if (j + 1 < n_it) {
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
    const int lb = (seg_start && qrow < T) ? seg_start[row_base + qrow] : 0;   // first visible key of this row
    for (int j = 0; j < n_it; ++j) {
      const int k0 = (j_begin + j) * AT_BC;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const bool need_mask = (CAUSAL && k0 + AT_BC - 1 > q0) || (k0 + AT_BC > T) || (k0 < lb);
      // the whole 128-key row of S lives in registers: ONE pass over TMEM (all four loads in flight, one wait)
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32(tS + lane_off + c * 32, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);       // S consumed: the next S MMA may overwrite it
      // cmax: last 32-key group of this tile the warp's rows can see (warp-uniform).  On the causal diagonal tile
      // (k0 == q0, BR == BC) the rows of TMEM quadrant q see groups 0..q-1 completely, group q up to their own index,
      // and nothing of the groups behind: those are neither compared, nor exponentiated (62 % of a tile's MUFU work on
      // average), they are written to P as zeros.
      int cmax = 3;
      if (need_mask) {
        if (CAUSAL && seg_start == nullptr && k0 == q0 && k0 + AT_BC <= T) {
          cmax = q;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c == q) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i > lane) v[c][i] = 0xff800000u;  // -inf
            }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int key = k0 + c * 32 + i;
              if (key >= T || (CAUSAL && key > qrow) || key < lb) v[c][i] = 0xff800000u;  // -inf
            }
        }
      }
      // row max: independent chains (one per 32-key group) instead of one 128-long dependent chain
      float mc[4] = {m_run, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c <= cmax) {
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            mc[c] = fmaxf(mc[c], fmaxf(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])));
        }
      const float m_new = fmaxf(fmaxf(mc[0], mc[1]), fmaxf(mc[2], mc[3]));
      const float mb = (m_new == -INFINITY) ? 0.f : m_new * sl2;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run * sl2 - mb);
      // the P buffer (and O) are still being read by the previous PV MMA until o_done: wait before overwriting P
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);
        tc_fence_after();
      }
      // p = 2^(s * scale*log2e - m * scale*log2e): packed FFMA2 for the argument, one MUFU.EX2 per element, packed FADD2
      // row sums in four independent chains
      const f32x2 sl22 = dup2(sl2), nmb2 = dup2(-mb);
      f32x2 rsum[4] = {dup2(0.f), dup2(0.f), dup2(0.f), dup2(0.f)};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
        if (c <= cmax) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0, x1;
            upk2(fma2(pk2(__uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1])), sl22, nmb2), x0, x1);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            rsum[i & 3] = add2(rsum[i & 3], pk2(p0, p1));
            pk[i] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
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
      float rs;
      {
        float a0, a1;
        upk2(add2(add2(rsum[0], rsum[1]), add2(rsum[2], rsum[3])), a0, a1);
        rs = a0 + a1;
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      // correction: rescale O when this row's max moved (skipped warp-wide when no lane needs it)
      if (j > 0) {
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
    mbar_wait(o_done, (n_it - 1) & 1);
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
                          float scale, cudaStream_t s, const int* seg_start) {
  SK_REQUIRE(seg_start == nullptr || causal, "attn_tc_fwd: document segments need the causal kernel");
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
  dim3 grid(H, B, (T + AT_BR - 1) / AT_BR);   // tile index slowest: see the kernel's note on dispatch order
  sk_prof_begin(1, s);
  if (causal) SK_CUDA_CHECK(sk_launch_pdl(attn_tc_fwd_kernel<true>, dim3(grid), dim3(AT_THREADS), (size_t)(AT_SMEM), s, tm, o, lse, T, ldo, H, KVH, scale, seg_start));
  else SK_CUDA_CHECK(sk_launch_pdl(attn_tc_fwd_kernel<false>, dim3(grid), dim3(AT_THREADS), (size_t)(AT_SMEM), s, tm, o, lse, T, ldo, H, KVH, scale, seg_start));
  sk_prof_end(s);
  SK_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// Split-precision (hi, lo) bidirectional forward for the HuBERT encoder on tcgen05: same pipeline as attn_tc_fwd_kernel,
// but every operand is a bf16 (hi, lo) pair and each product is three MMAs into the same TMEM accumulator
//   S = Ql Kh^T + Qh Kl^T + Qh Kh^T ,   O += Pl Vh + Ph Vl + Ph Vh        (fp32-grade, see hubert_kernels.cu)
// One CTA per SM (192 KB of operand tiles); S_{j+1} is issued while the softmax warps work on tile j.
// =================================================================================================
namespace {

constexpr uint32_t ATS_SMEM = 2 * SQ_BYTES + 4 * SKV_BYTES + 2 * SKV_BYTES + 2 * SP_BYTES + 256 + 1024;   // ~193 KB

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_fwd_split_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmL,
                         bf16* __restrict__ o_hi, bf16* __restrict__ o_lo, int T, int ldo, int H, float scale) {
  griddep_launch();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQh = smem_base, sQl = sQh + SQ_BYTES;
  const uint32_t sK = sQl + SQ_BYTES;            // stage s: Kh at sK + s*32K, Kl at +16K
  const uint32_t sVh = sK + 4 * SKV_BYTES, sVl = sVh + SKV_BYTES;
  const uint32_t sPh = sVl + SKV_BYTES, sPl = sPh + SP_BYTES;
  const uint32_t bar = sPl + SP_BYTES;
  const uint32_t q_full = bar, k_full = bar + 8 /*[2]*/, k_empty = bar + 24 /*[2]*/, v_full = bar + 40, v_empty = bar + 48,
                 s_full = bar + 56, s_empty = bar + 64, p_full = bar + 72, o_done = bar + 80, tmem_slot = bar + 88;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = (int)blockIdx.x * AT_BR;
  const int row_base = b * T;
  const int n_kv = (T + AT_BC - 1) / AT_BC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmL);
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
  griddep_wait();   // prologue (barriers, TMEM) may overlap the previous kernel's tail; its outputs are visible from here
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * SQ_BYTES);
      tma_load_2d(sQh, &tmH, q_full, h * 64, row_base + q0);
      tma_load_2d(sQl, &tmL, q_full, h * 64, row_base + q0);
      auto load_v = [&](int j) {
        mbar_wait_sleep(v_empty, (j & 1) ^ 1u);
        mbar_arrive_expect_tx(v_full, 2 * SKV_BYTES);
        tma_load_2d(sVh, &tmH, v_full, (2 * H + h) * 64, row_base + j * AT_BC);
        tma_load_2d(sVl, &tmL, v_full, (2 * H + h) * 64, row_base + j * AT_BC);
      };
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait_sleep(k_empty + 8 * st, ((j >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(k_full + 8 * st, 2 * SKV_BYTES);
        tma_load_2d(sK + st * 2 * SKV_BYTES, &tmH, k_full + 8 * st, (H + h) * 64, row_base + j * AT_BC);
        tma_load_2d(sK + st * 2 * SKV_BYTES + SKV_BYTES, &tmL, k_full + 8 * st, (H + h) * 64, row_base + j * AT_BC);
        if (j >= 1) load_v(j - 1);
      }
      load_v(n_kv - 1);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(1u, 0u, 0u, 128, 128);
      constexpr uint32_t idesc_o = umma_idesc(1u, 0u, 1u, 128, 64);
      auto issue_s = [&](int j) {
        const uint32_t kh = sK + (j & 1) * 2 * SKV_BYTES, kl = kh + SKV_BYTES;
        const uint32_t qa[3] = {sQl, sQh, sQh};      // small terms first
        const uint32_t kb[3] = {kh, kl, kh};
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tS, umma_desc_sw128(qa[t] + k * 32, 16, 1024), umma_desc_sw128(kb[t] + k * 32, 16, 1024), idesc_s,
                       (t > 0 || k > 0) ? 1u : 0u);
        tc_commit(s_full);
        tc_commit(k_empty + 8 * (j & 1));
      };
      mbar_wait_sleep(q_full, 0);
      mbar_wait_sleep(k_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          mbar_wait_sleep(k_full + 8 * ((j + 1) & 1), ((j + 1) >> 1) & 1);
          mbar_wait_sleep(s_empty, j & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait_sleep(p_full, j & 1);
        mbar_wait_sleep(v_full, j & 1);
        tc_fence_after();
        const uint32_t pa[3] = {sPl, sPh, sPh};
        const uint32_t vb[3] = {sVh, sVl, sVh};
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tc_mma_f16(tO, umma_desc_sw128(pa[t] + (k >> 2) * (AT_BR * 128) + (k & 3) * 32, 16, 1024),
                       umma_desc_sw128(vb[t] + k * 2048, 8192, 1024), idesc_o, (j > 0 || t > 0 || k > 0) ? 1u : 0u);
        tc_commit(v_empty);
        tc_commit(o_done);
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int qrow = q0 + r;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float sl2 = scale * 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int k0 = j * AT_BC;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32(tS + lane_off + c * 32, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      if (k0 + AT_BC > T) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (k0 + c * 32 + i >= T) v[c][i] = 0xff800000u;
      }
      float mx = m_run;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[c][i]));
      const float m_new = mx;
      const float mb = m_new * sl2;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run * sl2 - mb);
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);
        tc_fence_after();
      }
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // exp2f (not the approx form): this path must stay fp32-grade
          const float p0 = exp2f(fmaf(__uint_as_float(v[c][2 * i]), sl2, -mb));
          const float p1 = exp2f(fmaf(__uint_as_float(v[c][2 * i + 1]), sl2, -mb));
          rs0 += p0;
          rs1 += p1;
          const float h0 = bf16_round(p0), h1 = bf16_round(p1);
          ph[i] = pack_bf16(h0, h1);
          pl[i] = pack_bf16(p0 - h0, p1 - h1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int chunk = c * 4 + u;
          const uint32_t off = (chunk >> 3) * (AT_BR * 128) + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPh + off), "r"(ph[4 * u]), "r"(ph[4 * u + 1]),
                       "r"(ph[4 * u + 2]), "r"(ph[4 * u + 3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPl + off), "r"(pl[4 * u]), "r"(pl[4 * u + 1]),
                       "r"(pl[4 * u + 2]), "r"(pl[4 * u + 3]) : "memory");
        }
      }
      l_run = l_run * alpha + (rs0 + rs1);
      m_run = m_new;
      if (j > 0) {
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t w[32];
            tmem_ld_32(tO + lane_off + c * 32, w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * alpha);
            tmem_st_32(tO + lane_off + c * 32, w);
          }
          tmem_st_wait();
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_done, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const size_t off = ((size_t)(row_base + qrow)) * ldo + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t w[32];
      tmem_ld_32(tO + lane_off + c * 32, w);
      tmem_ld_wait();
      if (qrow < T) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t hh[4], ll[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a0 = __uint_as_float(w[8 * u + 2 * e]) * inv, a1 = __uint_as_float(w[8 * u + 2 * e + 1]) * inv;
            const float h0 = bf16_round(a0), h1 = bf16_round(a1);
            hh[e] = pack_bf16(h0, h1);
            ll[e] = pack_bf16(a0 - h0, a1 - h1);
          }
          stg128(o_hi + off + c * 32 + u * 8, make_uint4(hh[0], hh[1], hh[2], hh[3]));
          stg128(o_lo + off + c * 32 + u * 8, make_uint4(ll[0], ll[1], ll[2], ll[3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// qkv_hi / qkv_lo: [B*T, ld] with H q-heads, H k-heads, H v-heads (64 columns each); o_hi / o_lo: [B*T, ldo]
int sk_attn_tc_fwd_split_launch(const bf16* qkv_hi, const bf16* qkv_lo, bf16* o_hi, bf16* o_lo, int B, int T, int H, int ld,
                                int ldo, float scale, cudaStream_t s) {
  SK_REQUIRE(ld % 8 == 0 && ldo % 8 == 0, "attention: leading dims must be multiples of 8");
  CUtensorMap th, tl;
  int rc;
  if ((rc = sk_make_tmap_2d(&th, qkv_hi, 2, (uint64_t)3 * H * 64, (uint64_t)B * T, (uint64_t)ld, 64, 128))) return rc;
  if ((rc = sk_make_tmap_2d(&tl, qkv_lo, 2, (uint64_t)3 * H * 64, (uint64_t)B * T, (uint64_t)ld, 64, 128))) return rc;
  static bool init = false;
  if (!init) {
    SK_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_fwd_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATS_SMEM));
    init = true;
  }
  dim3 grid((T + AT_BR - 1) / AT_BR, H, B);
  sk_prof_begin(1, s);
  attn_tc_fwd_split_kernel<<<grid, AT_THREADS, ATS_SMEM, s>>>(th, tl, o_hi, o_lo, T, ldo, H, scale);
  sk_prof_end(s);
  SK_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// Backward on tcgen05.  Two deterministic kernels like the warp-level version (no atomics):
//   dQ    : CTA = (128-query tile, head, batch), loops over 64-key tiles; S and dP accumulate in TMEM, the thread that
//           owns a query row turns them into dS (bf16, swizzled smem), dQ += dS K on the tensor pipe.
//   dK/dV : CTA = (128-key tile, query head, batch), loops over 64-query tiles on transposed scores S^T = K Q^T and
//           dP^T = V dO^T (thread == key row); P^T and dS^T go to smem as A operands of dV += P^T dO, dK += dS^T Q.
//           Per-head fp32 partials are reduced over the GQA group in a fixed order by attn_tc_group_reduce_kernel.
// K / V / Q / dO tiles that act as the B operand of the second GEMMs are read MN-major straight from the row-major
// activations (no transposes anywhere).
// =================================================================================================
namespace {

constexpr int BWD_THREADS = 320;                            // TMA warp, MMA warp, 8 element-wise warps (2 per TMEM quadrant)
constexpr int BQ_BC = 64;                                   // keys per step in the dQ kernel
constexpr uint32_t T64_BYTES = 64 * 128;                    // [64 rows][64 dims] bf16 tile
constexpr uint32_t DQ_SMEM = 2 * SQ_BYTES + 4 * T64_BYTES + SQ_BYTES + SQ_BYTES + 1024 + 256 + 1024;   // Q, dO, K[2], V[2], dS, O, delta halves

template <bool CAUSAL>
__global__ void __launch_bounds__(BWD_THREADS, 2)
attn_tc_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                      const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmO,
                      const float* __restrict__ lse, float* __restrict__ delta_out, bf16* __restrict__ dq, int T, int ldg,
                      int H, int KVH, float scale, const int* __restrict__ seg_start, const bf16* __restrict__ rope_cos,
                      const bf16* __restrict__ rope_sin, const int* __restrict__ pos_ids, int max_pos) {
  // This kernel runs FIRST in the backward pass: it also produces delta[b,h,t] = sum_d dO*O for its 128 query rows (the
  // O tile rides along with Q and dO) and writes it for the dK/dV kernel, and -- when rope tables are given -- applies
  // the inverse rotary embedding to dQ in its epilogue (no separate delta / rope kernels, no extra pass over dqkv).
  griddep_launch();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sdO = sQ + SQ_BYTES, sK = sdO + SQ_BYTES, sV = sK + 2 * T64_BYTES,
                 sdS = sV + 2 * T64_BYTES, sO = sdS + SQ_BYTES, sDel = sO + SQ_BYTES, bar = sDel + 1024;
  float* del_ptr = reinterpret_cast<float*>(smem_raw + (sDel - smem_u32(smem_raw)));   // [2 halves][128 rows]
  const uint32_t q_full = bar, k_full = bar + 8, k_empty = bar + 24, v_full = bar + 40, v_empty = bar + 56,
                 sdp_full = bar + 72, sdp_empty = bar + 80, ds_full = bar + 88, ds_empty = bar + 96, dq_done = bar + 104,
                 tmem_slot = bar + 112;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (T + AT_BR - 1) / AT_BR;
  const int qt = n_qt - 1 - (int)blockIdx.z;   // heaviest query tiles first, tile index slowest (see the forward)
  const int h = blockIdx.x, b = blockIdx.y;
  const int g = h / (H / KVH);
  const int q0 = qt * AT_BR;
  const int row_base = b * T;
  const int n_kv = CAUSAL ? (min(T - 1, q0 + AT_BR - 1) / BQ_BC + 1) : (T + BQ_BC - 1) / BQ_BC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV128);
    tma_prefetch_desc(&tmQKV64);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full + 8 * s, 1);
      mbar_init(k_empty + 8 * s, 1);
      mbar_init(v_full + 8 * s, 1);
      mbar_init(v_empty + 8 * s, 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, 8);
    mbar_init(ds_full, 8);
    mbar_init(ds_empty, 1);
    mbar_init(dq_done, 1);
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
  griddep_wait();   // prologue (barriers, TMEM) may overlap the previous kernel's tail; its outputs are visible from here
  const uint32_t tS = tmem_base, tdP = tmem_base + 64, tdQ = tmem_base + 128;
  // packed batches: skip key tiles before the first document of this query tile, mask keys before each row's document
  const int j_begin = seg_start ? seg_start[row_base + q0] / BQ_BC : 0;
  const int n_it = n_kv - j_begin;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 3 * SQ_BYTES);
      tma_load_2d(sQ, &tmQKV128, q_full, h * 64, row_base + q0);
      tma_load_2d(sdO, &tmDO, q_full, h * 64, row_base + q0);
      tma_load_2d(sO, &tmO, q_full, h * 64, row_base + q0);
      for (int j = 0; j < n_it; ++j) {
        const int st = j & 1;
        const uint32_t par = ((j >> 1) & 1) ^ 1u;
        mbar_wait_sleep(k_empty + 8 * st, par);
        mbar_arrive_expect_tx(k_full + 8 * st, T64_BYTES);
        tma_load_2d(sK + st * T64_BYTES, &tmQKV64, k_full + 8 * st, (H + g) * 64, row_base + (j_begin + j) * BQ_BC);
        mbar_wait_sleep(v_empty + 8 * st, par);
        mbar_arrive_expect_tx(v_full + 8 * st, T64_BYTES);
        tma_load_2d(sV + st * T64_BYTES, &tmQKV64, v_full + 8 * st, (H + KVH + g) * 64, row_base + (j_begin + j) * BQ_BC);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(1u, 0u, 0u, 128, 64);    // S = Q K^T, dP = dO V^T (N = 64 keys)
      constexpr uint32_t idesc_q = umma_idesc(1u, 0u, 1u, 128, 64);    // dQ += dS K : B = K tile MN-major
      auto issue_sdp = [&](int j) {
        const uint32_t sKj = sK + (j & 1) * T64_BYTES, sVj = sV + (j & 1) * T64_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tS, umma_desc_sw128(sQ + k * 32, 16, 1024), umma_desc_sw128(sKj + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tdP, umma_desc_sw128(sdO + k * 32, 16, 1024), umma_desc_sw128(sVj + k * 32, 16, 1024), idesc_s, k > 0);
        tc_commit(sdp_full);
        tc_commit(v_empty + 8 * (j & 1));
      };
      mbar_wait_sleep(q_full, 0);
      mbar_wait_sleep(k_full, 0);
      mbar_wait_sleep(v_full, 0);
      tc_fence_after();
      issue_sdp(0);
      for (int j = 0; j < n_it; ++j) {
        if (j + 1 < n_it) {
          const int st = (j + 1) & 1;
          mbar_wait_sleep(k_full + 8 * st, ((j + 1) >> 1) & 1);
          mbar_wait_sleep(v_full + 8 * st, ((j + 1) >> 1) & 1);
          mbar_wait_sleep(sdp_empty, j & 1);
          tc_fence_after();
          issue_sdp(j + 1);
        }
        mbar_wait_sleep(ds_full, j & 1);
        tc_fence_after();
        const uint32_t sKj = sK + (j & 1) * T64_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tdQ, umma_desc_sw128(sdS + k * 32, 16, 1024), umma_desc_sw128(sKj + k * 2048, 8192, 1024), idesc_q,
                     (j > 0 || k > 0) ? 1u : 0u);
        tc_commit(k_empty + 8 * (j & 1));
        tc_commit(ds_empty);
        if (j == n_it - 1) tc_commit(dq_done);
      }
    }
  } else {
    // 8 element-wise warps: two per TMEM lane quadrant, each owning one 32-key half of the 64-key tile (P and dS are
    // purely element-wise given lse / delta, so the halves never need to talk to each other)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int qrow = q0 + r;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float sl2 = scale * 1.4426950408889634f;
    const bool row_ok = qrow < T;
    const size_t soff = ((size_t)b * H + h) * T + (row_ok ? qrow : 0);
    // packed-pair math (two keys per instruction): p = 2^(s*sl2 - lse2), dS' = p * (dP - delta); the 1/sqrt(d) factor of
    // dS is applied once to dQ in the epilogue instead of to every score
    const f32x2 nlse2 = dup2(row_ok ? -lse[soff] * 1.4426950408889634f : 0.f);
    // delta of this row: each of the row's two warps sums its 32 dims of dO*O from the swizzled tiles (8 dims per 16-byte
    // chunk, chunks paired up in a fixed tree), the halves meet through shared memory
    mbar_wait(q_full, 0);
    {
      float pc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((((half * 4 + c) ^ (r & 7))) << 4);
        uint32_t a[4], g[4];
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(sO + off));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(g[0]), "=r"(g[1]), "=r"(g[2]), "=r"(g[3]) : "r"(sdO + off));
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 x = unpack_bf16(a[k]), y = unpack_bf16(g[k]);
          acc += x.x * y.x + x.y * y.y;
        }
        pc[c] = acc;
      }
      del_ptr[half * 128 + r] = (pc[0] + pc[1]) + (pc[2] + pc[3]);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float delta_row = del_ptr[r] + del_ptr[128 + r];
    if (half == 0 && row_ok) delta_out[soff] = delta_row;
    const f32x2 del2 = dup2(row_ok ? delta_row : 0.f);
    const f32x2 sl22 = dup2(sl2);
    const int lb = (seg_start && row_ok) ? seg_start[row_base + qrow] : 0;   // first visible key of this row
    for (int j = 0; j < n_it; ++j) {
      const int k0 = (j_begin + j) * BQ_BC;
      mbar_wait(sdp_full, j & 1);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld_32(tS + lane_off + half * 32, sv);
      tmem_ld_32(tdP + lane_off + half * 32, dv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_empty);
      const bool need_mask = (CAUSAL && k0 + BQ_BC - 1 > q0) || (k0 + BQ_BC > T) || (q0 + AT_BR > T) || (k0 < lb);
      // mask mode of this warp's 32 keys (warp-uniform): 0 all visible, 1 causal diagonal group (key index <= the row's
      // lane), 2 all masked (nothing to compute: dS = 0), 3 general (sequence end / document bounds: per element)
      int mode = 0;
      if (need_mask) {
        mode = 3;
        if (CAUSAL && seg_start == nullptr && q0 + AT_BR <= T && k0 + BQ_BC <= T) {
          const int g = ((k0 - q0) >> 5) + half;       // this warp's key group, in 32-key units from the tile's first row
          mode = g < q ? 0 : (g == q ? 1 : 2);
        }
      }
      uint32_t pk[16];
      // one straight-line instance of the 16-pair loop per mode (a mode test inside the unrolled loop leaves a branch per
      // pair in the SASS and keeps the scheduler from interleaving the pairs' MUFU / FMA chains)
      auto body = [&](auto mode_c) {
        constexpr int MODE = decltype(mode_c)::value;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x0, x1;
          upk2(fma2(pk2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), sl22, nlse2), x0, x1);
          float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          if (MODE == 1) {
            if (2 * i > lane) p0 = 0.f;
            if (2 * i + 1 > lane) p1 = 0.f;
          } else if (MODE == 3) {
            const int key = k0 + half * 32 + 2 * i;
            if (!row_ok || key >= T || (CAUSAL && key > qrow) || key < lb) p0 = 0.f;
            if (!row_ok || key + 1 >= T || (CAUSAL && key + 1 > qrow) || key + 1 < lb) p1 = 0.f;
          }
          float d0, d1;
          upk2(mul2(pk2(p0, p1), sub2(pk2(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1])), del2)), d0, d1);
          pk[i] = pack_bf16(d0, d1);
        }
      };
      if (mode == 0) body(std::integral_constant<int, 0>{});
      else if (mode == 1) body(std::integral_constant<int, 1>{});
      else if (mode == 3) body(std::integral_constant<int, 3>{});
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      mbar_wait(ds_empty, (j & 1) ^ 1u);         // previous dQ MMA finished reading the dS buffer (the math above overlaps it)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int chunk = half * 4 + u;          // 8 chunks of 8 keys in the 64-key row
        const uint32_t dst = sdS + r * 128 + ((chunk ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * u]), "r"(pk[4 * u + 1]),
                     "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3])
                     : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(dq_done, 0);
    tc_fence_after();
    bf16* op = dq + ((size_t)(row_base + qrow)) * ldg + h * 64;
    if (rope_cos != nullptr) {
      // dims [16 half, 16 half + 16) pair with dims 32 further: this warp rotates its 16 pairs
      uint32_t a[16], b2[16];
      tmem_ld_16(tdQ + lane_off + half * 16, a);
      tmem_ld_16(tdQ + lane_off + 32 + half * 16, b2);
      tmem_ld_wait();
      if (row_ok) {
        int pos = pos_ids ? pos_ids[row_base + qrow] : qrow;
        pos = max(0, min(pos, max_pos - 1));
        uint32_t x1[8], x2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          x1[i] = pack_bf16(__uint_as_float(a[2 * i]) * scale, __uint_as_float(a[2 * i + 1]) * scale);
          x2[i] = pack_bf16(__uint_as_float(b2[2 * i]) * scale, __uint_as_float(b2[2 * i + 1]) * scale);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint4 cv = ldg128(rope_cos + (size_t)pos * 32 + half * 16 + u * 8), sv = ldg128(rope_sin + (size_t)pos * 32 + half * 16 + u * 8);
          const uint32_t p1[4] = {x1[4 * u], x1[4 * u + 1], x1[4 * u + 2], x1[4 * u + 3]};
          const uint32_t p2[4] = {x2[4 * u], x2[4 * u + 1], x2[4 * u + 2], x2[4 * u + 3]};
          uint4 o1, o2;
          rope_inv8(p1, p2, cv, sv, o1, o2);
          stg128(op + half * 16 + u * 8, o1);
          stg128(op + 32 + half * 16 + u * 8, o2);
        }
      }
    } else {
      uint32_t v[32];
      tmem_ld_32(tdQ + lane_off + half * 32, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(v[8 * u]) * scale, __uint_as_float(v[8 * u + 1]) * scale);
          w.y = pack_bf16(__uint_as_float(v[8 * u + 2]) * scale, __uint_as_float(v[8 * u + 3]) * scale);
          w.z = pack_bf16(__uint_as_float(v[8 * u + 4]) * scale, __uint_as_float(v[8 * u + 5]) * scale);
          w.w = pack_bf16(__uint_as_float(v[8 * u + 6]) * scale, __uint_as_float(v[8 * u + 7]) * scale);
          stg128(op + half * 32 + u * 8, w);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

constexpr int BK_BR = 64;                                   // query rows per step in the dK/dV kernel
constexpr uint32_t DKDV_SMEM = 2 * SKV_BYTES + 4 * T64_BYTES + 2 * SKV_BYTES + 2 * 3 * 64 * 4 + 256 + 1024;

template <bool CAUSAL>
__global__ void __launch_bounds__(BWD_THREADS, 2)
attn_tc_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                        const __grid_constant__ CUtensorMap tmDO64, const float* __restrict__ lse,
                        const float* __restrict__ delta, bf16* __restrict__ partial /*[B][H][T][128]*/, int T, int H,
                        int KVH, float scale, const int* __restrict__ seg_start, const int* __restrict__ seg_end) {
  griddep_launch();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = smem_base, sV = sK + SKV_BYTES, sQ = sV + SKV_BYTES /*[2]*/, sdO = sQ + 2 * T64_BYTES /*[2]*/,
                 sPT = sdO + 2 * T64_BYTES, sdST = sPT + SKV_BYTES, sStat = sdST + SKV_BYTES, bar = sStat + 2 * 3 * 64 * 4;
  const uint32_t kv_full = bar, q_full = bar + 8, q_empty = bar + 24, sdp_full = bar + 40, sdp_empty = bar + 48,
                 pds_full = bar + 56, pds_empty = bar + 64, acc_done = bar + 72, tmem_slot = bar + 80;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* stat_ptr = reinterpret_cast<float*>(smem_raw + (sStat - smem_u32(smem_raw)));   // [stage][lse|delta|seg][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.z;                   // key tile 0 has the most query tiles under the causal mask; tile index
  const int h = blockIdx.x, b = blockIdx.y;    // slowest so the heavy tiles of every (head, batch) are dispatched first
  const int g = h / (H / KVH);
  const int k0 = kt * AT_BC;
  const int row_base = b * T;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV128);
    tma_prefetch_desc(&tmQKV64);
    tma_prefetch_desc(&tmDO64);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full + 8 * s, 1);
      mbar_init(q_empty + 8 * s, 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, 8);
    mbar_init(pds_full, 8);
    mbar_init(pds_empty, 1);
    mbar_init(acc_done, 1);
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
  griddep_wait();   // prologue (barriers, TMEM) may overlap the previous kernel's tail; its outputs are visible from here
  const uint32_t tST = tmem_base, tdPT = tmem_base + 64, tdK = tmem_base + 128, tdV = tmem_base + 192;
  int n_qt = (T + BK_BR - 1) / BK_BR;
  // packed batches: queries of later documents never see these keys -- stop at the end of the document of the tile's
  // last key (seg_end[token] = in-row index one past its document).  Read AFTER griddep_wait: seg_end may have been
  // written by the kernel this launch overlaps with (the PDL rule of common.cuh: no global access before the wait).
  if (seg_end) n_qt = min(n_qt, (seg_end[row_base + min(k0 + AT_BC, T) - 1] + BK_BR - 1) / BK_BR);
  const int qt_begin = CAUSAL ? (k0 / BK_BR) : 0;
  const int n_it = max(0, n_qt - qt_begin);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * SKV_BYTES);
      tma_load_2d(sK, &tmQKV128, kv_full, (H + g) * 64, row_base + k0);
      tma_load_2d(sV, &tmQKV128, kv_full, (H + KVH + g) * 64, row_base + k0);
      for (int i = 0; i < n_it; ++i) {
        const int st = i & 1;
        mbar_wait_sleep(q_empty + 8 * st, ((i >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(q_full + 8 * st, 2 * T64_BYTES);
        tma_load_2d(sQ + st * T64_BYTES, &tmQKV64, q_full + 8 * st, h * 64, row_base + (qt_begin + i) * BK_BR);
        tma_load_2d(sdO + st * T64_BYTES, &tmDO64, q_full + 8 * st, h * 64, row_base + (qt_begin + i) * BK_BR);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(1u, 0u, 0u, 128, 64);    // S^T = K Q^T, dP^T = V dO^T (N = 64 queries)
      constexpr uint32_t idesc_a = umma_idesc(1u, 0u, 1u, 128, 64);    // dV += P^T dO, dK += dS^T Q (B MN-major)
      auto issue_sdp = [&](int i) {
        const uint32_t sQi = sQ + (i & 1) * T64_BYTES, sdOi = sdO + (i & 1) * T64_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tST, umma_desc_sw128(sK + k * 32, 16, 1024), umma_desc_sw128(sQi + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tdPT, umma_desc_sw128(sV + k * 32, 16, 1024), umma_desc_sw128(sdOi + k * 32, 16, 1024), idesc_s, k > 0);
        tc_commit(sdp_full);
      };
      if (n_it > 0) {
        mbar_wait_sleep(kv_full, 0);
        mbar_wait_sleep(q_full, 0);
        tc_fence_after();
        issue_sdp(0);
      }
      for (int i = 0; i < n_it; ++i) {
        if (i + 1 < n_it) {
          mbar_wait_sleep(q_full + 8 * ((i + 1) & 1), ((i + 1) >> 1) & 1);
          mbar_wait_sleep(sdp_empty, i & 1);
          tc_fence_after();
          issue_sdp(i + 1);
        }
        mbar_wait_sleep(pds_full, i & 1);
        tc_fence_after();
        const uint32_t sQi = sQ + (i & 1) * T64_BYTES, sdOi = sdO + (i & 1) * T64_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tdV, umma_desc_sw128(sPT + k * 32, 16, 1024), umma_desc_sw128(sdOi + k * 2048, 8192, 1024), idesc_a,
                     (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tdK, umma_desc_sw128(sdST + k * 32, 16, 1024), umma_desc_sw128(sQi + k * 2048, 8192, 1024), idesc_a,
                     (i > 0 || k > 0) ? 1u : 0u);
        tc_commit(q_empty + 8 * (i & 1));
        tc_commit(pds_empty);
        if (i == n_it - 1) tc_commit(acc_done);
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;            // which 32-query half of the 64-query tile this warp handles
    const int r = q * 32 + lane;                 // key row inside the tile == TMEM lane
    const int key = k0 + r;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const f32x2 sl22 = dup2(scale * 1.4426950408889634f);
    const int tid = threadIdx.x - 64;            // 0..255 among the element-wise warps
    // the 64 lse / delta values of a query tile are per-COLUMN quantities here: staged in smem, [0,64): -lse * log2e,
    // [64,128): delta.  Tile i+1's values are fetched into a register while tile i is processed, so the global-load
    // latency is off the per-iteration critical path.
    auto fetch_stat = [&](int it) -> float {
      const int qq = (qt_begin + it) * BK_BR + (tid & 63);
      if (tid >= 128) return __int_as_float((seg_start && qq < T) ? seg_start[row_base + qq] : 0);   // [128,192): seg_start
      const size_t off = ((size_t)b * H + h) * T + (qq < T ? qq : 0);
      return qq < T ? ((tid < 64) ? -lse[off] * 1.4426950408889634f : delta[off]) : 0.f;
    };
    if (n_it > 0 && tid < 192) stat_ptr[tid] = fetch_stat(0);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int i = 0; i < n_it; ++i) {
      const int q0 = (qt_begin + i) * BK_BR;
      const float* st_lse = stat_ptr + (i & 1) * 192;
      float stat_next = 0.f;
      if (i + 1 < n_it && tid < 192) stat_next = fetch_stat(i + 1);
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld_32(tST + lane_off + half * 32, sv);
      tmem_ld_32(tdPT + lane_off + half * 32, dv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_empty);
      const bool need_mask = (CAUSAL && q0 < k0 + AT_BC) || (q0 + BK_BR > T) || (k0 + AT_BC > T) || seg_start != nullptr;
      // mask mode of this warp's 32 queries against its 32 key rows (warp-uniform): 0 all visible, 1 causal diagonal
      // group (key lane <= query index), 2 all masked (P^T = dS^T = 0, nothing to compute), 3 general (per element)
      int mode = 0;
      if (need_mask) {
        mode = 3;
        if (CAUSAL && seg_start == nullptr && q0 + BK_BR <= T && k0 + AT_BC <= T) {
          const int gq = ((q0 - k0) >> 5) + half;      // this warp's query group, in 32-row units from the key tile's first row
          mode = gq > q ? 0 : (gq == q ? 1 : 2);
        }
      }
      uint32_t pp[16], pd[16];
      // packed-pair math (two queries per instruction); dS'^T omits the 1/sqrt(d) factor, applied to dK in the epilogue
      // one straight-line instance of the 16-pair loop per mode (see the dQ kernel)
      auto body = [&](auto mode_c) {
        constexpr int MODE = decltype(mode_c)::value;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int qi = half * 32 + 2 * e;
          const f32x2 nl2 = *reinterpret_cast<const f32x2*>(st_lse + qi);
          const f32x2 dl = *reinterpret_cast<const f32x2*>(st_lse + 64 + qi);
          float x0, x1;
          upk2(fma2(pk2(__uint_as_float(sv[2 * e]), __uint_as_float(sv[2 * e + 1])), sl22, nl2), x0, x1);
          float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          if (MODE == 1) {
            if (lane > 2 * e) p0 = 0.f;
            if (lane > 2 * e + 1) p1 = 0.f;
          } else if (MODE == 3) {
            const int qrow = q0 + qi;
            const int2 sg = *reinterpret_cast<const int2*>(st_lse + 128 + qi);     // first visible key of the two queries
            if (key >= T || qrow >= T || (CAUSAL && key > qrow) || key < sg.x) p0 = 0.f;
            if (key >= T || qrow + 1 >= T || (CAUSAL && key > qrow + 1) || key < sg.y) p1 = 0.f;
          }
          pp[e] = pack_bf16(p0, p1);
          float d0, d1;
          upk2(mul2(pk2(p0, p1), sub2(pk2(__uint_as_float(dv[2 * e]), __uint_as_float(dv[2 * e + 1])), dl)), d0, d1);
          pd[e] = pack_bf16(d0, d1);
        }
      };
      if (mode == 0) body(std::integral_constant<int, 0>{});
      else if (mode == 1) body(std::integral_constant<int, 1>{});
      else if (mode == 3) body(std::integral_constant<int, 3>{});
      else {
#pragma unroll
        for (int e = 0; e < 16; ++e) pp[e] = pd[e] = 0u;
      }
      mbar_wait(pds_empty, (i & 1) ^ 1u);        // previous dV / dK MMAs finished reading P^T / dS^T (overlapped by the math)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int chunk = half * 4 + u;
        const uint32_t off = r * 128 + ((chunk ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPT + off), "r"(pp[4 * u]), "r"(pp[4 * u + 1]),
                     "r"(pp[4 * u + 2]), "r"(pp[4 * u + 3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdST + off), "r"(pd[4 * u]), "r"(pd[4 * u + 1]),
                     "r"(pd[4 * u + 2]), "r"(pd[4 * u + 3]) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (i + 1 < n_it) {
        if (tid < 192) stat_ptr[((i + 1) & 1) * 192 + tid] = stat_next;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    // epilogue: fp32 partial dK | dV rows of this (batch, head, key)
    if (n_it > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
    }
    // per-(query head) partials in bf16: the reference's autograd also rounds every expanded head's dK / dV to bf16
    // before summing them over the GQA group (repeat_kv backward); half the bytes of fp32 partials
    bf16* pp = partial + (((size_t)b * H + h) * T + key) * 128;
#pragma unroll 1
    for (int c = half * 2; c < half * 2 + 2; ++c) {   // columns 0..63 = dK (TMEM 128..191), 64..127 = dV (192..255)
      uint32_t v[32];
      if (n_it > 0) {
        tmem_ld_32(tdK + lane_off + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0u;
      }
      if (key < T) {
        const float cs = c < 2 ? scale : 1.0f;     // dK columns carry the deferred 1/sqrt(d)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          stg128(pp + c * 32 + u * 8,
                 make_uint4(pack_bf16(__uint_as_float(v[8 * u]) * cs, __uint_as_float(v[8 * u + 1]) * cs),
                            pack_bf16(__uint_as_float(v[8 * u + 2]) * cs, __uint_as_float(v[8 * u + 3]) * cs),
                            pack_bf16(__uint_as_float(v[8 * u + 4]) * cs, __uint_as_float(v[8 * u + 5]) * cs),
                            pack_bf16(__uint_as_float(v[8 * u + 6]) * cs, __uint_as_float(v[8 * u + 7]) * cs)));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// dk / dv (bf16, column slices of the fused gradient buffer) = sum over the GQA group's query heads, fixed order; with
// rope tables the inverse rotary embedding is applied to dk on the way out.  16 threads per (b, g, t): threads 0..3 own
// dk dims [8j, 8j+8) AND [32+8j, 32+8j+8) (a rotation pair), threads 8..15 own 8 dv dims each (4..7 idle).
__global__ void attn_tc_group_reduce_kernel(const bf16* __restrict__ partial, bf16* __restrict__ dk, bf16* __restrict__ dv,
                                            int B, int T, int H, int KVH, int ldg, const bf16* __restrict__ rope_cos,
                                            const bf16* __restrict__ rope_sin, const int* __restrict__ pos_ids, int max_pos) {
  griddep_launch();
  griddep_wait();
  const int group = H / KVH;
  const long total = (long)B * KVH * T * 16;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i & 15);
    if (c8 >= 4 && c8 < 8) continue;
    const long rt = i >> 4;
    const int t = (int)(rt % T);
    const int gg = (int)((rt / T) % KVH);
    const int b = (int)(rt / ((long)T * KVH));
    auto sum8 = [&](int col, uint32_t (&out)[4]) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = 0; j < group; ++j) {
        const uint4 a = ldg128_stream(partial + (((size_t)b * H + gg * group + j) * T + t) * 128 + col);
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16(w[k]);
          acc[2 * k] += f.x;
          acc[2 * k + 1] += f.y;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k] = pack_bf16(acc[2 * k], acc[2 * k + 1]);
    };
    if (c8 < 4) {
      uint32_t x1[4], x2[4];
      sum8(c8 * 8, x1);
      sum8(32 + c8 * 8, x2);
      bf16* dst = dk + ((size_t)b * T + t) * ldg + gg * 64 + c8 * 8;
      uint4 o1 = make_uint4(x1[0], x1[1], x1[2], x1[3]), o2 = make_uint4(x2[0], x2[1], x2[2], x2[3]);
      if (rope_cos != nullptr) {
        int pos = pos_ids ? pos_ids[(size_t)b * T + t] : t;
        pos = max(0, min(pos, max_pos - 1));
        rope_inv8(x1, x2, ldg128(rope_cos + (size_t)pos * 32 + c8 * 8), ldg128(rope_sin + (size_t)pos * 32 + c8 * 8), o1, o2);
      }
      stg128(dst, o1);
      stg128(dst + 32, o2);
    } else {
      uint32_t x[4];
      sum8(64 + (c8 - 8) * 8, x);
      stg128(dv + ((size_t)b * T + t) * ldg + gg * 64 + (c8 - 8) * 8, make_uint4(x[0], x[1], x[2], x[3]));
    }
  }
}

}  // namespace

// Backward launcher.  qkv / dqkv share the fused [B*T, ld] layout; delta fp32 [B,H,T] and partial fp32 [B,H,T,128] are
// caller-provided scratch (delta is filled here).
int sk_attn_tc_bwd_launch(const bf16* qkv, const bf16* o, const bf16* d_o, const float* lse, float* delta, float* partial_f,
                          bf16* dqkv, int B, int T, int H, int KVH, int ld, int ldo, int ldg, int causal, float scale,
                          cudaStream_t s, const int* seg_start, const int* seg_end, const bf16* rope_cos,
                          const bf16* rope_sin, const int* pos_ids, int max_pos) {
  SK_REQUIRE((seg_start == nullptr) == (seg_end == nullptr), "attn_tc_bwd: seg_start and seg_end go together");
  SK_REQUIRE(seg_start == nullptr || causal, "attn_tc_bwd: document segments need the causal kernels");
  SK_REQUIRE(H % KVH == 0, "attention: H must be a multiple of KVH");
  SK_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr) && (rope_cos == nullptr || max_pos > 0), "attn_tc_bwd: rope tables go together");
  CUtensorMap tm128, tm64, tmdo128, tmdo64, tmo128;
  int rc;
  if ((rc = sk_make_tmap_2d(&tm128, qkv, 2, (uint64_t)(H + 2 * KVH) * 64, (uint64_t)B * T, (uint64_t)ld, 64, 128))) return rc;
  if ((rc = sk_make_tmap_2d(&tm64, qkv, 2, (uint64_t)(H + 2 * KVH) * 64, (uint64_t)B * T, (uint64_t)ld, 64, 64))) return rc;
  if ((rc = sk_make_tmap_2d(&tmdo128, d_o, 2, (uint64_t)H * 64, (uint64_t)B * T, (uint64_t)ldo, 64, 128))) return rc;
  if ((rc = sk_make_tmap_2d(&tmdo64, d_o, 2, (uint64_t)H * 64, (uint64_t)B * T, (uint64_t)ldo, 64, 64))) return rc;
  if ((rc = sk_make_tmap_2d(&tmo128, o, 2, (uint64_t)H * 64, (uint64_t)B * T, (uint64_t)ldo, 64, 128))) return rc;
  static bool init = false;
  if (!init) {
    SK_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    SK_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    SK_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_dkdv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKDV_SMEM));
    SK_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_dkdv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKDV_SMEM));
    init = true;
  }
  bf16* partial = reinterpret_cast<bf16*>(partial_f);   // bf16 [B,H,T,128]: uses the first half of the caller's fp32-sized scratch
  // dQ first (it also writes delta for the dK/dV kernel), then dK/dV, then the GQA group reduction (+ inverse RoPE on dk)
  sk_prof_begin(1, s);
  dim3 g1(H, B, (T + AT_BC - 1) / AT_BC), g2(H, B, (T + AT_BR - 1) / AT_BR);
  bf16* dq = dqkv;
  bf16* dk = dqkv + H * 64;
  bf16* dv = dqkv + (H + KVH) * 64;
  if (causal) {
    SK_CUDA_CHECK(sk_launch_pdl(attn_tc_bwd_dq_kernel<true>, dim3(g2), dim3(BWD_THREADS), (size_t)(DQ_SMEM), s, tm128, tm64, tmdo128, tmo128, lse, delta, dq, T, ldg, H, KVH, scale, seg_start, rope_cos, rope_sin, pos_ids, max_pos));
    SK_CUDA_CHECK(sk_launch_pdl(attn_tc_bwd_dkdv_kernel<true>, dim3(g1), dim3(BWD_THREADS), (size_t)(DKDV_SMEM), s, tm128, tm64, tmdo64, lse, (const float*)delta, partial, T, H, KVH, scale, seg_start, seg_end));
  } else {
    SK_CUDA_CHECK(sk_launch_pdl(attn_tc_bwd_dq_kernel<false>, dim3(g2), dim3(BWD_THREADS), (size_t)(DQ_SMEM), s, tm128, tm64, tmdo128, tmo128, lse, delta, dq, T, ldg, H, KVH, scale, seg_start, rope_cos, rope_sin, pos_ids, max_pos));
    SK_CUDA_CHECK(sk_launch_pdl(attn_tc_bwd_dkdv_kernel<false>, dim3(g1), dim3(BWD_THREADS), (size_t)(DKDV_SMEM), s, tm128, tm64, tmdo64, lse, (const float*)delta, partial, T, H, KVH, scale, seg_start, seg_end));
  }
  sk_count_launch();
  sk_count_launch();                               // two kernels above; the reduce below is counted by SK_LAUNCH_CHECK
  const long total = (long)B * KVH * T * 16;
  int blocks = (int)((total + 255) / 256);
  if (blocks > sk_num_sms() * 8) blocks = sk_num_sms() * 8;
  SK_CUDA_CHECK(sk_launch_pdl(attn_tc_group_reduce_kernel, dim3(blocks), dim3(256), (size_t)(0), s, (const bf16*)partial, dk, dv, B, T, H, KVH, ldg, rope_cos, rope_sin, pos_ids, max_pos));
  sk_prof_end(s);
  SK_LAUNCH_CHECK();
  return 0;
}

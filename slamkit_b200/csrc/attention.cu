// Flash-style attention for the causal-LM path (GQA, head_dim 64; HF:models/qwen2/modeling_qwen2.py:187-246) and the
// bidirectional HuBERT encoder (HF:models/hubert/modeling_hubert.py:262-345), forward and backward.
//
// Round-1 implementation: tiled online-softmax kernels on the warp-level tensor-core path (mma.sync m16n8k16 bf16,
// ldmatrix from XOR-swizzled shared memory, cp.async double buffering).  S/P never touch HBM; the backward is split
// in two deterministic kernels (dK/dV per key tile looping over the GQA group, dQ per query tile) so no atomics are
// needed.  The tcgen05/TMEM version of these kernels is the round-2 item (attention is 5.8 % of the step FLOPs).
//
// Layout: q/k/v are column slices of the fused projection output [B*T, ld] (q: H*64 cols, k/v: KVH*64 cols);
// o is [B*T, H*64]; lse is [B, H, T] fp32 (natural log of the scaled-score softmax denominator); delta likewise.
#include "common.cuh"

namespace {

constexpr int HD = 64;  // head dim

SK_DEVINL uint32_t tile_addr(uint32_t base, int r, int chunk) { return base + r * 128 + ((chunk ^ (r & 7)) << 4); }

SK_DEVINL void cp_async16(uint32_t saddr, const void* g, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
SK_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SK_DEVINL void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

SK_DEVINL void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
SK_DEVINL void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
SK_DEVINL void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// rows x 64 bf16 tile, global (row pitch ld) -> swizzled smem; rows >= limit are zero-filled
template <int ROWS, int NT>
SK_DEVINL void load_tile(uint32_t sbase, const bf16* g, int ld, int row0, int limit) {
  for (int i = threadIdx.x; i < ROWS * 8; i += NT) {
    const int r = i >> 3, c = i & 7;
    const int gr = row0 + r;
    const bool ok = gr < limit;
    cp_async16(tile_addr(sbase, r, c), g + (size_t)(ok ? gr : 0) * ld + c * 8, ok);
  }
}

// A fragments (16 rows x 64 k) of a [rows][64] smem tile starting at row r0: a[ks][4]
SK_DEVINL void load_a_frags(uint32_t sbase, int r0, uint32_t (&a)[4][4]) {
  const int lane = threadIdx.x & 31;
  const int r = r0 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(a[ks][0], a[ks][1], a[ks][2], a[ks][3], tile_addr(sbase, r, ks * 2 + (lane >> 4)));
}

// acc[16 x 64] (8 n-tiles) += A(16 x 64 k-frags) * B^T where B tile is [64 n-rows][64 k] in smem (non-transposed
// ldmatrix): used for S = Q K^T, S^T = K Q^T, dP = dO V^T, dP^T = V dO^T.
SK_DEVINL void gemm_a_bT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sB) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      const int nrow = np * 16 + (lane & 7) + 8 * (lane >> 4);
      ldsm_x4(b0, b1, b2, b3, tile_addr(sB, nrow, ks * 2 + ((lane >> 3) & 1)));
      mma16816(acc[2 * np], a[ks], b0, b1);
      mma16816(acc[2 * np + 1], a[ks], b2, b3);
    }
  }
}

// acc[16 x 64 dims] += P(16 x 64 k, packed bf16 A frags p[kk][4]) * B where B tile is [64 k-rows][64 n] in smem
// (transposed ldmatrix): used for O = P V, dV = P^T dO, dK = dS^T Q, dQ = dS K.
SK_DEVINL void gemm_p_b(float (&acc)[8][4], const uint32_t (&p)[4][4], uint32_t sB) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b0, b1, b2, b3;
      const int krow = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      ldsm_x4_t(b0, b1, b2, b3, tile_addr(sB, krow, dp * 2 + (lane >> 4)));
      mma16816(acc[2 * dp], p[kk], b0, b1);
      mma16816(acc[2 * dp + 1], p[kk], b2, b3);
    }
  }
}

// accumulator tile (16 x 64, fp32) -> bf16 A fragments for the next matmul
SK_DEVINL void acc_to_a(const float (&s)[8][4], uint32_t (&p)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    p[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    p[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    p[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    p[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}

SK_DEVINL void zero_acc(float (&a)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// forward: BR = 128 query rows per CTA (8 warps x 16 rows), BC = 64 keys per step
// ------------------------------------------------------------------------------------------------
template <bool CAUSAL>
__global__ void __launch_bounds__(256, 2)
attn_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, bf16* __restrict__ o,
                float* __restrict__ lse, int T, int ld, int ldo, int H, int group, float scale) {
  extern __shared__ __align__(128) uint8_t smem_attn[];
  const uint32_t sQ = smem_u32(smem_attn);
  const uint32_t sK = sQ + 128 * 128;
  const uint32_t sV = sK + 2 * 64 * 128;
  const int b = blockIdx.z, h = blockIdx.y;
  const int qt = gridDim.x - 1 - blockIdx.x;  // heaviest (last) query tiles first
  const int q0 = qt * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* qp = q + (size_t)b * T * ld + h * HD;
  const bf16* kp = k + (size_t)b * T * ld + (h / group) * HD;
  const bf16* vp = v + (size_t)b * T * ld + (h / group) * HD;
  const float sl2 = scale * 1.4426950408889634f;

  int n_kv = (T + 63) / 64;
  if (CAUSAL) {
    const int last = min(T - 1, q0 + 127);
    n_kv = last / 64 + 1;
  }
  load_tile<128, 256>(sQ, qp, ld, q0, T);
  load_tile<64, 256>(sK, kp, ld, 0, T);
  load_tile<64, 256>(sV, vp, ld, 0, T);
  cp_async_commit();

  uint32_t qa[4][4];
  float oacc[8][4];
  zero_acc(oacc);
  float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
  const int row_a = q0 + warp * 16 + (lane >> 2);  // this thread's rows: row_a, row_a + 8

  for (int j = 0; j < n_kv; ++j) {
    const int st = j & 1;
    if (j + 1 < n_kv) {
      load_tile<64, 256>(sK + (st ^ 1) * 8192, kp, ld, (j + 1) * 64, T);
      load_tile<64, 256>(sV + (st ^ 1) * 8192, vp, ld, (j + 1) * 64, T);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) load_a_frags(sQ, warp * 16, qa);

    float s[8][4];
    zero_acc(s);
    gemm_a_bT(s, qa, sK + st * 8192);

    const int k0 = j * 64;
    const bool need_mask = (CAUSAL && (k0 + 63 > q0 + warp * 16)) || (k0 + 64 > T);
    if (need_mask) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = k0 + nt * 8 + (lane & 3) * 2 + (e & 1);
          const int row = row_a + ((e >> 1) << 3);
          if (key >= T || (CAUSAL && key > row)) s[nt][e] = -INFINITY;
        }
      }
    }
    // online softmax
    float mx[2] = {m_i[0], m_i[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = (m_i[r] == -INFINITY) ? 0.f : exp2f((m_i[r] - mx[r]) * sl2);
      m_i[r] = mx[r];
    }
    const float mb0 = (mx[0] == -INFINITY) ? 0.f : mx[0] * sl2;
    const float mb1 = (mx[1] == -INFINITY) ? 0.f : mx[1] * sl2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] * sl2 - mb0);
      s[nt][1] = exp2f(s[nt][1] * sl2 - mb0);
      s[nt][2] = exp2f(s[nt][2] * sl2 - mb1);
      s[nt][3] = exp2f(s[nt][3] * sl2 - mb1);
      rs[0] += s[nt][0] + s[nt][1];
      rs[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_i[r] = l_i[r] * corr[r] + rs[r];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      oacc[nt][0] *= corr[0];
      oacc[nt][1] *= corr[0];
      oacc[nt][2] *= corr[1];
      oacc[nt][3] *= corr[1];
    }
    uint32_t pa[4][4];
    acc_to_a(s, pa);
    gemm_p_b(oacc, pa, sV + st * 8192);
    __syncthreads();
  }
  // finalize
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_a + 8 * r;
    if (row < T) {
      const float inv = l_i[r] > 0.f ? 1.0f / l_i[r] : 0.f;
      bf16* op = o + ((size_t)b * T + row) * ldo + h * HD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t pk = pack_bf16(oacc[nt][2 * r] * inv, oacc[nt][2 * r + 1] * inv);
        *reinterpret_cast<uint32_t*>(op + nt * 8 + (lane & 3) * 2) = pk;
      }
      if (lse && (lane & 3) == 0) lse[((size_t)b * H + h) * T + row] = m_i[r] * scale + logf(l_i[r]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward, split-bf16 precision (HuBERT path): q/k/v/o are (hi, lo) bf16 pairs representing fp32-grade values;
// S = Qh Kh^T + Qh Kl^T + Ql Kh^T and O = Ph Vh + Ph Vl + Pl Vh, all accumulated in fp32.  Bidirectional (the
// reference passes no mask to HuBERT, hubert_feature_extractor.py:42), forward only.
// ------------------------------------------------------------------------------------------------
SK_DEVINL void acc_to_a_split(const float (&s)[8][4], uint32_t (&ph)[4][4], uint32_t (&pl)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const float (&t)[4] = s[2 * kk + half];
      const float h0 = bf16_round(t[0]), h1 = bf16_round(t[1]), h2 = bf16_round(t[2]), h3 = bf16_round(t[3]);
      ph[kk][2 * half] = pack_bf16(h0, h1);
      ph[kk][2 * half + 1] = pack_bf16(h2, h3);
      pl[kk][2 * half] = pack_bf16(t[0] - h0, t[1] - h1);
      pl[kk][2 * half + 1] = pack_bf16(t[2] - h2, t[3] - h3);
    }
  }
}

__global__ void __launch_bounds__(256)
attn_fwd_split_kernel(const bf16* __restrict__ q_hi, const bf16* __restrict__ q_lo, const bf16* __restrict__ k_hi,
                      const bf16* __restrict__ k_lo, const bf16* __restrict__ v_hi, const bf16* __restrict__ v_lo,
                      bf16* __restrict__ o_hi, bf16* __restrict__ o_lo, int T, int ld, int ldo, float scale) {
  extern __shared__ __align__(128) uint8_t smem_attn[];
  const uint32_t sQh = smem_u32(smem_attn);
  const uint32_t sQl = sQh + 128 * 128;
  const uint32_t sKV = sQl + 128 * 128;  // per stage: Kh, Kl, Vh, Vl (4 x 8 KB)
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)b * T * ld + h * HD;
  const float sl2 = scale * 1.4426950408889634f;
  const int n_kv = (T + 63) / 64;

  auto load_kv = [&](int j, int st) {
    const uint32_t sb = sKV + st * 4 * 8192;
    load_tile<64, 256>(sb, k_hi + base, ld, j * 64, T);
    load_tile<64, 256>(sb + 8192, k_lo + base, ld, j * 64, T);
    load_tile<64, 256>(sb + 2 * 8192, v_hi + base, ld, j * 64, T);
    load_tile<64, 256>(sb + 3 * 8192, v_lo + base, ld, j * 64, T);
  };
  load_tile<128, 256>(sQh, q_hi + base, ld, q0, T);
  load_tile<128, 256>(sQl, q_lo + base, ld, q0, T);
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qh[4][4], ql[4][4];
  float oacc[8][4];
  zero_acc(oacc);
  float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
  const int row_a = q0 + warp * 16 + (lane >> 2);

  for (int j = 0; j < n_kv; ++j) {
    const int st = j & 1;
    if (j + 1 < n_kv) {
      load_kv(j + 1, st ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
      load_a_frags(sQh, warp * 16, qh);
      load_a_frags(sQl, warp * 16, ql);
    }
    const uint32_t sb = sKV + st * 4 * 8192;
    float s[8][4];
    zero_acc(s);
    gemm_a_bT(s, ql, sb);          // Ql Kh^T (small terms first)
    gemm_a_bT(s, qh, sb + 8192);   // Qh Kl^T
    gemm_a_bT(s, qh, sb);          // Qh Kh^T
    const int k0 = j * 64;
    if (k0 + 64 > T) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k0 + nt * 8 + (lane & 3) * 2 + (e & 1) >= T) s[nt][e] = -INFINITY;
    }
    float mx[2] = {m_i[0], m_i[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = (m_i[r] == -INFINITY) ? 0.f : exp2f((m_i[r] - mx[r]) * sl2);
      m_i[r] = mx[r];
    }
    const float mb0 = mx[0] * sl2, mb1 = mx[1] * sl2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] * sl2 - mb0);
      s[nt][1] = exp2f(s[nt][1] * sl2 - mb0);
      s[nt][2] = exp2f(s[nt][2] * sl2 - mb1);
      s[nt][3] = exp2f(s[nt][3] * sl2 - mb1);
      rs[0] += s[nt][0] + s[nt][1];
      rs[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_i[r] = l_i[r] * corr[r] + rs[r];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      oacc[nt][0] *= corr[0];
      oacc[nt][1] *= corr[0];
      oacc[nt][2] *= corr[1];
      oacc[nt][3] *= corr[1];
    }
    uint32_t ph[4][4], pl[4][4];
    acc_to_a_split(s, ph, pl);
    gemm_p_b(oacc, pl, sb + 2 * 8192);  // Pl Vh
    gemm_p_b(oacc, ph, sb + 3 * 8192);  // Ph Vl
    gemm_p_b(oacc, ph, sb + 2 * 8192);  // Ph Vh
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_a + 8 * r;
    if (row < T) {
      const float inv = 1.0f / l_i[r];
      const size_t off = ((size_t)b * T + row) * ldo + h * HD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float v0 = oacc[nt][2 * r] * inv, v1 = oacc[nt][2 * r + 1] * inv;
        const float h0 = bf16_round(v0), h1 = bf16_round(v1);
        *reinterpret_cast<uint32_t*>(o_hi + off + nt * 8 + (lane & 3) * 2) = pack_bf16(h0, h1);
        *reinterpret_cast<uint32_t*>(o_lo + off + nt * 8 + (lane & 3) * 2) = pack_bf16(v0 - h0, v1 - h1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward preprocess: delta[b,h,t] = sum_d dO*O
// ------------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta,
                                  int B, int T, int H, int ldo) {
  griddep_launch();
  griddep_wait();
  const long total = (long)B * T * H * 8;  // 8 threads per (row, head)
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const bool ok = i < total;
  float acc = 0.f;
  long rh = 0;
  if (ok) {
    const int part = (int)(i & 7);
    rh = i >> 3;
    const int h = (int)(rh % H);
    const long m = rh / H;
    const uint4 a = ldg128_stream(o + m * ldo + h * HD + part * 8);
    const uint4 g = ldg128_stream(d_o + m * ldo + h * HD + part * 8);
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      const float2 x = unpack_bf16(au[kq]), y = unpack_bf16(gu[kq]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (ok && (i & 7) == 0) {
    const int h = (int)(rh % H);
    const long m = rh / H;
    const long bb = m / T, t = m % T;
    delta[(bb * H + h) * T + t] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// backward dK/dV: one CTA per (64-key tile, kv head, batch); 4 warps x 16 keys; loops over the GQA group's query
// heads and their query tiles.  Works on transposed score tiles S^T = K Q^T so P^T/dS^T are directly A operands.
// ------------------------------------------------------------------------------------------------
template <bool CAUSAL>
__global__ void __launch_bounds__(128, 3)
attn_bwd_dkdv_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                     const bf16* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta,
                     bf16* __restrict__ dk, bf16* __restrict__ dv, int T, int ld, int ldo, int ldg, int H, int group,
                     float scale) {
  extern __shared__ __align__(128) uint8_t smem_attn[];
  const uint32_t sKV = smem_u32(smem_attn);            // K tile then V tile (each 8 KB), only used for the prologue
  const uint32_t sQ = sKV + 2 * 8192;                  // 2 stages
  const uint32_t sdO = sQ + 2 * 8192;                  // 2 stages
  float* sStat = reinterpret_cast<float*>(smem_attn + 6 * 8192);  // [2 stages][2 (lse, delta)][64]
  const int b = blockIdx.z, g = blockIdx.y;
  const int kt = blockIdx.x;  // tile 0 has the most work under the causal mask and is scheduled first
  const int k0 = kt * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float sl2 = scale * 1.4426950408889634f;
  const bf16* kp = k + (size_t)b * T * ld + g * HD;
  const bf16* vp = v + (size_t)b * T * ld + g * HD;

  load_tile<64, 128>(sKV, kp, ld, k0, T);
  load_tile<64, 128>(sKV + 8192, vp, ld, k0, T);
  cp_async_commit();

  const int n_qt = (T + 63) / 64;
  const int qt_begin = CAUSAL ? kt : 0;
  const int per_head = n_qt - qt_begin;
  const int n_iter = per_head * group;

  auto issue = [&](int it, int st) {
    const int h = g * group + it / per_head;
    const int qt = qt_begin + it % per_head;
    const bf16* qp = q + (size_t)b * T * ld + h * HD;
    const bf16* dop = d_o + (size_t)b * T * ldo + h * HD;
    load_tile<64, 128>(sQ + st * 8192, qp, ld, qt * 64, T);
    load_tile<64, 128>(sdO + st * 8192, dop, ldo, qt * 64, T);
    if (threadIdx.x < 64) {
      const int row = qt * 64 + threadIdx.x;
      const size_t off = ((size_t)b * H + h) * T + (row < T ? row : 0);
      sStat[(st * 2 + 0) * 64 + threadIdx.x] = row < T ? lse[off] : 0.f;
      sStat[(st * 2 + 1) * 64 + threadIdx.x] = row < T ? delta[off] : 0.f;
    }
    cp_async_commit();
  };
  if (n_iter > 0) issue(0, 0);

  float dkacc[8][4], dvacc[8][4];
  zero_acc(dkacc);
  zero_acc(dvacc);
  const int key_a = k0 + warp * 16 + (lane >> 2);  // this thread's keys: key_a, key_a + 8

  for (int it = 0; it < n_iter; ++it) {
    const int st = it & 1;
    if (it + 1 < n_iter) {
      issue(it + 1, st ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int qt = qt_begin + it % per_head;
    const int q0 = qt * 64;
    const float* s_lse = sStat + (st * 2 + 0) * 64;
    const float* s_del = sStat + (st * 2 + 1) * 64;

    // K / V A-fragments are re-read from their resident smem tiles each iteration (8 ldmatrix) instead of being
    // pinned in 32 registers: that keeps the kernel under 170 registers -> 3 CTAs per SM
    uint32_t fa[4][4];
    load_a_frags(sKV, warp * 16, fa);
    float st_acc[8][4];  // S^T tile: rows = keys (this warp's 16), cols = 64 query rows
    zero_acc(st_acc);
    gemm_a_bT(st_acc, fa, sQ + st * 8192);
    const bool need_mask = (CAUSAL && (q0 < k0 + 64)) || (q0 + 64 > T) || (k0 + 64 > T);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = nt * 8 + (lane & 3) * 2 + (e & 1);
        const int key = key_a + ((e >> 1) << 3);
        float pv = exp2f(st_acc[nt][e] * sl2 - s_lse[qi] * 1.4426950408889634f);
        if (need_mask) {
          const int qrow = q0 + qi;
          if (qrow >= T || key >= T || (CAUSAL && key > qrow)) pv = 0.f;
        }
        st_acc[nt][e] = pv;
      }
    }
    uint32_t pa[4][4];
    acc_to_a(st_acc, pa);
    gemm_p_b(dvacc, pa, sdO + st * 8192);  // dV += P^T dO

    float dp[8][4];
    zero_acc(dp);
    load_a_frags(sKV + 8192, warp * 16, fa);
    gemm_a_bT(dp, fa, sdO + st * 8192);  // dP^T = V dO^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = nt * 8 + (lane & 3) * 2 + (e & 1);
        dp[nt][e] = st_acc[nt][e] * (dp[nt][e] - s_del[qi]) * scale;  // dS^T
      }
    }
    acc_to_a(dp, pa);
    gemm_p_b(dkacc, pa, sQ + st * 8192);  // dK += dS^T Q
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = key_a + 8 * r;
    if (key < T) {
      bf16* dkp = dk + ((size_t)b * T + key) * ldg + g * HD;
      bf16* dvp = dv + ((size_t)b * T + key) * ldg + g * HD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(dkp + nt * 8 + (lane & 3) * 2) = pack_bf16(dkacc[nt][2 * r], dkacc[nt][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(dvp + nt * 8 + (lane & 3) * 2) = pack_bf16(dvacc[nt][2 * r], dvacc[nt][2 * r + 1]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward dQ: one CTA per (64-row query tile, head, batch); loops over key tiles up to the diagonal
// ------------------------------------------------------------------------------------------------
template <bool CAUSAL>
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                   const bf16* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta,
                   bf16* __restrict__ dq, int T, int ld, int ldo, int ldg, int H, int group, float scale) {
  extern __shared__ __align__(128) uint8_t smem_attn[];
  const uint32_t sQdO = smem_u32(smem_attn);  // Q tile, dO tile (prologue only)
  const uint32_t sK = sQdO + 2 * 8192;        // 2 stages
  const uint32_t sV = sK + 2 * 8192;          // 2 stages
  const int b = blockIdx.z, h = blockIdx.y;
  const int qt = gridDim.x - 1 - blockIdx.x;
  const int q0 = qt * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float sl2 = scale * 1.4426950408889634f;
  const bf16* qp = q + (size_t)b * T * ld + h * HD;
  const bf16* dop = d_o + (size_t)b * T * ldo + h * HD;
  const bf16* kp = k + (size_t)b * T * ld + (h / group) * HD;
  const bf16* vp = v + (size_t)b * T * ld + (h / group) * HD;

  int n_kv = (T + 63) / 64;
  if (CAUSAL) n_kv = min(T - 1, q0 + 63) / 64 + 1;
  load_tile<64, 128>(sQdO, qp, ld, q0, T);
  load_tile<64, 128>(sQdO + 8192, dop, ldo, q0, T);
  load_tile<64, 128>(sK, kp, ld, 0, T);
  load_tile<64, 128>(sV, vp, ld, 0, T);
  cp_async_commit();

  const int row_a = q0 + warp * 16 + (lane >> 2);
  float lse_r[2], del_r[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_a + 8 * r;
    const size_t off = ((size_t)b * H + h) * T + (row < T ? row : 0);
    lse_r[r] = row < T ? lse[off] * 1.4426950408889634f : 0.f;
    del_r[r] = row < T ? delta[off] : 0.f;
  }
  uint32_t qa[4][4], doa[4][4];
  float dqacc[8][4];
  zero_acc(dqacc);

  for (int j = 0; j < n_kv; ++j) {
    const int st = j & 1;
    if (j + 1 < n_kv) {
      load_tile<64, 128>(sK + (st ^ 1) * 8192, kp, ld, (j + 1) * 64, T);
      load_tile<64, 128>(sV + (st ^ 1) * 8192, vp, ld, (j + 1) * 64, T);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
      load_a_frags(sQdO, warp * 16, qa);
      load_a_frags(sQdO + 8192, warp * 16, doa);
    }
    const int k0 = j * 64;
    float s[8][4];
    zero_acc(s);
    gemm_a_bT(s, qa, sK + st * 8192);
    const bool need_mask = (CAUSAL && (k0 + 63 > q0 + warp * 16)) || (k0 + 64 > T) || (q0 + 64 > T);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float pv = exp2f(s[nt][e] * sl2 - lse_r[e >> 1]);
        if (need_mask) {
          const int key = k0 + nt * 8 + (lane & 3) * 2 + (e & 1);
          const int row = row_a + ((e >> 1) << 3);
          if (row >= T || key >= T || (CAUSAL && key > row)) pv = 0.f;
        }
        s[nt][e] = pv;
      }
    }
    float dp[8][4];
    zero_acc(dp);
    gemm_a_bT(dp, doa, sV + st * 8192);  // dP = dO V^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) dp[nt][e] = s[nt][e] * (dp[nt][e] - del_r[e >> 1]) * scale;
    }
    uint32_t pa[4][4];
    acc_to_a(dp, pa);
    gemm_p_b(dqacc, pa, sK + st * 8192);  // dQ += dS K
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_a + 8 * r;
    if (row < T) {
      bf16* dqp = dq + ((size_t)b * T + row) * ldg + h * HD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<uint32_t*>(dqp + nt * 8 + (lane & 3) * 2) = pack_bf16(dqacc[nt][2 * r], dqacc[nt][2 * r + 1]);
    }
  }
}

constexpr int FWD_SMEM = 128 * 128 + 4 * 8192;        // 48 KB
constexpr int DKDV_SMEM = 6 * 8192 + 2 * 2 * 64 * 4;  // 49 KB + stats
constexpr int DQ_SMEM = 6 * 8192;
constexpr int FWD_SPLIT_SMEM = 2 * 128 * 128 + 8 * 8192;  // 96 KB

template <typename K>
int set_smem(K kernel, int bytes) {
  SK_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

}  // namespace

int sk_attn_fwd_launch(const bf16* q, const bf16* k, const bf16* v, bf16* o, float* lse, int B, int T, int H, int KVH,
                       int ld, int ldo, int causal, float scale, cudaStream_t s) {
  SK_REQUIRE(H % KVH == 0, "attention: H must be a multiple of KVH");
  SK_REQUIRE(ld % 8 == 0 && ldo % 8 == 0, "attention: leading dims must be multiples of 8");
  static bool init = false;
  if (!init) {
    if (set_smem(attn_fwd_kernel<true>, FWD_SMEM)) return -2;
    if (set_smem(attn_fwd_kernel<false>, FWD_SMEM)) return -2;
    init = true;
  }
  dim3 grid((T + 127) / 128, H, B);
  sk_prof_begin(1, s);
  if (causal) attn_fwd_kernel<true><<<grid, 256, FWD_SMEM, s>>>(q, k, v, o, lse, T, ld, ldo, H, H / KVH, scale);
  else attn_fwd_kernel<false><<<grid, 256, FWD_SMEM, s>>>(q, k, v, o, lse, T, ld, ldo, H, H / KVH, scale);
  sk_prof_end(s);
  SK_LAUNCH_CHECK();
  return 0;
}

// dq/dk/dv are column slices of one gradient buffer with row pitch ldg (same layout as the fused qkv activation)
int sk_attn_bwd_launch(const bf16* q, const bf16* k, const bf16* v, const bf16* o, const bf16* d_o, const float* lse,
                       float* delta, bf16* dq, bf16* dk, bf16* dv, int B, int T, int H, int KVH, int ld, int ldo,
                       int ldg, int causal, float scale, cudaStream_t s) {
  SK_REQUIRE(H % KVH == 0, "attention: H must be a multiple of KVH");
  static bool init = false;
  if (!init) {
    if (set_smem(attn_bwd_dkdv_kernel<true>, DKDV_SMEM)) return -2;
    if (set_smem(attn_bwd_dkdv_kernel<false>, DKDV_SMEM)) return -2;
    if (set_smem(attn_bwd_dq_kernel<true>, DQ_SMEM)) return -2;
    if (set_smem(attn_bwd_dq_kernel<false>, DQ_SMEM)) return -2;
    init = true;
  }
  const long total = (long)B * T * H * 8;
  sk_prof_begin(1, s);
  SK_CUDA_CHECK(sk_launch_pdl(attn_delta_kernel, dim3((int)((total + 255) / 256)), dim3(256), (size_t)(0), s, o, d_o, delta, B, T, H, ldo));
  SK_LAUNCH_CHECK();
  const int group = H / KVH;
  dim3 g1((T + 63) / 64, KVH, B), g2((T + 63) / 64, H, B);
  if (causal) {
    attn_bwd_dkdv_kernel<true><<<g1, 128, DKDV_SMEM, s>>>(q, k, v, d_o, lse, delta, dk, dv, T, ld, ldo, ldg, H, group, scale);
    SK_LAUNCH_CHECK();
    attn_bwd_dq_kernel<true><<<g2, 128, DQ_SMEM, s>>>(q, k, v, d_o, lse, delta, dq, T, ld, ldo, ldg, H, group, scale);
  } else {
    attn_bwd_dkdv_kernel<false><<<g1, 128, DKDV_SMEM, s>>>(q, k, v, d_o, lse, delta, dk, dv, T, ld, ldo, ldg, H, group, scale);
    SK_LAUNCH_CHECK();
    attn_bwd_dq_kernel<false><<<g2, 128, DQ_SMEM, s>>>(q, k, v, d_o, lse, delta, dq, T, ld, ldo, ldg, H, group, scale);
  }
  sk_prof_end(s);
  SK_LAUNCH_CHECK();
  return 0;
}

// split-bf16 (hi, lo) bidirectional forward for the HuBERT encoder; all six inputs share the row pitch ld
int sk_attn_fwd_split_launch(const bf16* q_hi, const bf16* q_lo, const bf16* k_hi, const bf16* k_lo, const bf16* v_hi,
                             const bf16* v_lo, bf16* o_hi, bf16* o_lo, int B, int T, int H, int ld, int ldo, float scale,
                             cudaStream_t s) {
  SK_REQUIRE(ld % 8 == 0 && ldo % 8 == 0, "attention: leading dims must be multiples of 8");
  static bool init = false;
  if (!init) {
    if (set_smem(attn_fwd_split_kernel, FWD_SPLIT_SMEM)) return -2;
    init = true;
  }
  dim3 grid((T + 127) / 128, H, B);
  sk_prof_begin(1, s);
  attn_fwd_split_kernel<<<grid, 256, FWD_SPLIT_SMEM, s>>>(q_hi, q_lo, k_hi, k_lo, v_hi, v_lo, o_hi, o_lo, T, ld, ldo, scale);
  sk_prof_end(s);
  SK_LAUNCH_CHECK();
  return 0;
}

int sk_attn_delta_launch(const bf16* o, const bf16* d_o, float* delta, int B, int T, int H, int ldo, cudaStream_t s) {
  const long total = (long)B * T * H * 8;
  sk_prof_begin(1, s);
  SK_CUDA_CHECK(sk_launch_pdl(attn_delta_kernel, dim3((int)((total + 255) / 256)), dim3(256), (size_t)(0), s, o, d_o, delta, B, T, H, ldo));
  sk_prof_end(s);
  SK_LAUNCH_CHECK();
  return 0;
}

// tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T (+bias[N]) (+residual[M,N]),  bf16 in, fp32 accumulate
// in TMEM, bf16 (or fp32) out.
//
// One persistent CTA per SM, 6 warps, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction;
//                               tcgen05.commit frees smem slots and publishes accumulators)
//   warps 2..5  epilogue       (tcgen05.ld 32x32b.x32 from TMEM -> bias/residual -> 64 B per thread row stores)
// Accumulators are double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand majors.  "K-major" = the contraction index is contiguous in memory (A row-major [M,K], B row-major [N,K]).
// "MN-major" = the M (or N) index is contiguous (A stored as [K,M], B stored as [K,N]).  MN-major operands let the
// backward GEMMs (dgrad: dX = dY * W ; wgrad: dW = dY^T * X) read activations and weights in place, with no
// transposed copies in HBM.  Layouts follow cute::UMMA canonical SW128 forms:
//    K-major : ((8,m),(T,2)) : ((8T,SBO),(1,T))          rows of 128 B, 8-row groups SBO=1024 B apart
//    MN-major: ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO))    64-element MN atoms, LBO apart; 8-k-row groups SBO apart
#include "common.cuh"
#include <cudaTypedefs.h>
#include <stdio.h>
#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

struct GemmParams {
  int M, N, K;
  void* C;
  int ldc;
  const bf16* bias;
  const bf16* residual;
  int ldr;
  int tiles_m, tiles_n;
  int out_f32;          // 1: C is float
  int round_before_res; // 1: out = bf16(bf16(acc+bias) + res)  (matches an unfused bf16 linear followed by an add)
  int act;              // 0 none, 1 GELU(erf) applied to acc+bias
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int ACC_STRIDE = (BN <= 32) ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 for manual alignment
};

SK_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / p.tiles_n) * BM;
        const int n0 = (t % p.tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          const uint32_t fb = full_bar(stage);
          mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sA, &tmA, fb, kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * (BK * 128), &tmA, fb, m0 + 64 * j, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sB, &tmB, fb, kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * (BK * 128), &tmB, fb, n0 + 64 * j, kb * BK);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(1u, A_MN ? 1u : 0u, B_MN ? 1u : 0u, BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = A_MN ? umma_desc_sw128(sA + k * (UMMA_K * 128), BK * 128, 1024)
                                        : umma_desc_sw128(sA + k * (UMMA_K * 2), 16, 1024);
            const uint64_t bdesc = B_MN ? umma_desc_sw128(sB + k * (UMMA_K * 128), BK * 128, 1024)
                                        : umma_desc_sw128(sB + k * (UMMA_K * 2), 16, 1024);
            tc_mma_f16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(as));
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===== epilogue (warps 2..5); TMEM lane quadrant = warp % 4 =====
    const int q = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / p.tiles_n) * BM;
      const int n0 = (t % p.tiles_n) * BN;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * Cfg::ACC_STRIDE);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row_ok && col0 < p.N) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            if (col < p.N) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
              if (p.bias) {
                const uint4 bv = ldg128(p.bias + col);
                const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = unpack_bf16(bw[i]);
                  v[2 * i] += f.x;
                  v[2 * i + 1] += f.y;
                }
              }
              if (p.act == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
              }
              if (p.residual) {
                const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.ldr + col);
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = unpack_bf16(rw[i]);
                  if (p.round_before_res) {
                    v[2 * i] = bf16_round(v[2 * i]) + f.x;
                    v[2 * i + 1] = bf16_round(v[2 * i + 1]) + f.y;
                  } else {
                    v[2 * i] += f.x;
                    v[2 * i + 1] += f.y;
                  }
                }
              }
              if (p.out_f32) {
                float* dst = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col;
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else {
                uint4 o;
                o.x = pack_bf16(v[0], v[1]);
                o.y = pack_bf16(v[2], v[3]);
                o.z = pack_bf16(v[4], v[5]);
                o.w = pack_bf16(v[6], v[7]);
                stg128(reinterpret_cast<bf16*>(p.C) + (size_t)row * p.ldc + col, o);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

}  // namespace

// Build a 2-D bf16 (elem_bytes=2) / fp32 (elem_bytes=4) tensor map: `inner` contiguous elements, `outer` rows of pitch
// ld elements, box (box_inner x box_outer), 128B swizzle.
int sk_make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld,
                    uint32_t box_inner, uint32_t box_outer) {
  std::call_once(g_encode_once, load_encode);
  SK_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  SK_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  SK_REQUIRE((ld * elem_bytes) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%llu)",
             (unsigned long long)ld);
  SK_REQUIRE(box_inner * elem_bytes == 128, "SW128 box inner extent must be 128 bytes");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = g_encode(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu)", (int)r,
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return 0;
}

namespace {

template <int BN, bool A_MN, bool B_MN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    SK_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    attr_set = true;
  }
  sk_prof_begin(0, stream);
  gemm_tcgen05_kernel<BN, A_MN, B_MN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  sk_prof_end(stream);
  SK_LAUNCH_CHECK();
  return 0;
}

template <int BN>
int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid,
                   cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(tmA, tmB, p, grid, s);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true>(tmA, tmB, p, grid, s);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false>(tmA, tmB, p, grid, s);
  return launch_gemm<BN, true, true>(tmA, tmB, p, grid, s);
}

// Relative cost of one 128 x BN tile (BN=256 == 100), measured on B200 (profiles/r01_gemm_bench.txt): narrow tiles
// pay the per-tile pipeline fill / epilogue overhead and re-read the A tile from shared memory more often per FLOP.
inline int tile_cost(int bn) { return bn >= 256 ? 100 : (bn >= 128 ? 61 : 54); }

}  // namespace

int sk_pick_bn(int M, int N, int force_bn) {
  if (force_bn == 64 || force_bn == 128 || force_bn == 256) return force_bn;
  const int nsm = sk_num_sms();
  const int cands[3] = {256, 128, 64};
  int best = 256;
  long best_cost = -1;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long tiles = (long)((M + BM - 1) / BM) * ((N + bn - 1) / bn);
    const long waves = (tiles + nsm - 1) / nsm;
    const long cost = waves * tile_cost(bn);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// The general entry point.  A: [M,K] (a_mn=0, lda = row pitch of the [M,K] array) or stored [K,M] (a_mn=1, lda = row
// pitch of the [K,M] array).  B: [N,K] (b_mn=0) or stored [K,N] (b_mn=1).
int sk_gemm_launch(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                   int ldc, int out_f32, const void* bias, const void* residual, int ldr, int round_before_res, int act,
                   int force_bn, cudaStream_t stream) {
  SK_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  SK_REQUIRE(N % 8 == 0, "gemm: N must be a multiple of 8 (N=%d)", N);
  SK_REQUIRE(ldc % 8 == 0 && (residual == nullptr || ldr % 8 == 0), "gemm: ldc/ldr must be multiples of 8");
  SK_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm: C must be 16-byte aligned");
  const int BN = sk_pick_bn(M, N, force_bn);
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = sk_make_tmap_2d(&tmA, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  else       rc = sk_make_tmap_2d(&tmA, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  if (rc) return rc;
  if (!b_mn) rc = sk_make_tmap_2d(&tmB, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)BN);
  else       rc = sk_make_tmap_2d(&tmB, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc;
  p.bias = reinterpret_cast<const bf16*>(bias);
  p.residual = reinterpret_cast<const bf16*>(residual);
  p.ldr = ldr;
  p.tiles_m = (M + BM - 1) / BM;
  p.tiles_n = (N + BN - 1) / BN;
  p.out_f32 = out_f32;
  p.round_before_res = round_before_res;
  p.act = act;
  const int tiles = p.tiles_m * p.tiles_n;
  const int nsm = sk_num_sms();
  const int grid = tiles < nsm ? tiles : nsm;
  switch (BN) {
    case 256: return dispatch_major<256>(a_mn, b_mn, tmA, tmB, p, grid, stream);
    case 128: return dispatch_major<128>(a_mn, b_mn, tmA, tmB, p, grid, stream);
    default:  return dispatch_major<64>(a_mn, b_mn, tmA, tmB, p, grid, stream);
  }
}

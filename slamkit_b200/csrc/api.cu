// extern "C" surface of libslamkit_b200.so for the op-level entry points (see include/slamkit_b200.h), plus the
// error / device-info plumbing.  The handle-level entry points live in lm_step.cu and hubert_step.cu.
#include <stdlib.h>
#include "kernels.h"
#include "../../include/slamkit_b200.h"
#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <vector>

namespace {
thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void sk_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool sk_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("SK_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}
void sk_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-category device timing (bench.py's roofline leg): CUDA events recorded around the launches of one
// category on the launching stream.  Off by default; never active inside bench.py's timed region.
namespace {
constexpr int PROF_CATS = 4;
constexpr int PROF_MAX = 8192;
bool g_prof_on = false;
std::vector<cudaEvent_t> g_prof_ev;
std::vector<int> g_prof_cat;
int g_prof_used = 0;
}  // namespace
void sk_prof_begin(int cat, cudaStream_t s) {
  if (!g_prof_on || g_prof_used + 2 > PROF_MAX) return;
  cudaEventRecord(g_prof_ev[g_prof_used], s);
  g_prof_cat[g_prof_used / 2] = cat;
  g_prof_used += 1;
}
void sk_prof_end(cudaStream_t s) {
  if (!g_prof_on || (g_prof_used & 1) == 0) return;
  cudaEventRecord(g_prof_ev[g_prof_used], s);
  g_prof_used += 1;
}

int sk_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

#define S(x) ((cudaStream_t)(x))
#define BF(x) (reinterpret_cast<bf16*>(x))
#define CBF(x) (reinterpret_cast<const bf16*>(x))

extern "C" {

const char* sk_last_error(void) { return g_err; }
int sk_version(void) { return 1; }
int sk_device_sm_count(void) { return sk_num_sms(); }
int sk_device_cc(void) {
  int dev = 0, ma = 0, mi = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -2;
  cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev);
  return 10 * ma + mi;
}
int64_t sk_launch_count(void) { return (int64_t)g_launches.load(); }

int sk_prof_enable(int on) {
  if (on && g_prof_ev.empty()) {
    g_prof_ev.resize(PROF_MAX);
    g_prof_cat.resize(PROF_MAX / 2);
    for (int i = 0; i < PROF_MAX; ++i) SK_CUDA_CHECK(cudaEventCreate(&g_prof_ev[i]));
  }
  g_prof_on = on != 0;
  g_prof_used = 0;
  return 0;
}
int sk_prof_collect(double* ms_by_cat, int64_t* count_by_cat) {
  SK_CUDA_CHECK(cudaDeviceSynchronize());
  for (int c = 0; c < PROF_CATS; ++c) { ms_by_cat[c] = 0.0; count_by_cat[c] = 0; }
  for (int i = 0; i + 1 < g_prof_used; i += 2) {
    float ms = 0.f;
    SK_CUDA_CHECK(cudaEventElapsedTime(&ms, g_prof_ev[i], g_prof_ev[i + 1]));
    const int c = g_prof_cat[i / 2];
    if (c >= 0 && c < PROF_CATS) { ms_by_cat[c] += ms; count_by_cat[c] += 1; }
  }
  g_prof_used = 0;
  return 0;
}

int sk_gemm_bf16(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                 int ldc, int out_f32, const void* bias, const void* residual, int ldr, int round_before_res, int act,
                 int force_bn, void* stream) {
  SK_REQUIRE(A && B && C, "sk_gemm_bf16: null operand");
  return sk_gemm_launch(M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, out_f32, bias, residual, ldr, round_before_res, act,
                        force_bn, S(stream));
}
int sk_linear_swiglu_fwd(int M, int F, int K, const void* x, const void* w_gu, void* gu, void* act, void* stream) {
  SK_REQUIRE(x && w_gu && gu && act, "sk_linear_swiglu_fwd: null operand");
  return sk_linear_swiglu_fwd_launch(M, F, K, x, w_gu, gu, act, S(stream));
}
int sk_linear_swiglu_bwd(int M, int N, int F, const void* dy, const void* w_down, const void* gu, void* dgu, void* stream) {
  SK_REQUIRE(dy && w_down && gu && dgu, "sk_linear_swiglu_bwd: null operand");
  return sk_linear_swiglu_bwd_launch(M, N, F, dy, w_down, gu, dgu, S(stream));
}
int sk_linear_rope(int M, int N, int K, const void* x, const void* w, const void* bias, void* out, const void* cos_t,
                   const void* sin_t, const int32_t* pos_ids, int T, int rope_cols, int max_positions, void* stream) {
  SK_REQUIRE(x && w && out && cos_t && sin_t, "sk_linear_rope: null operand");
  return sk_linear_rope_launch(M, N, K, x, w, bias, out, cos_t, sin_t, pos_ids, T, rope_cols, max_positions, S(stream));
}
int sk_gemm_bf16_splitk(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                        int ldc, int accumulate, void* splitk_ws, int64_t splitk_ws_bytes, void* stream) {
  SK_REQUIRE(A && B && C, "sk_gemm_bf16_splitk: null operand");
  return sk_gemm_launch(M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, 0, nullptr, accumulate ? C : nullptr, ldc, 1, 0, 0,
                        S(stream), splitk_ws, (size_t)splitk_ws_bytes);
}
int sk_gemm_bf16_ws(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                    int ldc, int out_f32, const void* bias, const void* residual, int ldr, int round_before_res, int act,
                    int force_bn, void* ws, int64_t ws_bytes, void* stream) {
  SK_REQUIRE(A && B && C, "sk_gemm_bf16_ws: null operand");
  SK_REQUIRE(ws == nullptr || (((uintptr_t)ws & 15) == 0 && ws_bytes % 16 == 0), "sk_gemm_bf16_ws: scratch must be 16-byte aligned");
  return sk_gemm_launch(M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, out_f32, bias, residual, ldr, round_before_res, act,
                        force_bn, S(stream), ws, (size_t)ws_bytes);
}
int64_t sk_gemm_ws_bytes(void) { return (int64_t)sk_gemm_ws_min_bytes(); }
int sk_embed_fwd(const int64_t* ids, const void* table, void* out, int M, int D, int V, void* stream) {
  return sk_embed_fwd_launch(ids, CBF(table), BF(out), M, D, V, S(stream));
}
int sk_embed_bwd(const int64_t* ids, const void* dx, float* scratch, void* dtable, int M, int D, int V, int Vpad,
                 int accumulate, void* stream) {
  return sk_embed_bwd_launch(ids, CBF(dx), scratch, BF(dtable), M, D, V, Vpad, accumulate, S(stream));
}
int sk_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int M, int D, float eps, void* stream) {
  return sk_rmsnorm_fwd_launch(CBF(x), CBF(w), BF(y), rstd, M, D, eps, S(stream));
}
int sk_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx,
                   void* dw, float* dw_partial, int M, int D, int accumulate_dw, void* stream) {
  return sk_rmsnorm_bwd_launch(CBF(dy), CBF(x), CBF(w), rstd, CBF(dres), BF(dx), BF(dw), dw_partial, M, D,
                               accumulate_dw, S(stream));
}
int sk_colsum(const void* x, void* out, float* partial, int M, int N, int ld, int accumulate, void* stream) {
  return sk_colsum_launch(CBF(x), BF(out), partial, M, N, ld, accumulate, S(stream));
}
int sk_rope(void* qkv, const void* cos_t, const void* sin_t, const int32_t* pos_ids, int M, int T, int ld,
            int n_rot_heads, int head_dim, int inverse, int max_positions, void* stream) {
  return sk_rope_launch(BF(qkv), CBF(cos_t), CBF(sin_t), pos_ids, M, T, ld, n_rot_heads, head_dim, inverse, max_positions,
                        S(stream));
}
int sk_swiglu_fwd(const void* gu, void* act, int M, int F, void* stream) {
  return sk_swiglu_fwd_launch(CBF(gu), BF(act), M, F, S(stream));
}
int sk_swiglu_bwd(const void* gu, const void* dact, void* dgu, int M, int F, void* stream) {
  return sk_swiglu_bwd_launch(CBF(gu), CBF(dact), BF(dgu), M, F, S(stream));
}
int sk_ce_fwd_bwd(const void* logits, const int64_t* labels, void* dlogits, float* partial, float* row_nll,
                  float* stats, int M, int T, int V, int ldl, float num_items, float dloss, void* stream) {
  return sk_ce_launch(CBF(logits), labels, BF(dlogits), partial, row_nll, stats, M, T, V, ldl, num_items, dloss,
                      S(stream));
}
int sk_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int T, int H, int KVH, int ld,
                int ldo, int causal, float scale, void* stream) {
  return sk_attn_fwd_launch(CBF(q), CBF(k), CBF(v), BF(o), lse, B, T, H, KVH, ld, ldo, causal, scale, S(stream));
}
int sk_attn_tc_fwd(const void* qkv, void* o, float* lse, int B, int T, int H, int KVH, int ld, int ldo, int causal,
                   float scale, const int32_t* seg_start, void* stream) {
  SK_REQUIRE(qkv && o, "sk_attn_tc_fwd: null argument");
  return sk_attn_tc_fwd_launch(CBF(qkv), BF(o), lse, B, T, H, KVH, ld, ldo, causal, scale, S(stream), seg_start);
}
int sk_seg_bounds(const int32_t* pos_ids, int32_t* seg_start, int32_t* seg_end, int B, int T, void* stream) {
  return sk_seg_bounds_launch(pos_ids, seg_start, seg_end, B, T, S(stream));
}
int sk_attn_tc_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, float* delta, float* partial,
                   void* dqkv, int B, int T, int H, int KVH, int ld, int ldo, int ldg, int causal, float scale,
                   const int32_t* seg_start, const int32_t* seg_end, void* stream) {
  SK_REQUIRE(qkv && o && d_o && lse && delta && partial && dqkv, "sk_attn_tc_bwd: null argument");
  return sk_attn_tc_bwd_launch(CBF(qkv), CBF(o), CBF(d_o), lse, delta, partial, BF(dqkv), B, T, H, KVH, ld, ldo, ldg, causal,
                               scale, S(stream), seg_start, seg_end);
}
int sk_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                float* delta, void* dq, void* dk, void* dv, int B, int T, int H, int KVH, int ld, int ldo, int ldg,
                int causal, float scale, void* stream) {
  return sk_attn_bwd_launch(CBF(q), CBF(k), CBF(v), CBF(o), CBF(d_o), lse, delta, BF(dq), BF(dk), BF(dv), B, T, H, KVH,
                            ld, ldo, ldg, causal, scale, S(stream));
}
int sk_grad_norm(const void* grads, const int64_t* chunk_start, const int32_t* chunk_len, int n_chunks,
                 const int32_t* tensor_chunk_begin, int n_tensors, float* partial, float max_norm, int emulate_bf16,
                 float* stats, void* stream) {
  return sk_gradnorm_launch(CBF(grads), reinterpret_cast<const long*>(chunk_start), chunk_len, n_chunks,
                            tensor_chunk_begin, n_tensors, partial, max_norm, emulate_bf16, stats, S(stream));
}
int sk_adamw_step(void* params, const void* grads, void* exp_avg, void* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, const float* clip_stats, void* stream) {
  return sk_adamw_launch(BF(params), CBF(grads), BF(exp_avg), BF(exp_avg_sq), (long)n, lr, beta1, beta2, eps,
                         weight_decay, step, clip_stats, S(stream));
}

}  // extern "C"

// ---- peer-memory gradient all-reduce (p2p_comm.cu) ------------------------------------------------------------------
int64_t sk_p2p_flag_bytes(void) { return (int64_t)sk_p2p_flag_bytes_impl(); }
int sk_p2p_set_trace(void* buf) { return sk_p2p_set_trace_impl(buf); }
int sk_p2p_debug_hog(int ctas, int64_t ns, void* started_u32, void* stream) {
  return sk_p2p_hog_launch(ctas, (long long)ns, reinterpret_cast<unsigned*>(started_u32), S(stream));
}
int sk_p2p_alloc(int64_t bytes, void** out) { return sk_p2p_alloc_impl((size_t)bytes, out); }
int sk_p2p_free(void* p) { return sk_p2p_free_impl(p); }
int sk_p2p_export(const void* ptr, void* handle64, int64_t* offset) {
  size_t off = 0;
  const int rc = sk_p2p_export_impl(ptr, handle64, &off);
  if (offset) *offset = (int64_t)off;
  return rc;
}
int sk_p2p_open(const void* handle64, void** base) { return sk_p2p_open_impl(handle64, base); }
int sk_p2p_close(void* base) { return sk_p2p_close_impl(base); }
int sk_p2p_signal(void* const* flags, int rank, int world, int slot, uint32_t epoch, void* stream) {
  return sk_p2p_signal_launch(flags, rank, world, slot, epoch, S(stream));
}
int sk_p2p_wait(void* const* flags, int rank, int world, int slot_lo, int n_slots, uint32_t epoch, int* err_flag, void* stream) {
  return sk_p2p_wait_launch(flags, rank, world, slot_lo, n_slots, epoch, err_flag, S(stream));
}
int sk_p2p_allreduce_bf16(void* const* bufs, void* const* flags, int rank, int world, int64_t offset_elems, int64_t n_elems,
                          int slot, uint32_t epoch, int ctas, int* err_flag, void* stream) {
  SK_REQUIRE(offset_elems >= 0 && n_elems > 0, "p2p_allreduce: bad range");
  return sk_p2p_allreduce_launch(bufs, flags, rank, world, (size_t)offset_elems, (size_t)n_elems, slot, epoch, ctas, err_flag,
                                 S(stream));
}

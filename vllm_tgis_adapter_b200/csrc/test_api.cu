// C-ABI test hooks (include/tgis_kernels.h): every hot-path kernel callable on raw device pointers so that the GPU
// parity tests exercise exactly the code the engine runs, through the same shared library.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tgis_kernels.h"
#include "kernels.h"

using namespace tgis;
using bf16 = __nv_bfloat16;

namespace {
thread_local std::string t_err;
}  // namespace
extern "C" const char* tgis_k_last_error(void) { return t_err.c_str(); }

namespace {
int kfail(const std::string& m) {
  t_err = m;
  return -1;
}
#define KCK(expr)                                                                          \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) return kfail(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

int test_num_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <class T>
struct Tmp {
  T* p = nullptr;
  ~Tmp() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
  cudaError_t upload(const T* h, size_t n) {
    cudaError_t e = alloc(n);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice);
  }
};
}  // namespace

extern "C" {

int tgis_k_gemm_timeline(uint64_t* out64) { return gemm_timeline_read((unsigned long long*)out64); }

// Step timeline (debug builds, -DTGIS_STEP_TIMELINE): buffer = [8 header words: [0] = records written][4096 x 4 words]
// [4096 x 4 extra stamps of the GEMM records: first MMA issued, last TMA issued, last accumulator ready, rstd ready]
static unsigned long long* g_step_tl_buf = nullptr;
int tgis_k_step_timeline_enable(void) {
  if (!g_step_tl_buf) {
    if (cudaMalloc(&g_step_tl_buf, sizeof(unsigned long long) * (8 + 4096 * 8)) != cudaSuccess) return kfail("cudaMalloc");
  }
  KCK(cudaMemset(g_step_tl_buf, 0, sizeof(unsigned long long) * (8 + 4096 * 8)));
  int rc = gemm_set_step_timeline(g_step_tl_buf);
  if (rc != 0) return rc;  // -2: not a timeline build
  elementwise_set_step_timeline(g_step_tl_buf);
  attention_set_step_timeline(g_step_tl_buf);
  sampler_set_step_timeline(g_step_tl_buf);
  return 0;
}
// out: [0] = number of records since the last read, then 4096 x {kernel id, t_entry, t_waited, t_exit} (ns); resets
int tgis_k_step_timeline_read(uint64_t* out) {
  if (!g_step_tl_buf) return kfail("step timeline not enabled");
  KCK(cudaDeviceSynchronize());
  KCK(cudaMemcpy(out, g_step_tl_buf, sizeof(unsigned long long) * (8 + 4096 * 8), cudaMemcpyDeviceToHost));
  KCK(cudaMemset(g_step_tl_buf, 0, sizeof(unsigned long long) * 8));
  return 0;
}
int tgis_k_sizeof_sample_row(void) { return (int)sizeof(SampleRow); }
int tgis_k_sizeof_sample_out(void) { return (int)sizeof(SampleOut); }
int tgis_k_kv_block(void) { return KV_BLOCK; }

int tgis_k_gemm(const void* x_dev, const void* w_dev, void* y_dev, int32_t T, int32_t N, int32_t K,
                int32_t x_rows_alloc, int32_t impl, int32_t iters, float* ms_out, int32_t out_f32) {
  if (iters < 1) iters = 1;
  cudaStream_t st = 0;
  cudaEvent_t e0, e1;
  KCK(cudaEventCreate(&e0));
  KCK(cudaEventCreate(&e1));
  if (impl == 1) {
    KCK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i)
      KCK(gemm_bf16_ref_launch((const bf16*)x_dev, K, (const bf16*)w_dev, y_dev, out_f32 == 2 ? N / 2 : N, T, N, K, st, out_f32));
    KCK(cudaEventRecord(e1, st));
  } else {
    int dev = 0, sms = 148;
    KCK(cudaGetDevice(&dev));
    KCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CUtensorMap wm, xm;
    const int bt = gemm_pick_bt(T);
    if (x_rows_alloc < bt) return kfail("x must have at least one TMA box of rows allocated");
    if (make_tmap_bf16_2d(&wm, w_dev, N, K, K, 128, 64) != 0) return kfail("weight tensor map failed");
    if (make_tmap_bf16_2d(&xm, x_dev, x_rows_alloc, K, K, bt, 64) != 0) return kfail("activation tensor map failed");
    Tmp<float> ws;
    Tmp<int> ctr;
    KCK(ws.alloc(gemm_workspace_bytes(sms) / sizeof(float)));
    KCK(ctr.alloc(1 << 16));
    KCK(cudaMemset(ctr.p, 0, sizeof(int) << 16));
    KCK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i)
      KCK(gemm_bf16_launch(wm, xm, y_dev, out_f32 == 2 ? N / 2 : N, T, N, K, ws.p, ctr.p, sms, st, out_f32));
    KCK(cudaEventRecord(e1, st));
    KCK(cudaStreamSynchronize(st));
  }
  KCK(cudaDeviceSynchronize());
  float ms = 0.f;
  KCK(cudaEventElapsedTime(&ms, e0, e1));
  if (ms_out) *ms_out = ms / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

int tgis_k_rmsnorm(const void* x_dev, void* residual_dev, const void* w_dev, void* out_dev, int32_t T, int32_t hidden,
                   float eps) {
  if (residual_dev)
    KCK(add_rmsnorm_launch((const bf16*)x_dev, (bf16*)residual_dev, (const bf16*)w_dev, (bf16*)out_dev, T, hidden, eps, 0));
  else
    KCK(rmsnorm_launch((const bf16*)x_dev, (const bf16*)w_dev, (bf16*)out_dev, T, hidden, eps, 0));
  KCK(cudaDeviceSynchronize());
  return 0;
}

int tgis_k_opt_layernorm(const void* acc_dev, const void* acc_bias_dev, void* residual_dev, const void* w_dev,
                         const void* b_dev, void* out_dev, int32_t T, int32_t hidden, float eps) {
  KCK(opt_layernorm_launch((const float*)acc_dev, (const bf16*)acc_bias_dev, (bf16*)residual_dev, (const bf16*)w_dev,
                           (const bf16*)b_dev, (bf16*)out_dev, T, hidden, eps, 0));
  KCK(cudaDeviceSynchronize());
  return 0;
}

int tgis_k_opt_bias_act(const void* acc_dev, const void* bias_dev, void* out_dev, int32_t T, int32_t N, int32_t relu) {
  KCK(opt_bias_act_launch((const float*)acc_dev, N, (const bf16*)bias_dev, (bf16*)out_dev, N, T, N, relu, 148, 0));
  KCK(cudaDeviceSynchronize());
  return 0;
}

int tgis_k_opt_embed(const void* tok_dev, const void* pos_dev, const void* tok_table_dev, const void* pos_table_dev,
                     void* out_dev, int32_t T, int32_t hidden, int32_t vocab, int32_t n_pos_rows, int32_t offset) {
  KCK(opt_embed_launch((const int32_t*)tok_dev, (const int32_t*)pos_dev, (const bf16*)tok_table_dev,
                       (const bf16*)pos_table_dev, (bf16*)out_dev, T, hidden, vocab, n_pos_rows, offset, 0));
  KCK(cudaDeviceSynchronize());
  return 0;
}

int tgis_k_silu_mul(const void* gate_up_dev, void* act_dev, int32_t T, int32_t ffn) {
  KCK(silu_mul_launch((const bf16*)gate_up_dev, (bf16*)act_dev, T, ffn, 0));
  KCK(cudaDeviceSynchronize());
  return 0;
}

int tgis_k_rope_kv(void* qkv_dev, const int32_t* positions_host, const int32_t* slot_mapping_host,
                   const void* cos_sin_dev, void* k_cache_dev, void* v_cache_dev, int32_t T, int32_t n_q, int32_t n_kv) {
  Tmp<int32_t> pos, sm;
  KCK(pos.upload(positions_host, T));
  KCK(sm.upload(slot_mapping_host, T));
  KCK(rope_kvwrite_launch((bf16*)qkv_dev, pos.p, sm.p, (const bf16*)cos_sin_dev, (bf16*)k_cache_dev, (bf16*)v_cache_dev,
                          T, n_q, n_kv, 0));
  KCK(cudaDeviceSynchronize());
  return 0;
}

// qkv projection with the fused RoPE + KV-scatter epilogue (in the cluster or the global split-tile reduction).  Returns 1
// when the launch plan does not split every tile (nothing computed), 0 on success.
int tgis_k_gemm_rope(const void* x_dev, const void* w_dev, void* y_dev, int32_t T, int32_t n_q, int32_t n_kv, int32_t K,
                     int32_t x_rows_alloc, const int32_t* positions_host, const int32_t* slot_mapping_host,
                     const void* cos_sin_dev, void* k_cache_dev, void* v_cache_dev) {
  const int N = (n_q + 2 * n_kv) * HEAD_DIM;
  const int sms = test_num_sms();
  if (gemm_even_split(T, N, K, sms) < 2) return 1;
  CUtensorMap wm, xm;
  const int bt = gemm_pick_bt(T);
  if (x_rows_alloc < bt) return kfail("x must have at least one TMA box of rows allocated");
  if (make_tmap_bf16_2d(&wm, w_dev, N, K, K, 128, 64) != 0) return kfail("weight tensor map failed");
  if (make_tmap_bf16_2d(&xm, x_dev, x_rows_alloc, K, K, bt, 64) != 0) return kfail("activation tensor map failed");
  Tmp<float> ws;
  Tmp<int> ctr;
  Tmp<int32_t> pos, sm;
  KCK(ws.alloc(gemm_workspace_bytes(sms) / sizeof(float)));
  KCK(ctr.alloc(1 << 16));
  KCK(cudaMemset(ctr.p, 0, sizeof(int) << 16));
  KCK(pos.upload(positions_host, T));
  KCK(sm.upload(slot_mapping_host, T));
  const GemmRope rp{pos.p, sm.p, (const bf16*)cos_sin_dev, (bf16*)k_cache_dev, (bf16*)v_cache_dev, n_q, n_kv};
  KCK(gemm_bf16_launch(wm, xm, y_dev, N, T, N, K, ws.p, ctr.p, sms, 0, 0, nullptr, nullptr, &rp));
  KCK(cudaDeviceSynchronize());
  return 0;
}

// Layer-stack slice  a -> [GEMM W1] -> (+ residual, RMSNorm w_norm) -> [GEMM W2 (out_mode2: 0 bf16, 2 fused SwiGLU)] -> y2
// run either as three launches (fused = 0: gemm, add_rmsnorm_launch, gemm) or as two with the add + norm folded into the
// two GEMMs (fused = 1, GemmNorm).  residual [T, H] is updated in place either way.  Returns 1 when the launch plan cannot
// fuse this shape (nothing computed).  iters > 1 times the chain (residual keeps accumulating: timing only).
int tgis_k_gemm_norm_chain(const void* a_dev, int32_t a_rows_alloc, const void* w1_dev, void* residual_dev,
                           const void* w_norm_dev, const void* w2_dev, void* y2_dev, int32_t T, int32_t K1, int32_t H,
                           int32_t N2, float eps, int32_t out_mode2, int32_t fused, int32_t iters, float* us_out) {
  const int sms = test_num_sms();
  const int bt = gemm_pick_bt(T);
  if (fused && (T > GEMM_NORM_MAX_T || gemm_cluster_split(T, H, K1, sms) == 0 || H % 128 != 0)) return 1;
  if (a_rows_alloc < bt) return kfail("a must have at least one TMA box of rows allocated");
  Tmp<bf16> y1, xn;
  Tmp<float> ssq, ws;
  Tmp<int> ctr;
  const int rows = T > bt ? T : bt;
  KCK(y1.alloc((size_t)rows * H));
  KCK(xn.alloc((size_t)rows * H));
  KCK(cudaMemset(xn.p, 0, (size_t)rows * H * sizeof(bf16)));
  KCK(ssq.alloc((size_t)rows * (H / 128 + 1)));
  KCK(ws.alloc(gemm_workspace_bytes(sms) / sizeof(float)));
  KCK(ctr.alloc(1 << 16));
  KCK(cudaMemset(ctr.p, 0, sizeof(int) << 16));
  CUtensorMap wm1, am, wm2, xm;
  if (make_tmap_bf16_2d(&wm1, w1_dev, H, K1, K1, 128, 64) != 0) return kfail("w1 tensor map failed");
  if (make_tmap_bf16_2d(&am, a_dev, a_rows_alloc, K1, K1, bt, 64) != 0) return kfail("a tensor map failed");
  if (make_tmap_bf16_2d(&wm2, w2_dev, N2, H, H, 128, 64) != 0) return kfail("w2 tensor map failed");
  // the fused consumer stages the raw residual stream by TMA and normalises it in shared memory
  if (make_tmap_bf16_2d(&xm, fused ? residual_dev : (void*)xn.p, fused ? a_rows_alloc : rows, H, H, bt, 64) != 0)
    return kfail("xn tensor map failed");
  const int ldy2 = out_mode2 == 2 ? N2 / 2 : N2;
  cudaEvent_t e0, e1;
  KCK(cudaEventCreate(&e0));
  KCK(cudaEventCreate(&e1));
  if (iters < 1) iters = 1;
  KCK(cudaEventRecord(e0, 0));
  for (int i = 0; i < iters; ++i) {
    if (fused) {
      GemmNorm prod{};
      prod.residual = (bf16*)residual_dev;
      prod.sumsq_out = ssq.p;
      KCK(gemm_bf16_launch(wm1, am, y1.p, H, T, H, K1, ws.p, ctr.p, sms, 0, 0, nullptr, nullptr, nullptr, &prod));
      GemmNorm cons{};
      cons.h = (const bf16*)residual_dev;
      cons.sumsq_in = ssq.p;
      cons.w_norm = (const bf16*)w_norm_dev;
      cons.n_parts = H / 128;
      cons.eps = eps;
      KCK(gemm_bf16_launch(wm2, xm, y2_dev, ldy2, T, N2, H, ws.p, ctr.p, sms, 0, out_mode2, nullptr, nullptr, nullptr, &cons));
    } else {
      KCK(gemm_bf16_launch(wm1, am, y1.p, H, T, H, K1, ws.p, ctr.p, sms, 0, 0));
      KCK(add_rmsnorm_launch(y1.p, (bf16*)residual_dev, (const bf16*)w_norm_dev, xn.p, T, H, eps, 0));
      KCK(gemm_bf16_launch(wm2, xm, y2_dev, ldy2, T, N2, H, ws.p, ctr.p, sms, 0, out_mode2));
    }
  }
  KCK(cudaEventRecord(e1, 0));
  KCK(cudaDeviceSynchronize());
  float ms = 0.f;
  KCK(cudaEventElapsedTime(&ms, e0, e1));
  if (us_out) *us_out = ms * 1000.f / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

int tgis_k_attention(const void* qkv_dev, const void* k_cache_dev, const void* v_cache_dev, const int32_t* seqs_host,
                     int32_t n_seqs, const int32_t* block_table_host, int32_t bt_rows, int32_t bt_stride, void* out_dev,
                     int32_t n_q, int32_t n_kv, float scale) {
  if (n_kv <= 0 || n_q % n_kv) return kfail("bad head counts");
  const int G = n_q / n_kv;
  const int qkv_ld = (n_q + 2 * n_kv) * HEAD_DIM, out_ld = n_q * HEAD_DIM;
  std::vector<AttnSeq> seqs(n_seqs);
  std::vector<int32_t> dec, tseq, tq0;
  int max_kv = 1;
  for (int s = 0; s < n_seqs; ++s) {
    seqs[s] = AttnSeq{seqs_host[4 * s], seqs_host[4 * s + 1], seqs_host[4 * s + 2], seqs_host[4 * s + 3]};
    if (seqs[s].q_len == 1) {
      dec.push_back(s);
      max_kv = std::max(max_kv, seqs[s].kv_len);
    } else {
      for (int q0 = 0; q0 < seqs[s].q_len; q0 += 16) {
        tseq.push_back(s);
        tq0.push_back(q0);
      }
    }
  }
  Tmp<AttnSeq> d_seqs;
  Tmp<int32_t> d_dec, d_tseq, d_tq0, d_bt;
  Tmp<float> po, pml;
  KCK(d_seqs.upload(seqs.data(), n_seqs));
  KCK(d_dec.upload(dec.data(), dec.size()));
  KCK(d_tseq.upload(tseq.data(), tseq.size()));
  KCK(d_tq0.upload(tq0.data(), tq0.size()));
  KCK(d_bt.upload(block_table_host, (size_t)bt_rows * bt_stride));
  const int max_splits = (max_kv + DECODE_SPLIT - 1) / DECODE_SPLIT;
  const size_t nd = dec.size() ? dec.size() : 1;
  KCK(po.alloc(nd * n_kv * max_splits * G * HEAD_DIM));
  KCK(pml.alloc(nd * n_kv * max_splits * G * 2));
  std::vector<DecItem> items(1 + nd * max_splits);
  decode_items_build(items.data(), seqs.data(), dec.data(), (int)dec.size(), block_table_host, bt_stride);
  Tmp<DecItem> d_items;
  KCK(d_items.upload(items.data(), items.size()));
  Tmp<int> arrive;  // arrival counters of the in-kernel split merge (the launcher ignores them when that path is off)
  KCK(arrive.alloc(nd * n_kv));
  KCK(cudaMemset(arrive.p, 0, sizeof(int) * nd * n_kv));
  KCK(attn_decode_launch((const bf16*)qkv_dev, qkv_ld, (const bf16*)k_cache_dev, (const bf16*)v_cache_dev, d_items.p,
                         (int)dec.size() * max_splits, d_seqs.p, d_dec.p, (int)dec.size(), max_splits, po.p, pml.p,
                         (bf16*)out_dev, out_ld, n_q, n_kv, scale, test_num_sms(), 0, arrive.p));
  KCK(attn_prefill_launch((const bf16*)qkv_dev, qkv_ld, (const bf16*)k_cache_dev, (const bf16*)v_cache_dev, d_seqs.p,
                          d_tseq.p, d_tq0.p, (int)tseq.size(), d_bt.p, bt_stride, (bf16*)out_dev, out_ld, n_q, n_kv,
                          scale, 0));
  KCK(cudaDeviceSynchronize());
  return 0;
}

// decode attention only, timed: `iters` launches rotating over `n_layers` cache copies (layer_stride_bytes apart) so that
// a small batch does not simply sit in L2; us_out = average device time per launch
int tgis_k_attention_bench(const void* qkv_dev, const void* k_cache_dev, const void* v_cache_dev,
                           const int32_t* seqs_host, int32_t n_seqs, const int32_t* block_table_host, int32_t bt_rows,
                           int32_t bt_stride, void* out_dev, int32_t n_q, int32_t n_kv, float scale, int32_t n_layers,
                           int64_t layer_stride_bytes, int32_t iters, float* us_out) {
  if (n_kv <= 0 || n_q % n_kv) return kfail("bad head counts");
  const int G = n_q / n_kv;
  const int qkv_ld = (n_q + 2 * n_kv) * HEAD_DIM, out_ld = n_q * HEAD_DIM;
  std::vector<AttnSeq> seqs(n_seqs);
  int max_kv = 1;
  for (int s = 0; s < n_seqs; ++s) {
    seqs[s] = AttnSeq{seqs_host[4 * s], seqs_host[4 * s + 1], seqs_host[4 * s + 2], seqs_host[4 * s + 3]};
    if (seqs[s].q_len != 1) return kfail("decode sequences only");
    max_kv = std::max(max_kv, seqs[s].kv_len);
  }
  Tmp<AttnSeq> d_seqs;
  Tmp<int32_t> d_bt;
  Tmp<float> po, pml;
  KCK(d_seqs.upload(seqs.data(), n_seqs));
  KCK(d_bt.upload(block_table_host, (size_t)bt_rows * bt_stride));
  const int max_splits = (max_kv + DECODE_SPLIT - 1) / DECODE_SPLIT;
  KCK(po.alloc((size_t)n_seqs * n_kv * max_splits * G * HEAD_DIM));
  KCK(pml.alloc((size_t)n_seqs * n_kv * max_splits * G * 2));
  std::vector<DecItem> items(1 + (size_t)n_seqs * max_splits);
  decode_items_build(items.data(), seqs.data(), nullptr, n_seqs, block_table_host, bt_stride);
  Tmp<DecItem> d_items;
  KCK(d_items.upload(items.data(), items.size()));
  const int num_sms = test_num_sms();
  Tmp<int> arrive;
  KCK(arrive.alloc((size_t)n_seqs * n_kv));
  KCK(cudaMemset(arrive.p, 0, sizeof(int) * (size_t)n_seqs * n_kv));
  cudaEvent_t e0, e1;
  KCK(cudaEventCreate(&e0));
  KCK(cudaEventCreate(&e1));
  for (int it = -2; it < iters; ++it) {
    if (it == 0) KCK(cudaEventRecord(e0, 0));
    const int l = ((it % n_layers) + n_layers) % n_layers;
    const char* kc = (const char*)k_cache_dev + (size_t)l * layer_stride_bytes;
    const char* vc = (const char*)v_cache_dev + (size_t)l * layer_stride_bytes;
    KCK(attn_decode_launch((const bf16*)qkv_dev, qkv_ld, (const bf16*)kc, (const bf16*)vc, d_items.p,
                           n_seqs * max_splits, d_seqs.p, nullptr, n_seqs, max_splits, po.p, pml.p, (bf16*)out_dev,
                           out_ld, n_q, n_kv, scale, num_sms, 0, arrive.p));
  }
  KCK(cudaEventRecord(e1, 0));
  KCK(cudaDeviceSynchronize());
  float ms = 0.f;
  KCK(cudaEventElapsedTime(&ms, e0, e1));
  *us_out = ms * 1000.f / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

// Host-only: the launch plan of the GEMM for a shape (pure functions of the shape; no GPU needed): token-tile size,
// grid, and the even-split factor (0 = stream-K ranges may straddle tiles, reduced through global memory).
int tgis_k_gemm_plan(int32_t T, int32_t N, int32_t K, int32_t num_sms, int32_t* bt_out, int32_t* grid_out,
                     int32_t* even_split_out) {
  if (T <= 0 || N <= 0 || K <= 0 || num_sms <= 0) return kfail("bad shape");
  *bt_out = gemm_pick_bt(T);
  *grid_out = gemm_grid_size(T, N, K, num_sms);
  *even_split_out = gemm_even_split(T, N, K, num_sms);
  return 0;
}

// weight rows per GEMM unit (128 x the number of weight tiles that share one activation tile) for a T-token launch
int tgis_k_gemm_unit_rows(int32_t T) { return 128 * gemm_nw(T); }

// Host-only: the decode work-item list the scheduler builds for a step (no GPU needed).  seqs_host as in
// tgis_k_attention (all decode sequences); items_out: (1 + capacity) x 8 int32 records; returns the entry count or -1.
int tgis_k_decode_items(const int32_t* seqs_host, int32_t n_seqs, const int32_t* block_table_host, int32_t bt_stride,
                        int32_t* items_out, int32_t capacity) {
  std::vector<AttnSeq> seqs(n_seqs);
  long long need = 0;
  for (int s = 0; s < n_seqs; ++s) {
    seqs[s] = AttnSeq{seqs_host[4 * s], seqs_host[4 * s + 1], seqs_host[4 * s + 2], seqs_host[4 * s + 3]};
    need += (seqs[s].kv_len + DECODE_SPLIT - 1) / DECODE_SPLIT;
  }
  if (need > capacity) return kfail("capacity too small");
  static_assert(sizeof(DecItem) == 32, "8 int32 per record");
  return decode_items_build(reinterpret_cast<DecItem*>(items_out), seqs.data(), nullptr, n_seqs, block_table_host,
                            bt_stride);
}

// logits_bf16 = 0: fp32 logits, 1: bf16 logits.  iters > 1: the launch is repeated and us_out (optional) receives the
// average device time per launch (CUDA events on the launching stream).
int tgis_k_sampler_masked(const void* logits_dev, int32_t logits_bf16, int32_t ld, int32_t vocab, const void* rows_host,
                          int32_t n_rows, void* seen_bitmap_dev, const void* allow_bitmap_dev, void* out_host,
                          int32_t iters, float* us_out) {
  Tmp<SampleRow> rows;
  Tmp<SampleOut> outs;
  Tmp<float> scratch;
  Tmp<uint32_t> dummy_bm;
  KCK(rows.upload((const SampleRow*)rows_host, n_rows));
  KCK(outs.alloc(n_rows));
  KCK(scratch.alloc((size_t)n_rows * vocab));
  uint32_t* bm = (uint32_t*)seen_bitmap_dev;
  const int words = (vocab + 31) / 32;
  if (!bm) {
    // rows may still carry seq_slot >= 0: give them a scratch bitmap big enough for the largest slot
    int max_slot = 0;
    for (int i = 0; i < n_rows; ++i) max_slot = std::max(max_slot, ((const SampleRow*)rows_host)[i].seq_slot);
    KCK(dummy_bm.alloc((size_t)(max_slot + 1) * words));
    KCK(cudaMemset(dummy_bm.p, 0, sizeof(uint32_t) * (size_t)(max_slot + 1) * words));
    bm = dummy_bm.p;
  }
  if (iters < 1) iters = 1;
  int any_complex = 0;  // same rule as the engine: a row that is not plain greedy / forced needs selection passes
  for (int i = 0; i < n_rows; ++i) {
    const int f = ((const SampleRow*)rows_host)[i].flags;
    if (!(f & (SAMPLE_GREEDY | SAMPLE_FORCED)) || ((f & SAMPLE_TYPICAL) && !(f & SAMPLE_FORCED))) any_complex = 1;
  }
  cudaEvent_t e0, e1;
  KCK(cudaEventCreate(&e0));
  KCK(cudaEventCreate(&e1));
  if (iters > 1)  // warm-up launch outside the timed region
    KCK(sampler_launch(logits_dev, logits_bf16, ld, vocab, rows.p, n_rows, bm, words, scratch.p, outs.p, 0, any_complex, 148,
                       (const uint32_t*)allow_bitmap_dev));
  KCK(cudaEventRecord(e0, 0));
  for (int it = 0; it < iters; ++it)
    KCK(sampler_launch(logits_dev, logits_bf16, ld, vocab, rows.p, n_rows, bm, words, scratch.p, outs.p, 0, any_complex, 148,
                       (const uint32_t*)allow_bitmap_dev));
  KCK(cudaEventRecord(e1, 0));
  KCK(cudaMemcpy(out_host, outs.p, sizeof(SampleOut) * n_rows, cudaMemcpyDeviceToHost));
  KCK(cudaDeviceSynchronize());
  float ms = 0.f;
  KCK(cudaEventElapsedTime(&ms, e0, e1));
  if (us_out) *us_out = 1e3f * ms / (float)iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

int tgis_k_sampler_ex(const void* logits_dev, int32_t logits_bf16, int32_t ld, int32_t vocab, const void* rows_host,
                      int32_t n_rows, void* seen_bitmap_dev, void* out_host, int32_t iters, float* us_out) {
  return tgis_k_sampler_masked(logits_dev, logits_bf16, ld, vocab, rows_host, n_rows, seen_bitmap_dev, nullptr, out_host,
                               iters, us_out);
}

// one LoRA module through both kernels: y[t, col0 + n] += B_s[n, :] . (A_s x[t]) for tokens with tok_slot[t] = s >= 1
int tgis_k_lora(const void* x_dev, int32_t ldx, const int32_t* tok_slot_dev, const void* a_dev, const void* b_dev, int32_t K,
                int32_t N, int32_t Rm, int32_t col0, void* y_dev, int32_t ldy, int32_t T) {
  Tmp<float> v;
  KCK(v.alloc((size_t)T * Rm));
  LoraGroup g{};
  g.n_mods = 1;
  g.v_ld = Rm;
  g.mod[0] = LoraModule{(const __nv_bfloat16*)a_dev, (const __nv_bfloat16*)b_dev, K, N, Rm, col0, 0};
  KCK(lora_shrink_launch((const __nv_bfloat16*)x_dev, ldx, tok_slot_dev, g, v.p, T, 0));
  KCK(lora_expand_launch(v.p, tok_slot_dev, g, (__nv_bfloat16*)y_dev, ldy, T, 0));
  KCK(cudaDeviceSynchronize());
  return 0;
}

// timing variant: `iters` back-to-back launches of each kernel on the default stream, CUDA events around each group
int tgis_k_lora_bench(const void* x_dev, int32_t ldx, const int32_t* tok_slot_dev, const void* a_dev, const void* b_dev,
                      int32_t K, int32_t N, int32_t Rm, void* y_dev, int32_t ldy, int32_t T, int32_t iters,
                      float* us_shrink, float* us_expand) {
  Tmp<float> v;
  KCK(v.alloc((size_t)T * Rm));
  LoraGroup g{};
  g.n_mods = 1;
  g.v_ld = Rm;
  g.mod[0] = LoraModule{(const __nv_bfloat16*)a_dev, (const __nv_bfloat16*)b_dev, K, N, Rm, 0, 0};
  cudaEvent_t e0, e1, e2;
  KCK(cudaEventCreate(&e0));
  KCK(cudaEventCreate(&e1));
  KCK(cudaEventCreate(&e2));
  if (iters < 1) iters = 1;
  KCK(lora_shrink_launch((const __nv_bfloat16*)x_dev, ldx, tok_slot_dev, g, v.p, T, 0));
  KCK(lora_expand_launch(v.p, tok_slot_dev, g, (__nv_bfloat16*)y_dev, ldy, T, 0));
  KCK(cudaEventRecord(e0, 0));
  for (int i = 0; i < iters; ++i) KCK(lora_shrink_launch((const __nv_bfloat16*)x_dev, ldx, tok_slot_dev, g, v.p, T, 0));
  KCK(cudaEventRecord(e1, 0));
  for (int i = 0; i < iters; ++i) KCK(lora_expand_launch(v.p, tok_slot_dev, g, (__nv_bfloat16*)y_dev, ldy, T, 0));
  KCK(cudaEventRecord(e2, 0));
  KCK(cudaDeviceSynchronize());
  float ms = 0.f;
  KCK(cudaEventElapsedTime(&ms, e0, e1));
  if (us_shrink) *us_shrink = 1e3f * ms / (float)iters;
  KCK(cudaEventElapsedTime(&ms, e1, e2));
  if (us_expand) *us_expand = 1e3f * ms / (float)iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaEventDestroy(e2);
  return 0;
}

int tgis_k_silu_mul_interleaved(const void* gate_up_dev, void* act_dev, int32_t T, int32_t ffn) {
  KCK(silu_mul_interleaved_launch((const __nv_bfloat16*)gate_up_dev, (__nv_bfloat16*)act_dev, T, ffn, 0));
  KCK(cudaDeviceSynchronize());
  return 0;
}

int tgis_k_sampler(const void* logits_dev, int32_t ld, int32_t vocab, const void* rows_host, int32_t n_rows,
                   void* seen_bitmap_dev, void* out_host) {
  return tgis_k_sampler_ex(logits_dev, 0, ld, vocab, rows_host, n_rows, seen_bitmap_dev, out_host, 1, nullptr);
}

}  // extern "C"

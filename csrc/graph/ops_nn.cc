// Hot-path ops of transformer training, bound to the hand-written sm_100a kernels
// (tcgen05 GEMM / flash attention, fused norms, activations, embedding, losses),
// with ATen reference implementations for CPU / fp32 so the same graph runs in the
// CPU test-suite.  Every op has an explicit gradient (no recompute-VJP here).
// (capability parity: hetu/graph/ops/{Linear,matmul,LayerNorm,RMSNorm,Gelu,Relu,SwiGLU,
//  EmbeddingLookup,SoftmaxCrossEntropySparse,Attention,Rotary,Dropout}.cc)
#include <ATen/ATen.h>

#include <atomic>

#include "exec.h"
#include "ir.h"
#include "op_utils.h"

namespace hb {

using Ts = std::vector<at::Tensor>;

static std::atomic<int64_t> g_fallbacks{0};
bool strict_native() {
  static bool s = env_int("HETU_B200_STRICT", 0) != 0;
  return s;
}
void note_fallback(const char* op) {
  g_fallbacks.fetch_add(1);
  HB_CHECK(!strict_native()) << "op " << op << " fell back to ATen on a CUDA bf16 tensor (HETU_B200_STRICT=1)";
}
int64_t fallback_count() { return g_fallbacks.load(); }

at::Tensor native_add(const at::Tensor& a, const at::Tensor& b) {
  if (is_native(a) && is_native(b) && a.sizes() == b.sizes() && a.is_contiguous() && b.is_contiguous()) {
    at::Tensor out = at::empty_like(a);
    cuda_ok(add_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), cur_stream()), "add_bf16");
    return out;
  }
  return a + b;
}

// ------------------------------------------------------------------ DS rule for contractions
DistributedStates matmul_ds(const DistributedStates& a, int a_nd, int a_k, const DistributedStates& b, int b_k, int b_n,
                            int out_n_dim, const std::vector<int>& a_dim_to_out) {
  (void)a_nd;
  const int n = a.device_num();
  HB_CHECK(n == b.device_num()) << "matmul operands live on different numbers of devices";
  const int k_split = a.get_dim(a_k);
  HB_CHECK(k_split == b.get_dim(b_k)) << "contraction dim is split " << k_split << "-way on one operand and "
                                      << b.get_dim(b_k) << "-way on the other";
  std::map<int, int> st;
  int used = 1;
  for (auto& kv : a.states()) {
    if (kv.second <= 1 || kv.first < 0 || kv.first == a_k) continue;
    st[a_dim_to_out[kv.first]] = kv.second;
    used *= kv.second;
  }
  const int n_split = b.get_dim(b_n);
  if (n_split > 1) { st[out_n_dim] = (st.count(out_n_dim) ? st[out_n_dim] : 1) * n_split; used *= n_split; }
  const int partial = k_split * a.get_dim(kPartialDim) * b.get_dim(kPartialDim);
  if (partial > 1) { st[kPartialDim] = partial; used *= partial; }
  HB_CHECK(n % used == 0) << "inconsistent layouts in matmul: " << a << " x " << b;
  if (n / used > 1) st[kDupDim] = n / used;
  // device order: follow A's order, k -> partial, A's duplicate axis -> B's n split (or stays duplicate)
  std::vector<int> order;
  auto push = [&](int d) {
    if (st.count(d) && st[d] > 1 && std::find(order.begin(), order.end(), d) == order.end()) order.push_back(d);
  };
  for (int o : a.order()) {
    if (o == a_k || o == kPartialDim) push(kPartialDim);
    else if (o == kDupDim) { if (n_split > 1) push(out_n_dim); push(kDupDim); }
    else push(a_dim_to_out[o]);
  }
  for (auto& kv : st) push(kv.first);
  return DistributedStates(n, st, order);
}

// ------------------------------------------------------------------ GEMM helpers
static int act_code(const std::string& a) {
  if (a.empty() || a == "none") return ACT_NONE;
  if (a == "gelu") return ACT_GELU;
  if (a == "relu") return ACT_RELU;
  if (a == "gelu_tanh") return ACT_GELU_TANH;
  if (a == "silu") return ACT_SILU;
  HB_FAIL() << "unknown activation " << a;
}
static at::Tensor aten_act(const at::Tensor& x, int code) {
  switch (code) {
    case ACT_GELU: return at::gelu(x);
    case ACT_RELU: return at::relu(x);
    case ACT_GELU_TANH: return at::gelu(x, "tanh");
    case ACT_SILU: return at::silu(x);
    default: return x;
  }
}
static int act_bwd_mode(const std::string& k) {
  if (k == "gelu") return AUX_DGELU;
  if (k == "relu") return AUX_DRELU;
  if (k == "gelu_tanh") return AUX_DGELU_TANH;
  if (k == "silu") return AUX_DSILU;
  HB_FAIL() << "no fused backward for activation " << k;
}
// dy * act'(x) in fp32 with ATen (CPU path / reference numerics)
static at::Tensor aten_act_bwd(const at::Tensor& dy, const at::Tensor& x, const std::string& kind) {
  at::Tensor xf = x.to(at::kFloat), d;
  if (kind == "gelu") d = 0.5 * (1 + at::erf(xf * M_SQRT1_2)) + xf * at::exp(-0.5 * xf * xf) * 0.3989422804014327;
  else if (kind == "relu") d = (xf > 0).to(at::kFloat);
  else if (kind == "silu") { auto sg = at::sigmoid(xf); d = sg * (1 + xf * (1 - sg)); }
  else {
    auto u = 0.7978845608028654 * (xf + 0.044715 * xf * xf * xf);
    auto t = at::tanh(u);
    d = 0.5 * (1 + t) + 0.5 * xf * (1 - t * t) * 0.7978845608028654 * (1 + 0.134145 * xf * xf);
  }
  return (dy.to(at::kFloat) * d).to(x.scalar_type());
}

// C[M,N] = A * B with explicit operand majors on raw 2-D contiguous tensors
static void run_gemm(const at::Tensor& A, bool a_mn, const at::Tensor& B, bool b_mn, at::Tensor& C, int64_t M, int64_t N,
                     int64_t K, const at::Tensor* bias, const at::Tensor* aux_in, int aux_mode, at::Tensor* aux_out,
                     int act, bool accumulate, float alpha = 1.0f) {
  GemmCall c;
  c.A = A.data_ptr(); c.B = B.data_ptr(); c.C = C.data_ptr();
  c.M = (int)M; c.N = (int)N; c.K = (int)K;
  c.lda = A.stride(0); c.ldb = B.stride(0); c.ldc = C.stride(0);
  c.a_mn_major = a_mn; c.b_mn_major = b_mn;
  c.out = C.scalar_type() == at::kFloat ? GemmOut::FP32 : GemmOut::BF16;
  if (bias) c.bias = bias->data_ptr();
  if (aux_in) { c.aux_in = aux_in->data_ptr(); c.aux_mode = aux_mode; c.ld_aux = aux_in->stride(0); }
  if (aux_out) { c.aux_out = aux_out->data_ptr(); c.ld_aux = aux_out->stride(0); }
  c.act = act;
  c.accumulate = accumulate;
  c.alpha = alpha;
  GemmGather* ag = tls_gemm_gather;
  if (ag != nullptr && !ag->used && A.data_ptr() == ag->dst && !a_mn && M == ag->rows_per_rank * ag->world) {
    c.ag_src = ag->src; c.ag_flags = ag->flags; c.ag_world = ag->world; c.ag_rank = ag->my_rank; c.ag_rows_per_rank = (int)ag->rows_per_rank;
    ag->used = true;
  }
  GemmSink* sink = tls_gemm_sink;
  if (sink != nullptr && !sink->used && bias == nullptr && aux_in == nullptr && aux_out == nullptr && act == ACT_NONE && !accumulate &&
      c.out == GemmOut::BF16 && M == sink->rows_per_rank * sink->world && N == sink->cols) {
    c.C = sink->local; c.ldc = N;
    c.peer_c = sink->peers; c.world = sink->world; c.my_rank = sink->my_rank; c.rows_per_rank = (int)sink->rows_per_rank;
    sink->used = true;
  }
  cuda_ok(gemm_bf16(c, cur_stream()), "tcgen05 gemm");
}
// weight gradient dW[N,K] = dy^T x with the GEMM -> reduce-scatter epilogue: every 128-row block of dW is stored into
// the staging slot `my_rank` of the rank owning those rows (peer memory over NVLink); see graph/zero_fused.cc
void wgrad_to_peer_slots(const at::Tensor& dy, const at::Tensor& x, bool trans_b, void* local_c, void* const* peer_c, int world,
                         int my_rank, int64_t rows_per_rank) {
  HB_CHECK(trans_b) << "fused ZeRO weight gradients need [out, in] weights";
  at::Tensor d2 = flatten_rows(dy).contiguous(), x2 = flatten_rows(x).contiguous();
  const int64_t T = d2.size(0), N = d2.size(1), K = x2.size(1);
  GemmCall c;
  c.A = d2.data_ptr(); c.B = x2.data_ptr(); c.C = local_c;
  c.M = (int)N; c.N = (int)K; c.K = (int)T;
  c.lda = d2.stride(0); c.ldb = x2.stride(0); c.ldc = K;
  c.a_mn_major = true; c.b_mn_major = true;
  c.out = GemmOut::BF16;
  c.peer_c = peer_c; c.world = world; c.my_rank = my_rank; c.rows_per_rank = (int)rows_per_rank;
  cuda_ok(gemm_bf16(c, cur_stream()), "tcgen05 wgrad -> peer slots");
}
static bool gemm_ok(const at::Tensor& t) {
  return is_native(t) && t.dim() == 2 && t.stride(1) == 1 && (t.stride(0) % 8) == 0 &&
         (reinterpret_cast<uintptr_t>(t.data_ptr()) % 16) == 0;
}

// ------------------------------------------------------------------ linear
// inputs: x [..., K], w ([N, K] when trans_b else [K, N]), [bias [N]], [residual [..., N]]
// outputs: y (and the pre-activation when an activation is fused)
static Ts linear_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const bool trans_b = op.attrs.b("trans_b", true);
  const bool has_bias = op.attrs.b("has_bias"), has_res = op.attrs.b("has_residual");
  const int act = act_code(op.attrs.s("act"));
  const at::Tensor& x = in[0];
  const at::Tensor& w = in[1];
  const at::Tensor* bias = has_bias ? &in[2] : nullptr;
  const at::Tensor* res = has_res ? &in[has_bias ? 3 : 2] : nullptr;
  const int64_t N = trans_b ? w.size(0) : w.size(1);
  const int64_t K = x.size(-1);
  std::vector<int64_t> oshape = x.sizes().vec();
  oshape.back() = N;
  if (x.is_meta()) {
    Ts r = {at::empty(oshape, x.options())};
    if (act != ACT_NONE) r.push_back(at::empty(oshape, x.options()));
    return r;
  }
  at::Tensor x2 = flatten_rows(x).contiguous();
  const int64_t M = x2.size(0);
  if (gemm_ok(x2) && gemm_ok(w) && (!bias || is_native(*bias)) && (N % 8) == 0) {
    at::Tensor y = at::empty({M, N}, x.options());
    at::Tensor pre;
    at::Tensor r2;
    if (res) r2 = flatten_rows(*res).contiguous();
    if (act != ACT_NONE) pre = at::empty({M, N}, x.options());
    run_gemm(x2, false, w, !trans_b, y, M, N, K, bias, res ? &r2 : nullptr, AUX_ADD, act != ACT_NONE ? &pre : nullptr, act,
             false);
    Ts out = {y.reshape(oshape)};
    if (act != ACT_NONE) out.push_back(pre.reshape(oshape));
    return out;
  }
  if (is_native(x)) note_fallback("linear");
  at::Tensor y = trans_b ? at::matmul(x2, w.t()) : at::matmul(x2, w);
  if (bias) y = y + *bias;
  Ts out;
  if (act != ACT_NONE) {
    at::Tensor pre = y;
    y = aten_act(pre, act);
    if (res) y = y + flatten_rows(*res);
    out = {y.reshape(oshape), pre.reshape(oshape)};
  } else {
    if (res) y = y + flatten_rows(*res);
    out = {y.reshape(oshape)};
  }
  return out;
}

// dgrad: dx[..., K] = dy[..., N] * W    (W [N,K] when trans_b)
static Ts linear_dgrad_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const bool trans_b = op.attrs.b("trans_b", true);
  const at::Tensor& dy = in[0];
  const at::Tensor& w = in[1];
  const int64_t K = trans_b ? w.size(1) : w.size(0);
  std::vector<int64_t> oshape = dy.sizes().vec();
  oshape.back() = K;
  if (dy.is_meta()) return {at::empty(oshape, dy.options())};
  at::Tensor d2 = flatten_rows(dy).contiguous();
  const int64_t M = d2.size(0), N = d2.size(1);
  if (gemm_ok(d2) && gemm_ok(w) && (K % 8) == 0) {
    at::Tensor dx = at::empty({M, K}, dy.options());
    // dx[M,K] = dy[M,N] * W[N,K]: B must be [K(out-n), N(contract)]; W[N,K] row-major = MN-major B
    if (in.size() > 2) {
      // third operand: pre-activation (dx *= act'(pre)) or, with residual_add, a gradient to add (dx += other)
      at::Tensor aux = flatten_rows(in[2]).contiguous();
      const int mode = op.attrs.b("residual_add") ? (int)AUX_ADD : act_bwd_mode(op.attrs.s("act_bwd"));
      run_gemm(d2, false, w, trans_b, dx, M, K, N, nullptr, &aux, mode, nullptr, ACT_NONE, false);
    } else {
      run_gemm(d2, false, w, trans_b, dx, M, K, N, nullptr, nullptr, 0, nullptr, ACT_NONE, false);
    }
    return {dx.reshape(oshape)};
  }
  if (is_native(dy)) note_fallback("linear_dgrad");
  at::Tensor dxa = (trans_b ? at::matmul(d2, w) : at::matmul(d2, w.t())).reshape(oshape);
  if (in.size() > 2) dxa = op.attrs.b("residual_add") ? dxa + in[2].reshape(oshape) : aten_act_bwd(dxa, in[2], op.attrs.s("act_bwd"));
  return {dxa};
}
// wgrad: dw = dy^T * x  ([N,K] when trans_b else [K,N])
static Ts linear_wgrad_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const bool trans_b = op.attrs.b("trans_b", true);
  const at::Tensor& dy = in[0];
  const at::Tensor& x = in[1];
  const int64_t N = dy.size(-1), K = x.size(-1);
  std::vector<int64_t> wshape = trans_b ? std::vector<int64_t>{N, K} : std::vector<int64_t>{K, N};
  if (dy.is_meta()) return {at::empty(wshape, dy.options())};
  at::Tensor d2 = flatten_rows(dy).contiguous(), x2 = flatten_rows(x).contiguous();
  const int64_t T = d2.size(0);
  if (gemm_ok(d2) && gemm_ok(x2) && (N % 8) == 0 && (K % 8) == 0) {
    at::Tensor dw = at::empty(wshape, dy.options());
    if (trans_b) run_gemm(d2, true, x2, true, dw, N, K, T, nullptr, nullptr, 0, nullptr, ACT_NONE, false);
    else run_gemm(x2, true, d2, true, dw, K, N, T, nullptr, nullptr, 0, nullptr, ACT_NONE, false);
    return {dw};
  }
  if (is_native(dy)) note_fallback("linear_wgrad");
  return {trans_b ? at::matmul(d2.t(), x2) : at::matmul(x2.t(), d2)};
}
static Ts bias_grad_compute(const OpDef&, const Ts& in, RunCtx*) {
  const at::Tensor& dy = in[0];
  if (dy.is_meta()) return {at::empty({dy.size(-1)}, dy.options())};
  at::Tensor d2 = flatten_rows(dy).contiguous();
  if (is_native(d2)) {
    at::Tensor acc = at::empty({d2.size(1)}, d2.options().dtype(at::kFloat));
    cuda_ok(colsum_bf16(d2.data_ptr(), acc.data_ptr<float>(), d2.size(0), (int)d2.size(1), false, cur_stream()), "colsum");
    return {acc.to(dy.scalar_type())};
  }
  return {d2.sum(0)};
}

static void linear_deduce(OpDef& op, size_t s) {
  const Tensor& x = op.inputs[0];
  const Tensor& w = op.inputs[1];
  if (!x->has_ds(s) || !w->has_ds(s)) { deduce_states_like_input(op, s, 0); return; }
  const bool trans_b = op.attrs.b("trans_b", true);
  const int nd = x->ndim();
  std::vector<int> map(nd);
  for (int i = 0; i < nd; ++i) map[i] = i;
  DistributedStates out = matmul_ds(x->ds(s), nd, nd - 1, w->ds(s), trans_b ? 1 : 0, trans_b ? 0 : 1, nd - 1, map);
  for (size_t i = 0; i < op.outputs.size(); ++i) set_out_ds(op, i, s, out);
}
static void dgrad_deduce(OpDef& op, size_t s) {
  const Tensor& dy = op.inputs[0];
  const Tensor& w = op.inputs[1];
  if (!dy->has_ds(s) || !w->has_ds(s)) { deduce_states_like_input(op, s, 0); return; }
  const bool trans_b = op.attrs.b("trans_b", true);
  const int nd = dy->ndim();
  std::vector<int> map(nd);
  for (int i = 0; i < nd; ++i) map[i] = i;
  set_out_ds(op, 0, s, matmul_ds(dy->ds(s), nd, nd - 1, w->ds(s), trans_b ? 0 : 1, trans_b ? 1 : 0, nd - 1, map));
}
static void wgrad_deduce(OpDef& op, size_t s) {
  const Tensor& dy = op.inputs[0];
  const Tensor& x = op.inputs[1];
  if (!dy->has_ds(s) || !x->has_ds(s)) return;
  const bool trans_b = op.attrs.b("trans_b", true);
  // collapse leading (token) dims: they are all contracted.  Treat operands as 2-D [T, feat].
  auto collapse = [](const DistributedStates& ds, int nd) {
    std::map<int, int> st;
    int tok = 1;
    for (auto& kv : ds.states()) {
      if (kv.second <= 1) continue;
      if (kv.first < 0) st[kv.first] = kv.second;
      else if (kv.first == nd - 1) st[1] = kv.second;
      else tok *= kv.second;
    }
    if (tok > 1) st[0] = tok;
    std::vector<int> order;
    for (int o : ds.order()) {
      int m = o < 0 ? o : (o == nd - 1 ? 1 : 0);
      if (std::find(order.begin(), order.end(), m) == order.end()) order.push_back(m);
    }
    return DistributedStates(ds.device_num(), st, order);
  };
  DistributedStates a = collapse(dy->ds(s), dy->ndim());  // [T, N]
  DistributedStates b = collapse(x->ds(s), x->ndim());    // [T, K]
  DistributedStates out = trans_b ? matmul_ds(a, 2, 0, b, 0, 1, 1, {0, 0})   // dW[N,K]: rows from dy dim1
                                  : matmul_ds(b, 2, 0, a, 0, 1, 1, {0, 0});
  set_out_ds(op, 0, s, out);
}
static void bias_grad_deduce(OpDef& op, size_t s) {
  const Tensor& dy = op.inputs[0];
  if (!dy->has_ds(s)) return;
  const DistributedStates& ds = dy->ds(s);
  const int nd = dy->ndim();
  std::map<int, int> st;
  int partial = ds.get_dim(kPartialDim);
  for (auto& kv : ds.states()) {
    if (kv.second <= 1) continue;
    if (kv.first == kDupDim) st[kDupDim] = kv.second;
    else if (kv.first == nd - 1) st[0] = kv.second;
    else if (kv.first >= 0) partial *= kv.second;
  }
  if (partial > 1) st[kPartialDim] = partial;
  std::vector<int> order;
  for (int o : ds.order()) {
    int m = (o == nd - 1) ? 0 : (o >= 0 ? kPartialDim : o);
    if (std::find(order.begin(), order.end(), m) == order.end()) order.push_back(m);
  }
  set_out_ds(op, 0, s, DistributedStates(ds.device_num(), st, order));
}

static TensorList linear_grad(OpDef& op, const TensorList& g) {
  Graph* gr = op.graph;
  const bool has_bias = op.attrs.b("has_bias"), has_res = op.attrs.b("has_residual");
  const std::string act = op.attrs.s("act");
  Tensor dy = g[0];
  HB_CHECK(dy != nullptr) << "linear without an output gradient";
  TensorList res(op.inputs.size());
  if (has_res) res[has_bias ? 3 : 2] = dy;
  Tensor dpre = dy;
  if (!act.empty() && act != "none") {
    OpDef* p = dy->producer;
    if (p != nullptr && p->type == "linear_dgrad" && p->inputs.size() == 2 && dy->consumers.empty()) {
      // dy is the dgrad of the next linear and nobody else reads it: recompute it with act'(pre) applied in the GEMM
      // epilogue (AUX_DGELU...) -- the un-fused dgrad op becomes dead and is never scheduled
      AttrMap a = p->attrs;
      a.set("act_bwd", act);
      dpre = gr->make_op1("linear_dgrad", {p->inputs[0], p->inputs[1], op.outputs[1]}, a);
    } else {
      AttrMap a;
      a.set("kind", act);
      dpre = gr->make_op1("unary_act_bwd", {dy, op.outputs[1]}, a);
    }
  }
  AttrMap a;
  a.set("trans_b", op.attrs.b("trans_b", true));
  if (op.inputs[0]->requires_grad) res[0] = gr->make_op1("linear_dgrad", {dpre, op.inputs[1]}, a);
  if (op.inputs[1]->requires_grad) res[1] = gr->make_op1("linear_wgrad", {dpre, op.inputs[0]}, a);
  if (has_bias && op.inputs[2]->requires_grad) res[2] = gr->make_op1("bias_grad", {dpre});
  return res;
}
static void linear_infer(OpDef& op) { infer_meta_by_meta_exec(op); }
HB_REGISTER_OP(linear, "linear", -1, 0, linear_compute, linear_grad, linear_deduce, linear_infer);
HB_REGISTER_OP(linear_dgrad, "linear_dgrad", 1, 0, linear_dgrad_compute, nullptr, dgrad_deduce, nullptr);
HB_REGISTER_OP(linear_wgrad, "linear_wgrad", 1, 0, linear_wgrad_compute, nullptr, wgrad_deduce, nullptr);
HB_REGISTER_OP(bias_grad, "bias_grad", 1, 0, bias_grad_compute, nullptr, bias_grad_deduce, nullptr);

// ------------------------------------------------------------------ fp8 linear (block-scaled e4m3, fp32 accumulate)
// y = act(x W^T + b) with x and W quantised per 1 x K block (one fp32 scale per row of the K-major operand); the scales
// factor out of the contraction and are applied in the tcgen05 (kind::f8f6f4) GEMM epilogue.  Backward: the input
// gradient runs in fp8 as well (dy rows x W^T rows), the weight gradient stays in bf16.
static void run_gemm_fp8(const at::Tensor& qa, const at::Tensor& sa, const at::Tensor& qb, const at::Tensor& sb, at::Tensor& C,
                         int64_t M, int64_t N, int64_t K, const at::Tensor* bias, int act, at::Tensor* aux_out, const at::Tensor* aux_in,
                         int aux_mode) {
  GemmCall c;
  c.A = qa.data_ptr(); c.B = qb.data_ptr(); c.C = C.data_ptr();
  c.M = (int)M; c.N = (int)N; c.K = (int)K;
  c.lda = qa.stride(0); c.ldb = qb.stride(0); c.ldc = C.stride(0);
  c.fp8 = true; c.row_scale = sa.data_ptr<float>(); c.col_scale = sb.data_ptr<float>();
  c.out = C.scalar_type() == at::kFloat ? GemmOut::FP32 : GemmOut::BF16;
  if (bias) c.bias = bias->data_ptr();
  if (aux_in) { c.aux_in = aux_in->data_ptr(); c.aux_mode = aux_mode; c.ld_aux = aux_in->stride(0); }
  if (aux_out) { c.aux_out = aux_out->data_ptr(); c.ld_aux = aux_out->stride(0); }
  c.act = act;
  cuda_ok(gemm_bf16(c, cur_stream()), "tcgen05 fp8 gemm");
}
static std::pair<at::Tensor, at::Tensor> quant_rows(const at::Tensor& x2) {
  at::Tensor q = at::empty(x2.sizes(), x2.options().dtype(at::kByte));
  at::Tensor s = at::empty({x2.size(0)}, x2.options().dtype(at::kFloat));
  cuda_ok(quantize_rowwise_e4m3(x2.data_ptr(), q.data_ptr(), s.data_ptr<float>(), x2.size(0), (int)x2.size(1), x2.stride(0), q.stride(0),
                                cur_stream()), "quantize_rowwise_e4m3");
  return {q, s};
}
// emulation of the same numerics with ATen (CPU tests / reference): round operands through e4m3 with row scales
static at::Tensor fake_quant_rows(const at::Tensor& x2) {
  at::Tensor xf = x2.to(at::kFloat);
  at::Tensor sc = (xf.abs().amax(-1, true) / 448.0).clamp_min(1e-30);
  at::Tensor y = xf / sc;
  if (at::hasCUDA() || true) y = y.to(at::kFloat8_e4m3fn).to(at::kFloat);
  return y * sc;
}
static Ts linear_fp8_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const bool has_bias = op.attrs.b("has_bias");
  const int act = act_code(op.attrs.s("act"));
  const at::Tensor& x = in[0];
  const at::Tensor& w = in[1];
  const at::Tensor* bias = has_bias ? &in[2] : nullptr;
  const int64_t N = w.size(0), K = x.size(-1);
  std::vector<int64_t> oshape = x.sizes().vec();
  oshape.back() = N;
  if (x.is_meta()) {
    Ts out = {at::empty(oshape, x.options())};
    if (act != ACT_NONE) out.push_back(at::empty(oshape, x.options()));
    return out;
  }
  at::Tensor x2 = flatten_rows(x).contiguous();
  const int64_t M = x2.size(0);
  if (gemm_ok(x2) && gemm_ok(w) && K % 16 == 0 && N % 8 == 0) {
    auto qx = quant_rows(x2);
    auto qw = quant_rows(w);
    at::Tensor y = at::empty({M, N}, x.options()), pre;
    if (act != ACT_NONE) pre = at::empty({M, N}, x.options());
    run_gemm_fp8(qx.first, qx.second, qw.first, qw.second, y, M, N, K, bias, act, act != ACT_NONE ? &pre : nullptr, nullptr, 0);
    Ts out = {y.reshape(oshape)};
    if (act != ACT_NONE) out.push_back(pre.reshape(oshape));
    return out;
  }
  // shapes the e4m3 TMA maps cannot address (K not a multiple of 16 bytes): run the same product on the bf16 tensor cores
  if (gemm_ok(x2) && gemm_ok(w) && (K % 8) == 0 && (N % 8) == 0) return linear_compute(op, in, nullptr);
  if (is_native(x)) note_fallback("linear_fp8");
  at::Tensor y = at::matmul(fake_quant_rows(x2), fake_quant_rows(w).t());
  if (bias) y = y + bias->to(at::kFloat);
  Ts out;
  if (act != ACT_NONE) {
    at::Tensor pre = y.to(x.scalar_type());
    out = {aten_act(pre, act).reshape(oshape), pre.reshape(oshape)};
  } else out = {y.to(x.scalar_type()).reshape(oshape)};
  return out;
}
// dx = dy W  in fp8: dy [M, N] quantised per row, W^T [K, N] quantised per row of W^T (= per input feature)
static Ts linear_fp8_dgrad_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& dy = in[0];
  const at::Tensor& w = in[1];
  const int64_t N = w.size(0), K = w.size(1);
  std::vector<int64_t> oshape = dy.sizes().vec();
  oshape.back() = K;
  if (dy.is_meta()) return {at::empty(oshape, dy.options())};
  at::Tensor d2 = flatten_rows(dy).contiguous();
  const int64_t M = d2.size(0);
  if (gemm_ok(d2) && gemm_ok(w) && N % 16 == 0 && K % 8 == 0) {
    auto qd = quant_rows(d2);
    at::Tensor qwt = at::empty({K, N}, w.options().dtype(at::kByte));
    at::Tensor swt = at::empty({K}, w.options().dtype(at::kFloat));
    cuda_ok(quantize_transpose_e4m3(w.data_ptr(), qwt.data_ptr(), swt.data_ptr<float>(), N, (int)K, N, cur_stream()),
            "quantize_transpose_e4m3");
    at::Tensor dx = at::empty({M, K}, dy.options());
    if (in.size() > 2) {
      at::Tensor pre = flatten_rows(in[2]).contiguous();
      run_gemm_fp8(qd.first, qd.second, qwt, swt, dx, M, K, N, nullptr, ACT_NONE, nullptr, &pre, act_bwd_mode(op.attrs.s("act_bwd")));
    } else run_gemm_fp8(qd.first, qd.second, qwt, swt, dx, M, K, N, nullptr, ACT_NONE, nullptr, nullptr, 0);
    return {dx.reshape(oshape)};
  }
  if (gemm_ok(d2) && gemm_ok(w) && (K % 8) == 0) return linear_dgrad_compute(op, in, nullptr);     // bf16 tensor-core fallback
  if (is_native(dy)) note_fallback("linear_fp8_dgrad");
  at::Tensor dxa = at::matmul(fake_quant_rows(d2), fake_quant_rows(w.t().contiguous()).t()).to(dy.scalar_type()).reshape(oshape);
  if (in.size() > 2) dxa = aten_act_bwd(dxa, in[2], op.attrs.s("act_bwd"));
  return {dxa};
}
static TensorList linear_fp8_grad(OpDef& op, const TensorList& g) {
  Graph* gr = op.graph;
  const bool has_bias = op.attrs.b("has_bias");
  const std::string act = op.attrs.s("act");
  Tensor dy = g[0];
  HB_CHECK(dy != nullptr) << "linear_fp8 without an output gradient";
  TensorList res(op.inputs.size());
  Tensor dpre = dy;
  if (!act.empty() && act != "none") {
    OpDef* p = dy->producer;
    if (p != nullptr && (p->type == "linear_fp8_dgrad" || p->type == "linear_dgrad") && p->inputs.size() == 2 && dy->consumers.empty()) {
      AttrMap a = p->attrs;
      a.set("act_bwd", act);
      dpre = gr->make_op1(p->type, {p->inputs[0], p->inputs[1], op.outputs[1]}, a);
    } else {
      AttrMap a;
      a.set("kind", act);
      dpre = gr->make_op1("unary_act_bwd", {dy, op.outputs[1]}, a);
    }
  }
  AttrMap a;
  a.set("trans_b", true);
  if (op.inputs[0]->requires_grad) res[0] = gr->make_op1("linear_fp8_dgrad", {dpre, op.inputs[1]}, a);
  if (op.inputs[1]->requires_grad) res[1] = gr->make_op1("linear_wgrad", {dpre, op.inputs[0]}, a);
  if (has_bias && op.inputs[2]->requires_grad) res[2] = gr->make_op1("bias_grad", {dpre});
  return res;
}
HB_REGISTER_OP(linear_fp8, "linear_fp8", -1, 0, linear_fp8_compute, linear_fp8_grad, linear_deduce, linear_infer);
HB_REGISTER_OP(linear_fp8_dgrad, "linear_fp8_dgrad", 1, 0, linear_fp8_dgrad_compute, nullptr, dgrad_deduce, nullptr);

// ------------------------------------------------------------------ scaled masked softmax (Megatron fused softmax parity)
// x [B, H, Sq, Sk]; attrs: scale, causal; optional input 1: boolean mask [B, 1, Sq, Sk] (true = masked out)
static Ts scaled_masked_softmax_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& x = in[0];
  const double scale = op.attrs.f("scale", 1.0);
  const bool causal = op.attrs.b("causal");
  if (x.is_meta()) return {at::empty_like(x)};
  const int64_t Sk = x.size(-1), Sq = x.dim() >= 2 ? x.size(-2) : 1;
  if (is_native(x) && x.is_contiguous() && x.dim() == 4) {
    at::Tensor y = at::empty_like(x), m;
    if (in.size() > 1) m = in[1].to(at::kByte).expand({x.size(0), 1, Sq, Sk}).contiguous();
    cuda_ok(scaled_softmax_fwd(x.data_ptr(), m.defined() ? m.data_ptr<uint8_t>() : nullptr, y.data_ptr(), x.numel() / Sk, (int)Sk,
                               (int)(x.size(1) * Sq), (int)Sq, (float)scale, causal ? 2 : (m.defined() ? 1 : 0), cur_stream()),
            "scaled_softmax_fwd");
    return {y};
  }
  at::Tensor s = x.to(at::kFloat) * scale;
  if (in.size() > 1) s = s.masked_fill(in[1].to(at::kBool), -INFINITY);
  if (causal) s = s.masked_fill(at::ones({Sq, Sk}, s.options().dtype(at::kBool)).tril(Sk - Sq).logical_not(), -INFINITY);
  at::Tensor y = at::softmax(s, -1);
  y = at::where(at::isnan(y), at::zeros_like(y), y);
  return {y.to(x.scalar_type())};
}
static Ts scaled_masked_softmax_bwd_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& dy = in[0];
  const at::Tensor& y = in[1];
  const double scale = op.attrs.f("scale", 1.0);
  if (dy.is_meta()) return {at::empty_like(y)};
  if (is_native(y) && is_native(dy) && y.is_contiguous() && dy.is_contiguous()) {
    at::Tensor dx = at::empty_like(y);
    cuda_ok(scaled_softmax_bwd(dy.data_ptr(), y.data_ptr(), dx.data_ptr(), y.numel() / y.size(-1), (int)y.size(-1), (float)scale, cur_stream()),
            "scaled_softmax_bwd");
    return {dx};
  }
  at::Tensor yf = y.to(at::kFloat), df = dy.to(at::kFloat);
  return {(scale * yf * (df - (df * yf).sum(-1, true))).to(y.scalar_type())};
}
static TensorList scaled_masked_softmax_grad(OpDef& op, const TensorList& g) {
  TensorList r(op.inputs.size());
  r[0] = op.graph->make_op1("scaled_masked_softmax_bwd", {g[0], op.outputs[0]}, op.attrs);
  return r;
}
HB_REGISTER_OP(scaled_masked_softmax, "scaled_masked_softmax", 1, 0, scaled_masked_softmax_compute, scaled_masked_softmax_grad, nullptr, nullptr);
HB_REGISTER_OP(scaled_masked_softmax_bwd, "scaled_masked_softmax_bwd", 1, kFlagNondiff, scaled_masked_softmax_bwd_compute, nullptr, nullptr,
               nullptr);

// matmul(a, b, trans_a, trans_b): general 2-D matmul (autograd VJP for the long tail)
static Ts matmul_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const bool ta = op.attrs.b("trans_a"), tb = op.attrs.b("trans_b");
  const at::Tensor& a = in[0];
  const at::Tensor& b = in[1];
  if (!a.is_meta() && gemm_ok(a) && gemm_ok(b) && a.dim() == 2 && b.dim() == 2) {
    const int64_t M = ta ? a.size(1) : a.size(0), K = ta ? a.size(0) : a.size(1), N = tb ? b.size(0) : b.size(1);
    if (M % 8 == 0 && N % 8 == 0 && K % 8 == 0) {
      at::Tensor c = at::empty({M, N}, a.options());
      run_gemm(a, ta, b, !tb, c, M, N, K, nullptr, nullptr, 0, nullptr, ACT_NONE, false);
      return {c};
    }
  }
  return {at::matmul(ta ? a.transpose(-1, -2) : a, tb ? b.transpose(-1, -2) : b)};
}
static void matmul_deduce(OpDef& op, size_t s) {
  const Tensor& a = op.inputs[0];
  const Tensor& b = op.inputs[1];
  if (!a->has_ds(s) || !b->has_ds(s) || a->ndim() != 2 || b->ndim() != 2) { deduce_states_like_input(op, s, 0); return; }
  const bool ta = op.attrs.b("trans_a"), tb = op.attrs.b("trans_b");
  std::vector<int> map = ta ? std::vector<int>{0, 0} : std::vector<int>{0, 1};
  set_out_ds(op, 0, s, matmul_ds(a->ds(s), 2, ta ? 0 : 1, b->ds(s), tb ? 1 : 0, tb ? 0 : 1, 1, map));
}
HB_REGISTER_OP(matmul, "matmul", 1, 0, matmul_compute, nullptr, matmul_deduce, nullptr);

// ------------------------------------------------------------------ unary activations on the hot path
static int unary_code(const std::string& k) {
  if (k == "gelu") return U_GELU;
  if (k == "relu") return U_RELU;
  if (k == "silu") return U_SILU;
  if (k == "gelu_tanh") return U_GELU_TANH;
  HB_FAIL() << "unknown activation kind " << k;
}
static Ts unary_act_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const int code = unary_code(op.attrs.s("kind"));
  const at::Tensor& x = in[0];
  if (!x.is_meta() && is_native(x) && x.is_contiguous()) {
    at::Tensor y = at::empty_like(x);
    cuda_ok(unary_fwd(code, x.data_ptr(), y.data_ptr(), x.numel(), cur_stream()), "unary_fwd");
    return {y};
  }
  switch (code) {
    case U_GELU: return {at::gelu(x)};
    case U_RELU: return {at::relu(x)};
    case U_SILU: return {at::silu(x)};
    default: return {at::gelu(x, "tanh")};
  }
}
static Ts unary_act_bwd_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const int code = unary_code(op.attrs.s("kind"));
  const at::Tensor& dy = in[0];
  const at::Tensor& x = in[1];
  if (dy.is_meta()) return {at::empty_like(x)};
  if (is_native(x) && is_native(dy) && x.is_contiguous() && dy.is_contiguous()) {
    at::Tensor dx = at::empty_like(x);
    cuda_ok(unary_bwd(code, dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), cur_stream()), "unary_bwd");
    return {dx};
  }
  return {aten_act_bwd(dy, x, op.attrs.s("kind"))};
}
static TensorList unary_act_grad(OpDef& op, const TensorList& g) {
  AttrMap a;
  a.set("kind", op.attrs.s("kind"));
  return {op.graph->make_op1("unary_act_bwd", {g[0], op.inputs[0]}, a)};
}
HB_REGISTER_OP(unary_act, "unary_act", 1, 0, unary_act_compute, unary_act_grad, nullptr, nullptr);
HB_REGISTER_OP(unary_act_bwd, "unary_act_bwd", 1, 0, unary_act_bwd_compute, nullptr, nullptr, nullptr);

// swiglu: y = silu(x[..., :d]) * x[..., d:]
// attr interleaved: x[..., 2i] = gate_i, x[..., 2i+1] = up_i (the tensor-parallel friendly layout of the fused gate/up GEMM)
static Ts swiglu_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& x = in[0];
  const bool il = op.attrs.b("interleaved");
  const int64_t d = x.size(-1) / 2;
  std::vector<int64_t> oshape = x.sizes().vec();
  oshape.back() = d;
  if (x.is_meta()) return {at::empty(oshape, x.options())};
  if (is_native(x) && x.is_contiguous() && d % 8 == 0) {
    at::Tensor y = at::empty(oshape, x.options());
    if (il) cuda_ok(swiglu_interleaved_fwd(x.data_ptr(), y.data_ptr(), x.numel() / (2 * d), (int)d, cur_stream()), "swiglu_fwd");
    else cuda_ok(swiglu_fwd(x.data_ptr(), y.data_ptr(), x.numel() / (2 * d), (int)d, cur_stream()), "swiglu_fwd");
    return {y};
  }
  if (il) {
    std::vector<int64_t> ps = oshape;
    ps.push_back(2);
    at::Tensor xp = x.reshape(ps);
    return {at::silu(xp.select(-1, 0)) * xp.select(-1, 1)};
  }
  return {at::silu(x.narrow(-1, 0, d)) * x.narrow(-1, d, d)};
}
static Ts swiglu_bwd_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& dy = in[0];
  const at::Tensor& x = in[1];
  const bool il = op.attrs.b("interleaved");
  if (dy.is_meta()) return {at::empty_like(x)};
  const int64_t d = x.size(-1) / 2;
  if (is_native(x) && is_native(dy) && x.is_contiguous() && dy.is_contiguous() && d % 8 == 0) {
    at::Tensor dx = at::empty_like(x);
    if (il) cuda_ok(swiglu_interleaved_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel() / (2 * d), (int)d, cur_stream()), "swiglu_bwd");
    else cuda_ok(swiglu_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel() / (2 * d), (int)d, cur_stream()), "swiglu_bwd");
    return {dx};
  }
  at::Tensor a, b;
  if (il) {
    std::vector<int64_t> ps = dy.sizes().vec();
    ps.push_back(2);
    at::Tensor xp = x.reshape(ps).to(at::kFloat);
    a = xp.select(-1, 0);
    b = xp.select(-1, 1);
  } else {
    a = x.narrow(-1, 0, d).to(at::kFloat);
    b = x.narrow(-1, d, d).to(at::kFloat);
  }
  auto g = dy.to(at::kFloat);
  auto sg = at::sigmoid(a);
  auto da = g * b * sg * (1 + a * (1 - sg));
  auto db = g * a * sg;
  if (il) return {at::stack({da, db}, -1).reshape(x.sizes()).to(x.scalar_type())};
  return {at::cat({da, db}, -1).to(x.scalar_type())};
}
static TensorList swiglu_grad(OpDef& op, const TensorList& g) {
  return {op.graph->make_op1("swiglu_bwd", {g[0], op.inputs[0]}, op.attrs)};
}
HB_REGISTER_OP(swiglu, "swiglu", 1, 0, swiglu_compute, swiglu_grad, nullptr, nullptr);
HB_REGISTER_OP(swiglu_bwd, "swiglu_bwd", 1, 0, swiglu_bwd_compute, nullptr,
               [](OpDef& op, size_t s) { copy_out_ds(op, 0, s, op.inputs[1]); }, nullptr);

// ------------------------------------------------------------------ layer / rms norm
// outputs: y, mean (layernorm only), rstd
static Ts norm_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const bool rms = op.attrs.b("rms");
  const double eps = op.attrs.f("eps", 1e-5);
  const at::Tensor& x = in[0];
  const at::Tensor& gamma = in[1];
  const int64_t cols = x.size(-1);
  std::vector<int64_t> sshape(x.sizes().begin(), x.sizes().end() - 1);
  auto fopt = x.options().dtype(at::kFloat);
  if (x.is_meta()) return {at::empty_like(x), at::empty(sshape, fopt), at::empty(sshape, fopt)};
  const int64_t rows = x.numel() / cols;
  if (is_native(x) && x.is_contiguous() && is_native(gamma)) {
    at::Tensor y = at::empty_like(x), mean = at::empty(sshape, fopt), rstd = at::empty(sshape, fopt);
    if (rms) {
      cuda_ok(rmsnorm_fwd(x.data_ptr(), gamma.data_ptr(), y.data_ptr(), rstd.data_ptr<float>(), rows, (int)cols,
                          (float)eps, cur_stream()), "rmsnorm_fwd");
    } else {
      const void* beta = in.size() > 2 ? in[2].data_ptr() : nullptr;
      cuda_ok(layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta, y.data_ptr(), mean.data_ptr<float>(),
                            rstd.data_ptr<float>(), rows, (int)cols, (float)eps, cur_stream()), "layernorm_fwd");
    }
    return {y, mean, rstd};
  }
  at::Tensor xf = x.to(at::kFloat);
  at::Tensor mean, rstd, y;
  if (rms) {
    mean = at::zeros(sshape, fopt);
    rstd = at::rsqrt(xf.pow(2).mean(-1) + eps);
    y = xf * rstd.unsqueeze(-1) * gamma.to(at::kFloat);
  } else {
    mean = xf.mean(-1);
    at::Tensor var = (xf - mean.unsqueeze(-1)).pow(2).mean(-1);
    rstd = at::rsqrt(var + eps);
    y = (xf - mean.unsqueeze(-1)) * rstd.unsqueeze(-1) * gamma.to(at::kFloat);
    if (in.size() > 2) y = y + in[2].to(at::kFloat);
  }
  return {y.to(x.scalar_type()), mean, rstd};
}
// inputs: dy, x, gamma, mean, rstd, [dx_add] -> dx (+ dx_add), dgamma, (dbeta)
// dx_add is the gradient that reaches the norm's input through the residual stream; Graph::gradients folds that addition
// in here (see sum_grads) so the backward pass has no separate elementwise add per norm
static Ts norm_bwd_compute(const OpDef& op, const Ts& in, RunCtx* rc) {
  const bool rms = op.attrs.b("rms");
  const at::Tensor* dx_add = in.size() > 5 ? &in[5] : nullptr;
  const at::Tensor& dy = in[0];
  const at::Tensor& x = in[1];
  const at::Tensor& gamma = in[2];
  const at::Tensor& mean = in[3];
  const at::Tensor& rstd = in[4];
  const int64_t cols = x.size(-1);
  if (dy.is_meta()) {
    Ts r = {at::empty_like(x), at::empty_like(gamma)};
    if (!rms) r.push_back(at::empty_like(gamma));
    return r;
  }
  const int64_t rows = x.numel() / cols;
  if (is_native(x) && is_native(dy) && x.is_contiguous() && dy.is_contiguous() &&
      (!dx_add || (is_native(*dx_add) && dx_add->is_contiguous() && dx_add->sizes() == x.sizes()))) {
    auto fopt = x.options().dtype(at::kFloat);
    const void* addp = dx_add ? dx_add->data_ptr() : nullptr;
    // parameter gradients come out in the parameter's dtype straight from the fold kernel (no cast pass)
    const bool pbf = gamma.scalar_type() == at::kBFloat16;
    at::Tensor dx = at::empty_like(x), dg = at::empty({cols}, pbf ? gamma.options() : fopt), db = at::empty({cols}, pbf ? gamma.options() : fopt);
    at::Tensor ws = rc && rc->workspace ? rc->scratch("norm_bwd_ws", {2 * (int64_t)ln_bwd_parts() * cols}, at::kFloat, x.device())
                                        : at::empty({2 * (int64_t)ln_bwd_parts() * cols}, fopt);
    if (rms) {
      cuda_ok(rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), rstd.data_ptr<float>(), dx.data_ptr(),
                          pbf ? nullptr : dg.data_ptr<float>(), ws.data_ptr<float>(), rows, (int)cols, false, cur_stream(), addp,
                          pbf ? dg.data_ptr() : nullptr), "rmsnorm_bwd");
      return {dx, pbf ? dg : dg.to(gamma.scalar_type())};
    }
    cuda_ok(layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr<float>(), rstd.data_ptr<float>(),
                          dx.data_ptr(), pbf ? nullptr : dg.data_ptr<float>(), pbf ? nullptr : db.data_ptr<float>(), ws.data_ptr<float>(), rows,
                          (int)cols, false, cur_stream(), addp, pbf ? dg.data_ptr() : nullptr, pbf ? db.data_ptr() : nullptr), "layernorm_bwd");
    return {dx, pbf ? dg : dg.to(gamma.scalar_type()), pbf ? db : db.to(gamma.scalar_type())};
  }
  at::Tensor xf = x.to(at::kFloat), g = dy.to(at::kFloat), gm = gamma.to(at::kFloat);
  at::Tensor xhat = rms ? xf * rstd.unsqueeze(-1) : (xf - mean.unsqueeze(-1)) * rstd.unsqueeze(-1);
  at::Tensor dyg = g * gm;
  at::Tensor m2 = (dyg * xhat).mean(-1, true);
  at::Tensor dx = rms ? rstd.unsqueeze(-1) * (dyg - xhat * m2) : rstd.unsqueeze(-1) * (dyg - dyg.mean(-1, true) - xhat * m2);
  at::Tensor dgm = (g * xhat).reshape({-1, cols}).sum(0);
  if (dx_add) dx = dx + dx_add->to(at::kFloat);
  Ts r = {dx.to(x.scalar_type()), dgm.to(gamma.scalar_type())};
  if (!rms) r.push_back(g.reshape({-1, cols}).sum(0).to(gamma.scalar_type()));
  return r;
}
static void norm_infer(OpDef& op) { infer_meta_by_meta_exec(op); }
static TensorList norm_grad(OpDef& op, const TensorList& g) {
  AttrMap a;
  a.set("rms", op.attrs.b("rms"));
  TensorList outs = op.graph->make_op("norm_bwd", {g[0], op.inputs[0], op.inputs[1], op.outputs[1], op.outputs[2]}, a);
  TensorList r(op.inputs.size());
  r[0] = outs[0];
  r[1] = outs[1];
  if (op.inputs.size() > 2 && outs.size() > 2) r[2] = outs[2];
  return r;
}
static void norm_bwd_deduce(OpDef& op, size_t s) {
  copy_out_ds(op, 0, s, op.inputs[1]);
  // dgamma / dbeta: token dims are reduced -> partial over every split of x
  const Tensor& x = op.inputs[1];
  if (!x->has_ds(s)) return;
  const DistributedStates& ds = x->ds(s);
  int partial = ds.get_dim(kPartialDim);
  std::map<int, int> st;
  for (auto& kv : ds.states()) {
    if (kv.second <= 1) continue;
    if (kv.first >= 0) partial *= kv.second;
    else if (kv.first == kDupDim) st[kDupDim] = kv.second;
  }
  if (partial > 1) st[kPartialDim] = partial;
  std::vector<int> order;
  for (int o : ds.order()) {
    int m = o >= 0 ? kPartialDim : o;
    if (std::find(order.begin(), order.end(), m) == order.end()) order.push_back(m);
  }
  DistributedStates gds(ds.device_num(), st, order);
  for (size_t i = 1; i < op.outputs.size(); ++i) set_out_ds(op, i, s, gds);
}
HB_REGISTER_OP(norm_op, "fused_norm", 3, 0, norm_compute, norm_grad, nullptr, norm_infer);
HB_REGISTER_OP(norm_bwd, "norm_bwd", -1, 0, norm_bwd_compute, nullptr, norm_bwd_deduce, nullptr);

// ------------------------------------------------------------------ embedding
// inputs: table [V, H], ids [...]   (ids outside [0, V) produce zeros: vocab-parallel shards rely on it)
// vocab-parallel shards: first row of this rank's shard, per strategy (hot switching changes the tensor-parallel degree)
static int64_t vocab_offset_of(const OpDef& op) {
  const std::vector<int64_t> per = op.attrs.ints("vocab_offsets");
  const int s = op.graph ? op.graph->cur_strategy() : 0;
  if (!per.empty() && s >= 0 && (size_t)s < per.size()) return per[s];
  return op.attrs.i("vocab_offset", 0);
}
static Ts embedding_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& table = in[0];
  const at::Tensor& ids = in[1];
  std::vector<int64_t> oshape = ids.sizes().vec();
  oshape.push_back(table.size(1));
  if (table.is_meta()) return {at::empty(oshape, table.options())};
  const int64_t offset = vocab_offset_of(op);
  at::Tensor idl = ids.to(at::kLong).contiguous();
  if (offset != 0) idl = idl - offset;
  if (is_native(table) && table.is_contiguous() && table.size(1) % 8 == 0) {
    at::Tensor y = at::empty(oshape, table.options());
    cuda_ok(embedding_fwd(idl.data_ptr<int64_t>(), nullptr, table.data_ptr(), nullptr, y.data_ptr(), idl.numel(),
                          (int)table.size(1), table.size(0), cur_stream()), "embedding_fwd");
    return {y};
  }
  at::Tensor valid = (idl >= 0).logical_and(idl < table.size(0));
  at::Tensor safe = at::where(valid, idl, at::zeros_like(idl));
  at::Tensor y = at::embedding(table, safe) * valid.unsqueeze(-1).to(table.scalar_type());
  return {y};
}
static Ts embedding_grad_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& dy = in[0];
  const at::Tensor& ids = in[1];
  // the table (input 2) gives the local vocabulary size of the CURRENT strategy; the attribute is the build-time value
  const int64_t V = in.size() > 2 ? in[2].size(0) : op.attrs.i("vocab"), H = dy.size(-1);
  if (dy.is_meta()) return {at::empty({V, H}, dy.options())};
  const int64_t offset = vocab_offset_of(op);
  at::Tensor idl = ids.to(at::kLong).contiguous();
  if (offset != 0) idl = idl - offset;
  at::Tensor d2 = dy.reshape({-1, H}).contiguous();
  if (is_native(d2) && H % 8 == 0) {
    at::Tensor acc = at::zeros({V, H}, dy.options().dtype(at::kFloat));
    cuda_ok(embedding_bwd(idl.data_ptr<int64_t>(), nullptr, d2.data_ptr(), acc.data_ptr<float>(), nullptr, idl.numel(),
                          (int)H, V, cur_stream()), "embedding_bwd");
    return {acc.to(dy.scalar_type())};
  }
  at::Tensor flat = idl.reshape({-1});
  at::Tensor valid = (flat >= 0).logical_and(flat < V);
  at::Tensor safe = at::where(valid, flat, at::zeros_like(flat));
  at::Tensor acc = at::zeros({V, H}, dy.options().dtype(at::kFloat));
  acc.index_add_(0, safe, d2.to(at::kFloat) * valid.unsqueeze(-1).to(at::kFloat));
  return {acc.to(dy.scalar_type())};
}
static TensorList embedding_grad(OpDef& op, const TensorList& g) {
  AttrMap a;
  a.set("vocab", op.inputs[0]->shape[0]);
  a.set("vocab_offset", op.attrs.i("vocab_offset", 0));
  a.set("vocab_offsets", op.attrs.ints("vocab_offsets"));
  return {op.graph->make_op1("embedding_grad", {g[0], op.inputs[1], op.inputs[0]}, a), nullptr};
}
static void embedding_deduce(OpDef& op, size_t s) {
  const Tensor& table = op.inputs[0];
  const Tensor& ids = op.inputs[1];
  if (!ids->has_ds(s)) return;
  const DistributedStates& ids_ds = ids->ds(s);
  if (!table->has_ds(s) || table->ds(s).get_dim(0) <= 1) { copy_out_ds(op, 0, s, ids); return; }
  // vocab-parallel table: the replicas of ids along the table's vocab split hold partial sums
  const int vsplit = table->ds(s).get_dim(0);
  std::map<int, int> st;
  for (auto& kv : ids_ds.states()) if (kv.second > 1) st[kv.first] = kv.second;
  HB_CHECK(st.count(kDupDim) && st[kDupDim] % vsplit == 0) << "vocab-parallel embedding needs ids duplicated over the vocab split";
  st[kDupDim] /= vsplit;
  st[kPartialDim] = (st.count(kPartialDim) ? st[kPartialDim] : 1) * vsplit;
  std::vector<int> order;
  for (int o : ids_ds.order()) {
    if (o == kDupDim) { order.push_back(kPartialDim); if (st[kDupDim] > 1) order.push_back(kDupDim); }
    else order.push_back(o);
  }
  set_out_ds(op, 0, s, DistributedStates(ids_ds.device_num(), st, order));
}
static void embedding_grad_deduce(OpDef& op, size_t s) {
  // dense table gradient [V_local, H]: keeps the table's own (vocab) split; the table's replicas that saw different
  // tokens (token-split of dy) hold partial sums
  const Tensor& dy = op.inputs[0];
  if (!dy->has_ds(s)) return;
  const DistributedStates& ds = dy->ds(s);
  const int nd = dy->ndim();
  int token_split = 1;
  for (auto& kv : ds.states()) if (kv.first >= 0 && kv.first != nd - 1 && kv.second > 1) token_split *= kv.second;
  if (op.inputs.size() > 2 && op.inputs[2]->has_ds(s)) {
    const DistributedStates& tds = op.inputs[2]->ds(s);
    std::map<int, int> st;
    for (auto& kv : tds.states()) if (kv.second > 1) st[kv.first] = kv.second;
    std::vector<int> order = tds.order();
    const int dup = tds.get_dim(kDupDim);
    if (token_split > 1 && dup > 1 && dup % token_split == 0) {
      st[kPartialDim] = (st.count(kPartialDim) ? st[kPartialDim] : 1) * token_split;
      if (dup / token_split > 1) st[kDupDim] = dup / token_split; else st.erase(kDupDim);
      std::vector<int> no;
      for (int o : order) {
        if (o == kDupDim) { no.push_back(kPartialDim); if (st.count(kDupDim)) no.push_back(kDupDim); }
        else no.push_back(o);
      }
      order = no;
    }
    set_out_ds(op, 0, s, DistributedStates(tds.device_num(), st, order));
    return;
  }
  int partial = ds.get_dim(kPartialDim) * token_split;
  std::map<int, int> st;
  for (auto& kv : ds.states()) {
    if (kv.second <= 1) continue;
    if (kv.first == kDupDim) st[kDupDim] = kv.second;
    else if (kv.first == nd - 1) st[1] = kv.second;
  }
  if (partial > 1) st[kPartialDim] = partial;
  std::vector<int> order;
  for (int o : ds.order()) {
    int m = (o == nd - 1) ? 1 : (o >= 0 ? kPartialDim : o);
    if (std::find(order.begin(), order.end(), m) == order.end()) order.push_back(m);
  }
  set_out_ds(op, 0, s, DistributedStates(ds.device_num(), st, order));
}
HB_REGISTER_OP(embedding_lookup, "embedding_lookup", 1, 0, embedding_compute, embedding_grad, embedding_deduce, nullptr);
HB_REGISTER_OP(embedding_grad, "embedding_grad", 1, 0, embedding_grad_compute, nullptr, embedding_grad_deduce, nullptr);

// ------------------------------------------------------------------ sparse softmax cross entropy
// inputs: logits [..., V], labels [...]; outputs: loss (reduced per attrs), dlogits_unit = softmax - onehot (saved)
static Ts ce_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& logits = in[0];
  const at::Tensor& labels = in[1];
  const int64_t ignore = op.attrs.i("ignore_index", -1);
  const std::string red = op.attrs.s("reduction", "mean");
  const int64_t V = logits.size(-1);
  std::vector<int64_t> lshape(logits.sizes().begin(), logits.sizes().end() - 1);
  auto fopt = logits.options().dtype(at::kFloat);
  if (logits.is_meta()) {
    at::Tensor l = red == "none" ? at::empty(lshape, fopt) : at::empty({}, fopt);
    return {l, at::empty_like(logits)};
  }
  at::Tensor lab = labels.to(at::kLong).reshape({-1}).contiguous();
  const int64_t rows = lab.numel();
  at::Tensor per_tok, unit;
  at::Tensor inv_cnt;   // mean reduction: 1 / #valid tokens, folded into the saved gradient (native path)
  if (is_native(logits) && logits.is_contiguous() && V % 8 == 0) {
    // the kernel overwrites its input with (softmax - onehot) [/ #valid]: work on a copy only when the logits are
    // still needed elsewhere (the executor marks single-consumer inputs as donatable)
    unit = op.attrs.b("donate_logits") ? logits : logits.clone();
    per_tok = at::empty({rows}, fopt);
    if (red == "mean") inv_cnt = (lab != ignore).sum().to(at::kFloat).clamp_min(1.0).reciprocal();
    cuda_ok(softmax_ce_fwd_bwd(unit.data_ptr(), lab.data_ptr<int64_t>(), per_tok.data_ptr<float>(), nullptr, rows, (int)V, V,
                               ignore, 1.0f, true, cur_stream(), inv_cnt.defined() ? inv_cnt.data_ptr<float>() : nullptr),
            "softmax_ce");
  } else {
    at::Tensor lf = logits.to(at::kFloat).reshape({rows, V});
    at::Tensor lsm = at::log_softmax(lf, -1);
    at::Tensor valid = (lab != ignore);
    at::Tensor safe = at::where(valid, lab, at::zeros_like(lab));
    per_tok = -lsm.gather(1, safe.unsqueeze(1)).squeeze(1) * valid.to(at::kFloat);
    at::Tensor u = at::exp(lsm);
    u.scatter_add_(1, safe.unsqueeze(1), -at::ones({rows, 1}, fopt));
    unit = (u * valid.unsqueeze(1).to(at::kFloat)).to(logits.scalar_type()).reshape(logits.sizes());
  }
  at::Tensor loss;
  if (red == "none") loss = per_tok.reshape(lshape);
  else if (red == "sum") loss = per_tok.sum();
  else if (inv_cnt.defined()) loss = per_tok.sum() * inv_cnt;
  else {
    at::Tensor cnt = (lab != ignore).sum().to(at::kFloat).clamp_min(1.0);
    loss = per_tok.sum() / cnt;
  }
  return {loss, unit.reshape(logits.sizes())};
}
// inputs: dloss, unit, labels -> dlogits
static Ts ce_bwd_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& dl = in[0];
  const at::Tensor& unit = in[1];
  const at::Tensor& labels = in[2];
  if (unit.is_meta()) return {at::empty_like(unit)};
  const std::string red = op.attrs.s("reduction", "mean");
  const int64_t ignore = op.attrs.i("ignore_index", -1);
  const int64_t V = unit.size(-1);
  if (is_native(unit) && unit.is_contiguous() && V % 8 == 0) {
    // native forward already folded 1/#valid into `unit` (mean); a unit seed (dloss == 1 by construction) is free
    if (op.attrs.b("unit_dloss") && red != "none") return {unit};
    at::Tensor sc = dl.to(at::kFloat).contiguous();
    at::Tensor out = at::empty_like(unit);
    cuda_ok(scale_rows_bf16(unit.data_ptr(), out.data_ptr(), sc.data_ptr<float>(), red == "none", unit.numel() / V, V, cur_stream()),
            "ce_bwd scale");
    return {out};
  }
  at::Tensor scale;
  if (red == "none") scale = dl.to(at::kFloat).unsqueeze(-1);
  else if (red == "sum") scale = dl.to(at::kFloat);
  else scale = dl.to(at::kFloat) / (labels.to(at::kLong) != ignore).sum().to(at::kFloat).clamp_min(1.0);
  return {(unit.to(at::kFloat) * scale).to(unit.scalar_type())};
}
static TensorList ce_grad(OpDef& op, const TensorList& g) {
  AttrMap a;
  a.set("reduction", op.attrs.s("reduction", "mean"));
  a.set("ignore_index", op.attrs.i("ignore_index", -1));
  // loss seeded with ones_like(loss): the scale is the constant 1, the saved (softmax - onehot)/count IS the gradient
  if (g[0]->producer != nullptr && g[0]->producer->type == "ones_like") a.set("unit_dloss", true);
  return {op.graph->make_op1("softmax_ce_sparse_bwd", {g[0], op.outputs[1], op.inputs[1]}, a), nullptr};
}
static void ce_deduce(OpDef& op, size_t s) {
  const Tensor& lg = op.inputs[0];
  if (!lg->has_ds(s)) return;
  copy_out_ds(op, 1, s, lg);
  const std::string red = op.attrs.s("reduction", "mean");
  const DistributedStates& ds = lg->ds(s);
  const int nd = lg->ndim();
  std::map<int, int> st;
  std::vector<int> order;
  if (red == "none") {
    for (auto& kv : ds.states()) if (kv.second > 1 && kv.first != nd - 1) st[kv.first] = kv.second;
    for (int o : ds.order()) if (o != nd - 1) order.push_back(o);
  } else {
    // scalar loss: every token split holds the loss of its own tokens (kept as "duplicate" like the
    // reference: the trainer averages per-rank losses explicitly)
    st[kDupDim] = ds.device_num();
    order = {kDupDim};
  }
  set_out_ds(op, 0, s, DistributedStates(ds.device_num(), st, order));
}
HB_REGISTER_OP(softmax_ce_sparse, "softmax_cross_entropy_sparse", 2, 0, ce_compute, ce_grad, ce_deduce, nullptr);
HB_REGISTER_OP(softmax_ce_sparse_bwd, "softmax_ce_sparse_bwd", 1, 0, ce_bwd_compute, nullptr,
               [](OpDef& op, size_t s) { copy_out_ds(op, 0, s, op.inputs[1]); }, nullptr);

// ------------------------------------------------------------------ attention
// inputs: q [B,Sq,Hq,D], k [B,Sk,Hkv,D], v [B,Sk,Hkv,D]  (any strides with D contiguous) -> o, lse [B,Hq,Sq]
static AttnTensor as_attn(const at::Tensor& t) {
  AttnTensor a;
  a.ptr = t.data_ptr();
  a.stride_b = t.stride(0); a.stride_s = t.stride(1); a.stride_h = t.stride(2);
  if (t.dim() == 5) {   // grouped [B, S, G, rep, D]: slots of stride(3) elements, group stride(2)
    a.stride_h = t.stride(3);
    a.h_div = (int)t.size(3);
    a.h_mul = (int)(t.stride(2) / t.stride(3));
    a.h_slots = (int)(t.size(2) * a.h_mul);
  }
  return a;
}
static bool attn5_ok(const at::Tensor& t) {
  return is_native(t) && t.dim() == 5 && t.stride(4) == 1 && t.stride(3) == t.size(4) && t.stride(2) % t.stride(3) == 0 &&
         t.stride(0) % 8 == 0 && t.stride(1) % 8 == 0 && (reinterpret_cast<uintptr_t>(t.data_ptr()) % 16) == 0;
}
static bool attn_ok(const at::Tensor& t) {
  return is_native(t) && t.dim() == 4 && t.stride(3) == 1 && t.stride(0) % 8 == 0 && t.stride(1) % 8 == 0 &&
         t.stride(2) % 8 == 0 && (reinterpret_cast<uintptr_t>(t.data_ptr()) % 16) == 0;
}
static std::pair<at::Tensor, at::Tensor> aten_attention(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v,
                                                        double scale, bool causal) {
  // reference math in fp32: returns (o [B,Sq,Hq,D], lse [B,Hq,Sq])
  const int64_t Hq = q.size(2), Hkv = k.size(2), Sq = q.size(1), Sk = k.size(1);
  at::Tensor qf = q.to(at::kFloat).permute({0, 2, 1, 3});
  at::Tensor kf = k.to(at::kFloat).permute({0, 2, 1, 3});
  at::Tensor vf = v.to(at::kFloat).permute({0, 2, 1, 3});
  if (Hq != Hkv) {
    kf = kf.repeat_interleave(Hq / Hkv, 1);
    vf = vf.repeat_interleave(Hq / Hkv, 1);
  }
  at::Tensor s = at::matmul(qf, kf.transpose(-1, -2)) * scale;
  if (causal) {
    at::Tensor mask = at::ones({Sq, Sk}, s.options().dtype(at::kBool)).tril(Sk - Sq);
    s = s.masked_fill(mask.logical_not(), -INFINITY);
  }
  at::Tensor lse = at::logsumexp(s, -1);
  at::Tensor p = at::exp(s - lse.unsqueeze(-1));
  p = at::where(at::isfinite(lse).unsqueeze(-1), p, at::zeros_like(p));
  at::Tensor o = at::matmul(p, vf).permute({0, 2, 1, 3}).contiguous();
  return {o, lse};
}
static Ts attn_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& q = in[0];
  const at::Tensor& k = in[1];
  const at::Tensor& v = in[2];
  const bool causal = op.attrs.b("causal", true);
  const double scale = op.attrs.f("softmax_scale", 0.0) > 0 ? op.attrs.f("softmax_scale") : 1.0 / std::sqrt((double)q.size(3));
  auto fopt = q.options().dtype(at::kFloat);
  if (q.is_meta()) return {at::empty(q.sizes(), q.options()), at::empty({q.size(0), q.size(2), q.size(1)}, fopt)};
  if (attn_ok(q) && attn_ok(k) && attn_ok(v) && (q.size(3) == 64 || q.size(3) == 128)) {
    at::Tensor o = at::empty(q.sizes(), q.options());
    at::Tensor lse = at::empty({q.size(0), q.size(2), q.size(1)}, fopt);
    AttnFwdCall c;
    c.q = as_attn(q); c.k = as_attn(k); c.v = as_attn(v); c.o = as_attn(o);
    c.lse = lse.data_ptr<float>();
    c.B = (int)q.size(0); c.Sq = (int)q.size(1); c.Hq = (int)q.size(2); c.D = (int)q.size(3);
    c.Sk = (int)k.size(1); c.Hkv = (int)k.size(2);
    c.softmax_scale = (float)scale; c.causal = causal;
    cuda_ok(attn_fwd(c, cur_stream()), "attn_fwd");
    return {o, lse};
  }
  // other head sizes below 128 (32, 40, 80, 96, 112, ...): zero-pad the head dimension to the next supported size; the
  // scores are unchanged (padding contributes 0 to q.k), the padded output columns are dropped.  The softmax scale of the
  // ORIGINAL head size is passed explicitly.
  if (is_native(q) && is_native(k) && is_native(v) && q.size(3) < 128 && q.size(3) != 64 && !op.attrs.b("_padded")) {
    const int64_t D = q.size(3), Dp = D < 64 ? 64 : 128;
    auto pad = [&](const at::Tensor& t) { return at::constant_pad_nd(t, {0, Dp - D}).contiguous(); };
    OpDef tmp;
    tmp.attrs = op.attrs;
    tmp.attrs.set("softmax_scale", scale);
    tmp.attrs.set("_padded", true);
    Ts r = attn_compute(tmp, {pad(q), pad(k), pad(v)}, nullptr);
    return {r[0].narrow(3, 0, D).contiguous(), r[1]};
  }
  if (is_native(q)) note_fallback("attn");
  auto r = aten_attention(q, k, v, scale, causal);
  return {r.first.to(q.scalar_type()), r.second};
}
// inputs: do, q, k, v, o, lse -> dq, dk, dv
static Ts attn_bwd_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& d_o = in[0];
  const at::Tensor& q = in[1];
  const at::Tensor& k = in[2];
  const at::Tensor& v = in[3];
  const at::Tensor& o = in[4];
  const at::Tensor& lse = in[5];
  const bool causal = op.attrs.b("causal", true);
  const double scale = op.attrs.f("softmax_scale", 0.0) > 0 ? op.attrs.f("softmax_scale") : 1.0 / std::sqrt((double)q.size(3));
  if (q.is_meta()) return {at::empty(q.sizes(), q.options()), at::empty(k.sizes(), k.options()), at::empty(v.sizes(), v.options())};
  at::Tensor doc = d_o.stride(3) == 1 ? d_o : d_o.contiguous();
  if (attn_ok(q) && attn_ok(k) && attn_ok(v) && attn_ok(o) && attn_ok(doc) && (q.size(3) == 64 || q.size(3) == 128)) {
    at::Tensor dq = at::empty(q.sizes(), q.options()), dk = at::empty(k.sizes(), k.options()), dv = at::empty(v.sizes(), v.options());
    at::Tensor delta = at::empty_like(lse);
    AttnBwdCall c;
    c.q = as_attn(q); c.k = as_attn(k); c.v = as_attn(v); c.o = as_attn(o); c.d_o = as_attn(doc);
    c.dq = as_attn(dq); c.dk = as_attn(dk); c.dv = as_attn(dv);
    c.lse = lse.data_ptr<float>(); c.delta = delta.data_ptr<float>();
    c.B = (int)q.size(0); c.Sq = (int)q.size(1); c.Hq = (int)q.size(2); c.D = (int)q.size(3);
    c.Sk = (int)k.size(1); c.Hkv = (int)k.size(2);
    c.softmax_scale = (float)scale; c.causal = causal;
    cuda_ok(attn_bwd(c, cur_stream()), "attn_bwd");
    return {dq, dk, dv};
  }
  if (is_native(q) && is_native(k) && is_native(v) && q.size(3) < 128 && q.size(3) != 64 && !op.attrs.b("_padded")) {
    const int64_t D = q.size(3), Dp = D < 64 ? 64 : 128;
    auto pad = [&](const at::Tensor& t) { return at::constant_pad_nd(t, {0, Dp - D}).contiguous(); };
    OpDef tmp;
    tmp.attrs = op.attrs;
    tmp.attrs.set("softmax_scale", scale);
    tmp.attrs.set("_padded", true);
    Ts r = attn_bwd_compute(tmp, {pad(d_o), pad(q), pad(k), pad(v), pad(o), lse}, nullptr);
    return {r[0].narrow(3, 0, D).contiguous(), r[1].narrow(3, 0, D).contiguous(), r[2].narrow(3, 0, D).contiguous()};
  }
  if (is_native(q)) note_fallback("attn_bwd");
  // reference backward in fp32 from the SAVED statistics (o, lse), exactly like the kernels: P is rebuilt as
  // exp(S - lse), so a call on one (q-block, kv-block) pair with the global lse composes into ring attention
  const int64_t Hq = q.size(2), Hkv = k.size(2), Sq = q.size(1), Sk = k.size(1), rep = Hq / Hkv;
  at::Tensor qf = q.to(at::kFloat).permute({0, 2, 1, 3}), kf = k.to(at::kFloat).permute({0, 2, 1, 3}), vf = v.to(at::kFloat).permute({0, 2, 1, 3});
  at::Tensor of = in[4].to(at::kFloat).permute({0, 2, 1, 3}), dof = d_o.to(at::kFloat).permute({0, 2, 1, 3});
  if (rep > 1) { kf = kf.repeat_interleave(rep, 1); vf = vf.repeat_interleave(rep, 1); }
  at::Tensor sc = at::matmul(qf, kf.transpose(-1, -2)) * scale;                     // [B, Hq, Sq, Sk]
  if (causal) sc = sc.masked_fill(at::ones({Sq, Sk}, sc.options().dtype(at::kBool)).tril(Sk - Sq).logical_not(), -INFINITY);
  at::Tensor pr = at::exp(sc - lse.to(at::kFloat).unsqueeze(-1));
  pr = at::where(at::isfinite(pr), pr, at::zeros_like(pr));
  at::Tensor dv = at::matmul(pr.transpose(-1, -2), dof);
  at::Tensor dp = at::matmul(dof, vf.transpose(-1, -2));
  at::Tensor delta = (dof * of).sum(-1, true);
  at::Tensor ds = pr * (dp - delta);
  at::Tensor dq = at::matmul(ds, kf) * scale;
  at::Tensor dk = at::matmul(ds.transpose(-1, -2), qf) * scale;
  if (rep > 1) {
    dk = dk.reshape({dk.size(0), Hkv, rep, Sk, dk.size(3)}).sum(2);
    dv = dv.reshape({dv.size(0), Hkv, rep, Sk, dv.size(3)}).sum(2);
  }
  return {dq.permute({0, 2, 1, 3}).contiguous().to(q.scalar_type()), dk.permute({0, 2, 1, 3}).contiguous().to(k.scalar_type()),
          dv.permute({0, 2, 1, 3}).contiguous().to(v.scalar_type())};
}
static TensorList attn_grad(OpDef& op, const TensorList& g) {
  AttrMap a;
  a.set("causal", op.attrs.b("causal", true));
  a.set("softmax_scale", op.attrs.f("softmax_scale", 0.0));
  TensorList r = op.graph->make_op("attn_bwd", {g[0], op.inputs[0], op.inputs[1], op.inputs[2], op.outputs[0], op.outputs[1]}, a);
  return {r[0], r[1], r[2]};
}
static void attn_deduce(OpDef& op, size_t s) {
  copy_out_ds(op, 0, s, op.inputs[0]);
  const Tensor& q = op.inputs[0];
  if (!q->has_ds(s)) return;
  // lse [B,H,S]: batch split 0 -> 0, head split 2 -> 1, seq split 1 -> 2
  const DistributedStates& ds = q->ds(s);
  std::map<int, int> st;
  auto m = [](int d) { return d == 2 ? 1 : (d == 1 ? 2 : d); };
  for (auto& kv : ds.states()) if (kv.second > 1 && kv.first != 3) st[m(kv.first)] = kv.second;
  std::vector<int> order;
  for (int o : ds.order()) if (o != 3) order.push_back(m(o));
  set_out_ds(op, 1, s, DistributedStates(ds.device_num(), st, order));
}
static void attn_bwd_deduce(OpDef& op, size_t s) {
  copy_out_ds(op, 0, s, op.inputs[1]);
  copy_out_ds(op, 1, s, op.inputs[2]);
  copy_out_ds(op, 2, s, op.inputs[3]);
}
HB_REGISTER_OP(attn, "attn", 2, kFlagAttention, attn_compute, attn_grad, attn_deduce, nullptr);
HB_REGISTER_OP(attn_bwd, "attn_bwd", 3, kFlagAttention, attn_bwd_compute, nullptr, attn_bwd_deduce, nullptr);

// ------------------------------------------------------------------ rotary
// x [T..., H, D] with positions [T...]; half-split convention; inverse rotation is the gradient
static Ts rotary_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& x = in[0];
  if (x.is_meta()) return {at::empty_like(x)};
  const bool inverse = op.attrs.b("inverse");
  const double base = op.attrs.f("base", 10000.0);
  const int64_t D = x.size(-1), H = x.size(-2);
  const int64_t rot = op.attrs.i("rot_dim", 0) > 0 ? op.attrs.i("rot_dim") : D;
  const int64_t tokens = x.numel() / (H * D);
  at::Tensor pos;
  if (in.size() > 1) pos = in[1].to(at::kInt).reshape({-1}).contiguous();
  else {
    // default positions: index inside the second-to-last token dim (sequence), plus an offset (CP slot)
    const int64_t S = x.dim() >= 3 ? x.size(-3) : tokens;
    pos = (at::arange(tokens, x.options().dtype(at::kInt)) % S + (int64_t)op.attrs.i("pos_offset", 0)).contiguous();
  }
  if (is_native(x) && x.is_contiguous()) {
    at::Tensor y = at::empty_like(x);
    cuda_ok(rotary_apply(x.data_ptr(), y.data_ptr(), pos.data_ptr<int32_t>(), tokens, (int)H, (int)D, (int)rot,
                         (float)base, inverse, H * D, cur_stream()), "rotary");
    return {y};
  }
  at::Tensor xf = x.to(at::kFloat).reshape({tokens, H, D});
  const int64_t half = rot / 2;
  at::Tensor inv_freq = at::pow(base, -at::arange(0, half, xf.options()) * 2.0 / (double)rot);
  at::Tensor ang = pos.to(at::kFloat).unsqueeze(1) * inv_freq.unsqueeze(0);  // [T, half]
  at::Tensor cs = at::cos(ang).unsqueeze(1), sn = at::sin(ang).unsqueeze(1);
  if (inverse) sn = -sn;
  at::Tensor a = xf.narrow(-1, 0, half), b = xf.narrow(-1, half, half);
  at::Tensor ya = a * cs - b * sn, yb = b * cs + a * sn;
  at::Tensor y = xf.clone();
  y.narrow(-1, 0, half).copy_(ya);
  y.narrow(-1, half, half).copy_(yb);
  return {y.reshape(x.sizes()).to(x.scalar_type())};
}
static TensorList rotary_grad(OpDef& op, const TensorList& g) {
  AttrMap a = op.attrs;
  a.set("inverse", !op.attrs.b("inverse"));
  TensorList ins = {g[0]};
  if (op.inputs.size() > 1) ins.push_back(op.inputs[1]);
  TensorList r = {op.graph->make_op1("rotary", ins, a)};
  if (op.inputs.size() > 1) r.push_back(nullptr);
  return r;
}
HB_REGISTER_OP(rotary, "rotary", 1, 0, rotary_compute, rotary_grad, nullptr, nullptr);

// ------------------------------------------------------------------ dropout (mask recomputed from seed in bwd)
static Ts dropout_compute(const OpDef& op, const Ts& in, RunCtx* rc) {
  const at::Tensor& x = in[0];
  const double p = op.attrs.f("p", 0.0);
  if (x.is_meta() || p <= 0.0 || (rc && !rc->training)) return {x.is_meta() ? at::empty_like(x) : x};
  const uint64_t seed = (rc ? rc->seed : 0) + 0x9E3779B97F4A7C15ull * (uint64_t)(op.attrs.i("seed_op", op.id) + 1);
  const uint64_t offset = rc ? (uint64_t)rc->micro_batch << 40 : 0;
  if (is_native(x) && x.numel() % 8 == 0) {
    // (an expanded / strided gradient is materialised first: the mask is defined on the flattened element index and must
    // be the one the forward -- possibly the fused dropout_add_norm kernel -- used)
    at::Tensor xc = x.contiguous();
    at::Tensor y = at::empty_like(xc);
    cuda_ok(dropout_fwd(xc.data_ptr(), y.data_ptr(), xc.numel(), (float)p, seed, offset, cur_stream()), "dropout");
    return {y};
  }
  auto gen = at::detail::createCPUGenerator(seed ^ offset);
  at::Tensor mask = at::empty(x.sizes(), at::TensorOptions().dtype(at::kFloat)).uniform_(0, 1, gen).to(x.device()) >= p;
  return {(x * mask.to(x.scalar_type())) / (1.0 - p)};
}
static TensorList dropout_grad(OpDef& op, const TensorList& g) {
  AttrMap a = op.attrs;
  a.set("seed_op", op.attrs.i("seed_op", op.id));
  return {op.graph->make_op1("dropout", {g[0]}, a)};
}
HB_REGISTER_OP(dropout, "dropout", 1, 0, dropout_compute, dropout_grad, nullptr, nullptr);

// ------------------------------------------------------------------ fused dropout + residual add + norm
// inputs: x, [residual], gamma, [beta]      attrs: rms, eps, p, has_residual
// outputs: y = norm(z), z = residual + dropout(x) (the new residual stream), mean, rstd
// One kernel in forward (csrc/kernels/norm.cu dropout_add_norm_fwd_kernel); the backward is composed of the norm
// backward on z, the add of the residual-stream gradient and the dropout mask re-created from the same Philox counters.
// (ref: hetu/impl/kernel/RMSNorm.cu:90 DropoutAddLnFwdCuda / :257 DropoutAddLnBwdCuda)
static Ts dropout_add_norm_compute(const OpDef& op, const Ts& in, RunCtx* rc) {
  const bool rms = op.attrs.b("rms"), has_res = op.attrs.b("has_residual");
  const double eps = op.attrs.f("eps", 1e-5);
  double p = op.attrs.f("p", 0.0);
  if (rc && !rc->training) p = 0.0;
  const at::Tensor& x = in[0];
  const at::Tensor* res = has_res ? &in[1] : nullptr;
  const at::Tensor& gamma = in[has_res ? 2 : 1];
  const at::Tensor* beta = (!rms && in.size() > (size_t)(has_res ? 3 : 2)) ? &in[has_res ? 3 : 2] : nullptr;
  const int64_t cols = x.size(-1);
  std::vector<int64_t> sshape(x.sizes().begin(), x.sizes().end() - 1);
  auto fopt = x.options().dtype(at::kFloat);
  if (x.is_meta()) return {at::empty_like(x), at::empty_like(x), at::empty(sshape, fopt), at::empty(sshape, fopt)};
  const int64_t rows = x.numel() / cols;
  const uint64_t seed = (rc ? rc->seed : 0) + 0x9E3779B97F4A7C15ull * (uint64_t)(op.attrs.i("seed_op", op.id) + 1);
  const uint64_t offset = rc ? (uint64_t)rc->micro_batch << 40 : 0;
  if (is_native(x) && x.is_contiguous() && is_native(gamma) && cols % 8 == 0 && cols <= 8192 && (!res || (is_native(*res) && res->is_contiguous()))) {
    at::Tensor y = at::empty_like(x), z = at::empty_like(x), mean = at::empty(sshape, fopt), rstd = at::empty(sshape, fopt);
    if (rms)
      cuda_ok(dropout_add_rmsnorm_fwd(x.data_ptr(), res ? res->data_ptr() : nullptr, gamma.data_ptr(), y.data_ptr(), z.data_ptr(),
                                      rstd.data_ptr<float>(), rows, (int)cols, (float)eps, (float)p, seed, offset, cur_stream()),
              "dropout_add_rmsnorm_fwd");
    else
      cuda_ok(dropout_add_layernorm_fwd(x.data_ptr(), res ? res->data_ptr() : nullptr, gamma.data_ptr(), beta ? beta->data_ptr() : nullptr,
                                        y.data_ptr(), z.data_ptr(), mean.data_ptr<float>(), rstd.data_ptr<float>(), rows, (int)cols,
                                        (float)eps, (float)p, seed, offset, cur_stream()), "dropout_add_layernorm_fwd");
    if (rms) mean.zero_();
    return {y, z, mean, rstd};
  }
  if (is_native(x)) note_fallback("dropout_add_norm");
  at::Tensor d = x;
  if (p > 0.0) {
    auto gen = at::detail::createCPUGenerator(seed ^ offset);
    at::Tensor mask = at::empty(x.sizes(), at::TensorOptions().dtype(at::kFloat)).uniform_(0, 1, gen).to(x.device()) >= p;
    d = (x * mask.to(x.scalar_type())) / (1.0 - p);
  }
  at::Tensor z = (res ? d + *res : d).to(x.scalar_type());
  at::Tensor zf = z.to(at::kFloat), mean, rstd, y;
  if (rms) {
    mean = at::zeros(sshape, fopt);
    rstd = at::rsqrt(zf.pow(2).mean(-1) + eps);
    y = zf * rstd.unsqueeze(-1) * gamma.to(at::kFloat);
  } else {
    mean = zf.mean(-1);
    rstd = at::rsqrt((zf - mean.unsqueeze(-1)).pow(2).mean(-1) + eps);
    y = (zf - mean.unsqueeze(-1)) * rstd.unsqueeze(-1) * gamma.to(at::kFloat);
    if (beta) y = y + beta->to(at::kFloat);
  }
  return {y.to(x.scalar_type()), z, mean, rstd};
}
static TensorList dropout_add_norm_grad(OpDef& op, const TensorList& g) {
  const bool rms = op.attrs.b("rms"), has_res = op.attrs.b("has_residual");
  const Tensor& gamma = op.inputs[has_res ? 2 : 1];
  TensorList r(op.inputs.size());
  Tensor dz = g.size() > 1 ? g[1] : Tensor();
  if (g[0]) {
    AttrMap a;
    a.set("rms", rms);
    TensorList nb = op.graph->make_op("norm_bwd", {g[0], op.outputs[1], gamma, op.outputs[2], op.outputs[3]}, a);
    dz = dz ? op.graph->make_op1("add", {nb[0], dz}, AttrMap()) : nb[0];
    r[has_res ? 2 : 1] = nb[1];
    if (!rms && op.inputs.size() > (size_t)(has_res ? 3 : 2) && nb.size() > 2) r[has_res ? 3 : 2] = nb[2];
  }
  if (!dz) return r;
  if (has_res) r[1] = dz;
  if (op.attrs.f("p", 0.0) > 0.0) {
    AttrMap a;
    a.set("p", op.attrs.f("p", 0.0));
    a.set("seed_op", op.attrs.i("seed_op", op.id));
    r[0] = op.graph->make_op1("dropout", {dz}, a);
  } else r[0] = dz;
  return r;
}
static void dropout_add_norm_infer(OpDef& op) { infer_meta_by_meta_exec(op); }
HB_REGISTER_OP(dropout_add_norm, "dropout_add_norm", 4, 0, dropout_add_norm_compute, dropout_add_norm_grad, nullptr, dropout_add_norm_infer);

}  // namespace hb

namespace hb {
// ------------------------------------------------------------------ packed-QKV attention
// qkv [T, (Hq + 2*Hkv) * D] straight out of the fused projection ([q heads | k heads | v heads] per row);
// the kernels read q/k/v through strides (TMA), and the backward writes dq/dk/dv into one packed buffer that
// feeds the projection's backward GEMMs -- no split / concat / transpose copies anywhere.
using TsP = std::vector<at::Tensor>;
static void packed_views(const at::Tensor& qkv, int64_t S, int64_t Hq, int64_t Hkv, int64_t D, at::Tensor* q, at::Tensor* k, bool interleaved,
                         at::Tensor* v) {
  const int64_t T = qkv.size(0);
  HB_CHECK(T % S == 0) << "attn_packed: " << T << " tokens are not a multiple of seq_len " << S;
  if (interleaved) {
    // per-head interleaved layout [h0: q k v | h1: q k v | ...] (Megatron's): any TP degree that divides the head
    // count owns complete heads, so the same global weight means the same model under every strategy
    // with grouped-query attention the unit is a kv head: [g0: q x rep, k, v | g1: ...]; q is then a 5-D view
    // [B, S, G, rep, D] (two head strides) that the kernels address through their head-slot mapping
    const int64_t rep = Hq / Hkv;
    at::Tensor y = qkv.view({T / S, S, Hkv, rep + 2, D});
    *q = rep == 1 ? y.select(3, 0) : y.narrow(3, 0, rep);
    *k = y.select(3, rep);
    *v = y.select(3, rep + 1);
    return;
  }
  at::Tensor x = qkv.view({T / S, S, Hq + 2 * Hkv, D});
  *q = x.narrow(2, 0, Hq);
  *k = x.narrow(2, Hq, Hkv);
  *v = x.narrow(2, Hq + Hkv, Hkv);
}
// single-launch variable-length attention: per-token document start / end (device int32 [T]), set by attn_packed_compute /
// attn_packed_bwd_compute around ONE call of the row helpers over the whole packed buffer (B = 1, S = T)
struct VarlenCtx { const int* row_start = nullptr; const int* row_end = nullptr; };
static thread_local VarlenCtx tls_varlen;
static bool varlen_fused_enabled() { return env_int("HETU_ATTN_VARLEN_FUSED", 1) != 0; }
// (row_start, row_end) tensors of a cu_seqlens list, cached for the current step like the host copy
static std::pair<at::Tensor, at::Tensor> varlen_rows(const std::vector<int64_t>& cu, int64_t T, const at::Device& dev) {
  static thread_local std::vector<int64_t> last_cu;
  static thread_local at::Tensor last_start, last_end;
  if (last_cu == cu && last_start.defined() && last_start.size(0) == T && last_start.device() == dev) return {last_start, last_end};
  at::Tensor hs = at::empty({T}, at::TensorOptions().dtype(at::kInt)), he = at::empty({T}, at::TensorOptions().dtype(at::kInt));
  int32_t* ps = hs.data_ptr<int32_t>();
  int32_t* pe = he.data_ptr<int32_t>();
  for (int64_t t = 0; t < T; ++t) { ps[t] = (int32_t)t; pe[t] = (int32_t)(t + 1); }      // tokens outside every document: alone
  for (size_t i = 0; i + 1 < cu.size(); ++i)
    for (int64_t t = cu[i]; t < cu[i + 1] && t < T; ++t) { ps[t] = (int32_t)cu[i]; pe[t] = (int32_t)cu[i + 1]; }
  last_cu = cu;
  last_start = hs.to(dev);
  last_end = he.to(dev);
  return {last_start, last_end};
}
static TsP attn_packed_rows(const OpDef& op, const at::Tensor& qkv, int64_t S, RunCtx* rc) {
  // head counts follow the width of the packed projection actually fed (the tensor-parallel degree may change between
  // strategies); the attributes fix the q : kv ratio and the head size
  const int64_t D = op.attrs.i("head_dim");
  const int64_t rep_ = std::max<int64_t>(1, op.attrs.i("num_heads") / std::max<int64_t>(1, op.attrs.i("num_kv_heads", op.attrs.i("num_heads"))));
  const int64_t Hkv = qkv.size(-1) / ((rep_ + 2) * D), Hq = Hkv * rep_;
  const int64_t T = qkv.size(0);
  auto fopt = qkv.options().dtype(at::kFloat);
  if (qkv.is_meta()) return {at::empty({T, Hq * D}, qkv.options()), at::empty({T / std::max<int64_t>(S, 1), Hq, S}, fopt)};
  at::Tensor q, k, v;
  at::Tensor src = qkv.is_contiguous() ? qkv : qkv.contiguous();
  packed_views(src, S, Hq, Hkv, D, &q, &k, op.attrs.s("layout", "qkv") == "hqkv", &v);
  if (q.dim() == 5) {
    const int64_t B = T / S;
    if (attn5_ok(q) && attn_ok(k) && attn_ok(v) && (D == 64 || D == 128)) {
      at::Tensor o = at::empty({B, S, Hq, D}, qkv.options());
      at::Tensor lse = at::empty({B, Hq, S}, fopt);
      AttnFwdCall c;
      c.q = as_attn(q); c.k = as_attn(k); c.v = as_attn(v); c.o = as_attn(o);
      c.lse = lse.data_ptr<float>();
      c.B = (int)B; c.Sq = (int)S; c.Sk = (int)S; c.Hq = (int)Hq; c.Hkv = (int)Hkv; c.D = (int)D;
      c.softmax_scale = (float)(op.attrs.f("softmax_scale", 0.0) > 0 ? op.attrs.f("softmax_scale") : 1.0 / std::sqrt((double)D));
      c.causal = op.attrs.b("causal", true);
      c.row_start = tls_varlen.row_start;
      cuda_ok(attn_fwd(c, cur_stream()), "attn_fwd");
      return {o.reshape({T, Hq * D}), lse};
    }
    q = q.reshape({B, S, Hq, D});   // fallback: gather the heads
  } else if (tls_varlen.row_start != nullptr && attn_ok(q) && attn_ok(k) && attn_ok(v) && (D == 64 || D == 128)) {
    const int64_t B = T / S;
    at::Tensor o = at::empty({B, S, Hq, D}, qkv.options());
    at::Tensor lse = at::empty({B, Hq, S}, fopt);
    AttnFwdCall c;
    c.q = as_attn(q); c.k = as_attn(k); c.v = as_attn(v); c.o = as_attn(o);
    c.lse = lse.data_ptr<float>();
    c.B = (int)B; c.Sq = (int)S; c.Sk = (int)S; c.Hq = (int)Hq; c.Hkv = (int)Hkv; c.D = (int)D;
    c.softmax_scale = (float)(op.attrs.f("softmax_scale", 0.0) > 0 ? op.attrs.f("softmax_scale") : 1.0 / std::sqrt((double)D));
    c.causal = true;
    c.row_start = tls_varlen.row_start;
    cuda_ok(attn_fwd(c, cur_stream()), "attn_fwd");
    return {o.reshape({T, Hq * D}), lse};
  }
  HB_CHECK(tls_varlen.row_start == nullptr) << "single-launch varlen attention needs the native kernel path";
  static const OpKernel* kern = OpRegistry::get().find("attn");
  OpDef tmp;
  tmp.attrs = op.attrs;
  tmp.kernel = kern;
  auto r = kern->compute(tmp, {q, k, v}, rc);
  return {r[0].reshape({T, Hq * D}), r[1]};
}
// inputs: do [T, Hq*D], qkv, o [T, Hq*D], lse -> dqkv
// document boundaries of a packed batch (varlen attention): host copy of cu_seqlens, cached for the current step
static std::vector<int64_t> host_cu_seqlens(const at::Tensor& cu, RunCtx* rc) {
  static thread_local std::unordered_map<const void*, std::pair<uint64_t, std::vector<int64_t>>> cache;
  const uint64_t stamp = rc ? rc->seed * 131 + (uint64_t)rc->micro_batch : 0;
  auto it = cache.find(cu.data_ptr());
  if (rc != nullptr && it != cache.end() && it->second.first == stamp && (int64_t)it->second.second.size() == cu.numel()) return it->second.second;
  at::Tensor h = cu.to(at::kCPU, at::kLong).contiguous();
  std::vector<int64_t> v(h.data_ptr<int64_t>(), h.data_ptr<int64_t>() + h.numel());
  if (cache.size() > 64) cache.clear();
  cache[cu.data_ptr()] = {stamp, v};
  return v;
}
// forward.  inputs: qkv [T, W], [cu_seqlens int32 [n + 1]] -- with cu_seqlens every document [cu[i], cu[i+1]) attends
// (causally) only to itself: one kernel launch per document over strided views of the packed buffer; the saved lse is the
// concatenation of the per-document [Hq, len] blocks (shape [1, Hq, T])
static TsP attn_packed_compute(const OpDef& op, const TsP& in, RunCtx* rc) {
  const at::Tensor& qkv = in[0];
  const int64_t S = op.sy_shape.empty() ? op.attrs.i("seq_len") : op.sy_shape[0].get_val();
  if (in.size() < 2) return attn_packed_rows(op, qkv, S, rc);
  const int64_t T = qkv.size(0), D = op.attrs.i("head_dim");
  const int64_t rep_ = std::max<int64_t>(1, op.attrs.i("num_heads") / std::max<int64_t>(1, op.attrs.i("num_kv_heads", op.attrs.i("num_heads"))));
  const int64_t Hq = qkv.size(-1) / ((rep_ + 2) * D) * rep_;
  auto fopt = qkv.options().dtype(at::kFloat);
  if (qkv.is_meta()) return {at::empty({T, Hq * D}, qkv.options()), at::empty({1, Hq, T}, fopt)};
  const std::vector<int64_t> cu = host_cu_seqlens(in[1], rc);
  HB_CHECK(cu.size() >= 2 && cu.front() == 0 && cu.back() <= T) << "attn_packed: bad cu_seqlens";
  at::Tensor src = qkv.is_contiguous() ? qkv : qkv.contiguous();
  if (is_native(src) && (D == 64 || D == 128) && op.attrs.b("causal", true) && varlen_fused_enabled()) {
    // ONE launch over the whole packed buffer: block-diagonal causal mask from the per-token document starts, KV tiles
    // outside a query tile's documents are skipped inside the kernel; lse comes back in the natural [1, Hq, T] layout
    auto rows = varlen_rows(cu, T, src.device());
    tls_varlen.row_start = rows.first.data_ptr<int32_t>();
    tls_varlen.row_end = rows.second.data_ptr<int32_t>();
    TsP r;
    bool ok = true;
    try { r = attn_packed_rows(op, src, T, rc); } catch (const Error&) { ok = false; }     // layout not addressable: per-document path
    tls_varlen = VarlenCtx();
    if (ok) return {r[0], r[1].reshape({1, Hq, T})};
  }
  at::Tensor o = at::zeros({T, Hq * D}, qkv.options());
  at::Tensor lse = at::zeros({Hq * T}, fopt);
  for (size_t i = 0; i + 1 < cu.size(); ++i) {
    const int64_t lo = cu[i], len = cu[i + 1] - cu[i];
    if (len <= 0) continue;
    TsP r = attn_packed_rows(op, src.narrow(0, lo, len), len, rc);
    o.narrow(0, lo, len).copy_(r[0]);
    lse.narrow(0, Hq * lo, Hq * len).copy_(r[1].reshape({-1}));
  }
  return {o, lse.reshape({1, Hq, T})};
}
static TsP attn_packed_bwd_rows(const OpDef& op, const at::Tensor& d_o, const at::Tensor& qkv, const at::Tensor& o, const at::Tensor& lse,
                                int64_t S, RunCtx* rc);
// inputs: do [T, Hq*D], qkv, o, lse, [cu_seqlens] -> dqkv
static TsP attn_packed_bwd_compute(const OpDef& op, const TsP& in, RunCtx* rc) {
  const at::Tensor& qkv = in[1];
  if (qkv.is_meta()) return {at::empty_like(qkv)};
  const int64_t S = op.sy_shape.empty() ? op.attrs.i("seq_len") : op.sy_shape[0].get_val();
  if (in.size() < 5) return attn_packed_bwd_rows(op, in[0], qkv, in[2], in[3], S, rc);
  const int64_t T = qkv.size(0), D = op.attrs.i("head_dim");
  const int64_t rep_ = std::max<int64_t>(1, op.attrs.i("num_heads") / std::max<int64_t>(1, op.attrs.i("num_kv_heads", op.attrs.i("num_heads"))));
  const int64_t Hq = qkv.size(-1) / ((rep_ + 2) * D) * rep_;
  const std::vector<int64_t> cu = host_cu_seqlens(in[4], rc);
  at::Tensor src = qkv.is_contiguous() ? qkv : qkv.contiguous();
  if (is_native(src) && (D == 64 || D == 128) && op.attrs.b("causal", true) && varlen_fused_enabled()) {
    auto rows = varlen_rows(cu, T, src.device());
    tls_varlen.row_start = rows.first.data_ptr<int32_t>();
    tls_varlen.row_end = rows.second.data_ptr<int32_t>();
    TsP r;
    bool ok = true;
    try { r = attn_packed_bwd_rows(op, in[0], src, in[2], in[3].reshape({1, Hq, T}), T, rc); } catch (const Error&) { ok = false; }
    tls_varlen = VarlenCtx();
    if (ok) return r;
  }
  at::Tensor dqkv = at::zeros_like(src);
  at::Tensor lflat = in[3].reshape({-1}), dof = in[0].contiguous(), of = in[2].contiguous();
  for (size_t i = 0; i + 1 < cu.size(); ++i) {
    const int64_t lo = cu[i], len = cu[i + 1] - cu[i];
    if (len <= 0) continue;
    TsP r = attn_packed_bwd_rows(op, dof.narrow(0, lo, len), src.narrow(0, lo, len), of.narrow(0, lo, len),
                                 lflat.narrow(0, Hq * lo, Hq * len).reshape({1, Hq, len}), len, rc);
    dqkv.narrow(0, lo, len).copy_(r[0]);
  }
  (void)T;
  return {dqkv};
}
static TsP attn_packed_bwd_rows(const OpDef& op, const at::Tensor& d_o, const at::Tensor& qkv, const at::Tensor& o, const at::Tensor& lse,
                                int64_t S, RunCtx* rc) {
  // head counts follow the width of the packed projection actually fed (the tensor-parallel degree may change between
  // strategies); the attributes fix the q : kv ratio and the head size
  const int64_t D = op.attrs.i("head_dim");
  const int64_t rep_ = std::max<int64_t>(1, op.attrs.i("num_heads") / std::max<int64_t>(1, op.attrs.i("num_kv_heads", op.attrs.i("num_heads"))));
  const int64_t Hkv = qkv.size(-1) / ((rep_ + 2) * D), Hq = Hkv * rep_;
  const int64_t T = qkv.size(0), B = T / S;
  at::Tensor src = qkv.is_contiguous() ? qkv : qkv.contiguous();
  at::Tensor q, k, v, dq, dk, dv;
  packed_views(src, S, Hq, Hkv, D, &q, &k, op.attrs.s("layout", "qkv") == "hqkv", &v);
  at::Tensor dqkv = at::empty_like(src);
  packed_views(dqkv, S, Hq, Hkv, D, &dq, &dk, op.attrs.s("layout", "qkv") == "hqkv", &dv);
  at::Tensor o4 = o.contiguous().view({B, S, Hq, D}), do4 = d_o.contiguous().view({B, S, Hq, D});
  const bool causal = op.attrs.b("causal", true);
  const double scale = op.attrs.f("softmax_scale", 0.0) > 0 ? op.attrs.f("softmax_scale") : 1.0 / std::sqrt((double)D);
  const bool grouped = q.dim() == 5;
  if ((grouped ? attn5_ok(q) : attn_ok(q)) && attn_ok(k) && attn_ok(v) && attn_ok(o4) && attn_ok(do4) && (D == 64 || D == 128)) {
    at::Tensor delta = at::empty_like(lse);
    AttnBwdCall c;
    c.q = as_attn(q); c.k = as_attn(k); c.v = as_attn(v); c.o = as_attn(o4); c.d_o = as_attn(do4);
    c.dq = as_attn(dq); c.dk = as_attn(dk); c.dv = as_attn(dv);
    c.lse = lse.data_ptr<float>(); c.delta = delta.data_ptr<float>();
    c.B = (int)B; c.Sq = (int)S; c.Sk = (int)S; c.Hq = (int)Hq; c.Hkv = (int)Hkv; c.D = (int)D;
    c.softmax_scale = (float)scale; c.causal = causal;
    c.row_start = tls_varlen.row_start; c.row_end = tls_varlen.row_end;
    cuda_ok(attn_bwd(c, cur_stream()), "attn_bwd");
    return {dqkv};
  }
  HB_CHECK(tls_varlen.row_start == nullptr) << "single-launch varlen attention backward needs the native kernel path";
  static const OpKernel* bwd = OpRegistry::get().find("attn_bwd");
  OpDef tmp;
  tmp.attrs = op.attrs;
  tmp.kernel = bwd;
  auto r = bwd->compute(tmp, {do4, grouped ? q.reshape({B, S, Hq, D}) : q, k, v, o4, lse}, rc);
  dq.copy_(grouped ? r[0].reshape(dq.sizes()) : r[0]); dk.copy_(r[1]); dv.copy_(r[2]);
  return {dqkv};
}
static TensorList attn_packed_grad(OpDef& op, const TensorList& g) {
  OpDef* fw = &op;
  TensorList ins = {g[0], op.inputs[0], op.outputs[0], op.outputs[1]};
  if (op.inputs.size() > 1) ins.push_back(op.inputs[1]);      // cu_seqlens
  TensorList r = {op.graph->make_op1("attn_packed_bwd", ins, op.attrs, {}, [fw](OpDef& o) { o.sy_shape = fw->sy_shape; })};
  if (op.inputs.size() > 1) r.push_back(nullptr);
  return r;
}
static void attn_packed_deduce(OpDef& op, size_t s) {
  copy_out_ds(op, 0, s, op.inputs[0]);
  const Tensor& x = op.inputs[0];
  if (!x->has_ds(s)) return;
  // lse [B, H, S]: token split -> batch split, feature (head) split -> head split
  const DistributedStates& ds = x->ds(s);
  std::map<int, int> st;
  for (auto& kv : ds.states()) if (kv.second > 1) st[kv.first] = kv.second;
  set_out_ds(op, 1, s, DistributedStates(ds.device_num(), st, ds.order()));
}
HB_REGISTER_OP(attn_packed, "attn_packed", 2, kFlagAttention, attn_packed_compute, attn_packed_grad, attn_packed_deduce, nullptr);
HB_REGISTER_OP(attn_packed_bwd, "attn_packed_bwd", 1, kFlagAttention, attn_packed_bwd_compute, nullptr,
               [](OpDef& op, size_t s) { copy_out_ds(op, 0, s, op.inputs[1]); }, nullptr);

// rotary on the q and k heads of a packed qkv buffer (one launch, v untouched)
static TsP rotary_packed_compute(const OpDef& op, const TsP& in, RunCtx*) {
  const at::Tensor& qkv = in[0];
  if (qkv.is_meta()) return {at::empty_like(qkv)};
  const int64_t S = op.sy_shape.empty() ? op.attrs.i("seq_len") : op.sy_shape[0].get_val();
  // head counts follow the width of the packed projection actually fed (the tensor-parallel degree may change between
  // strategies); the attributes fix the q : kv ratio and the head size
  const int64_t D = op.attrs.i("head_dim");
  const int64_t rep_ = std::max<int64_t>(1, op.attrs.i("num_heads") / std::max<int64_t>(1, op.attrs.i("num_kv_heads", op.attrs.i("num_heads"))));
  const int64_t Hkv = qkv.size(-1) / ((rep_ + 2) * D), Hq = Hkv * rep_;
  const bool inverse = op.attrs.b("inverse");
  const double base = op.attrs.f("base", 10000.0);
  const int64_t T = qkv.size(0);
  at::Tensor pos;
  if (in.size() > 1) pos = in[1].to(at::kInt).reshape({-1}).contiguous();
  else pos = (at::arange(T, qkv.options().dtype(at::kInt)) % S + (int64_t)op.attrs.i("pos_offset", 0)).contiguous();
  at::Tensor out = qkv.contiguous().clone();
  const bool grouped = op.attrs.s("layout", "qkv") == "hqkv";   // [g: q x rep, k, v] per kv head: rotate rep + 1 heads of each group
  const int64_t rep = Hq / Hkv;
  if (is_native(out)) {
    cuda_ok(rotary_apply(out.data_ptr(), out.data_ptr(), pos.data_ptr<int32_t>(), T, (int)(Hq + Hkv), (int)D, (int)D, (float)base,
                         inverse, (Hq + 2 * Hkv) * D, cur_stream(), grouped ? (int)(rep + 1) : 0, grouped ? (int)((rep + 2) * D) : 0),
            "rotary_packed");
    return {out};
  }
  at::Tensor rot_view = grouped ? out.view({T, Hkv, rep + 2, D}).narrow(2, 0, rep + 1)
                                : out.view({T, 1, Hq + 2 * Hkv, D}).narrow(2, 0, Hq + Hkv);
  at::Tensor x = rot_view.reshape({T, Hq + Hkv, D}).to(at::kFloat);
  const int64_t half = D / 2;
  at::Tensor inv_freq = at::pow(base, -at::arange(0, half, x.options()) * 2.0 / (double)D);
  at::Tensor ang = pos.to(at::kFloat).unsqueeze(1) * inv_freq.unsqueeze(0);
  at::Tensor cs = at::cos(ang).unsqueeze(1), sn = at::sin(ang).unsqueeze(1);
  if (inverse) sn = -sn;
  at::Tensor a = x.narrow(-1, 0, half), b = x.narrow(-1, half, half);
  at::Tensor y = at::cat({a * cs - b * sn, b * cs + a * sn}, -1).to(out.scalar_type());
  rot_view.copy_(y.reshape(rot_view.sizes()));
  return {out};
}
static TensorList rotary_packed_grad(OpDef& op, const TensorList& g) {
  AttrMap a = op.attrs;
  a.set("inverse", !op.attrs.b("inverse"));
  TensorList ins = {g[0]};
  if (op.inputs.size() > 1) ins.push_back(op.inputs[1]);
  OpDef* fw = &op;
  TensorList r = {op.graph->make_op1("rotary_packed", ins, a, {}, [fw](OpDef& o) { o.sy_shape = fw->sy_shape; })};
  if (op.inputs.size() > 1) r.push_back(nullptr);
  return r;
}
HB_REGISTER_OP(rotary_packed, "rotary_packed", 1, 0, rotary_packed_compute, rotary_packed_grad, nullptr, nullptr);
}  // namespace hb

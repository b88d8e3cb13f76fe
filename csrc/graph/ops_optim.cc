// Optimizer update ops, loss-scaling helpers and MoE ops.
// The update ops declare (param, grad, states, hyper-parameters); on the training hot
// path the executor fuses all of them into one flat multi-tensor Adam launch per
// parameter buffer (see exec.cc), the per-tensor compute below serves eager graphs
// and parameters that are not part of a flat buffer.
// (capability parity: hetu/graph/ops/optimizer_update.cc:21-83, graph/optim/optimizer.cc,
//  graph/autocast/gradscaler; hetu/v1/python/hetu/layers/{moe_layer,TopGate}.py)
#include <ATen/ATen.h>

#include <algorithm>

#include "exec.h"
#include "ir.h"
#include "op_utils.h"
#include "../runtime/symm_mem.h"
#include "exec.h"

namespace hb {

using Ts = std::vector<at::Tensor>;

static void scalar_out_infer(OpDef& op) { make_out(op, 0, {1}, DataType::FLOAT32); }

// inputs: param, grad, m, v, step, [master]
static Ts adam_update_compute(const OpDef& op, const Ts& in, RunCtx* rc) {
  at::Tensor param = in[0];
  const at::Tensor& grad = in[1];
  at::Tensor m = in[2], v = in[3], step = in[4];
  at::Tensor master = in.size() > 5 ? in[5] : param;
  if (param.is_meta()) return {at::empty({1}, param.options().dtype(at::kFloat))};
  const double lr = op.attrs.f("lr", 1e-3), b1 = op.attrs.f("beta1", 0.9), b2 = op.attrs.f("beta2", 0.999),
               eps = op.attrs.f("eps", 1e-8), wd = op.attrs.f("weight_decay", 0.0);
  const bool native = master.is_cuda() && master.scalar_type() == at::kFloat && master.is_contiguous() && m.is_contiguous() &&
                      v.is_contiguous() && grad.is_contiguous() && (grad.scalar_type() == at::kFloat || grad.scalar_type() == at::kBFloat16);
  const bool defer_step = native && rc != nullptr && rc->deferred_steps != nullptr && step.is_cuda();
  if (!defer_step) step.add_(1);
  if (master.is_cuda() && master.scalar_type() == at::kFloat && master.is_contiguous() && m.is_contiguous() &&
      v.is_contiguous() && grad.is_contiguous() && (grad.scalar_type() == at::kFloat || grad.scalar_type() == at::kBFloat16)) {
    AdamArgs a;
    a.master = master.data_ptr<float>(); a.m = m.data_ptr<float>(); a.v = v.data_ptr<float>();
    a.grad = grad.data_ptr(); a.grad_is_bf16 = grad.scalar_type() == at::kBFloat16;
    a.param_bf16 = (param.scalar_type() == at::kBFloat16 && param.data_ptr() != master.data_ptr()) ? param.data_ptr() : nullptr;
    a.n = master.numel();
    a.lr = (float)lr; a.beta1 = (float)b1; a.beta2 = (float)b2; a.eps = (float)eps; a.weight_decay = (float)wd;
    at::Tensor step_dev = step.is_cuda() ? step : step.to(master.device());
    a.step_ptr = step_dev.data_ptr<int64_t>();
    if (defer_step) { a.step_add = 1; rc->deferred_steps->push_back(step); }
    cuda_ok(adam_update(a, cur_stream()), "adam_update");
    if (rc != nullptr && rc->workspace != nullptr) return {rc->scratch("__update_done", {1}, at::kFloat, master.device())};
    return {at::zeros({1}, master.options())};
  }
  const double t = (double)step.item<int64_t>();
  at::Tensor g = grad.to(at::kFloat).to(master.device());
  m.mul_(b1).add_(g, 1 - b1);
  v.mul_(b2).addcmul_(g, g, 1 - b2);
  const double bc1 = 1 - std::pow(b1, t), bc2 = 1 - std::pow(b2, t);
  at::Tensor denom = (v.sqrt() / std::sqrt(bc2)).add_(eps);
  at::Tensor mf = master.to(at::kFloat);
  mf = mf - (lr / bc1) * (m / denom) - lr * wd * mf;
  master.copy_(mf);
  if (param.data_ptr() != master.data_ptr()) param.copy_(mf);
  return {at::zeros({1}, at::TensorOptions().dtype(at::kFloat))};
}
HB_REGISTER_OP(adam_update, "adam_update", 1, kFlagOptimizerUpdate | kFlagNondiff | kFlagNoMetaExec | kFlagInplace,
               adam_update_compute, nullptr, nullptr, scalar_out_infer);

// inputs: param, grad, [velocity]
static Ts sgd_update_compute(const OpDef& op, const Ts& in, RunCtx*) {
  at::Tensor param = in[0];
  const at::Tensor& grad = in[1];
  if (param.is_meta()) return {at::empty({1}, param.options().dtype(at::kFloat))};
  const double lr = op.attrs.f("lr", 0.01), mom = op.attrs.f("momentum", 0.0), wd = op.attrs.f("weight_decay", 0.0);
  const bool nesterov = op.attrs.b("nesterov");
  at::Tensor g = grad.to(param.scalar_type());
  if (wd != 0.0) g = g + wd * param;
  if (mom != 0.0 && in.size() > 2) {
    at::Tensor vel = in[2];
    vel.mul_(mom).add_(g);
    g = nesterov ? g + mom * vel : vel;
  }
  param.sub_(g * lr);
  return {at::zeros({1}, at::TensorOptions().dtype(at::kFloat))};
}
HB_REGISTER_OP(sgd_update, "sgd_update", 1, kFlagOptimizerUpdate | kFlagNondiff | kFlagNoMetaExec | kFlagInplace,
               sgd_update_compute, nullptr, nullptr, scalar_out_infer);

// rule_update: the rest of the optimizer family, one op with a `rule` attribute (ref: hetu/v1/python/hetu/optimizer.py AdaGrad /
// AMSGrad / Lamb update ops and hetu/v1/src/ops/Optimizer*.cu).  inputs by rule:
//   adagrad: param, grad, accumulator                       p -= lr * g / (sqrt(acc += g^2) + eps)
//   amsgrad: param, grad, m, v, vhat, step                  Adam whose denominator uses the running maximum of v
//   lamb   : param, grad, m, v, step                        Adam direction (+ decoupled decay) scaled by |p| / |update| per tensor
// `l2` adds l2 * p to the gradient first (v1's l2reg); `weight_decay` is the decoupled form.
static Ts rule_update_compute(const OpDef& op, const Ts& in, RunCtx*) {
  at::Tensor param = in[0];
  if (param.is_meta()) return {at::empty({1}, param.options().dtype(at::kFloat))};
  const std::string rule = op.attrs.s("rule", "adagrad");
  const double lr = op.attrs.f("lr", 0.01), eps = op.attrs.f("eps", 1e-7), l2 = op.attrs.f("l2", 0.0),
               wd = op.attrs.f("weight_decay", 0.0), b1 = op.attrs.f("beta1", 0.9), b2 = op.attrs.f("beta2", 0.999);
  at::NoGradGuard ng;
  at::Tensor p = param.to(at::kFloat);
  at::Tensor g = in[1].to(at::kFloat).to(param.device());
  if (l2 != 0.0) g = g + l2 * p;
  if (rule == "adagrad") {
    at::Tensor acc = in[2];
    acc.addcmul_(g, g);
    p = p - lr * g / (acc.sqrt() + eps);
  } else if (rule == "amsgrad") {
    at::Tensor m = in[2], v = in[3], vhat = in[4], step = in[5];
    step.add_(1);
    const double t = (double)step.item<int64_t>();
    m.mul_(b1).add_(g, 1 - b1);
    v.mul_(b2).addcmul_(g, g, 1 - b2);
    at::Tensor vc = v / (1 - std::pow(b2, t));
    vhat.copy_(at::maximum(vhat, vc));
    p = p - lr * (m / (1 - std::pow(b1, t))) / (vhat.sqrt() + eps) - lr * wd * p;
  } else if (rule == "lamb") {
    at::Tensor m = in[2], v = in[3], step = in[4];
    step.add_(1);
    const double t = (double)step.item<int64_t>();
    m.mul_(b1).add_(g, 1 - b1);
    v.mul_(b2).addcmul_(g, g, 1 - b2);
    at::Tensor upd = (m / (1 - std::pow(b1, t))) / ((v / (1 - std::pow(b2, t))).sqrt() + eps) + wd * p;
    const double pn = p.norm().item<double>(), un = upd.norm().item<double>();
    const double trust = (pn > 0 && un > 0) ? pn / un : 1.0;
    p = p - lr * trust * upd;
  } else {
    HB_CHECK(false) << "rule_update: unknown rule '" << rule << "'";
  }
  param.copy_(p);
  return {at::zeros({1}, at::TensorOptions().dtype(at::kFloat))};
}
HB_REGISTER_OP(rule_update, "rule_update", 1, kFlagOptimizerUpdate | kFlagNondiff | kFlagNoMetaExec | kFlagInplace,
               rule_update_compute, nullptr, nullptr, scalar_out_infer);

// update_scale(scale, growth_tracker, found_inf): dynamic loss scaling (GradScaler)
static Ts update_scale_compute(const OpDef& op, const Ts& in, RunCtx*) {
  at::Tensor scale = in[0], tracker = in[1];
  const at::Tensor& found = in[2];
  if (scale.is_meta()) return {at::empty({1}, scale.options())};
  const double growth = op.attrs.f("growth_factor", 2.0), backoff = op.attrs.f("backoff_factor", 0.5);
  const int64_t interval = op.attrs.i("growth_interval", 2000);
  if (found.item<float>() > 0) {
    scale.mul_(backoff);
    tracker.zero_();
  } else {
    tracker.add_(1);
    if (tracker.item<int64_t>() >= interval) {
      scale.mul_(growth);
      tracker.zero_();
    }
  }
  return {scale.clone()};
}
HB_REGISTER_OP(update_scale, "update_scale", 1, kFlagNondiff | kFlagNoMetaExec | kFlagInplace, update_scale_compute, nullptr,
               nullptr, scalar_out_infer);

// ------------------------------------------------------------------ MoE (HetuMoE)
// moe_gate: logits [T, E] -> (gates [T,k] fp32, topk_idx [T,k] int32, location [T,k] int32, aux_loss [1])
static Ts moe_gate_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& logits = in[0];
  const int64_t T = logits.size(0), E = logits.size(1);
  const int64_t k = op.attrs.i("k", 1), capacity = op.attrs.i("capacity");
  auto fopt = logits.options().dtype(at::kFloat);
  auto iopt = logits.options().dtype(at::kInt);
  if (logits.is_meta()) return {at::empty({T, k}, fopt), at::empty({T, k}, iopt), at::empty({T, k}, iopt), at::empty({1}, fopt)};
  at::Tensor probs, idx, val, loc = at::empty({T, k}, iopt);
  if (is_native(logits) && logits.is_contiguous() && E <= 256 && k <= 8) {
    probs = at::empty({T, E}, fopt);
    idx = at::empty({T, k}, iopt);
    val = at::empty({T, k}, fopt);
    cuda_ok(moe_gate_topk(logits.data_ptr(), probs.data_ptr<float>(), idx.data_ptr<int32_t>(), val.data_ptr<float>(), T, (int)E,
                          (int)k, cur_stream()), "moe_gate_topk");
    cuda_ok(moe_assign_slots(idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(), nullptr, T, (int)E, (int)k, (int)capacity,
                             cur_stream()), "moe_assign_slots");
  } else {
    probs = at::softmax(logits.to(at::kFloat), -1);
    auto tk = at::topk(probs, k, -1);
    val = std::get<0>(tk).contiguous();
    idx = std::get<1>(tk).to(at::kInt).contiguous();
    // GShard ordering: all first choices before any second choice, token order inside a choice
    at::Tensor idx_cpu = idx.cpu();
    at::Tensor loc_cpu = at::empty({T, k}, at::TensorOptions().dtype(at::kInt));
    std::vector<int> count(E, 0);
    auto ia = idx_cpu.accessor<int, 2>();
    auto la = loc_cpu.accessor<int, 2>();
    for (int64_t kk = 0; kk < k; ++kk)
      for (int64_t t = 0; t < T; ++t) {
        const int e = ia[t][kk];
        la[t][kk] = count[e] < capacity ? count[e] : -1;
        count[e]++;
      }
    loc = loc_cpu.to(logits.device());
  }
  // normalise the kept gates (dropped tokens contribute zero), load-balancing aux loss as in GShard
  at::Tensor kept = (loc >= 0).to(at::kFloat);
  at::Tensor gates = val * kept;
  if (k > 1) gates = gates / gates.sum(-1, true).clamp_min(1e-9);
  at::Tensor me = probs.mean(0);
  at::Tensor ce = at::one_hot(idx.select(1, 0).to(at::kLong), E).to(at::kFloat).mean(0);
  at::Tensor aux = (me * ce).sum().reshape({1}) * (double)E;
  return {gates, idx, loc, aux};
}
// routing decisions are per token: they inherit the token layout of the logits; the auxiliary loss is a per-rank scalar
static void moe_gate_deduce(OpDef& op, size_t s) {
  const Tensor& lg = op.inputs[0];
  if (!lg->has_ds(s)) return;
  for (size_t i = 0; i < 3 && i < op.outputs.size(); ++i) copy_out_ds(op, i, s, lg);
  if (op.outputs.size() > 3) {
    const int n = lg->ds(s).device_num();
    set_out_ds(op, 3, s, DistributedStates(n, {{kDupDim, n}}, {kDupDim}));
  }
}
HB_REGISTER_OP(moe_gate, "moe_gate", 4, kFlagNondiff, moe_gate_compute, nullptr, moe_gate_deduce, nullptr);


// moe_balance_assign: BASE-layer balanced assignment.  scores [T, E] -> (gate prob of the assigned expert [T,1] fp32,
// expert [T,1] int32, slot [T,1] int32, aux [1] = 0).  Every expert receives at most `capacity` tokens.  Rounds of
// (propose to the best expert with room, experts keep their best proposers); the CUDA path runs the two kernels of
// moe.cu for E rounds without touching the host, the CPU path is the same algorithm in plain loops.
// (ref: hetu/v1/python/hetu/layers/BalanceGate.py:42 balance_assignment_op, gpu_ops/BalanceAssignment.py)
static Ts moe_balance_assign_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& logits = in[0];
  const int64_t T = logits.size(0), E = logits.size(1);
  int64_t capacity = op.attrs.i("capacity", 0);
  if (capacity <= 0) capacity = (T + E - 1) / E;
  HB_CHECK(capacity * E >= T) << "balanced assignment needs capacity * experts >= tokens (" << capacity << " * " << E << " < " << T << ")";
  auto fopt = logits.options().dtype(at::kFloat);
  auto iopt = logits.options().dtype(at::kInt);
  if (logits.is_meta()) return {at::empty({T, 1}, fopt), at::empty({T, 1}, iopt), at::empty({T, 1}, iopt), at::empty({1}, fopt)};
  at::Tensor scores = logits.to(at::kFloat).contiguous();
  at::Tensor idx = at::empty({T, 1}, iopt), loc = at::empty({T, 1}, iopt);
  if (scores.is_cuda() && env_int("HETU_B200_FORCE_CPU", 0) == 0) {
    at::Tensor filled = at::empty({E}, iopt), choice = at::empty({T}, iopt);
    cuda_ok(moe_balance_assign(scores.data_ptr<float>(), idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(), filled.data_ptr<int32_t>(),
                               choice.data_ptr<int32_t>(), T, (int)E, (int)capacity, cur_stream()), "moe_balance_assign");
  } else {
    at::Tensor sc = scores.cpu();
    const float* sp = sc.data_ptr<float>();
    std::vector<int> ia(T, -1), la(T, -1), filled(E, 0), choice(T, -1);
    for (int64_t round = 0; round < E; ++round) {
      bool any = false;
      for (int64_t t = 0; t < T; ++t) {
        choice[t] = -1;
        if (ia[t] >= 0) continue;
        float best = 0.f;
        for (int64_t e = 0; e < E; ++e) {
          if (filled[e] >= capacity) continue;
          if (choice[t] < 0 || sp[t * E + e] > best) { best = sp[t * E + e]; choice[t] = (int)e; }
        }
        any = true;
      }
      if (!any) break;
      for (int64_t e = 0; e < E; ++e) {
        const int room = (int)capacity - filled[e];
        if (room <= 0) continue;
        std::vector<int64_t> cand;
        for (int64_t t = 0; t < T; ++t) if (choice[t] == e) cand.push_back(t);
        if ((int)cand.size() > room) {
          // keep the `room` best scores, ties by token order; then place in token order
          std::stable_sort(cand.begin(), cand.end(), [&](int64_t a, int64_t b) { return sp[a * E + e] > sp[b * E + e]; });
          cand.resize(room);
          std::sort(cand.begin(), cand.end());
        }
        for (int64_t t : cand) { ia[t] = (int)e; la[t] = filled[e]++; }
      }
    }
    at::Tensor ic = at::empty({T, 1}, at::TensorOptions().dtype(at::kInt)), lc = at::empty({T, 1}, at::TensorOptions().dtype(at::kInt));
    std::copy(ia.begin(), ia.end(), ic.data_ptr<int32_t>());
    std::copy(la.begin(), la.end(), lc.data_ptr<int32_t>());
    idx = ic.to(logits.device());
    loc = lc.to(logits.device());
  }
  at::Tensor probs = at::softmax(scores, -1).gather(1, idx.to(at::kLong));
  return {probs, idx, loc, at::zeros({1}, fopt)};
}
HB_REGISTER_OP(moe_balance_assign, "moe_balance_assign", 4, kFlagNondiff, moe_balance_assign_compute, nullptr, moe_gate_deduce, nullptr);


// ------------------------------------------------------------------ expert parallelism over peer memory
// Symmetric buffer of an expert-parallel MoE op (persistent per op and micro-batch, so the forward activations stay
// valid for backward): every rank of `ranks` allocates [experts_per_rank, ep * capacity, hidden] and maps its peers.
struct EpBuffer {
  SymmBuffer* buf = nullptr;
  MoePeers peers;
  at::Tensor local;   // [epr, ep * C, H] view of this rank's copy
};
static bool ep_native(const at::Tensor& x, const std::vector<int64_t>& ranks) {
  return is_native(x) && ranks.size() > 1 && ranks.size() <= 8 && CommRuntime::get().initialized() && env_int("HETU_EP_FUSED", 1) != 0;
}
static EpBuffer ep_buffer(const OpDef& op, RunCtx* rc, const char* tag, const std::vector<int64_t>& ranks64, int64_t epr, int64_t C, int64_t H,
                          const at::TensorOptions& opt) {
  std::vector<int> ranks(ranks64.begin(), ranks64.end());
  const int ep = (int)ranks.size();
  int pos = -1;
  for (int i = 0; i < ep; ++i) if (ranks[i] == CommRuntime::get().rank()) pos = i;
  HB_CHECK(pos >= 0) << "rank " << CommRuntime::get().rank() << " is not in the expert-parallel group of " << op.name();
  const size_t bytes = (size_t)epr * ep * C * H * 2;
  const std::string name = std::string("ep_") + tag + "_" + std::to_string(op.id) + "_mb" + std::to_string(rc ? rc->micro_batch : 0);
  auto& sm = SymmMem::get();
  if (!sm.has(name) || sm.buffer(name).bytes < bytes) symm_exchange_and_open(name, bytes, ranks, pos);
  EpBuffer b;
  b.buf = &sm.buffer(name);
  for (int r = 0; r < ep; ++r) b.peers.base[r] = b.buf->peer[r];
  b.peers.ep = ep; b.peers.experts_per_rank = (int)epr; b.peers.src_rank = pos;
  b.local = at::from_blob(b.buf->local, {epr, (int64_t)ep * C, H}, opt);
  return b;
}
// moe_dispatch: x [T,H], idx, loc -> [E, capacity, H]
static Ts moe_dispatch_compute(const OpDef& op, const Ts& in, RunCtx* rc) {
  const at::Tensor& x = in[0];
  const at::Tensor& idx = in[1];
  const at::Tensor& loc = in[2];
  const int64_t E = op.attrs.i("experts"), C = op.attrs.i("capacity"), H = x.size(1), T = x.size(0), k = idx.size(1);
  const std::vector<int64_t> ep_ranks = op.attrs.ints("ep_ranks");
  const int64_t ep = std::max<int64_t>(1, (int64_t)ep_ranks.size());
  if (x.is_meta()) return {ep > 1 ? at::empty({E / ep, ep * C, H}, x.options()) : at::empty({E, C, H}, x.options())};
  const bool scaled = in.size() > 3;
  if (ep > 1 && ep_native(x, ep_ranks) && x.is_contiguous() && H % 8 == 0) {
    // fused layout transform + all-to-all: every token row is stored straight into its expert's rank over NVLink
    EpBuffer b = ep_buffer(op, rc, "disp", ep_ranks, E / ep, C, H, x.options());
    cudaStream_t st = cur_stream();
    cuda_ok(cudaMemsetAsync(b.buf->local, 0, (size_t)(E / ep) * ep * C * H * 2, st), "memset");
    cuda_ok(symm_barrier(*b.buf, st), "ep barrier");      // every destination buffer is cleared
    cuda_ok(moe_dispatch_peers(x.data_ptr(), idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(),
                               scaled ? in[3].contiguous().data_ptr<float>() : nullptr, b.peers, T, (int)H, (int)k, (int)C, st),
            "moe_dispatch_peers");
    cuda_ok(symm_barrier(*b.buf, st), "ep barrier");      // every rank's tokens have landed
    return {b.local};
  }
  if (ep > 1) {
    // unfused: local layout transform, then an all-to-all of the expert blocks
    OpDef local = op;
    local.attrs.set("ep_ranks", std::vector<int64_t>{});
    at::Tensor disp = moe_dispatch_compute(local, in, rc)[0];
    std::vector<int> ranks(ep_ranks.begin(), ep_ranks.end());
    return {CommRuntime::get().all_to_all(disp, ranks, 0, 1)};
  }
  if (is_native(x) && x.is_contiguous() && H % 8 == 0) {
    at::Tensor out = at::empty({E, C, H}, x.options());
    cuda_ok(moe_dispatch(x.data_ptr(), idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(),
                         scaled ? in[3].contiguous().data_ptr<float>() : nullptr, out.data_ptr(), T, (int)H, (int)E, (int)k,
                         (int)C, cur_stream()), "moe_dispatch");
    return {out};
  }
  at::Tensor out = at::zeros({E * C, H}, x.options());
  at::Tensor flat = idx.to(at::kLong) * C + loc.to(at::kLong);  // [T,k]
  at::Tensor valid = loc >= 0;
  for (int64_t kk = 0; kk < k; ++kk) {
    at::Tensor sel = valid.select(1, kk).nonzero().squeeze(1);
    at::Tensor rows = x.index_select(0, sel);
    if (scaled) rows = rows * in[3].select(1, kk).index_select(0, sel).unsqueeze(1).to(x.scalar_type());
    out.index_copy_(0, flat.select(1, kk).index_select(0, sel), rows);
  }
  return {out.reshape({E, C, H})};
}
// moe_combine: expert_out [E,C,H], idx, loc, [gates] -> [T,H]
static Ts moe_combine_compute(const OpDef& op, const Ts& in, RunCtx* rc) {
  const at::Tensor& idx = in[1];
  const at::Tensor& loc = in[2];
  const std::vector<int64_t> ep_ranks = op.attrs.ints("ep_ranks");
  const int64_t ep = std::max<int64_t>(1, (int64_t)ep_ranks.size());
  const bool gated = in.size() > 3;
  if (ep > 1 && !in[0].is_meta()) {
    const at::Tensor& src = in[0];                      // [E / ep, ep * C, H] on the experts' rank
    const int64_t epr = src.size(0), C = src.size(1) / ep, H = src.size(2), T = idx.size(0), k = idx.size(1);
    if (ep_native(src, ep_ranks) && H % 8 == 0) {
      // fused all-to-all + reverse layout transform: expert outputs are read from their rank over NVLink
      EpBuffer b = ep_buffer(op, rc, "comb", ep_ranks, epr, C, H, src.options());
      cudaStream_t st = cur_stream();
      if (src.data_ptr() != b.buf->local) b.local.copy_(src);
      cuda_ok(symm_barrier(*b.buf, st), "ep barrier");    // every rank published its expert outputs
      at::Tensor y = at::empty({T, H}, src.options());
      cuda_ok(moe_combine_peers(b.peers, idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(),
                                gated ? in[3].contiguous().data_ptr<float>() : nullptr, y.data_ptr(), T, (int)H, (int)k, (int)C, st),
              "moe_combine_peers");
      return {y};
    }
    std::vector<int> ranks(ep_ranks.begin(), ep_ranks.end());
    at::Tensor back = CommRuntime::get().all_to_all(src.contiguous(), ranks, 1, 0);    // [E, C, H]
    OpDef local = op;
    local.attrs.set("ep_ranks", std::vector<int64_t>{});
    Ts lin = in;
    lin[0] = back;
    return moe_combine_compute(local, lin, rc);
  }
  const at::Tensor& eo = in[0];
  const int64_t E = eo.size(0), C = eo.size(1), H = eo.size(2), T = idx.size(0), k = idx.size(1);
  if (eo.is_meta()) return {at::empty({T, H}, eo.options())};
  if (is_native(eo) && eo.is_contiguous() && H % 8 == 0) {
    at::Tensor y = at::empty({T, H}, eo.options());
    cuda_ok(moe_combine(eo.data_ptr(), idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(),
                        gated ? in[3].contiguous().data_ptr<float>() : nullptr, y.data_ptr(), T, (int)H, (int)E, (int)k, (int)C,
                        cur_stream()), "moe_combine");
    return {y};
  }
  at::Tensor flat = (idx.to(at::kLong) * C + loc.to(at::kLong).clamp_min(0)).reshape({-1});
  at::Tensor rows = eo.reshape({E * C, H}).index_select(0, flat).reshape({T, k, H}).to(at::kFloat);
  at::Tensor w = (loc >= 0).to(at::kFloat);
  if (gated) w = w * in[3];
  return {(rows * w.unsqueeze(-1)).sum(1).to(eo.scalar_type())};
}
// d gates of combine: <dy[t], expert_out[e_k, slot_k]>
static Ts moe_combine_gate_grad_compute(const OpDef& op, const Ts& in, RunCtx* rc) {
  const at::Tensor& dy = in[0];
  const at::Tensor& idx = in[2];
  const at::Tensor& loc = in[3];
  const std::vector<int64_t> ep_ranks = op.attrs.ints("ep_ranks");
  const int64_t ep = std::max<int64_t>(1, (int64_t)ep_ranks.size());
  if (ep > 1 && !dy.is_meta()) {
    const at::Tensor& src = in[1];
    const int64_t epr = src.size(0), C = src.size(1) / ep, H = src.size(2), T = idx.size(0), k = idx.size(1);
    if (ep_native(src, ep_ranks) && is_native(dy) && dy.is_contiguous() && H % 8 == 0) {
      EpBuffer b = ep_buffer(op, rc, "gateg", ep_ranks, epr, C, H, src.options());
      cudaStream_t st = cur_stream();
      if (src.data_ptr() != b.buf->local) b.local.copy_(src);
      cuda_ok(symm_barrier(*b.buf, st), "ep barrier");
      at::Tensor dg = at::empty({T, k}, dy.options().dtype(at::kFloat));
      cuda_ok(moe_combine_bwd_gate_peers(dy.data_ptr(), b.peers, idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(), dg.data_ptr<float>(), T,
                                         (int)H, (int)k, (int)C, st), "moe_combine_bwd_gate_peers");
      return {dg};
    }
    std::vector<int> ranks(ep_ranks.begin(), ep_ranks.end());
    at::Tensor back = CommRuntime::get().all_to_all(src.contiguous(), ranks, 1, 0);
    OpDef local = op;
    local.attrs.set("ep_ranks", std::vector<int64_t>{});
    Ts lin = in;
    lin[1] = back;
    return moe_combine_gate_grad_compute(local, lin, rc);
  }
  const at::Tensor& eo = in[1];
  const int64_t E = eo.size(0), C = eo.size(1), H = eo.size(2), T = idx.size(0), k = idx.size(1);
  if (dy.is_meta()) return {at::empty({T, k}, dy.options().dtype(at::kFloat))};
  if (is_native(eo) && is_native(dy) && eo.is_contiguous() && dy.is_contiguous() && H % 8 == 0) {
    at::Tensor dg = at::empty({T, k}, dy.options().dtype(at::kFloat));
    cuda_ok(moe_combine_bwd_gate(dy.data_ptr(), eo.data_ptr(), idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(),
                                 dg.data_ptr<float>(), T, (int)H, (int)E, (int)k, (int)C, cur_stream()), "moe_combine_bwd_gate");
    return {dg};
  }
  at::Tensor flat = (idx.to(at::kLong) * C + loc.to(at::kLong).clamp_min(0)).reshape({-1});
  at::Tensor rows = eo.reshape({E * C, H}).index_select(0, flat).reshape({T, k, H}).to(at::kFloat);
  return {(rows * dy.to(at::kFloat).unsqueeze(1)).sum(-1) * (loc >= 0).to(at::kFloat)};
}
static TensorList moe_dispatch_grad(OpDef& op, const TensorList& g) {
  // dx = combine(d_dispatched) with the same scale; d scale is not propagated (gates flow through combine)
  TensorList ins = {g[0], op.inputs[1], op.inputs[2]};
  if (op.inputs.size() > 3) ins.push_back(op.inputs[3]);
  TensorList r(op.inputs.size());
  AttrMap a;
  a.set("ep_ranks", op.attrs.ints("ep_ranks"));
  r[0] = op.graph->make_op1("moe_combine", ins, a);
  return r;
}
static TensorList moe_combine_grad(OpDef& op, const TensorList& g) {
  AttrMap a;
  const std::vector<int64_t> epr = op.attrs.ints("ep_ranks");
  const int64_t ep = std::max<int64_t>(1, (int64_t)epr.size());
  a.set("experts", op.inputs[0]->shape[0] * ep);
  a.set("capacity", op.inputs[0]->shape[1] / ep);
  a.set("ep_ranks", epr);
  TensorList ins = {g[0], op.inputs[1], op.inputs[2]};
  if (op.inputs.size() > 3) ins.push_back(op.inputs[3]);
  TensorList r(op.inputs.size());
  r[0] = op.graph->make_op1("moe_dispatch", ins, a);
  if (op.inputs.size() > 3) {
    AttrMap ga;
    ga.set("ep_ranks", epr);
    r[3] = op.graph->make_op1("moe_combine_gate_grad", {g[0], op.inputs[0], op.inputs[1], op.inputs[2]}, ga);
  }
  return r;
}
// expert-major buffers are rank-local (no global layout); token-major results take the layout of the routing tensors
static void moe_no_ds(OpDef&, size_t) {}
HB_REGISTER_OP(moe_dispatch, "moe_dispatch", 1, 0, moe_dispatch_compute, moe_dispatch_grad, moe_no_ds, nullptr);
HB_REGISTER_OP(moe_combine, "moe_combine", 1, 0, moe_combine_compute, moe_combine_grad,
               [](OpDef& op, size_t s) { copy_out_ds(op, 0, s, op.inputs[1]); if (op.outputs[0]->has_ds(s) && op.inputs[1]->has_ds(s)) {} }, nullptr);
HB_REGISTER_OP(moe_combine_gate_grad, "moe_combine_gate_grad", 1, kFlagNondiff, moe_combine_gate_grad_compute, nullptr,
               [](OpDef& op, size_t s) { copy_out_ds(op, 0, s, op.inputs[2]); }, nullptr);

// differentiable gate values: gates = normalise(softmax(logits)[topk]) for fixed routing decisions
static Ts moe_gate_values_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& logits = in[0];
  const at::Tensor& idx = in[1];
  const at::Tensor& loc = in[2];
  (void)op;
  if (logits.is_meta()) return {at::empty(idx.sizes(), logits.options().dtype(at::kFloat))};
  at::Tensor probs = at::softmax(logits.to(at::kFloat), -1);
  at::Tensor val = probs.gather(1, idx.to(at::kLong)) * (loc >= 0).to(at::kFloat);
  if (idx.size(1) > 1) val = val / val.sum(-1, true).clamp_min(1e-9);
  return {val};
}
HB_REGISTER_OP(moe_gate_values, "moe_gate_values", 1, 0, moe_gate_values_compute, nullptr,
               [](OpDef& op, size_t s) { copy_out_ds(op, 0, s, op.inputs[0]); }, nullptr);

}  // namespace hb

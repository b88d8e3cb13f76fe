#include "ps_server.h"

#include <algorithm>
#include <chrono>
#include <cmath>

namespace hb {

ParameterServer::ParameterServer(int num_workers) : num_workers_(num_workers), clocks_(num_workers, 0) {
  HB_CHECK(num_workers > 0) << "parameter server needs at least one worker";
}

void ParameterServer::apply(const PsParamConfig& c, int64_t step, float* w, float* s1, float* s2, const float* g, int64_t n) {
  switch (c.opt) {
    case PsOptimizer::NONE:
      for (int64_t i = 0; i < n; ++i) w[i] += g[i];     // raw accumulation (the worker already scaled the update)
      break;
    case PsOptimizer::SGD:
      for (int64_t i = 0; i < n; ++i) w[i] -= c.lr * g[i];
      break;
    case PsOptimizer::MOMENTUM:
      for (int64_t i = 0; i < n; ++i) { s1[i] = c.momentum * s1[i] - c.lr * g[i]; w[i] += s1[i]; }
      break;
    case PsOptimizer::ADAGRAD:
      for (int64_t i = 0; i < n; ++i) { s1[i] += g[i] * g[i]; w[i] -= c.lr * g[i] / (std::sqrt(s1[i]) + c.eps); }
      break;
    case PsOptimizer::ADAM: {
      const double bc1 = 1.0 - std::pow((double)c.beta1, (double)step), bc2 = 1.0 - std::pow((double)c.beta2, (double)step);
      for (int64_t i = 0; i < n; ++i) {
        s1[i] = c.beta1 * s1[i] + (1 - c.beta1) * g[i];
        s2[i] = c.beta2 * s2[i] + (1 - c.beta2) * g[i] * g[i];
        w[i] -= (float)(c.lr * (s1[i] / bc1) / (std::sqrt(s2[i] / bc2) + c.eps));
      }
      break;
    }
  }
}

void ParameterServer::init_dense(int64_t key, const std::vector<float>& value, const PsParamConfig& cfg) {
  std::lock_guard<std::mutex> lk(mu_);
  if (dense_.count(key)) return;          // first initialiser wins (all workers call init)
  Dense d;
  d.w = value; d.cfg = cfg;
  if (cfg.opt != PsOptimizer::SGD && cfg.opt != PsOptimizer::NONE) d.s1.assign(value.size(), 0.f);
  if (cfg.opt == PsOptimizer::ADAM) d.s2.assign(value.size(), 0.f);
  dense_[key] = std::move(d);
}
void ParameterServer::push_dense(int64_t key, const std::vector<float>& grad) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = dense_.find(key);
  HB_CHECK(it != dense_.end()) << "PS: dense key " << key << " not initialised";
  Dense& d = it->second;
  HB_CHECK(grad.size() == d.w.size()) << "PS: gradient size " << grad.size() << " != parameter size " << d.w.size();
  ++d.step; ++n_push_;
  apply(d.cfg, d.step, d.w.data(), d.s1.data(), d.s2.data(), grad.data(), (int64_t)grad.size());
}
std::vector<float> ParameterServer::pull_dense(int64_t key) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = dense_.find(key);
  HB_CHECK(it != dense_.end()) << "PS: dense key " << key << " not initialised";
  ++n_pull_;
  return it->second.w;
}
std::vector<float> ParameterServer::push_pull_dense(int64_t key, const std::vector<float>& grad) {
  push_dense(key, grad);
  return pull_dense(key);
}

void ParameterServer::init_sparse(int64_t key, int64_t rows, int width, const std::vector<float>& value, const PsParamConfig& cfg) {
  std::lock_guard<std::mutex> lk(mu_);
  if (sparse_.count(key)) return;
  HB_CHECK((int64_t)value.size() == rows * width) << "PS: sparse initial value has the wrong size";
  Sparse s;
  s.w = value; s.rows = rows; s.width = width; s.cfg = cfg; s.version.assign(rows, 0);
  if (cfg.opt != PsOptimizer::SGD && cfg.opt != PsOptimizer::NONE) s.s1.assign(value.size(), 0.f);
  if (cfg.opt == PsOptimizer::ADAM) s.s2.assign(value.size(), 0.f);
  sparse_[key] = std::move(s);
}
void ParameterServer::push_sparse(int64_t key, const std::vector<int64_t>& rows, const std::vector<float>& grads) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = sparse_.find(key);
  HB_CHECK(it != sparse_.end()) << "PS: sparse key " << key << " not initialised";
  Sparse& s = it->second;
  HB_CHECK((int64_t)grads.size() == (int64_t)rows.size() * s.width) << "PS: sparse gradient has the wrong size";
  ++s.step; ++n_push_;
  for (size_t i = 0; i < rows.size(); ++i) {
    const int64_t r = rows[i];
    HB_CHECK(r >= 0 && r < s.rows) << "PS: row " << r << " out of range";
    const int64_t o = r * s.width;
    apply(s.cfg, s.step, s.w.data() + o, s.s1.empty() ? nullptr : s.s1.data() + o, s.s2.empty() ? nullptr : s.s2.data() + o,
          grads.data() + i * s.width, s.width);
    ++s.version[r];
  }
}
std::vector<float> ParameterServer::pull_sparse(int64_t key, const std::vector<int64_t>& rows) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = sparse_.find(key);
  HB_CHECK(it != sparse_.end()) << "PS: sparse key " << key << " not initialised";
  Sparse& s = it->second;
  ++n_pull_;
  std::vector<float> out(rows.size() * s.width);
  for (size_t i = 0; i < rows.size(); ++i) {
    HB_CHECK(rows[i] >= 0 && rows[i] < s.rows) << "PS: row " << rows[i] << " out of range";
    std::copy_n(s.w.data() + rows[i] * s.width, s.width, out.data() + i * s.width);
  }
  return out;
}
std::vector<int64_t> ParameterServer::row_versions(int64_t key, const std::vector<int64_t>& rows) {
  std::lock_guard<std::mutex> lk(mu_);
  Sparse& s = sparse_.at(key);
  std::vector<int64_t> v(rows.size());
  for (size_t i = 0; i < rows.size(); ++i) v[i] = s.version[rows[i]];
  return v;
}
void ParameterServer::sync_cache(int64_t key, const std::vector<int64_t>& rows, const std::vector<int64_t>& client_versions, int64_t bound,
                                 std::vector<int64_t>* stale_rows, std::vector<float>* fresh_values, std::vector<int64_t>* fresh_versions) {
  std::lock_guard<std::mutex> lk(mu_);
  Sparse& s = sparse_.at(key);
  for (size_t i = 0; i < rows.size(); ++i) {
    const int64_t r = rows[i];
    if (s.version[r] - client_versions[i] > bound) {
      stale_rows->push_back(r);
      fresh_versions->push_back(s.version[r]);
      fresh_values->insert(fresh_values->end(), s.w.begin() + r * s.width, s.w.begin() + (r + 1) * s.width);
    }
  }
}

void ParameterServer::barrier(int) {
  std::unique_lock<std::mutex> lk(mu_);
  const int64_t gen = bar_gen_;
  if (++bar_count_ == num_workers_) {
    bar_count_ = 0;
    ++bar_gen_;
    bar_cv_.notify_all();
  } else bar_cv_.wait(lk, [&] { return bar_gen_ != gen; });
}
void ParameterServer::ssp_init(int staleness) {
  std::lock_guard<std::mutex> lk(mu_);
  staleness_ = staleness;
  std::fill(clocks_.begin(), clocks_.end(), 0);
}
void ParameterServer::ssp_sync(int worker, int clock) {
  std::unique_lock<std::mutex> lk(mu_);
  clocks_[worker] = clock;
  ssp_cv_.notify_all();
  ssp_cv_.wait(lk, [&] { return clock - *std::min_element(clocks_.begin(), clocks_.end()) <= staleness_; });
}

std::vector<float> ParameterServer::preduce(int worker, int64_t key, const std::vector<float>& value, int min_workers, int wait_ms,
                                            std::vector<int>* partners) {
  std::unique_lock<std::mutex> lk(mu_);
  ++n_preduce_;
  PGroup& g = open_[key];
  if (g.members.empty()) { g.sum.assign(value.size(), 0.f); g.id = next_group_++; }
  HB_CHECK(g.sum.size() == value.size()) << "PS: preduce size mismatch";
  g.members.push_back(worker);
  for (size_t i = 0; i < value.size(); ++i) g.sum[i] += value[i];
  const int64_t gid = g.id;
  auto close = [&](int64_t k) {
    PGroup fin = std::move(open_[k]);
    open_.erase(k);
    fin.closed = true;
    for (float& v : fin.sum) v /= (float)fin.members.size();
    done_[fin.id] = std::move(fin);
    pr_cv_.notify_all();
  };
  if ((int)g.members.size() >= std::min(min_workers, num_workers_) && (int)g.members.size() == num_workers_) close(key);
  else if (g.members.size() == 1) {
    // the first arrival waits out the window, then closes the group with whoever joined (at least min_workers if they come)
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(wait_ms);
    pr_cv_.wait_until(lk, deadline, [&] { return done_.count(gid) > 0; });
    if (!done_.count(gid)) {
      // keep waiting for the minimum group size, but no longer than 20 windows
      const auto hard = std::chrono::steady_clock::now() + std::chrono::milliseconds(wait_ms * 20);
      pr_cv_.wait_until(lk, hard, [&] { return done_.count(gid) > 0 || (open_.count(key) && (int)open_[key].members.size() >= min_workers); });
      if (!done_.count(gid)) close(key);
    }
  } else if ((int)g.members.size() >= num_workers_) close(key);
  else pr_cv_.notify_all();
  pr_cv_.wait(lk, [&] { return done_.count(gid) > 0; });
  PGroup& d = done_[gid];
  std::vector<float> out = d.sum;
  if (partners) *partners = d.members;
  if (++d.taken == (int)d.members.size()) done_.erase(gid);
  return out;
}

std::map<std::string, int64_t> ParameterServer::stats() const {
  std::lock_guard<std::mutex> lk(mu_);
  return {{"pushes", n_push_}, {"pulls", n_pull_}, {"preduces", n_preduce_}, {"dense_keys", (int64_t)dense_.size()},
          {"sparse_keys", (int64_t)sparse_.size()}};
}

}  // namespace hb

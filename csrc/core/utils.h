// Small host-side utilities of the runtime:
//   ContextStore -- typed key/value scratch space an op's forward leaves for its backward (or one executor phase for the
//                   next), with pop / migrate semantics            (ref: hetu/utils/context_store.h)
//   TaskQueue    -- bounded multi-producer queue drained by N worker threads; used for work that must not block the
//                   step loop (checkpoint shards, log writers, deferred frees)   (ref: hetu/utils/task_queue.h)
#pragma once
#include <ATen/ATen.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <variant>
#include <vector>

#include "base.h"

namespace hb {

class ContextStore {
 public:
  using Value = std::variant<bool, int64_t, double, std::string, std::vector<int64_t>, std::vector<double>, at::Tensor>;

  template <typename T>
  void put(const std::string& key, T value) { m_[key] = Value(std::move(value)); }
  template <typename T>
  bool contains(const std::string& key) const {
    auto it = m_.find(key);
    return it != m_.end() && std::holds_alternative<T>(it->second);
  }
  bool has(const std::string& key) const { return m_.count(key) > 0; }
  template <typename T>
  T get(const std::string& key) const {
    auto it = m_.find(key);
    HB_CHECK(it != m_.end()) << "context store: no entry '" << key << "'";
    const T* p = std::get_if<T>(&it->second);
    HB_CHECK(p != nullptr) << "context store: entry '" << key << "' holds another type";
    return *p;
  }
  template <typename T>
  T get_or(const std::string& key, T dflt) const {
    auto it = m_.find(key);
    if (it == m_.end()) return dflt;
    const T* p = std::get_if<T>(&it->second);
    return p ? *p : dflt;
  }
  // read and remove: saved activations are released the moment the backward consumed them
  template <typename T>
  T pop(const std::string& key) {
    T v = get<T>(key);
    m_.erase(key);
    return v;
  }
  bool erase(const std::string& key) { return m_.erase(key) > 0; }
  // move an entry over from `src` (optionally under a new key); an existing entry is only overwritten by a present source
  void migrate_from(ContextStore& src, const std::string& key, const std::string& new_key = "") {
    auto it = src.m_.find(key);
    if (it == src.m_.end()) return;
    m_[new_key.empty() ? key : new_key] = std::move(it->second);
    src.m_.erase(it);
  }
  std::vector<std::string> keys() const {
    std::vector<std::string> k;
    for (auto& kv : m_) k.push_back(kv.first);
    return k;
  }
  const Value& raw(const std::string& key) const {
    auto it = m_.find(key);
    HB_CHECK(it != m_.end()) << "context store: no entry '" << key << "'";
    return it->second;
  }
  size_t size() const { return m_.size(); }
  void clear() { m_.clear(); }

 private:
  std::map<std::string, Value> m_;
};

class TaskQueue {
 public:
  TaskQueue(std::string name, int num_workers, size_t max_pending = 1024)
      : name_(std::move(name)), max_pending_(std::max<size_t>(1, max_pending)) {
    HB_CHECK(num_workers > 0) << "task queue '" << name_ << "' needs at least one worker";
    for (int i = 0; i < num_workers; ++i) workers_.emplace_back([this] { worker(); });
  }
  ~TaskQueue() { shutdown(); }
  TaskQueue(const TaskQueue&) = delete;
  TaskQueue& operator=(const TaskQueue&) = delete;

  // blocks while `max_pending` tasks are queued (back-pressure on the producer)
  void add(std::function<void()> task) {
    std::unique_lock<std::mutex> lk(mu_);
    HB_CHECK(!stop_) << "task queue '" << name_ << "' was shut down";
    cv_space_.wait(lk, [&] { return stop_ || tasks_.size() < max_pending_; });
    HB_CHECK(!stop_) << "task queue '" << name_ << "' was shut down";
    tasks_.push_back(std::move(task));
    ++submitted_;
    cv_task_.notify_one();
  }
  // all tasks submitted so far have finished; the first exception a task raised is re-thrown here
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return finished_ == submitted_; });
    if (!error_.empty()) {
      std::string e;
      e.swap(error_);
      throw Error("task queue '" + name_ + "': " + e);
    }
  }
  // drains the queue, then joins the workers
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (stop_) return;
      stop_ = true;
    }
    cv_task_.notify_all();
    cv_space_.notify_all();
    for (auto& t : workers_)
      if (t.joinable()) t.join();
  }
  int num_workers() const { return (int)workers_.size(); }
  bool running() const { return !stop_; }
  size_t pending() const {
    std::lock_guard<std::mutex> lk(mu_);
    return tasks_.size();
  }
  int64_t completed() const { return finished_; }
  const std::string& name() const { return name_; }

 private:
  void worker() {
    for (;;) {
      std::function<void()> task;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_task_.wait(lk, [&] { return stop_ || !tasks_.empty(); });
        if (tasks_.empty()) return;          // stop requested and nothing left
        task = std::move(tasks_.front());
        tasks_.pop_front();
        cv_space_.notify_one();
      }
      std::string err;
      try {
        task();
      } catch (const std::exception& e) {
        err = e.what();
      } catch (...) {
        err = "unknown exception";
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!err.empty() && error_.empty()) error_ = err;
        ++finished_;
      }
      cv_done_.notify_all();
    }
  }
  std::string name_;
  size_t max_pending_;
  mutable std::mutex mu_;
  std::condition_variable cv_task_, cv_space_, cv_done_;
  std::deque<std::function<void()>> tasks_;
  std::vector<std::thread> workers_;
  int64_t submitted_ = 0, finished_ = 0;
  std::string error_;
  bool stop_ = false;
};

}  // namespace hb

// Lazy integer expressions for symbolic shapes (per-micro-batch sequence lengths,
// batch sizes, CP group ids ...).  A leaf holds a value that the trainer updates
// before each run; derived symbols re-evaluate on demand.
// (capability parity: hetu/core/symbol.h:10-178)
#pragma once
#include <memory>
#include <vector>

#include "base.h"

namespace hb {

enum class SymOp : int8_t { LEAF = 0, ADD, SUB, MUL, DIV, REM };

class IntSymbolDef {
 public:
  IntSymbolDef() : op_(SymOp::LEAF), has_val_(false), val_(0) {}
  explicit IntSymbolDef(int64_t v) : op_(SymOp::LEAF), has_val_(true), val_(v) {}
  IntSymbolDef(SymOp op, std::shared_ptr<IntSymbolDef> a, std::shared_ptr<IntSymbolDef> b)
      : op_(op), has_val_(false), val_(0), a_(std::move(a)), b_(std::move(b)) {}
  bool is_leaf() const { return op_ == SymOp::LEAF; }
  bool is_instantiated() const {
    if (is_leaf()) return has_val_;
    return a_->is_instantiated() && b_->is_instantiated();
  }
  int64_t get_val() const {
    if (is_leaf()) {
      HB_CHECK(has_val_) << "symbol has no value yet";
      return val_;
    }
    const int64_t x = a_->get_val(), y = b_->get_val();
    switch (op_) {
      case SymOp::ADD: return x + y;
      case SymOp::SUB: return x - y;
      case SymOp::MUL: return x * y;
      case SymOp::DIV: HB_CHECK(y != 0) << "symbolic division by zero"; return x / y;
      case SymOp::REM: HB_CHECK(y != 0) << "symbolic modulo by zero"; return x % y;
      default: HB_FAIL() << "bad symbol op";
    }
  }
  void set_val(int64_t v) {
    HB_CHECK(is_leaf()) << "only leaf symbols can be assigned";
    has_val_ = true;
    val_ = v;
  }
  void reset() {
    HB_CHECK(is_leaf()) << "only leaf symbols can be reset";
    has_val_ = false;
  }

 private:
  SymOp op_;
  bool has_val_;
  int64_t val_;
  std::shared_ptr<IntSymbolDef> a_, b_;
};

class IntSymbol {
 public:
  IntSymbol() : p_(std::make_shared<IntSymbolDef>()) {}
  IntSymbol(int64_t v) : p_(std::make_shared<IntSymbolDef>(v)) {}  // NOLINT implicit
  explicit IntSymbol(std::shared_ptr<IntSymbolDef> p) : p_(std::move(p)) {}
  int64_t get_val() const { return p_->get_val(); }
  void set_val(int64_t v) { p_->set_val(v); }
  void reset() { p_->reset(); }
  bool is_leaf() const { return p_->is_leaf(); }
  bool is_instantiated() const { return p_->is_instantiated(); }
  const std::shared_ptr<IntSymbolDef>& ptr() const { return p_; }
  IntSymbol operator+(const IntSymbol& o) const { return IntSymbol(std::make_shared<IntSymbolDef>(SymOp::ADD, p_, o.p_)); }
  IntSymbol operator-(const IntSymbol& o) const { return IntSymbol(std::make_shared<IntSymbolDef>(SymOp::SUB, p_, o.p_)); }
  IntSymbol operator*(const IntSymbol& o) const { return IntSymbol(std::make_shared<IntSymbolDef>(SymOp::MUL, p_, o.p_)); }
  IntSymbol operator/(const IntSymbol& o) const { return IntSymbol(std::make_shared<IntSymbolDef>(SymOp::DIV, p_, o.p_)); }
  IntSymbol operator%(const IntSymbol& o) const { return IntSymbol(std::make_shared<IntSymbolDef>(SymOp::REM, p_, o.p_)); }

 private:
  std::shared_ptr<IntSymbolDef> p_;
};

using SyShape = std::vector<IntSymbol>;
inline std::vector<int64_t> sy_shape_values(const SyShape& s) {
  std::vector<int64_t> v;
  v.reserve(s.size());
  for (auto& e : s) v.push_back(e.get_val());
  return v;
}

}  // namespace hb

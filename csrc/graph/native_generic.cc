// at::Tensor front-ends of the generic device kernels (csrc/kernels/generic.cu).  Every function returns an undefined tensor
// when the native kernel does not apply (CPU tensor, integer dtype, rank > 8, HETU_NATIVE_GENERIC=0 ...); the caller then takes
// the ATen path, so the op semantics never depend on which one ran.
#include <ATen/core/grad_mode.h>

#include <algorithm>

#include "../kernels/kernels.h"
#include "op_utils.h"

namespace hb {

static bool generic_enabled() {
  static const bool on = env_int("HETU_NATIVE_GENERIC", 1) != 0;
  return on;
}
static int gd_of(at::ScalarType t) {
  switch (t) {
    case at::kFloat: return GD_F32;
    case at::kBFloat16: return GD_BF16;
    case at::kHalf: return GD_F16;
    case at::kLong: return GD_I64;
    case at::kInt: return GD_I32;
    case at::kByte: case at::kBool: return GD_U8;
    default: return -1;
  }
}
// Ops without a hand-written gradient are differentiated by re-running their compute function under ATen autograd
// (graph.cc `autograd_vjp`): there the result must carry an autograd history, so the native kernels step aside.
static bool traced(const at::Tensor& t) { return at::GradMode::is_enabled() && t.defined() && t.requires_grad(); }
static bool float_cuda(const at::Tensor& t) {
  return generic_enabled() && t.defined() && t.is_cuda() && t.numel() > 0 && !traced(t) &&
         (t.scalar_type() == at::kFloat || t.scalar_type() == at::kBFloat16 || t.scalar_type() == at::kHalf);
}

at::Tensor g_contiguous(const at::Tensor& x) {
  if (!generic_enabled() || !x.is_cuda() || x.is_contiguous() || x.numel() == 0 || x.dim() > 8 || x.numel() >= (1ll << 31) || traced(x))
    return at::Tensor();
  at::Tensor out = at::empty(x.sizes(), x.options());
  const auto dst = out.strides();
  if (strided_copy((int)x.element_size(), x.data_ptr(), out.data_ptr(), (int)x.dim(), x.sizes().data(), x.strides().data(), dst.data(),
                   cur_stream()) != cudaSuccess) {
    cudaGetLastError();
    return at::Tensor();
  }
  return out;
}

at::Tensor g_unary(int op, const at::Tensor& x, float p0, float p1) {
  if (!float_cuda(x)) return at::Tensor();
  at::Tensor xc = x;
  if (!x.is_contiguous()) {
    xc = g_contiguous(x);
    if (!xc.defined()) return at::Tensor();
  }
  at::Tensor out = at::empty(xc.sizes(), xc.options());
  if (generic_unary(op, gd_of(xc.scalar_type()), xc.data_ptr(), out.data_ptr(), xc.numel(), p0, p1, cur_stream()) != cudaSuccess) {
    cudaGetLastError();
    return at::Tensor();
  }
  return out;
}

at::Tensor g_binary(int op, const at::Tensor& a, const at::Tensor& b) {
  if (!float_cuda(a) || !float_cuda(b) || a.scalar_type() != b.scalar_type()) return at::Tensor();
  std::vector<int64_t> shape;
  try {
    shape = at::infer_size(a.sizes(), b.sizes());
  } catch (...) {
    return at::Tensor();
  }
  if (shape.size() > 8) return at::Tensor();
  const int nd = (int)shape.size();
  auto bstrides = [&](const at::Tensor& t) {
    std::vector<int64_t> st(nd, 0);
    const int lead = nd - (int)t.dim();
    for (int d = 0; d < (int)t.dim(); ++d) st[lead + d] = t.size(d) == 1 ? 0 : t.stride(d);
    return st;
  };
  const auto as = bstrides(a), bs = bstrides(b);
  at::Tensor out = at::empty(shape, a.options());
  if (out.numel() == 0) return out;
  if (generic_binary(op, gd_of(a.scalar_type()), a.data_ptr(), b.data_ptr(), out.data_ptr(), nd, shape.data(), as.data(), bs.data(),
                     cur_stream()) != cudaSuccess) {
    cudaGetLastError();
    return at::Tensor();
  }
  return out;
}

at::Tensor g_reduce(int mode, const at::Tensor& x, std::vector<int64_t> axes, bool keepdims) {
  if (!float_cuda(x) || x.dim() == 0 || x.dim() > 8) return at::Tensor();
  const int nd = (int)x.dim();
  if (axes.empty()) for (int i = 0; i < nd; ++i) axes.push_back(i);
  for (auto& a : axes) if (a < 0) a += nd;
  std::sort(axes.begin(), axes.end());
  axes.erase(std::unique(axes.begin(), axes.end()), axes.end());
  for (auto a : axes) if (a < 0 || a >= nd) return at::Tensor();
  // the reduced axes must form one contiguous run [lo, hi] of a contiguous tensor; otherwise permute them to the back first
  at::Tensor xc = x;
  bool run = true;
  for (size_t i = 1; i < axes.size(); ++i) run = run && axes[i] == axes[i - 1] + 1;
  int64_t outer = 1, red = 1, inner = 1;
  if (run) {
    if (!xc.is_contiguous()) {
      xc = g_contiguous(x);
      if (!xc.defined()) return at::Tensor();
    }
    for (int d = 0; d < nd; ++d) {
      if (d < axes.front()) outer *= xc.size(d);
      else if (d > axes.back()) inner *= xc.size(d);
      else red *= xc.size(d);
    }
  } else {
    std::vector<int64_t> perm;
    for (int d = 0; d < nd; ++d) if (!std::binary_search(axes.begin(), axes.end(), (int64_t)d)) perm.push_back(d);
    for (auto a : axes) perm.push_back(a);
    at::Tensor p = x.permute(perm);
    xc = p.is_contiguous() ? p : g_contiguous(p);
    if (!xc.defined()) return at::Tensor();
    for (auto a : axes) red *= x.size(a);
    outer = x.numel() / std::max<int64_t>(red, 1);
  }
  if (red == 0) return at::Tensor();
  std::vector<int64_t> oshape;
  for (int d = 0; d < nd; ++d) {
    const bool r = std::binary_search(axes.begin(), axes.end(), (int64_t)d);
    if (!r) oshape.push_back(x.size(d));
    else if (keepdims) oshape.push_back(1);
  }
  at::Tensor out = at::empty(oshape, x.options());
  if (out.numel() == 0) return out;
  at::Tensor ws;
  const int64_t ws_floats = generic_reduce_workspace_floats(outer, red, inner);
  if (ws_floats > 0) ws = at::empty({ws_floats}, x.options().dtype(at::kFloat));
  if (generic_reduce(mode, gd_of(x.scalar_type()), xc.data_ptr(), out.data_ptr(), ws.defined() ? ws.data_ptr<float>() : nullptr, outer, red, inner,
                     cur_stream()) != cudaSuccess) {
    cudaGetLastError();
    return at::Tensor();
  }
  return out;
}

at::Tensor g_softmax(bool log, const at::Tensor& x, int64_t dim) {
  if (!float_cuda(x) || x.dim() == 0) return at::Tensor();
  const int nd = (int)x.dim();
  if (dim < 0) dim += nd;
  if (dim < 0 || dim >= nd) return at::Tensor();
  at::Tensor xc = x.is_contiguous() ? x : g_contiguous(x);
  if (!xc.defined()) return at::Tensor();
  int64_t outer = 1, inner = 1;
  for (int d = 0; d < dim; ++d) outer *= xc.size(d);
  for (int d = (int)dim + 1; d < nd; ++d) inner *= xc.size(d);
  at::Tensor out = at::empty(xc.sizes(), xc.options());
  if (generic_softmax(log, gd_of(x.scalar_type()), xc.data_ptr(), out.data_ptr(), outer, xc.size(dim), inner, cur_stream()) != cudaSuccess) {
    cudaGetLastError();
    return at::Tensor();
  }
  return out;
}

at::Tensor g_concat(const std::vector<at::Tensor>& in, int64_t dim) {
  if (!generic_enabled() || in.empty() || !in[0].is_cuda() || in[0].dim() == 0 || in[0].dim() > 8) return at::Tensor();
  const int nd = (int)in[0].dim();
  if (dim < 0) dim += nd;
  if (dim < 0 || dim >= nd) return at::Tensor();
  std::vector<int64_t> shape = in[0].sizes().vec();
  int64_t total = 0;
  for (auto& t : in) {
    if (!t.is_cuda() || t.scalar_type() != in[0].scalar_type() || t.dim() != nd || t.device() != in[0].device() || traced(t)) return at::Tensor();
    for (int d = 0; d < nd; ++d) if (d != dim && t.size(d) != shape[d]) return at::Tensor();
    total += t.size(dim);
  }
  shape[dim] = total;
  at::Tensor out = at::empty(shape, in[0].options());
  if (out.numel() >= (1ll << 31)) return at::Tensor();
  int64_t at_dim = 0;
  for (auto& t : in) {
    if (t.numel() > 0) {
      at::Tensor view = out.narrow(dim, at_dim, t.size(dim));
      if (strided_copy((int)t.element_size(), t.data_ptr(), view.data_ptr(), nd, t.sizes().data(), t.strides().data(), view.strides().data(),
                       cur_stream()) != cudaSuccess) {
        cudaGetLastError();
        return at::Tensor();
      }
    }
    at_dim += t.size(dim);
  }
  return out;
}

at::Tensor g_cast(const at::Tensor& x, at::ScalarType to) {
  if (!generic_enabled() || !x.is_cuda() || x.numel() == 0 || x.scalar_type() == to || traced(x)) return at::Tensor();
  const int s = gd_of(x.scalar_type()), d = gd_of(to);
  if (s < 0 || d < 0 || d == GD_U8 || x.scalar_type() == at::kBool) return at::Tensor();
  at::Tensor xc = x.is_contiguous() ? x : g_contiguous(x);
  if (!xc.defined()) return at::Tensor();
  at::Tensor out = at::empty(xc.sizes(), xc.options().dtype(to));
  if (generic_cast(s, d, xc.data_ptr(), out.data_ptr(), xc.numel(), cur_stream()) != cudaSuccess) {
    cudaGetLastError();
    return at::Tensor();
  }
  return out;
}

at::Tensor g_full(at::IntArrayRef shape, const at::TensorOptions& opt, double value) {
  if (!generic_enabled() || !opt.device().is_cuda()) return at::Tensor();
  const int d = gd_of(c10::typeMetaToScalarType(opt.dtype()));
  if (d < 0) return at::Tensor();
  at::Tensor out = at::empty(shape, opt);
  if (generic_fill(d, out.data_ptr(), out.numel(), value, cur_stream()) != cudaSuccess) {
    cudaGetLastError();
    return at::Tensor();
  }
  return out;
}

}  // namespace hb

// Test-infrastructure shim: read-only/read-write absl::Span over contiguous memory.
#pragma once
#include <cstddef>
#include <type_traits>
namespace absl {
template <typename T>
class Span {
 public:
  using size_type = std::size_t;
  using value_type = std::remove_cv_t<T>;
  Span() = default;
  Span(T* p, size_type n) : p_(p), n_(n) {}
  template <std::size_t N>
  Span(T (&a)[N]) : p_(a), n_(N) {}  // NOLINT
  T* data() const { return p_; }
  size_type size() const { return n_; }
  bool empty() const { return n_ == 0; }
  T& operator[](size_type i) const { return p_[i]; }
  T* begin() const { return p_; }
  T* end() const { return p_ + n_; }
  T& back() const { return p_[n_ - 1]; }
  Span subspan(size_type pos) const { return Span(p_ + pos, n_ - pos); }
 private:
  T* p_ = nullptr;
  size_type n_ = 0;
};
}  // namespace absl

// Test-infrastructure shim: absl::StatusOr<T> with ok() / status() / operator*, and the value-type conversion
// cc/lib/bit_coder.cc relies on (`return bit;` from StatusOr<uint64_t> to StatusOr<int32_t>).
#pragma once
#include <type_traits>
#include "absl/status/status.h"
namespace absl {
template <typename T>
class StatusOr {
 public:
  StatusOr(const T& value) : value_(value) {}                                  // NOLINT: implicit, as in absl
  StatusOr(const Status& status) : status_(status), value_() {}                // NOLINT
  template <typename U, typename = std::enable_if_t<!std::is_same<U, T>::value>>
  StatusOr(const StatusOr<U>& other)                                           // NOLINT
      : status_(other.status()), value_(other.ok() ? static_cast<T>(*other) : T()) {}
  bool ok() const { return status_.ok(); }
  const Status& status() const { return status_; }
  const T& operator*() const { return value_; }
 private:
  Status status_;
  T value_;
};
}  // namespace absl

// Test-infrastructure shim: absl::Status with just ok()/message().
#pragma once
#include <string>
#include <utility>
namespace absl {
class Status {
 public:
  Status() = default;
  explicit Status(std::string m) : ok_(false), msg_(std::move(m)) {}
  bool ok() const { return ok_; }
  const std::string& message() const { return msg_; }
 private:
  bool ok_ = true;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }
inline Status InvalidArgumentError(const std::string& m) { return Status(m); }
inline Status DataLossError(const std::string& m) { return Status(m); }
}  // namespace absl

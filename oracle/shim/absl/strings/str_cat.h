// Test-infrastructure shim: absl::StrCat over streamable arguments.
#pragma once
#include <sstream>
#include <string>
namespace absl {
template <typename... A>
std::string StrCat(const A&... a) {
  std::ostringstream os;
  (void)std::initializer_list<int>{((os << a), 0)...};
  return os.str();
}
}  // namespace absl

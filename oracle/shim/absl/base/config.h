// Test-infrastructure shim (NOT product code): what cc/lib/bit_coder.cc takes from absl/base/config.h (and, through
// it, absl/numeric/bits.h) so that the reference's bit coder compiles in place.
#pragma once
#include <cstdint>
#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
#define ABSL_IS_LITTLE_ENDIAN 1
#endif
namespace absl {
inline int bit_width(uint32_t x) { return x == 0 ? 0 : 32 - __builtin_clz(x); }
}  // namespace absl

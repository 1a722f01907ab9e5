// Test-infrastructure shim: CHECK/DCHECK macros (DCHECKs compiled out, as in -c opt).
#pragma once
#include <cstdio>
#include <cstdlib>
#define TFCB_SHIM_CHECK(cond)                                                   \
  do {                                                                          \
    if (!(cond)) {                                                              \
      std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      std::abort();                                                             \
    }                                                                           \
  } while (0)
#ifndef CHECK
#define CHECK(c) TFCB_SHIM_CHECK(c)
#define CHECK_EQ(a, b) TFCB_SHIM_CHECK((a) == (b))
#define CHECK_NE(a, b) TFCB_SHIM_CHECK((a) != (b))
#define CHECK_LT(a, b) TFCB_SHIM_CHECK((a) < (b))
#define CHECK_LE(a, b) TFCB_SHIM_CHECK((a) <= (b))
#define CHECK_GT(a, b) TFCB_SHIM_CHECK((a) > (b))
#define CHECK_GE(a, b) TFCB_SHIM_CHECK((a) >= (b))
#define DCHECK(c) do { } while (0)
#define DCHECK_EQ(a, b) do { } while (0)
#define DCHECK_NE(a, b) do { } while (0)
#define DCHECK_LT(a, b) do { } while (0)
#define DCHECK_LE(a, b) do { } while (0)
#define DCHECK_GT(a, b) do { } while (0)
#define DCHECK_GE(a, b) do { } while (0)
#endif

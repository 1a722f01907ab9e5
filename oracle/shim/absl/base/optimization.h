// Test-infrastructure shim (NOT product code): the minimum of absl needed so that
// /root/reference/tensorflow_compression/cc/lib/range_coder.{h,cc} compile in place.
#pragma once
#define ABSL_PREDICT_FALSE(x) (__builtin_expect(!!(x), 0))
#define ABSL_PREDICT_TRUE(x) (__builtin_expect(!!(x), 1))

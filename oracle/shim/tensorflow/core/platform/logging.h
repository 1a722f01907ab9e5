// Test-infrastructure shim: stands in for TF's logging header (CHECK/DCHECK only).
#pragma once
#include <algorithm>
#include "absl/log/check.h"

// GDN forward on the 5th-generation tensor cores (tcgen05 + TMEM) -- see DESIGN.md.
// Placeholder until the tcgen05 kernel lands: reports "not handled" so that tfcb_gdn_forward uses the
// fp32 CUDA-core kernel of gdn.cu.
#include "common.cuh"

namespace tfcb {

int gdn_tc_forward(const float*, const float*, const float*, float*, long long, int, int, float, float,
                   cudaStream_t, bool* handled) {
  *handled = false;
  return TFCB_OK;
}

}  // namespace tfcb

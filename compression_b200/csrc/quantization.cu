// StochasticRound (tensorflow_compression/cc/kernels/quantization_kernels.cc:48-95, op contract
// cc/ops/quantization_ops.cc:28-53) on the GPU, bit-identical to the reference's single sequential stream.
//
// The reference draws one xoshiro256+ number per element, in element order, from a state seeded through
// std::seed_seq.  The generator's transition is linear over GF(2) (xor / shift / rotate only), so the state after k
// steps is T^k s0 with T a 256 x 256 bit matrix.  Each thread handles kRun consecutive elements; thread i jumps to
// s_{i * kRun} = T^(i * kRun) s0 by applying the precomputed matrices J_b = T^(kRun * 2^b) for the set bits b of i
// (each stored as 256 columns of 256 bits; a matrix-vector product is the xor of the columns selected by the
// state's bits), then walks its run exactly like the reference loop: floor, fraction, (next >> 40) * 2^-24 < fraction.
#include <chrono>
#include <memory>
#include <mutex>
#include <random>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace tfcb {
namespace {

constexpr int kRun = 32;        // consecutive elements per thread (one sequential sub-stream)
constexpr int kJumpBits = 40;   // thread index bits covered: 2^40 * 32 elements

struct State {
  uint64_t s[4];
};

// quantization_kernels.cc:35-46
__host__ __device__ __forceinline__ uint64_t next_random(uint64_t* s) {
  const uint64_t result = s[0] + s[3];
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = (s[3] << 45) | (s[3] >> (64 - 45));
  return result;
}

// 256 x 256 bit matrix as 256 columns; column k = image of the k-th basis state.
struct BitMatrix {
  State col[256];
};

State apply(const BitMatrix& m, const State& v) {
  State r = {{0, 0, 0, 0}};
  for (int k = 0; k < 256; ++k)
    if ((v.s[k >> 6] >> (k & 63)) & 1ull)
      for (int w = 0; w < 4; ++w) r.s[w] ^= m.col[k].s[w];
  return r;
}

void multiply(const BitMatrix& a, const BitMatrix& b, BitMatrix* out) {  // out = a * b (apply b, then a)
  for (int k = 0; k < 256; ++k) out->col[k] = apply(a, b.col[k]);
}

// Device copy of J_b, b = 0 .. kJumpBits - 1 (320 KB), built once per process and device.
const State* jump_tables(cudaStream_t s, int* rc) {
  static std::mutex mu;
  static State* dev_tables[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  *rc = TFCB_OK;
  if (dev < 0 || dev >= 64) {
    *rc = fail(TFCB_CUDA_ERROR, "unexpected device index %d", dev);
    return nullptr;
  }
  if (dev_tables[dev]) return dev_tables[dev];
  static std::vector<State> host;
  if (host.empty()) {
    auto t = std::make_unique<BitMatrix>();
    for (int k = 0; k < 256; ++k) {
      State e = {{0, 0, 0, 0}};
      e.s[k >> 6] = 1ull << (k & 63);
      next_random(e.s);
      t->col[k] = e;
    }
    auto j = std::make_unique<BitMatrix>(*t), tmp = std::make_unique<BitMatrix>();
    for (int b = 1; b < kRun; b <<= 1) {  // T^kRun by squaring (kRun is a power of two)
      multiply(*j, *j, tmp.get());
      std::swap(j, tmp);
    }
    host.resize((size_t)kJumpBits * 256);
    for (int b = 0; b < kJumpBits; ++b) {
      for (int k = 0; k < 256; ++k) host[(size_t)b * 256 + k] = j->col[k];
      multiply(*j, *j, tmp.get());
      std::swap(j, tmp);
    }
  }
  State* d = nullptr;
  if (cudaMalloc(&d, host.size() * sizeof(State)) != cudaSuccess) {
    (void)cudaGetLastError();
    *rc = fail(TFCB_OUT_OF_MEMORY, "cannot allocate the StochasticRound jump tables");
    return nullptr;
  }
  if (cudaMemcpyAsync(d, host.data(), host.size() * sizeof(State), cudaMemcpyHostToDevice, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(d);
    *rc = fail(TFCB_CUDA_ERROR, "cannot upload the StochasticRound jump tables");
    return nullptr;
  }
  dev_tables[dev] = d;
  return d;
}

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256) stochastic_round_kernel(const T* __restrict__ in, long long n, float step_size,
                                                               State s0, const State* __restrict__ jump,
                                                               int32_t* __restrict__ out) {
  const long long run = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long begin = run * kRun;
  if (begin >= n) return;
  State st = s0;
  for (int b = 0; b < kJumpBits && (run >> b) != 0; ++b) {
    if (!((run >> b) & 1ll)) continue;
    const State* m = jump + (size_t)b * 256;
    State r = {{0ull, 0ull, 0ull, 0ull}};
#pragma unroll 1
    for (int w = 0; w < 4; ++w) {
      uint64_t bits = st.s[w];
      while (bits) {
        const int k = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(m + w * 64 + k));
        const ulonglong2 c = __ldg(reinterpret_cast<const ulonglong2*>(m + w * 64 + k) + 1);
        r.s[0] ^= a.x;
        r.s[1] ^= a.y;
        r.s[2] ^= c.x;
        r.s[3] ^= c.y;
      }
    }
    st = r;
  }
  const long long end = min(n, begin + kRun);
  for (long long i = begin; i < end; ++i) {
    // quantization_kernels.cc:80-91: float32 divide, floor, compare in float32
    const float number = __fdiv_rn(to_float<T>(in[i]), step_size);
    const float integral = floorf(number);
    int32_t o = (int32_t)integral;
    const float fractional = __fsub_rn(number, integral);
    const float random = (float)(next_random(st.s) >> 40) * 0x1.0p-24f;
    if (random < fractional) ++o;
    out[i] = o;
  }
}

}  // namespace
}  // namespace tfcb

extern "C" int tfcb_stochastic_round(const void* inputs_dev, int dtype, int64_t n, float step_size,
                                     const int32_t* seed_host, int64_t seed_len, int32_t* outputs_dev, void* stream) {
  using namespace tfcb;
  if (n < 0 || seed_len < 0 || (seed_len > 0 && !seed_host))
    return fail(TFCB_INVALID_ARGUMENT, "StochasticRound: bad sizes");
  if (dtype < 0 || dtype > 2) return fail(TFCB_INVALID_ARGUMENT, "StochasticRound: dtype must be 0 (f32), 1 (f16) or 2 (bf16)");
  if (n == 0) return TFCB_OK;
  if (!inputs_dev || !outputs_dev) return fail(TFCB_INVALID_ARGUMENT, "StochasticRound: null tensor");
  if (((long long)n + kRun - 1) / kRun >= (1ll << kJumpBits)) return fail(TFCB_INVALID_ARGUMENT, "StochasticRound: too many elements");
  // quantization_kernels.cc:66-78: std::seed_seq over the seed words, or over the clock when the seed is empty
  State s0;
  if (seed_len > 0) {
    std::seed_seq seq(seed_host, seed_host + seed_len);
    seq.generate(reinterpret_cast<uint32_t*>(s0.s), reinterpret_cast<uint32_t*>(s0.s + 4));
  } else {
    const uint64_t seed = std::chrono::high_resolution_clock::now().time_since_epoch().count();
    std::seed_seq seq{seed, seed >> 32};
    seq.generate(reinterpret_cast<uint32_t*>(s0.s), reinterpret_cast<uint32_t*>(s0.s + 4));
  }
  cudaStream_t s = as_stream(stream);
  int rc = TFCB_OK;
  const State* jump = jump_tables(s, &rc);
  if (!jump) return rc;
  const long long runs = ((long long)n + kRun - 1) / kRun;
  const unsigned grid = (unsigned)((runs + 255) / 256);
  if (dtype == 0)
    stochastic_round_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(inputs_dev), n, step_size, s0, jump, outputs_dev);
  else if (dtype == 1)
    stochastic_round_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(inputs_dev), n, step_size, s0, jump, outputs_dev);
  else
    stochastic_round_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(inputs_dev), n, step_size, s0, jump,
                                                               outputs_dev);
  TFCB_LAUNCHED();
  TFCB_CUDA_TRY(cudaGetLastError());
  return TFCB_OK;
}

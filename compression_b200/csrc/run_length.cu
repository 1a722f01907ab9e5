// RunLengthEncode / RunLengthDecode (tensorflow_compression/cc/kernels/run_length_kernels.cc:52-262, op contract
// cc/ops/run_length_ops.cc:28-84, bit packing cc/lib/bit_coder.cc:50-191; RunLengthGammaEncode/Decode are the
// (-1, -1, false) special case, run_length_gamma_kernels.cc).
//
// The reference walks the tensor once and appends variable-length codes (Elias gamma or Rice) LSB-first to a bit
// string.  Every code is a function of one element and of the run it belongs to, so the ENCODER is data parallel:
//   1. run starts by a forward max-scan (a run = maximal stretch of zeros or of non-zeros), run lengths scattered to
//      the run starts by each run's last element;
//   2. the bit length of the token each element contributes (zero for most zeros);
//   3. an exclusive 64-bit prefix sum of the lengths = every token's bit offset;
//   4. every token ORs its few set bits into the zero-initialised output (the long unary zero prefixes are never
//      touched), two 32-bit atomics per field.
// The DECODER is inherently serial (a code's position depends on all previous codes): one thread walks the string,
// as the reference does; it exists for completeness of the op pair, not for speed.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace tfcb {
namespace {

struct RlParams {
  int rl_code, mag_code, rl_nz;
};

__device__ __forceinline__ int bit_width_dev(uint32_t v) { return 32 - __clz(v); }

// bits of WriteRunLength(v), run_length_kernels.cc:66-72
__device__ __forceinline__ unsigned long long len_rl(uint32_t v, const RlParams& P) {
  if (P.rl_code >= 0) return (unsigned long long)(v >> P.rl_code) + 1ull + (unsigned long long)P.rl_code;
  return 2ull * (unsigned long long)bit_width_dev(v + 1u) - 1ull;
}
// magnitude payload of WriteNonZero(sample), :74-89
__device__ __forceinline__ uint32_t mag_value(int32_t s, const RlParams& P) {
  if (P.mag_code >= 0) return (uint32_t)(s > 0 ? s - 1 : -(s + 1));
  if (s == INT32_MIN) return (uint32_t)INT32_MAX;  // "We can't encode int32 minimum. Encode closest value instead."
  return (uint32_t)(s > 0 ? s : -s);
}
__device__ __forceinline__ unsigned long long len_mag(uint32_t m, const RlParams& P) {
  if (P.mag_code >= 0) return (unsigned long long)(m >> P.mag_code) + 1ull + (unsigned long long)P.mag_code;
  return 2ull * (unsigned long long)bit_width_dev(m) - 1ull;
}

struct MaxOp {
  __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

// start candidates: i where the zero-ness changes, else -1
__global__ void rl_boundaries_kernel(const int32_t* __restrict__ data, long long n, int* __restrict__ cand) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool nz = data[i] != 0;
  cand[i] = (i == 0 || (data[i - 1] != 0) != nz) ? (int)i : -1;
}

__global__ void rl_runlen_kernel(const int32_t* __restrict__ data, long long n, const int* __restrict__ start,
                                 int* __restrict__ runlen) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool nz = data[i] != 0;
  if (i == n - 1 || (data[i + 1] != 0) != nz) runlen[start[i]] = (int)(i - start[i] + 1);  // the run's last element
}

// The zero run that ends right before element i (0 if none) and whether a non-zero run precedes it.
__device__ __forceinline__ uint32_t zero_run_before(const int32_t* data, const int* start, long long i, bool* later_run) {
  *later_run = false;
  if (i == 0 || data[i - 1] != 0) return 0u;
  const int s = start[i - 1];
  *later_run = s > 0;
  return (uint32_t)(i - s);
}

__global__ void rl_lengths_kernel(const int32_t* __restrict__ data, long long n, const int* __restrict__ start,
                                  const int* __restrict__ runlen, RlParams P, unsigned long long* __restrict__ len) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t s = data[i];
  unsigned long long L = 0;
  if (s != 0) {
    L = 1ull + len_mag(mag_value(s, P), P);
    bool later;
    const uint32_t zr = zero_run_before(data, start, i, &later);
    if (!P.rl_nz) {
      L += len_rl(zr, P);
    } else if (i == 0 || data[i - 1] == 0) {  // first element of a non-zero run: both run lengths precede it
      L += len_rl(zr - (later ? 1u : 0u), P) + len_rl((uint32_t)runlen[i] - 1u, P);
    }
  } else if (i == n - 1) {  // trailing zeros: one last run length
    const int st = start[i];
    const uint32_t zr = (uint32_t)(n - st);
    L = len_rl(zr - ((P.rl_nz && st > 0) ? 1u : 0u), P);
  }
  len[i] = L;
}

// ORs the low `nbits` (<= 32) bits of `value` into the bit string at bit position `pos` (LSB first, bit_coder.cc:50-66)
__device__ __forceinline__ void put_bits(uint32_t* out, unsigned long long pos, int nbits, uint32_t value) {
  if (nbits <= 0) return;
  const unsigned long long v = (unsigned long long)(nbits == 32 ? value : (value & ((1u << nbits) - 1u))) << (pos & 31ull);
  const unsigned long long w = pos >> 5;
  if ((uint32_t)v) atomicOr(out + w, (uint32_t)v);
  if ((uint32_t)(v >> 32)) atomicOr(out + w + 1, (uint32_t)(v >> 32));
}
// WriteRice / WriteGamma at `pos`; returns the position after the code
__device__ __forceinline__ unsigned long long put_code(uint32_t* out, unsigned long long pos, uint32_t value, int code) {
  if (code >= 0) {  // Rice: value >> code zeros, a one, `code` low bits
    pos += (unsigned long long)(value >> code);
    put_bits(out, pos, 1, 1u);
    put_bits(out, pos + 1, code, value);
    return pos + 1ull + (unsigned long long)code;
  }
  const int bw = bit_width_dev(value);  // gamma of value > 0: bw - 1 zeros, a one, bw - 1 low bits
  pos += (unsigned long long)(bw - 1);
  put_bits(out, pos, 1, 1u);
  put_bits(out, pos + 1, bw - 1, value);
  return pos + (unsigned long long)bw;
}
__device__ __forceinline__ unsigned long long put_rl(uint32_t* out, unsigned long long pos, uint32_t v, const RlParams& P) {
  return put_code(out, pos, P.rl_code >= 0 ? v : v + 1u, P.rl_code);
}

__global__ void rl_emit_kernel(const int32_t* __restrict__ data, long long n, const int* __restrict__ start,
                               const int* __restrict__ runlen, RlParams P, const unsigned long long* __restrict__ off,
                               uint32_t* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t s = data[i];
  unsigned long long pos = off[i];
  if (s != 0) {
    bool later;
    const uint32_t zr = zero_run_before(data, start, i, &later);
    if (!P.rl_nz) {
      pos = put_rl(out, pos, zr, P);
    } else if (i == 0 || data[i - 1] == 0) {
      pos = put_rl(out, pos, zr - (later ? 1u : 0u), P);
      pos = put_rl(out, pos, (uint32_t)runlen[i] - 1u, P);
    }
    put_bits(out, pos, 1, s > 0 ? 1u : 0u);
    put_code(out, pos + 1, mag_value(s, P), P.mag_code);
  } else if (i == n - 1) {
    const int st = start[i];
    put_rl(out, pos, (uint32_t)(n - st) - ((P.rl_nz && st > 0) ? 1u : 0u), P);
  }
}

// ---- decoder: one thread, BitReader semantics of bit_coder.cc:92-189 (errors -> flag) ----
struct BitRd {
  const uint8_t* p;
  unsigned long long nbits, pos;
  __device__ __forceinline__ bool bit(uint32_t* b) {
    if (pos >= nbits) return false;
    *b = (p[pos >> 3] >> (pos & 7)) & 1u;
    ++pos;
    return true;
  }
  __device__ __forceinline__ bool bits(int count, uint32_t* v) {
    if (pos + (unsigned long long)count > nbits) return false;
    uint32_t r = 0;
    for (int k = 0; k < count; ++k, ++pos) r |= (uint32_t)((p[pos >> 3] >> (pos & 7)) & 1u) << k;
    *v = r;
    return true;
  }
  // 0 ok, 1 out of bits, 2 gamma too wide
  __device__ __forceinline__ int gamma(uint32_t* v) {
    int bw = 1;
    for (;;) {
      uint32_t b;
      if (!bit(&b)) return 1;
      if (b) break;
      ++bw;
    }
    if (bw > 31) return 2;
    uint32_t lsbs;
    if (!bits(bw - 1, &lsbs)) return 1;
    *v = (1u << (bw - 1)) | lsbs;
    return 0;
  }
  __device__ __forceinline__ int rice(int k, uint32_t* v) {
    uint32_t msbs = 0;
    for (;;) {
      uint32_t b;
      if (!bit(&b)) return 1;
      if (b) break;
      ++msbs;
    }
    uint32_t lsbs;
    if (!bits(k, &lsbs)) return 1;
    *v = (msbs << k) | lsbs;
    return 0;
  }
};

// err: 0 ok, 1 "Out of bits to read.", 2 "Exceeded maximum gamma bit width.", 3 "Decoded past end of tensor."
__global__ void rl_decode_kernel(const uint8_t* __restrict__ code, long long n_bytes, int32_t* __restrict__ data,
                                 long long n, RlParams P, int* __restrict__ err) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  BitRd rd = {code, (unsigned long long)n_bytes * 8ull, 0ull};
  auto read_rl = [&](uint32_t* v) -> int {
    if (P.rl_code >= 0) return rd.rice(P.rl_code, v);
    const int e = rd.gamma(v);
    if (e == 0) *v -= 1u;
    return e;
  };
  auto read_nz = [&](int32_t* s) -> int {
    uint32_t pos_bit, m;
    if (!rd.bit(&pos_bit)) return 1;
    const int e = (P.mag_code >= 0) ? rd.rice(P.mag_code, &m) : rd.gamma(&m);
    if (e) return e;
    if (P.mag_code >= 0) *s = pos_bit ? (int32_t)m + 1 : -(int32_t)m - 1;
    else *s = pos_bit ? (int32_t)m : -(int32_t)m;
    return 0;
  };
  long long p = 0;
  uint32_t offset = 0;
  int e = 0;
  while (p < n) {
    uint32_t run;
    if ((e = read_rl(&run))) break;
    p += (long long)run + offset;
    if (!(p < n)) {
      if (p != n) e = 3;
      break;
    }
    if (P.rl_nz) {
      if ((e = read_rl(&run))) break;
      const long long next_zero = p + (long long)run + 1;
      if (next_zero > n) {
        e = 3;
        break;
      }
      while (p < next_zero) {
        int32_t s;
        if ((e = read_nz(&s))) break;
        data[p++] = s;
      }
      if (e) break;
      offset = 1;
    } else {
      int32_t s;
      if ((e = read_nz(&s))) break;
      data[p++] = s;
    }
  }
  *err = e;
}

}  // namespace
}  // namespace tfcb

extern "C" {

int tfcb_run_length_encode(const int32_t* data_dev, int64_t n, int run_length_code, int magnitude_code,
                           int use_run_length_for_non_zeros, uint8_t* code_dev, int64_t capacity, int64_t* n_bytes_host,
                           void* stream) {
  using namespace tfcb;
  if (n < 0 || n >= (1ll << 31) || capacity < 0 || !n_bytes_host) return fail(TFCB_INVALID_ARGUMENT, "RunLengthEncode: bad sizes");
  if (run_length_code > 31 || magnitude_code > 31) return fail(TFCB_INVALID_ARGUMENT, "RunLengthEncode: Rice parameter > 31");
  *n_bytes_host = 0;
  if (n == 0) return TFCB_OK;
  if (!data_dev || (!code_dev && capacity > 0)) return fail(TFCB_INVALID_ARGUMENT, "RunLengthEncode: null tensor");
  if (capacity & 3) capacity &= ~3ll;  // the bit string is assembled in 32-bit words
  cudaStream_t s = as_stream(stream);
  const RlParams P = {run_length_code, magnitude_code, use_run_length_for_non_zeros ? 1 : 0};
  int *cand = nullptr, *start = nullptr, *runlen = nullptr;
  unsigned long long *len = nullptr, *off = nullptr;
  void* tmp = nullptr;
  size_t tmp1 = 0, tmp2 = 0;
  cub::DeviceScan::InclusiveScan(nullptr, tmp1, (int*)nullptr, (int*)nullptr, MaxOp(), (int)n, s);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp2, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)n, s);
  const size_t tmp_bytes = std::max(tmp1, tmp2) + 16;
  int rc = dev_alloc((void**)&cand, (size_t)n * sizeof(int), s);
  if (rc == TFCB_OK) rc = dev_alloc((void**)&start, (size_t)n * sizeof(int), s);
  if (rc == TFCB_OK) rc = dev_alloc((void**)&runlen, (size_t)n * sizeof(int), s);
  if (rc == TFCB_OK) rc = dev_alloc((void**)&len, (size_t)(n + 1) * sizeof(unsigned long long), s);
  if (rc == TFCB_OK) rc = dev_alloc((void**)&off, (size_t)(n + 1) * sizeof(unsigned long long), s);
  if (rc == TFCB_OK) rc = dev_alloc(&tmp, tmp_bytes, s);
  auto cleanup = [&]() {
    dev_free(cand, s); dev_free(start, s); dev_free(runlen, s); dev_free(len, s); dev_free(off, s); dev_free(tmp, s);
  };
  if (rc != TFCB_OK) {
    cleanup();
    return rc;
  }
  const unsigned grid = (unsigned)((n + 255) / 256);
  rl_boundaries_kernel<<<grid, 256, 0, s>>>(data_dev, n, cand);
  size_t t1 = tmp_bytes;
  cub::DeviceScan::InclusiveScan(tmp, t1, cand, start, MaxOp(), (int)n, s);
  rl_runlen_kernel<<<grid, 256, 0, s>>>(data_dev, n, start, runlen);
  rl_lengths_kernel<<<grid, 256, 0, s>>>(data_dev, n, start, runlen, P, len);
  cudaMemsetAsync(len + n, 0, sizeof(unsigned long long), s);
  size_t t2 = tmp_bytes;
  cub::DeviceScan::ExclusiveSum(tmp, t2, len, off, (int)(n + 1), s);  // off[n] = total bits
  for (int k = 0; k < 6; ++k) TFCB_LAUNCHED();
  unsigned long long total_bits = 0;
  cudaError_t e = cudaMemcpyAsync(&total_bits, off + n, sizeof total_bits, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    cleanup();
    return fail(TFCB_CUDA_ERROR, "RunLengthEncode: %s", cudaGetErrorString(e));
  }
  const int64_t n_bytes = (int64_t)((total_bits + 7ull) / 8ull);
  *n_bytes_host = n_bytes;
  const int64_t words = (n_bytes + 3) / 4;
  if (words * 4 > capacity) {
    cleanup();
    return fail(TFCB_INVALID_ARGUMENT, "RunLengthEncode: the code needs %lld bytes (capacity %lld)", (long long)(words * 4),
                (long long)capacity);
  }
  cudaMemsetAsync(code_dev, 0, (size_t)words * 4, s);
  rl_emit_kernel<<<grid, 256, 0, s>>>(data_dev, n, start, runlen, P, off, reinterpret_cast<uint32_t*>(code_dev));
  TFCB_LAUNCHED();
  e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) return fail(TFCB_CUDA_ERROR, "RunLengthEncode launch failed: %s", cudaGetErrorString(e));
  return TFCB_OK;
}

int tfcb_run_length_decode(const uint8_t* code_dev, int64_t n_bytes, int run_length_code, int magnitude_code,
                           int use_run_length_for_non_zeros, int32_t* data_dev, int64_t n, void* stream) {
  using namespace tfcb;
  if (n < 0 || n_bytes < 0) return fail(TFCB_INVALID_ARGUMENT, "RunLengthDecode: bad sizes");
  if (run_length_code > 31 || magnitude_code > 31) return fail(TFCB_INVALID_ARGUMENT, "RunLengthDecode: Rice parameter > 31");
  if (n == 0) return TFCB_OK;
  if (!data_dev || (!code_dev && n_bytes > 0)) return fail(TFCB_INVALID_ARGUMENT, "RunLengthDecode: null tensor");
  cudaStream_t s = as_stream(stream);
  const RlParams P = {run_length_code, magnitude_code, use_run_length_for_non_zeros ? 1 : 0};
  int* err = nullptr;
  TFCB_TRY(dev_alloc((void**)&err, sizeof(int), s));
  cudaMemsetAsync(data_dev, 0, (size_t)n * sizeof(int32_t), s);  // "Fill data tensor with zeros."
  rl_decode_kernel<<<1, 32, 0, s>>>(code_dev, n_bytes, data_dev, n, P, err);
  TFCB_LAUNCHED();
  int h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, err, sizeof h, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  dev_free(err, s);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(TFCB_CUDA_ERROR, "RunLengthDecode: %s", cudaGetErrorString(e));
  }
  if (h == 1) return fail(TFCB_INVALID_ARGUMENT, "Out of bits to read.");
  if (h == 2) return fail(TFCB_INVALID_ARGUMENT, "Exceeded maximum gamma bit width.");
  if (h == 3) return fail(TFCB_INVALID_ARGUMENT, "Decoded past end of tensor.");
  return TFCB_OK;
}

}  // extern "C"

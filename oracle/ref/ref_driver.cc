// ORACLE / TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// Links the reference's own range coder (compiled IN PLACE from
// /root/reference/tensorflow_compression/cc/lib/range_coder.cc, see oracle/Makefile)
// and restates, on the CPU, the TensorFlow-entangled glue that the reference's op
// kernels put around it (those files need TF headers and cannot be compiled here):
//
//   lookup parsing        cc/kernels/range_coder_kernels.cc:110-164  (ScanCDF, IndexCDFVector/Matrix)
//   channel/index loops   cc/kernels/range_coder_kernels.cc:191-272, 360-429
//   overflow (Elias-gamma) cc/kernels/range_coder_kernels.cc:290-322, 449-471
//   finalize              cc/kernels/range_coder_kernels.cc:274-287, 431-446
//   legacy RangeEncode/RangeDecode addressing
//                         cc/kernels/range_coding_kernels.cc:60-132, 134-173, 232-269, 345-373
//                         cc/kernels/range_coding_kernels_util.cc:34-91 (MergeAxes)
//   PmfToQuantizedCdf     cc/kernels/pmf_to_cdf_kernels.cc:58-101, 104-208
//   RunLengthEncode/Decode op loops  cc/kernels/run_length_kernels.cc:69-135, 158-250  (around the reference's own
//                         BitWriter / BitReader, cc/lib/bit_coder.cc, compiled in place like the range coder)
//
// Every arithmetic step of the coder itself is executed by the reference's
// tensorflow_compression::RangeEncoder / RangeDecoder objects.  Streams are sharded
// over a PERSISTENT pool of std::thread workers (created once, reused by every call) to
// emulate the reference's ParallelFor over streams on TF's intra-op pool
// (range_coder_kernels.cc:212-218).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the library built from this file.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <random>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "tensorflow_compression/cc/lib/bit_coder.h"
#include "tensorflow_compression/cc/lib/range_coder.h"

namespace tfc = tensorflow_compression;

namespace {

thread_local std::string g_error;

int Fail(const std::string& msg) {
  g_error = msg;
  return 1;
}

struct Row {
  const int32_t* p;  // points at the precision entry
  int64_t size;      // precision entry + cdf entries (incl. the final 2^|P|)
};

// Lookup grammar, range_coder_kernels.cc:110-137.
bool ScanRow(const int32_t* end, const int32_t** cur, std::vector<Row>* rows, std::string* err) {
  const int32_t* p = *cur;
  if (end < p + 3) {
    *err = "CDF ended prematurely.";
    return false;
  }
  const int32_t* first = p;
  const int64_t aprec = std::abs(static_cast<int64_t>(*first));
  if (aprec < 1 || aprec >= 17) {
    *err = "precision=" + std::to_string(aprec) + " not in range [1, 17)";
    return false;
  }
  const int32_t last_value = 1 << aprec;
  if (*(++p) != 0) {
    *err = "CDF must start with 0.";
    return false;
  }
  do {
    if (++p == end) {
      *err = "CDF must end with 1 << precision.";
      return false;
    }
    if (p[0] < p[-1]) {
      *err = "CDF must be monotonically increasing.";
      return false;
    }
  } while (*p != last_value);
  ++p;
  rows->push_back(Row{first, p - first});
  while (p != end && *p == last_value) ++p;
  *cur = p;
  return true;
}

bool ParseLookup(const int32_t* lookup, int64_t len, int64_t cols, std::vector<Row>* rows,
                 std::string* err) {
  rows->clear();
  const int32_t* const end = lookup + len;
  if (cols <= 0) {  // 1-D concatenated form, :139-148
    for (const int32_t* cur = lookup; cur != end;) {
      if (!ScanRow(end, &cur, rows, err)) return false;
    }
  } else {  // 2-D stacked form, :150-164
    for (const int32_t* cur = lookup; cur != end;) {
      const int32_t* row_end = cur + cols;
      if (!ScanRow(row_end, &cur, rows, err)) return false;
      if (cur != row_end) {
        *err = "CDF must end with 1 << precision.";
        return false;
      }
    }
  }
  return true;
}

// range_coder_kernels.cc:290-322
void OverflowEncode(tfc::RangeEncoder& enc, std::string* sink, const Row& row, int32_t value) {
  const int32_t max_value = static_cast<int32_t>(row.size) - 3;
  const int32_t sign = value < 0;
  int32_t gamma = 0;
  if (sign) {
    gamma = -value;
    value = max_value;
  } else if (value >= max_value) {
    gamma = value - max_value + 1;
    value = max_value;
  }
  enc.Encode(row.p[value + 1], row.p[value + 2], -row.p[0], sink);
  if (value != max_value) return;
  int32_t n = 1;
  while (gamma >= (1 << n)) {
    enc.Encode(0, 1, 1, sink);
    ++n;
  }
  while (--n >= 0) {
    const int32_t bit = (gamma >> n) & 1;
    enc.Encode(bit, bit + 1, 1, sink);
  }
  enc.Encode(sign, sign + 1, 1, sink);
}

// range_coder_kernels.cc:449-471
int32_t OverflowDecode(tfc::RangeDecoder& dec, const Row& row) {
  static const int32_t kBinary[] = {0, 1, 2};
  const int32_t max_value = static_cast<int32_t>(row.size) - 3;
  int32_t value = dec.Decode(absl::Span<const int32_t>(row.p + 1, row.size - 1), -row.p[0]);
  if (value != max_value) return value;
  int32_t n = 0;
  while (dec.DecodeLinearly(absl::Span<const int32_t>(kBinary, 3), 1) == 0) ++n;
  value = 1 << n;
  while (--n >= 0) value |= dec.DecodeLinearly(absl::Span<const int32_t>(kBinary, 3), 1) << n;
  const int32_t sign = dec.DecodeLinearly(absl::Span<const int32_t>(kBinary, 3), 1);
  return sign ? -value : value + max_value - 1;
}

// Persistent worker pool: the reference runs its per-stream loop on TensorFlow's intra-op pool through
// ParallelFor (range_coder_kernels.cc:212-218) -- threads that already exist when the op is called.  The
// workers here are created once, on the first multi-threaded call, and sleep between calls; a call hands
// out contiguous blocks of streams through an atomic cursor (the analogue of ParallelFor's cost-based
// sharding: enough blocks per thread to balance streams of unequal length).
class StreamPool {
 public:
  static StreamPool& Get() {
    static StreamPool* pool = new StreamPool;  // never destroyed: workers outlive every caller
    return *pool;
  }

  void Run(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& fn) {
    if (threads <= 1 || n <= 1) {
      fn(0, n);
      return;
    }
    std::lock_guard<std::mutex> serial(run_mu_);  // one parallel region at a time
    const int want = static_cast<int>(std::min<int64_t>(threads, n)) - 1;  // the caller works too
    {
      std::unique_lock<std::mutex> lk(mu_);
      while (static_cast<int>(workers_.size()) < want) {
        const int id = static_cast<int>(workers_.size());
        workers_.emplace_back([this, id] { Worker(id); });
      }
      fn_ = &fn;
      n_ = n;
      block_ = std::max<int64_t>(1, n / (4 * (want + 1)));
      next_.store(0, std::memory_order_relaxed);
      active_ = want;
      pending_ = want;
      ++epoch_;
    }
    wake_.notify_all();
    Drain();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

  int size() {
    std::lock_guard<std::mutex> lk(mu_);
    return static_cast<int>(workers_.size());
  }

 private:
  void Drain() {
    for (;;) {
      const int64_t lo = next_.fetch_add(block_, std::memory_order_relaxed);
      if (lo >= n_) return;
      (*fn_)(lo, std::min(n_, lo + block_));
    }
  }
  void Worker(int id) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        wake_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (id >= active_) continue;  // this region uses fewer threads than the pool holds
      }
      Drain();
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_one();
    }
  }

  std::mutex run_mu_, mu_;
  std::condition_variable wake_, done_;
  std::vector<std::thread> workers_;
  const std::function<void(int64_t, int64_t)>* fn_ = nullptr;
  int64_t n_ = 0, block_ = 1;
  std::atomic<int64_t> next_{0};
  int active_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
};

void ParallelOverStreams(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& fn) {
  StreamPool::Get().Run(n, threads, fn);
}

struct Encoder {
  std::vector<int32_t> lookup;
  std::vector<Row> rows;
  std::vector<tfc::RangeEncoder> enc;
  std::vector<std::string> sink;
};

struct Decoder {
  std::vector<int32_t> lookup;
  std::vector<Row> rows;
  std::vector<std::string> data;
  std::vector<tfc::RangeDecoder> dec;
};

}  // namespace

extern "C" {

const char* tfcref_last_error() { return g_error.c_str(); }

// ---- raw triples: drives RangeEncoder::Encode/Finalize directly (range_coder.cc:37-307) ----
int64_t tfcref_encode_triples(const int32_t* lower, const int32_t* upper, const int32_t* precision,
                              int64_t n, uint8_t* out, int64_t out_cap) {
  tfc::RangeEncoder enc;
  std::string sink;
  for (int64_t i = 0; i < n; ++i) enc.Encode(lower[i], upper[i], precision[i], &sink);
  enc.Finalize(&sink);
  if (static_cast<int64_t>(sink.size()) > out_cap) return -static_cast<int64_t>(sink.size());
  std::memcpy(out, sink.data(), sink.size());
  return static_cast<int64_t>(sink.size());
}

// ---- CreateRangeEncoder / EntropyEncode{Channel,Index} / EntropyEncodeFinalize ----
void* tfcref_encoder_create(const int32_t* lookup, int64_t len, int64_t cols, int64_t n_streams) {
  auto* e = new Encoder;
  e->lookup.assign(lookup, lookup + len);
  std::string err;
  if (!ParseLookup(e->lookup.data(), len, cols, &e->rows, &err)) {
    g_error = err;
    delete e;
    return nullptr;
  }
  e->enc.resize(n_streams);
  e->sink.resize(n_streams);
  return e;
}

int tfcref_encoder_encode(void* h, const int32_t* index, const int32_t* value, int64_t n_per_stream,
                          int threads) {
  auto* e = static_cast<Encoder*>(h);
  const int64_t S = static_cast<int64_t>(e->enc.size());
  const int64_t R = static_cast<int64_t>(e->rows.size());
  std::mutex mu;
  std::string err;
  ParallelOverStreams(S, threads, [&](int64_t lo, int64_t hi) {
    for (int64_t s = lo; s < hi; ++s) {
      tfc::RangeEncoder& enc = e->enc[s];
      std::string* sink = &e->sink[s];
      const int32_t* pv = value + s * n_per_stream;
      const int32_t* pi = index ? index + s * n_per_stream : nullptr;
      for (int64_t ind = 0, j = 0; j < n_per_stream; ++ind, ++j) {
        const int32_t val = pv[j];
        int64_t r;
        if (pi) {
          r = pi[j];
          if (r < 0 || r >= R) {
            std::lock_guard<std::mutex> g(mu);
            err = "index=" + std::to_string(r) + " not in range [0, " + std::to_string(R) + ")";
            return;
          }
        } else {
          if (ind >= R) ind = 0;  // :253
          r = ind;
        }
        const Row& row = e->rows[r];
        if (row.p[0] > 0) {
          if (val < 0 || val >= row.size - 2) {
            std::lock_guard<std::mutex> g(mu);
            err = "value=" + std::to_string(val) + " not in range [0, " +
                  std::to_string(row.size - 2) + ")";
            return;
          }
          enc.Encode(row.p[val + 1], row.p[val + 2], row.p[0], sink);
        } else {
          // The reference's Elias-gamma width loop (`while (gamma >= (1 << n))`, :310-315, "TODO Clamp gamma")
          // never ends once the payload reaches 2^30: `1 << 31` is negative and wider shifts wrap.  A checker
          // must not hang on such input, so it is refused here; below 2^30 the loop is the reference's.
          const int64_t payload = val < 0 ? -static_cast<int64_t>(val)
                                          : static_cast<int64_t>(val) - (static_cast<int64_t>(row.size) - 3) + 1;
          if (payload >= (int64_t{1} << 30)) {
            std::lock_guard<std::mutex> g(mu);
            err = "value=" + std::to_string(val) +
                  " has an Elias-gamma payload >= 2^30; the reference's width loop does not terminate on it";
            return;
          }
          OverflowEncode(enc, sink, row, val);
        }
      }
    }
  });
  if (!err.empty()) return Fail(err);
  return 0;
}

// Finalizes every stream (idempotence is NOT provided, as in the reference) and reports sizes.
int tfcref_encoder_finalize(void* h, int64_t* offsets /* n_streams + 1 */) {
  auto* e = static_cast<Encoder*>(h);
  offsets[0] = 0;
  for (size_t s = 0; s < e->enc.size(); ++s) {
    e->enc[s].Finalize(&e->sink[s]);
    offsets[s + 1] = offsets[s] + static_cast<int64_t>(e->sink[s].size());
  }
  return 0;
}

void tfcref_encoder_bytes(void* h, uint8_t* out) {
  auto* e = static_cast<Encoder*>(h);
  for (auto& s : e->sink) {
    std::memcpy(out, s.data(), s.size());
    out += s.size();
  }
}

void tfcref_encoder_free(void* h) { delete static_cast<Encoder*>(h); }

// ---- CreateRangeDecoder / EntropyDecode{Channel,Index} / EntropyDecodeFinalize ----
void* tfcref_decoder_create(const uint8_t* bytes, const int64_t* offsets, int64_t n_streams,
                            const int32_t* lookup, int64_t len, int64_t cols) {
  auto* d = new Decoder;
  d->lookup.assign(lookup, lookup + len);
  std::string err;
  if (!ParseLookup(d->lookup.data(), len, cols, &d->rows, &err)) {
    g_error = err;
    delete d;
    return nullptr;
  }
  d->data.resize(n_streams);
  d->dec.reserve(n_streams);
  for (int64_t s = 0; s < n_streams; ++s) {
    d->data[s].assign(reinterpret_cast<const char*>(bytes) + offsets[s], offsets[s + 1] - offsets[s]);
  }
  for (int64_t s = 0; s < n_streams; ++s) d->dec.emplace_back(absl::string_view(d->data[s]));
  return d;
}

int tfcref_decoder_decode(void* h, const int32_t* index, int32_t* out, int64_t n_per_stream,
                          int threads) {
  auto* d = static_cast<Decoder*>(h);
  const int64_t S = static_cast<int64_t>(d->dec.size());
  const int64_t R = static_cast<int64_t>(d->rows.size());
  std::mutex mu;
  std::string err;
  ParallelOverStreams(S, threads, [&](int64_t lo, int64_t hi) {
    for (int64_t s = lo; s < hi; ++s) {
      tfc::RangeDecoder& dec = d->dec[s];
      int32_t* po = out + s * n_per_stream;
      const int32_t* pi = index ? index + s * n_per_stream : nullptr;
      for (int64_t ind = 0, j = 0; j < n_per_stream; ++ind, ++j) {
        int64_t r;
        if (pi) {
          r = pi[j];
          if (r < 0 || r >= R) {
            std::lock_guard<std::mutex> g(mu);
            err = "index=" + std::to_string(r) + " not in range [0, " + std::to_string(R) + ")";
            return;
          }
        } else {
          if (R <= ind) ind = 0;
          r = ind;
        }
        const Row& row = d->rows[r];
        if (row.p[0] > 0) {
          po[j] = dec.Decode(absl::Span<const int32_t>(row.p + 1, row.size - 1), row.p[0]);
        } else {
          po[j] = OverflowDecode(dec, row);
        }
      }
    }
  });
  if (!err.empty()) return Fail(err);
  return 0;
}

int tfcref_decoder_finalize(void* h, uint8_t* ok /* n_streams */) {
  auto* d = static_cast<Decoder*>(h);
  for (size_t s = 0; s < d->dec.size(); ++s) ok[s] = d->dec[s].Finalize() ? 1 : 0;
  return 0;
}

void tfcref_decoder_free(void* h) { delete static_cast<Decoder*>(h); }

}  // extern "C"

// ------------------------------------------------------------------------------------------
// Legacy RangeEncode / RangeDecode (one stream, broadcastable N-D CDF).
// ------------------------------------------------------------------------------------------
namespace {

// range_coding_kernels_util.cc:34-91
bool MergeAxes(const std::vector<int64_t>& bshape, const std::vector<int64_t>& sshape,
               std::vector<int64_t>* mb, std::vector<int64_t>* ms, std::string* err) {
  auto fmt = [](const std::vector<int64_t>& v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); ++i) s += (i ? "," : "") + std::to_string(v[i]);
    return s + "]";
  };
  mb->assign(1, 1);
  ms->assign(1, 1);
  const int rank = static_cast<int>(bshape.size());
  for (int i = 0, j = 0; j < rank; ++j) {
    if (bshape[j] != sshape[j] && sshape[j] != 1) {
      *err = "Cannot broadcast shape " + fmt(sshape) + " to " + fmt(bshape);
      return false;
    }
    const bool was_b = ((*ms)[i] == 1);
    const bool is_b = (sshape[j] == 1);
    const bool merge = (was_b == is_b) || (bshape[j] <= 1) || ((*mb)[i] <= 1);
    if (merge) {
      (*mb)[i] *= bshape[j];
      (*ms)[i] *= sshape[j];
    } else {
      mb->push_back(bshape[j]);
      ms->push_back(sshape[j]);
      ++i;
    }
  }
  int64_t stride = 1;
  for (size_t i = bshape.size(); i < sshape.size(); ++i) stride *= sshape[i];
  ms->push_back(stride);
  return true;
}

// Walks data linearly and yields the CDF strip for each element,
// range_coding_kernels.cc:60-132 (arbitrary N instead of a template).
struct BroadcastWalk {
  std::vector<int64_t> dshape, displace, idx;
  const int32_t* cdf;
  BroadcastWalk(const std::vector<int64_t>& d, const std::vector<int64_t>& c, const int32_t* p)
      : dshape(d), displace(d.size(), c.back()), idx(d.size(), 0), cdf(p) {
    int64_t stride = c.back();
    for (int i = static_cast<int>(d.size()) - 1; i >= 0; --i) {
      if (c[i] <= 1) displace[i] -= stride;
      stride *= c[i];
    }
  }
  const int32_t* Next() {
    const int32_t* ret = cdf;
    int i = static_cast<int>(dshape.size()) - 1;
    for (; i > 0; --i) {
      if (++idx[i] < dshape[i]) break;
      idx[i] = 0;
    }
    cdf += displace[i];
    return ret;
  }
};

bool CheckLegacy(const std::vector<int64_t>& dshape, const std::vector<int64_t>& cshape,
                 const int32_t* cdf, int precision, int debug_level, std::string* err) {
  if (!(0 < precision && precision <= 16)) {
    *err = "`precision` must be in [1, 16]: " + std::to_string(precision);
    return false;
  }
  if (cshape.size() != dshape.size() + 1) {
    *err = "`cdf` should have one more axis than `data`";
    return false;
  }
  if (cshape.back() <= 1) {
    *err = "The last dimension of `cdf` should be > 1";
    return false;
  }
  if (debug_level > 0) {  // CheckCdfValues, :150-173
    const int64_t size = cshape.back();
    if (size <= 2) {
      *err = "CDF size should be > 2: " + std::to_string(size);
      return false;
    }
    int64_t rows = 1;
    for (size_t i = 0; i + 1 < cshape.size(); ++i) rows *= cshape[i];
    const int32_t upper = 1 << precision;
    for (int64_t r = 0; r < rows; ++r) {
      const int32_t* s = cdf + r * size;
      if (s[0] != 0 || s[size - 1] != upper) {
        *err = "CDF should start from 0 and end at " + std::to_string(upper) +
               ": cdf[0]=" + std::to_string(s[0]) + ", cdf[^1]=" + std::to_string(s[size - 1]);
        return false;
      }
      for (int64_t j = 0; j + 1 < size; ++j) {
        if (s[j + 1] <= s[j]) {
          *err = "CDF is not monotonic";
          return false;
        }
      }
    }
  }
  return true;
}

}  // namespace

extern "C" {

// data int16 with shape dshape[rank]; cdf int32 with shape cshape[rank+1].
int64_t tfcref_range_encode(const int16_t* data, const int64_t* dshape_, int rank, const int32_t* cdf,
                            const int64_t* cshape_, int crank, int precision, int debug_level,
                            uint8_t* out, int64_t out_cap) {
  std::vector<int64_t> dshape(dshape_, dshape_ + rank), cshape(cshape_, cshape_ + crank);
  std::string err;
  if (!CheckLegacy(dshape, cshape, cdf, precision, debug_level, &err)) return -Fail(err);
  std::vector<int64_t> mb, ms;
  if (!MergeAxes(dshape, cshape, &mb, &ms, &err)) return -Fail(err);
  if (mb.size() > 6) return -Fail("Irregular broadcast pattern");
  int64_t n = 1;
  for (auto d : dshape) n *= d;
  const int64_t chip = cshape.back();
  BroadcastWalk walk(mb, ms, cdf);
  tfc::RangeEncoder enc;
  std::string sink;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* strip = walk.Next();
    const int64_t v = data[i];
    if (debug_level > 0 && (v < 0 || chip <= v + 1)) {
      return -Fail("'data' value not in [0, " + std::to_string(chip - 1) + "): value=" +
                   std::to_string(v));
    }
    enc.Encode(strip[v], strip[v + 1], precision, &sink);
  }
  enc.Finalize(&sink);
  if (static_cast<int64_t>(sink.size()) > out_cap) return -Fail("output buffer too small");
  std::memcpy(out, sink.data(), sink.size());
  g_error.clear();
  return static_cast<int64_t>(sink.size());
}

int tfcref_range_decode(const uint8_t* bytes, int64_t nbytes, const int64_t* dshape_, int rank,
                        const int32_t* cdf, const int64_t* cshape_, int crank, int precision,
                        int debug_level, int16_t* out) {
  std::vector<int64_t> dshape(dshape_, dshape_ + rank), cshape(cshape_, cshape_ + crank);
  std::string err;
  if (!CheckLegacy(dshape, cshape, cdf, precision, debug_level, &err)) return Fail(err);
  std::vector<int64_t> mb, ms;
  if (!MergeAxes(dshape, cshape, &mb, &ms, &err)) return Fail(err);
  if (mb.size() > 6) return Fail("Irregular broadcast pattern");
  int64_t n = 1;
  for (auto d : dshape) n *= d;
  const int64_t chip = cshape.back();
  BroadcastWalk walk(mb, ms, cdf);
  std::string src(reinterpret_cast<const char*>(bytes), nbytes);
  tfc::RangeDecoder dec{absl::string_view(src)};
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* strip = walk.Next();
    out[i] = static_cast<int16_t>(dec.Decode(absl::Span<const int32_t>(strip, chip), precision));
  }
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// PmfToQuantizedCdf, pmf_to_cdf_kernels.cc:58-101,104-208.  Uses std::sort / std::rotate like the
// reference so that tie order matches what the reference would produce with this libstdc++.
// ------------------------------------------------------------------------------------------
namespace {

struct Penalty {
  int32_t* v;
  double mass, key;
  double Next() const {
    if (*v <= 1) return std::numeric_limits<double>::infinity();
    return mass * (std::log2(*v) - std::log2(*v - 1));
  }
  Penalty(int32_t* p, double m) : v(p), mass(m) { key = Next(); }
  void Step() {
    --*v;
    key = Next();
  }
  friend bool operator<(const Penalty& a, const Penalty& b) { return a.key < b.key; }
};

struct Gain {
  int32_t* v;
  double mass, key;
  double Next() const {
    if (*v < 1) return -std::numeric_limits<double>::infinity();
    return mass * (std::log2(*v + 1) - std::log2(*v));
  }
  Gain(int32_t* p, double m) : v(p), mass(m) { key = Next(); }
  void Step() {
    ++*v;
    key = Next();
  }
  friend bool operator>(const Gain& a, const Gain& b) { return a.key > b.key; }
};

void PmfRow(const float* pmf, int64_t n, int precision, int32_t* cdf /* n+1 */) {
  const int32_t normalizer = 1 << precision;
  int32_t* q = cdf + 1;
  cdf[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t value = static_cast<int32_t>(std::rint(pmf[i] * normalizer));
    q[i] = std::max(value, 1);
  }
  int32_t sum = std::accumulate(q, q + n, 0);
  if (sum > normalizer) {
    std::vector<Penalty> queue;
    queue.reserve(n);
    for (int64_t i = 0; i < n; ++i) queue.emplace_back(&q[i], pmf[i]);
    std::sort(queue.begin(), queue.end());
    while (sum-- > normalizer) {
      queue[0].Step();
      auto it = std::find_if(std::next(queue.begin()), queue.end(),
                             [&queue](const Penalty& rhs) { return queue[0] < rhs; });
      std::rotate(queue.begin(), std::next(queue.begin()), it);
    }
  } else if (sum < normalizer) {
    std::vector<Gain> queue;
    queue.reserve(n);
    for (int64_t i = 0; i < n; ++i) queue.emplace_back(&q[i], pmf[i]);
    std::sort(queue.begin(), queue.end(), std::greater<Gain>());
    while (sum++ < normalizer) {
      queue[0].Step();
      auto it = std::find_if(std::next(queue.begin()), queue.end(),
                             [&queue](const Gain& rhs) { return queue[0] > rhs; });
      std::rotate(queue.begin(), std::next(queue.begin()), it);
    }
  }
  std::partial_sum(q, q + n, q);
}

}  // namespace

extern "C" {

int tfcref_pmf_to_cdf(const float* pmf, int64_t rows, int64_t n, int precision, int32_t* cdf) {
  if (!(0 < precision && precision <= 16))
    return Fail("`precision` must be in [1, 16]: " + std::to_string(precision));
  if (n <= 1) return Fail("`pmf` size should be at least 2 in the last axis.");
  for (int64_t i = 0; i < rows * n; ++i) {
    if (!(std::isfinite(pmf[i]) && pmf[i] >= 0))
      return Fail("`pmf` has non-finite or negative element: " + std::to_string(pmf[i]));
  }
  for (int64_t r = 0; r < rows; ++r) PmfRow(pmf + r * n, n, precision, cdf + r * (n + 1));
  return 0;
}

// StochasticRound, quantization_kernels.cc:35-95 restated: one xoshiro256+ stream seeded through std::seed_seq,
// one draw per element in element order.  `inputs` are the op's inputs already promoted to float32 (the op does
// static_cast<float>, exact for float16 / bfloat16).  An empty seed (clock seeding) is not reproducible and is
// rejected here.
int tfcref_stochastic_round(const float* inputs, int64_t n, float step_size, const int32_t* seed, int64_t seed_len,
                            int32_t* outputs) {
  if (seed_len <= 0) return Fail("the oracle needs an explicit seed");
  uint64_t state[4];
  std::seed_seq seq(seed, seed + seed_len);
  seq.generate(reinterpret_cast<uint32_t*>(state), reinterpret_cast<uint32_t*>(state + 4));
  auto next = [&state]() {
    const uint64_t result = state[0] + state[3];
    const uint64_t t = state[1] << 17;
    state[2] ^= state[0];
    state[3] ^= state[1];
    state[1] ^= state[2];
    state[0] ^= state[3];
    state[2] ^= t;
    state[3] = (state[3] << 45) | (state[3] >> (64 - 45));
    return result;
  };
  for (int64_t i = 0; i < n; ++i) {
    float number = inputs[i] / step_size;
    float integral = std::floor(number);
    outputs[i] = integral;
    float fractional = number - integral;
    float random = (next() >> 40) * 0x1.0p-24f;
    if (random < fractional) ++outputs[i];
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// RunLengthEncode / RunLengthDecode (cc/kernels/run_length_kernels.cc:52-262): the op loops restated around the
// reference's BitWriter / BitReader objects, which do every bit of the packing.
// ------------------------------------------------------------------------------------------
}  // extern "C"

namespace {
struct RunLengthCodes {
  int run_length_code, magnitude_code;
  bool non_zero_runs;
};

// :69-75
void WriteRunLength(const RunLengthCodes& c, tfc::BitWriter& enc, int32_t run_length) {
  if (c.run_length_code >= 0) {
    enc.WriteRice(run_length, c.run_length_code);
  } else {
    enc.WriteGamma(run_length + 1);
  }
}

// :77-92
void WriteNonZero(const RunLengthCodes& c, tfc::BitWriter& enc, int32_t sample) {
  const int32_t sign = sample > 0;
  enc.WriteOneBit(sign);
  if (c.magnitude_code >= 0) {
    enc.WriteRice(sign ? sample - 1 : -(sample + 1), c.magnitude_code);
  } else if (sample == std::numeric_limits<int32_t>::min()) {
    enc.WriteGamma(-(std::numeric_limits<int32_t>::min() + 1));   // not representable: the closest value instead
  } else {
    enc.WriteGamma(sign ? sample : -sample);
  }
}

// :158-168
absl::StatusOr<int32_t> ReadRunLength(const RunLengthCodes& c, tfc::BitReader& dec) {
  if (c.run_length_code >= 0) return dec.ReadRice(c.run_length_code);
  auto gamma = dec.ReadGamma();
  if (!gamma.ok()) return gamma;
  return *gamma - 1;
}

// :170-184
absl::StatusOr<int32_t> ReadNonZero(const RunLengthCodes& c, tfc::BitReader& dec) {
  auto positive = dec.ReadOneBit();
  if (!positive.ok()) return positive;
  if (c.magnitude_code >= 0) {
    auto rice = dec.ReadRice(c.magnitude_code);
    if (!rice.ok()) return rice;
    return *positive ? *rice + 1 : -*rice - 1;
  }
  auto gamma = dec.ReadGamma();
  if (!gamma.ok()) return gamma;
  return *positive ? *gamma : -*gamma;
}
}  // namespace

extern "C" {

// Returns the number of code bytes, or -(needed) if `cap` is too small.  :94-135
int64_t tfcref_run_length_encode(const int32_t* data, int64_t n, int rl_code, int mag_code, int rl_nz, uint8_t* out,
                                 int64_t cap) {
  const RunLengthCodes c{rl_code, mag_code, rl_nz != 0};
  tfc::BitWriter enc;
  const int32_t* const end = data + n;
  const int32_t* p = data;
  int32_t run_length_offset = 0;   // with runs of non-zeros too, only the first zero run can be empty
  while (p < end) {
    const int32_t* q = std::find_if_not(p, end, [](int32_t x) { return x == 0; });
    WriteRunLength(c, enc, static_cast<int32_t>(q - p) - run_length_offset);
    p = q;
    if (!(p < end)) break;
    if (c.non_zero_runs) {
      q = std::find_if(p, end, [](int32_t x) { return x == 0; });
      WriteRunLength(c, enc, static_cast<int32_t>(q - p) - 1);
      while (p < q) WriteNonZero(c, enc, *p++);
      run_length_offset = 1;
    } else {
      WriteNonZero(c, enc, *p++);
    }
  }
  const auto encoded = enc.GetData();
  const int64_t nb = static_cast<int64_t>(encoded.size());
  if (nb > cap) return -nb;
  std::memcpy(out, encoded.data(), encoded.size());
  return nb;
}

// 0 ok; 1 "Out of bits to read."; 2 "Exceeded maximum gamma bit width."; 3 "Decoded past end of tensor."
// (the same codes as the C port; the message is also left in tfcref_last_error()).  :186-250
int tfcref_run_length_decode(const uint8_t* code, int64_t n_bytes, int rl_code, int mag_code, int rl_nz, int32_t* data,
                             int64_t n) {
  const RunLengthCodes c{rl_code, mag_code, rl_nz != 0};
  auto fail = [](const absl::Status& st) {
    g_error = st.message();
    return g_error == "Out of bits to read." ? 1 : (g_error == "Exceeded maximum gamma bit width." ? 2 : 3);
  };
  const absl::Status past_end = absl::DataLossError("Decoded past end of tensor.");
  tfc::BitReader dec(absl::string_view(reinterpret_cast<const char*>(code), static_cast<size_t>(n_bytes)));
  std::memset(data, 0, static_cast<size_t>(n) * sizeof(int32_t));
  int32_t* const end = data + n;
  int32_t* p = data;
  int32_t run_length_offset = 0;
  while (p < end) {
    auto run_length = ReadRunLength(c, dec);
    if (!run_length.ok()) return fail(run_length.status());
    // the reference advances the pointer itself (p += ...); an index keeps a corrupt run length inside the language
    const int64_t at = (p - data) + static_cast<int64_t>(*run_length) + run_length_offset;
    if (!(at < n)) {
      if (at != n) return fail(past_end);
      break;
    }
    p = data + at;
    if (c.non_zero_runs) {
      run_length = ReadRunLength(c, dec);
      if (!run_length.ok()) return fail(run_length.status());
      const int64_t next_zero = at + static_cast<int64_t>(*run_length) + 1;
      if (next_zero > n) return fail(past_end);
      while (p < data + next_zero) {
        auto nonzero = ReadNonZero(c, dec);
        if (!nonzero.ok()) return fail(nonzero.status());
        *p++ = *nonzero;
      }
      run_length_offset = 1;
    } else {
      auto nonzero = ReadNonZero(c, dec);
      if (!nonzero.ok()) return fail(nonzero.status());
      *p++ = *nonzero;
    }
  }
  return 0;
}

int tfcref_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }
// Workers currently parked in the persistent pool (0 until the first multi-threaded call).
int tfcref_pool_threads() { return StreamPool::Get().size(); }

}  // extern "C"

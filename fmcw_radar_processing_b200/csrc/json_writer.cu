// Host-side streaming writer for the large numeric arrays of the reference's JSON payloads (RP:307-318,
// 358-367, 576-587).  MATLAB jsonencode conventions: matrices nest row-major, NaN / Inf become null.  Numbers are
// written in their shortest round-trip form (std::to_chars).  Rows are formatted by a pool of threads; at the
// reference's hop 1 the spectrogram payload is gigabytes of text.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fmcw_cuda.h"

namespace {

template <class T>
inline char* put_number(char* p, char* end, T v) {
  if (!std::isfinite(v)) { std::memcpy(p, "null", 4); return p + 4; }
  auto r = std::to_chars(p, end, v);
  return r.ptr;
}

template <class T>
void format_row(const T* a, uint64_t cols, int64_t col_stride, std::string& out) {
  out.resize(cols * 26 + 4);
  char* p = &out[0];
  char* end = p + out.size();
  *p++ = '[';
  for (uint64_t c = 0; c < cols; ++c) {
    if (c) *p++ = ',';
    p = put_number(p, end, a[(int64_t)c * col_stride]);
  }
  *p++ = ']';
  out.resize(p - out.data());
}

template <class T>
fmcw_status append_matrix(const char* path, const T* a, uint64_t rows, uint64_t cols, int64_t row_stride, int64_t col_stride,
                          int flatten_vectors) {
  if (!path || (!a && rows && cols)) return FMCW_ERR_POINTER;
  FILE* f = std::fopen(path, "ab");
  if (!f) return FMCW_ERR_SIZE;
  const bool flat = flatten_vectors && (rows == 1 || cols == 1);      // jsonencode flattens row and column vectors
  if (flat) {
    std::string s;
    if (rows == 1) format_row(a, cols, col_stride, s);
    else format_row(a, rows, row_stride, s);
    std::fwrite(s.data(), 1, s.size(), f);
    std::fclose(f);
    return FMCW_OK;
  }
  unsigned nthr = std::thread::hardware_concurrency();
  if (nthr == 0) nthr = 4;
  if (nthr > 32) nthr = 32;
  const uint64_t batch = nthr * 4;                                    // rows formatted per round
  std::vector<std::string> bufs(batch);
  std::fputc('[', f);
  for (uint64_t r0 = 0; r0 < rows; r0 += batch) {
    const uint64_t nr = rows - r0 < batch ? rows - r0 : batch;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthr; ++t)
      pool.emplace_back([&, t]() {
        for (uint64_t i = t; i < nr; i += nthr) format_row(a + (int64_t)(r0 + i) * row_stride, cols, col_stride, bufs[i]);
      });
    for (auto& th : pool) th.join();
    for (uint64_t i = 0; i < nr; ++i) {
      if (r0 + i) std::fputc(',', f);
      std::fwrite(bufs[i].data(), 1, bufs[i].size(), f);
    }
  }
  std::fputc(']', f);
  const bool bad = std::ferror(f) != 0;
  std::fclose(f);
  return bad ? FMCW_ERR_SIZE : FMCW_OK;
}

}  // namespace

extern "C" {

fmcw_status fmcw_json_append_f32(const char* path, const float* a, uint64_t rows, uint64_t cols, int64_t row_stride,
                                 int64_t col_stride, int flatten_vectors) {
  return append_matrix(path, a, rows, cols, row_stride, col_stride, flatten_vectors);
}

fmcw_status fmcw_json_append_f64(const char* path, const double* a, uint64_t rows, uint64_t cols, int64_t row_stride,
                                 int64_t col_stride, int flatten_vectors) {
  return append_matrix(path, a, rows, cols, row_stride, col_stride, flatten_vectors);
}

}  // extern "C"

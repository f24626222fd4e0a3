// Interchange formats, host side of the C ABI.
//
//  * NumPy .npy (format 1.0, little-endian float64, C order) for the batched arrays of the ABI --
//    coefficients [B][K][D][N], samples [B][M][n_deriv][D], times [B][K]: the binary replacement for the
//    reference's only serialisation, and what numpy.load / numpy.save read and write directly.
//  * The reference's sampled-trajectory table (ref printMatlabSampledTrajectory, NL.i:2567-2662):
//    rows [t, position(D), velocity(D), acceleration(D), jerk(D), snap(D), t_vertex] at a fixed time
//    increment per segment (t advances by repeated addition from 0 while t < T_i, the row time is t plus the
//    start of the segment), sum_i (ceil(T_i / dt) + 1) rows of which the unused ones stay zero, and the
//    cumulative vertex times in rows 0..K-1 of the last column.  The values come from the GPU sampler
//    (minsnap_sample_at, every segment as its own one-segment trajectory so that the local time decides).
//    The text form writes 17 significant digits (the reference streams an Eigen matrix at 6).
#include "../../include/minsnap_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

bool parse_npy_header(FILE* f, std::string* descr, bool* fortran, std::vector<int64_t>* shape) {
  unsigned char magic[10];
  if (std::fread(magic, 1, 10, f) != 10 || std::memcmp(magic, "\x93NUMPY", 6) != 0) return false;
  size_t hlen = 0;
  if (magic[6] == 1) {
    hlen = magic[8] | (size_t(magic[9]) << 8);
  } else {
    unsigned char more[2];
    if (std::fread(more, 1, 2, f) != 2) return false;
    hlen = magic[8] | (size_t(magic[9]) << 8) | (size_t(more[0]) << 16) | (size_t(more[1]) << 24);
  }
  std::string h(hlen, '\0');
  if (std::fread(&h[0], 1, hlen, f) != hlen) return false;
  size_t p = h.find("'descr'");
  if (p == std::string::npos) return false;
  p = h.find('\'', h.find(':', p));
  const size_t q = h.find('\'', p + 1);
  *descr = h.substr(p + 1, q - p - 1);
  p = h.find("'fortran_order'");
  if (p == std::string::npos) return false;
  *fortran = h.compare(h.find_first_not_of(" ", h.find(':', p) + 1), 4, "True") == 0;
  p = h.find("'shape'");
  if (p == std::string::npos) return false;
  p = h.find('(', p);
  const size_t e = h.find(')', p);
  shape->clear();
  std::string inner = h.substr(p + 1, e - p - 1);
  size_t pos = 0;
  while (pos < inner.size()) {
    while (pos < inner.size() && (inner[pos] == ' ' || inner[pos] == ',')) ++pos;
    if (pos >= inner.size()) break;
    size_t used = 0;
    shape->push_back(std::stoll(inner.substr(pos), &used));
    pos += used;
  }
  return true;
}

}  // namespace

extern "C" {

int minsnap_npy_write_f64(const char* path, const double* h_data, int ndim, const int64_t* shape) {
  if (!path || ndim < 0 || ndim > 8 || (ndim > 0 && !shape)) return MINSNAP_ERR_ARG;
  size_t count = 1;
  std::string dims = "(";
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 0) return MINSNAP_ERR_ARG;
    count *= (size_t)shape[i];
    dims += std::to_string(shape[i]);
    dims += (ndim == 1 || i + 1 < ndim) ? ", " : "";
  }
  if (!dims.empty() && dims.back() == ' ' && ndim > 1) dims.pop_back(), dims.pop_back();
  dims += ")";
  if (count > 0 && !h_data) return MINSNAP_ERR_ARG;
  std::string header = "{'descr': '<f8', 'fortran_order': False, 'shape': " + dims + ", }";
  const size_t unpadded = 10 + header.size() + 1;
  header.append((64 - unpadded % 64) % 64, ' ');
  header.push_back('\n');
  FILE* f = std::fopen(path, "wb");
  if (!f) return MINSNAP_ERR_ARG;
  const unsigned char magic[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(header.size() & 0xff),
                                   (unsigned char)(header.size() >> 8)};
  bool ok = std::fwrite(magic, 1, 10, f) == 10 && std::fwrite(header.data(), 1, header.size(), f) == header.size();
  if (ok && count > 0) ok = std::fwrite(h_data, sizeof(double), count, f) == count;   // little-endian host assumed (x86-64, aarch64)
  ok = (std::fclose(f) == 0) && ok;
  return ok ? MINSNAP_OK : MINSNAP_ERR_ARG;
}

int minsnap_npy_read_f64(const char* path, double* h_data, size_t capacity, int* ndim, int64_t* shape) {
  if (!path || !ndim || !shape) return MINSNAP_ERR_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) return MINSNAP_ERR_ARG;
  std::string descr;
  bool fortran = false;
  std::vector<int64_t> shp;
  if (!parse_npy_header(f, &descr, &fortran, &shp) || shp.size() > 8) {
    std::fclose(f);
    return MINSNAP_ERR_ARG;
  }
  if ((descr != "<f8" && descr != "=f8" && descr != "|f8") || fortran) {
    std::fclose(f);
    return MINSNAP_ERR_UNSUPPORTED;   // only little-endian float64 in C order
  }
  *ndim = (int)shp.size();
  size_t count = 1;
  for (size_t i = 0; i < shp.size(); ++i) {
    shape[i] = shp[i];
    count *= (size_t)shp[i];
  }
  int rc = MINSNAP_OK;
  if (h_data) {
    if (capacity < count) rc = MINSNAP_ERR_WORKSPACE;
    else if (count > 0 && std::fread(h_data, sizeof(double), count, f) != count) rc = MINSNAP_ERR_ARG;
  }
  std::fclose(f);
  return rc;
}

int minsnap_sampled_table_rows(int K, const double* h_times, double dt) {
  if (K < 1 || !h_times || !(dt > 0.0)) return -1;
  long rows = 0;
  for (int i = 0; i < K; ++i) rows += static_cast<int>(std::ceil(h_times[i] / dt)) + 1;   // ref NL.i:2580-2582
  return rows > 0x7fffffffL ? -1 : (int)rows;
}

int minsnap_sampled_table_host(int K, int D, int N, const double* h_coeffs, const double* h_times, double dt,
                               double* h_table, int capacity_rows, int* rows_out, int* cols_out) {
  if (K < 1 || D < 1 || !h_coeffs || !h_times || !(dt > 0.0) || !h_table) return MINSNAP_ERR_ARG;
  const int rows = minsnap_sampled_table_rows(K, h_times, dt);
  const int cols = 5 * D + 2;
  if (rows < 0) return MINSNAP_ERR_ARG;
  if (rows_out) *rows_out = rows;
  if (cols_out) *cols_out = cols;
  if (capacity_rows < rows) return MINSNAP_ERR_WORKSPACE;
  // local sample times per segment: t = 0; t < T_i; t += dt (repeated addition, ref NL.i:2609)
  std::vector<std::vector<double> > local(K);
  size_t m_max = 1;
  for (int i = 0; i < K; ++i) {
    for (double t = 0.0; t < h_times[i]; t += dt) local[i].push_back(t);
    m_max = std::max(m_max, local[i].size());
  }
  // every segment as its own one-segment trajectory: instants past the end of a shorter segment are padded with
  // its duration (outside: zeros, never used)
  std::vector<double> t_grid((size_t)K * m_max), out((size_t)K * m_max * 5 * D);
  std::vector<int32_t> seg((size_t)K * m_max);
  for (int i = 0; i < K; ++i)
    for (size_t m = 0; m < m_max; ++m) t_grid[(size_t)i * m_max + m] = m < local[i].size() ? local[i][m] : h_times[i];
  const int rc = minsnap_sample_at_host(K, 1, D, N, h_coeffs, h_times, (int)m_max, t_grid.data(), (long)m_max, 5, out.data(),
                                        seg.data());
  if (rc != MINSNAP_OK) return rc;
  std::memset(h_table, 0, sizeof(double) * (size_t)rows * cols);
  int j = 0;
  double current_segment_time = 0.0;
  for (int i = 0; i < K; ++i) {
    for (size_t m = 0; m < local[i].size(); ++m) {
      if (j < rows) {
        double* row = h_table + (size_t)j * cols;
        row[0] = local[i][m] + current_segment_time;
        const double* v = out.data() + ((size_t)i * m_max + m) * 5 * D;
        for (int e = 0; e < 5 * D; ++e) row[1 + e] = v[e];   // [derivative][dimension], as the reference lays the row out
        ++j;
      }
    }
    current_segment_time += h_times[i];
    h_table[(size_t)i * cols + 1 + 5 * D] = current_segment_time;   // ref NL.i:2653
  }
  return MINSNAP_OK;
}

int minsnap_table_write_text(const char* path, const double* h_table, int rows, int cols) {
  if (!path || !h_table || rows < 0 || cols < 1) return MINSNAP_ERR_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) return MINSNAP_ERR_ARG;
  for (int r = 0; r < rows; ++r) {
    for (int c = 0; c < cols; ++c) std::fprintf(f, c ? " %.17g" : "%.17g", h_table[(size_t)r * cols + c]);
    if (r + 1 < rows) std::fputc('\n', f);
  }
  return std::fclose(f) == 0 ? MINSNAP_OK : MINSNAP_ERR_ARG;
}

}  // extern "C"

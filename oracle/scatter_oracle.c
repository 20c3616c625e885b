/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the reference's
 * scatter-add pullback and gather at sizes NumPy's np.add.at is too slow for (C5:
 * 4 194 304 source columns of 256 Float32).
 *
 * Follows /root/reference/src/helpers.jl:7-11 (ungetindex! -> NNlib.scatter!(+, dx, dy, idx))
 * whose CPU method, like ChainRules' ∇getindex!, is the sequential loop
 *     for k in eachindex(dy) (column-major):  dx[I_k] += dy[k]
 * (SURVEY.md Appendix A.8): per destination a left fold in source order from +0.0f.
 * Compiled with -O2 and NO -ffast-math so every add is an IEEE single add. */
#include <stdint.h>
#include <string.h>

/* dx[rows x n_dst] = 0;  dx[:, idx[k]-1] += dy[:, k]  for k = 0..n_src-1 in order */
void oracle_scatter_add_cols_f32(float* dx, int64_t rows, int64_t n_dst, const float* dy,
                                 const int64_t* idx, int64_t n_src) {
  memset(dx, 0, (size_t)rows * (size_t)n_dst * sizeof(float));
  for (int64_t k = 0; k < n_src; k++) {
    float* d = dx + (idx[k] - 1) * rows;
    const float* s = dy + k * rows;
    for (int64_t r = 0; r < rows; r++) d[r] = d[r] + s[r];
  }
}

/* y[:, k] = x[:, idx[k]-1] */
void oracle_gather_cols_f32(float* y, const float* x, int64_t rows, const int64_t* idx, int64_t n_src) {
  for (int64_t k = 0; k < n_src; k++) memcpy(y + k * rows, x + (idx[k] - 1) * rows, (size_t)rows * sizeof(float));
}

/* C5 forward value: sum(E[:, idx] .* V) accumulated in double (order-independent check value) */
double oracle_gather_dot_f64(const float* E, const float* V, int64_t rows, const int64_t* idx, int64_t n_src) {
  double acc = 0.0;
  for (int64_t k = 0; k < n_src; k++) {
    const float* e = E + (idx[k] - 1) * rows;
    const float* v = V + k * rows;
    for (int64_t r = 0; r < rows; r++) acc += (double)(e[r] * v[r]);
  }
  return acc;
}

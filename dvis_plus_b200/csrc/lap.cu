// Batched linear-assignment (Hungarian / shortest augmenting path) on the GPU, one CTA per problem, plus the
// composition of the per-frame assignments into the tracker's index chain.
//
// Replaces the per-frame `C.cpu()` + scipy.optimize.linear_sum_assignment in Noiser.match_embds
// (P/dvis_Plus/noiser.py:43-56, called once per frame from P/dvis_Plus/tracker.py:224,285) -- a device->host sync that
// serialises the tracker (SURVEY.md section 3 "hot loop 4", section 8f rank 3).  The matching chain depends only on
// the segmenter's frame embeddings: last_frame_embeds_t = cur_t[idx_t] and idx_t = LSA(cost(last_{t-1}, cur_t)).
// With sigma_t = LSA against the UN-permuted previous frame, idx_t = sigma_t o idx_{t-1} (the optimum of a row-permuted
// assignment problem is the permuted optimum), so all T problems are independent and run concurrently.
//
// Algorithm: the classic O(n^3) potentials formulation (rows are inserted one at a time; each insertion grows an
// alternating tree column by column with a Dijkstra-like scan).  The scan over columns and the arg-min are parallel
// over the CTA's threads (one column per thread); potentials in double like SciPy's implementation.  For float costs
// without exact ties the optimal assignment is unique, so the result equals SciPy's.
#include <algorithm>

#include "common.cuh"

namespace dvis {
namespace {

constexpr int kMaxN = 1024;

struct MinIdx {
  double v;
  int j;
};
__device__ __forceinline__ MinIdx min2(MinIdx a, MinIdx b) { return (b.v < a.v || (b.v == a.v && b.j < a.j)) ? b : a; }

// One problem per CTA: n "rows" are inserted one at a time into m >= n "columns"; element (row i, column j) is
// a[i * si + j * sj] (so a cost matrix with more rows than columns is solved through its transpose without being copied).  The assignment is left in shared memory: p[j] = row matched to column j (1-based, 0 = free).
__device__ __forceinline__ int *lap_solve(const float *__restrict__ a, int n, int m, int64_t si, int64_t sj, double *sm) {
  double *u = sm;                        // row potentials      [n+1]
  double *v = u + (n + 1);               // column potentials   [m+1]
  double *minv = v + (m + 1);            // [m+1]
  int *p = reinterpret_cast<int *>(minv + (m + 1));   // p[j] = row assigned to column j (1-based, 0 = none)  [m+1]
  int *way = p + (m + 1);                // [m+1]
  int *used = way + (m + 1);             // [m+1]
  __shared__ MinIdx red[32];
  __shared__ int s_j0;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int j = tid; j <= n; j += nthr) u[j] = 0;
  for (int j = tid; j <= m; j += nthr) { v[j] = 0; p[j] = 0; way[j] = 0; }
  __syncthreads();
  for (int i = 1; i <= n; ++i) {
    if (tid == 0) { p[0] = i; s_j0 = 0; }
    for (int j = tid; j <= m; j += nthr) { minv[j] = INFINITY; used[j] = 0; }
    __syncthreads();
    while (true) {
      const int j0 = s_j0;
      const int i0 = p[j0];
      const double ui0 = u[i0];
      // scan the unused columns: relax minv through row i0, find the closest one
      MinIdx best{INFINITY, m + 1};
      for (int j = tid + 1; j <= m; j += nthr) {
        if (j == j0 || used[j]) continue;
        float c = a[(int64_t)(i0 - 1) * si + (int64_t)(j - 1) * sj];
        if (c != c) c = 0.f;                                  // NaN -> 0 like noiser.py:52
        const double cur = double(c) - ui0 - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        best = min2(best, MinIdx{minv[j], j});
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        MinIdx other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.j, o)};
        best = min2(best, other);
      }
      if ((tid & 31) == 0) red[tid >> 5] = best;
      __syncthreads();
      if (tid < 32) {
        MinIdx b = tid < (nthr + 31) / 32 ? red[tid] : MinIdx{INFINITY, m + 1};
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          MinIdx other{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.j, o)};
          b = min2(b, other);
        }
        if (tid == 0) red[0] = b;
      }
      __syncthreads();
      const double delta = red[0].v;
      const int j1 = red[0].j;
      // read the exit condition BEFORE the barrier below: once a thread leaves the loop, thread 0 starts rewriting p[]
      // (augmentation) and a slower warp must not see the updated p[j1]
      const bool reached_free_column = p[j1] == 0;
      // update potentials: tree columns (incl. j0, which joins the tree now) move with their rows
      for (int j = tid; j <= m; j += nthr) {
        if (j == j0 || used[j]) { u[p[j]] += delta; v[j] -= delta; }   // j0 first: thread 0 sets used[j0] below, unsynchronised
        else minv[j] -= delta;
      }
      if (tid == 0) { used[j0] = 1; s_j0 = j1; }
      __syncthreads();
      if (reached_free_column) break;
    }
    // augment along the alternating path
    if (tid == 0) {
      int j0 = s_j0;
      do {
        const int j1 = way[j0];
        p[j0] = p[j1];
        j0 = j1;
      } while (j0);
    }
    __syncthreads();
  }
  return p;
}

// cost: (T, n, n) row-major, rows = reference items, cols = current items.  sigma[t][row] = assigned column.
__global__ void __launch_bounds__(1024) lap_kernel(const float *__restrict__ cost, int n, int64_t *__restrict__ sigma) {
  extern __shared__ double sm[];
  const int *p = lap_solve(cost + (size_t)blockIdx.x * n * n, n, n, n, 1, sm);
  for (int j = threadIdx.x + 1; j <= n; j += blockDim.x) sigma[(size_t)blockIdx.x * n + (p[j] - 1)] = j - 1;
}

// Rectangular problems, cost (B, rows, cols) row-major: row_to_col[b][r] = column assigned to row r, -1 if r stays
// unmatched (possible only when rows > cols) -- scipy.optimize.linear_sum_assignment's (row_ind, col_ind) as a dense map.
__global__ void __launch_bounds__(1024) lap_rect_kernel(const float *__restrict__ cost, int rows, int cols,
                                                        int64_t *__restrict__ row_to_col) {
  extern __shared__ double sm[];
  const float *a = cost + (size_t)blockIdx.x * rows * cols;
  int64_t *out = row_to_col + (size_t)blockIdx.x * rows;
  if (rows <= cols) {
    const int *p = lap_solve(a, rows, cols, cols, 1, sm);               // insert the rows, choose among the columns
    for (int j = threadIdx.x + 1; j <= cols; j += blockDim.x)
      if (p[j]) out[p[j] - 1] = j - 1;
  } else {
    const int *p = lap_solve(a, cols, rows, 1, cols, sm);               // transpose: insert the columns, choose among the rows
    for (int j = threadIdx.x + 1; j <= rows; j += blockDim.x) out[j - 1] = p[j] ? p[j] - 1 : -1;
  }
}

// idx[0] = sigma[0] o idx_init (or sigma[0] when idx_init is null); idx[t] = sigma[t] o idx[t-1]
__global__ void __launch_bounds__(1024) lap_chain_kernel(const int64_t *__restrict__ sigma, const int64_t *__restrict__ idx_init,
                                                         int T, int n, int64_t *__restrict__ idx) {
  for (int t = 0; t < T; ++t) {
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
      const int64_t prev = t == 0 ? (idx_init ? idx_init[r] : r) : idx[(size_t)(t - 1) * n + r];
      idx[(size_t)t * n + r] = sigma[(size_t)t * n + prev];
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_lap_chain(const float *cost, int T, int n, const int64_t *idx_init, int64_t *sigma, int64_t *idx,
                              void *stream) {
  DVIS_REQUIRE(cost && sigma && idx, "lap_chain: null pointer argument");
  DVIS_REQUIRE(T > 0 && n > 0 && n <= kMaxN, "lap_chain: need T > 0 and 0 < n <= %d (n=%d)", kMaxN, n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int threads = 32;
  while (threads < n && threads < 1024) threads <<= 1;
  const size_t smem = size_t(n + 1) * (3 * sizeof(double) + 3 * sizeof(int)) + 16;
  lap_kernel<<<T, threads, smem, s>>>(cost, n, sigma);
  if (int rc = check_launch("lap_kernel")) return rc;
  lap_chain_kernel<<<1, threads, 0, s>>>(sigma, idx_init, T, n, idx);
  return check_launch("lap_chain_kernel");
}

extern "C" int dvis_lap_rect(const float *cost, int B, int rows, int cols, int64_t *row_to_col, void *stream) {
  DVIS_REQUIRE(cost && row_to_col, "lap_rect: null pointer argument");
  DVIS_REQUIRE(B > 0 && rows > 0 && cols > 0 && rows <= kMaxN && cols <= kMaxN, "lap_rect: need B > 0 and 1 <= rows, cols <= %d (%d x %d)",
               kMaxN, rows, cols);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = std::min(rows, cols), m = std::max(rows, cols);
  int threads = 32;
  while (threads < m && threads < 1024) threads <<= 1;
  const size_t smem = size_t(n + 1) * sizeof(double) + size_t(m + 1) * (2 * sizeof(double) + 3 * sizeof(int)) + 16;
  lap_rect_kernel<<<B, threads, smem, s>>>(cost, rows, cols, row_to_col);
  return check_launch("lap_rect_kernel");
}

// Synthetic matrices of BASELINE.json generated directly in device memory.
// The same definitions are restated on the CPU in oracle/cask_oracle.c (oracle_gen_*) and the two
// are compared bit for bit in tests/test_gpu_synth.py.
#include "ctx.cuh"

namespace caskb200 {
namespace {

// number of in-grid neighbours (incl. centre) of a 27-point stencil along one axis
__host__ __device__ inline int span3(int c, int N) { return 1 + (c > 0) + (c < N - 1); }

__host__ __device__ inline int row_len(int kind, int N, int64_t row) {
  if (kind == CASK_B200_SYNTH_POISSON2D) {
    const int i = (int)(row / N), j = (int)(row % N);
    return 1 + (i > 0) + (j > 0) + (j < N - 1) + (i < N - 1);
  }
  const int x = (int)(row % N), y = (int)((row / N) % N), z = (int)(row / ((int64_t)N * N));
  if (kind == CASK_B200_SYNTH_POISSON3D27) return span3(x, N) * span3(y, N) * span3(z, N);
  return 1 + (x > 0) + (x < N - 1) + (y > 0) + (y < N - 1) + (z > 0) + (z < N - 1);
}

// nnz in rows [0, row) — closed forms, so any row stripe can be generated independently
__host__ __device__ inline int64_t nnz_before(int kind, int N, int64_t row) {
  const int64_t n1 = N;
  if (kind == CASK_B200_SYNTH_POISSON2D) {
    // full grid rows i' < i contribute 5N - 2 - [i'==0] - [i'==N-1]; partial row handled per column
    const int64_t i = row / N, j = row % N;
    int64_t total = 0;
    for (int64_t ii = 0; ii < i; ii++) total += 5 * n1 - 2 - (ii == 0 ? n1 : 0) - (ii == N - 1 ? n1 : 0);
    for (int64_t jj = 0; jj < j; jj++) total += row_len(kind, N, i * N + jj);
    return total;
  }
  const int64_t plane = n1 * n1;
  const int64_t z = row / plane, rem = row % plane, y = rem / N, x = rem % N;
  int64_t total = 0;
  if (kind == CASK_B200_SYNTH_POISSON3D27) {
    const int64_t line = 3 * n1 - 2;        // sum over x of span3
    const int64_t pl = line * line;         // sum over (x,y) of span3(x)span3(y)
    for (int64_t zz = 0; zz < z; zz++) total += pl * span3((int)zz, N);
    for (int64_t yy = 0; yy < y; yy++) total += line * span3((int)yy, N) * span3((int)z, N);
    for (int64_t xx = 0; xx < x; xx++) total += (int64_t)span3((int)xx, N) * span3((int)y, N) * span3((int)z, N);
    return total;
  }
  // 7-point: row length = 1 + axis neighbours
  const int64_t line_x = 3 * n1 - 2;  // sum over x of (1 + [x>0] + [x<N-1])
  for (int64_t zz = 0; zz < z; zz++) {
    const int64_t zn = (zz > 0) + (zz < N - 1);
    total += n1 * (line_x + 2 * n1 - 2) + zn * plane;   // sum over plane of (1+xn+yn) + zn per point
  }
  const int64_t zn = (z > 0) + (z < N - 1);
  for (int64_t yy = 0; yy < y; yy++) total += line_x + ((yy > 0) + (yy < N - 1) + zn) * n1;
  for (int64_t xx = 0; xx < x; xx++) total += 1 + (xx > 0) + (xx < N - 1) + (y > 0) + (y < N - 1) + zn;
  return total;
}

// Pass 1: row lengths -> row_ptr (rebased so that row_ptr[0] = 0) via the closed form at CTA starts.
__global__ void synth_fill_kernel(int kind, int N, int64_t row0, int64_t nrows, int64_t base_nnz,
                                  int32_t* __restrict__ row_ptr, int32_t* __restrict__ col, double* __restrict__ val) {
  // each CTA handles 256 consecutive rows; thread 0 computes the CTA's starting offset in closed form,
  // then a CTA scan of the row lengths positions every row.
  __shared__ int32_t scan[256];
  __shared__ int64_t cta_base;
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t row = row0 + r;
  const int len = r < nrows ? row_len(kind, N, row) : 0;
  scan[threadIdx.x] = len;
  if (threadIdx.x == 0) cta_base = nnz_before(kind, N, row0 + (int64_t)blockIdx.x * 256) - base_nnz;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    int32_t v = threadIdx.x >= d ? scan[threadIdx.x - d] : 0;
    __syncthreads();
    scan[threadIdx.x] += v;
    __syncthreads();
  }
  if (r >= nrows) return;
  int64_t k = cta_base + scan[threadIdx.x] - len;
  row_ptr[r] = (int32_t)k;
  if (r == nrows - 1) row_ptr[nrows] = (int32_t)(k + len);
  if (kind == CASK_B200_SYNTH_POISSON2D) {
    const int i = (int)(row / N), j = (int)(row % N);
    if (i > 0) { col[k] = (int32_t)(row - N); val[k++] = -1.0; }
    if (j > 0) { col[k] = (int32_t)(row - 1); val[k++] = -1.0; }
    col[k] = (int32_t)row; val[k++] = 4.0;
    if (j < N - 1) { col[k] = (int32_t)(row + 1); val[k++] = -1.0; }
    if (i < N - 1) { col[k] = (int32_t)(row + N); val[k++] = -1.0; }
    return;
  }
  const int x = (int)(row % N), y = (int)((row / N) % N), z = (int)(row / ((int64_t)N * N));
  if (kind == CASK_B200_SYNTH_POISSON3D27) {
    for (int dz = -1; dz <= 1; dz++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          const int zz = z + dz, yy = y + dy, xx = x + dx;
          if (zz < 0 || zz >= N || yy < 0 || yy >= N || xx < 0 || xx >= N) continue;
          col[k] = (int32_t)(((int64_t)zz * N + yy) * N + xx);
          val[k++] = (dz == 0 && dy == 0 && dx == 0) ? 26.0 : -1.0;
        }
    return;
  }
  const double cx = 0.5, cy = 0.25, cz = 0.125;  // cell Peclet 0.5 * (1, 0.5, 0.25), upwind
  const int64_t NN = (int64_t)N * N;
  if (z > 0) { col[k] = (int32_t)(row - NN); val[k++] = -1.0 - cz; }
  if (y > 0) { col[k] = (int32_t)(row - N); val[k++] = -1.0 - cy; }
  if (x > 0) { col[k] = (int32_t)(row - 1); val[k++] = -1.0 - cx; }
  col[k] = (int32_t)row; val[k++] = 6.0 + cx + cy + cz;
  if (x < N - 1) { col[k] = (int32_t)(row + 1); val[k++] = -1.0; }
  if (y < N - 1) { col[k] = (int32_t)(row + N); val[k++] = -1.0; }
  if (z < N - 1) { col[k] = (int32_t)(row + NN); val[k++] = -1.0; }
}

}  // namespace
}  // namespace caskb200

using namespace caskb200;

extern "C" int cask_b200_synth_rows(int32_t kind, int32_t N, int64_t* n) {
  if (!n || N <= 0 || kind < 0 || kind > 2) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "synth_rows: bad arguments");
  *n = kind == CASK_B200_SYNTH_POISSON2D ? (int64_t)N * N : (int64_t)N * N * N;
  return CASK_B200_OK;
}

extern "C" int cask_b200_synth_nnz(int32_t kind, int32_t N, int64_t row0, int64_t nrows, int64_t* nnz) {
  int64_t n = 0;
  CB_TRY(cask_b200_synth_rows(kind, N, &n));
  if (!nnz || row0 < 0 || nrows < 0 || row0 + nrows > n) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "synth_nnz: bad row range");
  *nnz = nnz_before(kind, N, row0 + nrows) - nnz_before(kind, N, row0);
  return CASK_B200_OK;
}

extern "C" int cask_b200_synth_device(int32_t kind, int32_t N, int64_t row0, int64_t nrows, int32_t* d_row_ptr,
                                      int32_t* d_col_ind, double* d_values, void* cuda_stream) {
  int64_t nnz = 0;
  CB_TRY(cask_b200_synth_nnz(kind, N, row0, nrows, &nnz));
  if (nnz > INT32_MAX) return fail(CASK_B200_ERR_UNSUPPORTED, "synth: stripe has more than INT32_MAX nonzeros");
  if (nrows == 0) return CASK_B200_OK;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  synth_fill_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, s>>>(kind, N, row0, nrows, nnz_before(kind, N, row0),
                                                                   d_row_ptr, d_col_ind, d_values);
  CB_CUDA(cudaGetLastError());
  return CASK_B200_OK;
}

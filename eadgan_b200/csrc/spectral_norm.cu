// Legacy torch.nn.utils.spectral_norm (SpectralNorm.compute_weight in
// torch/nn/utils/spectral_norm.py) as HBM-bound mat-vec passes over W[rows, cols]:
//   v <- normalize(W^T u);  u <- normalize(W v);  sigma = u . (W v);  W_sn = W / sigma
// with normalize(x) = x / max(||x||_2, eps), eps = 1e-12, one power iteration per
// training-mode forward, u/v updated in place (SURVEY.md sections 2.3, 7.3-6).
// Reference call sites: spectral_norm(...) at celebA/EAD-GAN_celebA.py:110-119,
// dSprites/rp.py:95-109,165-183, MNIST/EAD-GAN_rpqmnxy.py:107,124,143,161-163.
// Algorithmic bytes: 3 reads + 1 write of W per forward (16 B / weight).
#include "common.cuh"

namespace {

// t[split][j] = sum_{i in row range of split} W[i,j] * u[i];  block = 128 threads = 128 columns
__global__ void __launch_bounds__(128) sn_wtu_kernel(const float* __restrict__ W, const float* __restrict__ u,
                                                     float* __restrict__ t, int rows, int cols, int rows_per_blk) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_blk;
  const int r1 = min(rows, r0 + rows_per_blk);
  if (j >= cols) return;
  float acc = 0.f;
  int i = r0;
  for (; i + 4 <= r1; i += 4) {
    const float a0 = W[(int64_t)i * cols + j], a1 = W[(int64_t)(i + 1) * cols + j];
    const float a2 = W[(int64_t)(i + 2) * cols + j], a3 = W[(int64_t)(i + 3) * cols + j];
    acc = fmaf(a0, u[i], acc); acc = fmaf(a1, u[i + 1], acc);
    acc = fmaf(a2, u[i + 2], acc); acc = fmaf(a3, u[i + 3], acc);
  }
  for (; i < r1; ++i) acc = fmaf(W[(int64_t)i * cols + j], u[i], acc);
  t[(int64_t)blockIdx.y * cols + j] = acc;   // per-split partial: summed in a fixed order by sn_vsum_kernel
}

// x = sum over `parts` partial vectors (fixed order: deterministic), stored over the first slab; per-block partial
// of ||x||^2 -> nrm[block].  Replaces a single-block pass over parts x n floats (24 us at 19 x 8192).
__global__ void __launch_bounds__(256) sn_vsum_kernel(float* __restrict__ t, int n, int parts, float* __restrict__ nrm) {
  __shared__ float red[32];
  const int i = blockIdx.x * 256 + threadIdx.x;
  float x = 0.f;
  if (i < n) {
    x = t[i];
    for (int p = 1; p < parts; ++p) x += t[(int64_t)p * n + i];
    t[i] = x;
  }
  const float s = eg_block_sum(x * x, red);
  if (threadIdx.x == 0) nrm[blockIdx.x] = s;
}

// s[i] = sum_j W[i,j] v[j] with v = x / max(||x||, eps) formed on the fly (x: un-normalised W^T u, nrm: the
// partials of ||x||^2, re-summed by every block in the same order); block i also stores its slice of v.
// x == nullptr: v is taken as is (no power iteration).  One block per row.
__global__ void __launch_bounds__(256) sn_wv_kernel(const float* __restrict__ W, const float* __restrict__ x,
                                                    const float* __restrict__ nrm, int nparts, float eps,
                                                    float* __restrict__ v, float* __restrict__ s, int rows, int cols) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  const float* wr = W + (int64_t)i * cols;
  float inv = 1.f;
  if (x != nullptr) {
    float q = 0.f;
    for (int p = 0; p < nparts; ++p) q += nrm[p];
    inv = 1.f / fmaxf(sqrtf(q), eps);
    const int chunk = (cols + rows - 1) / rows;
    for (int j = i * chunk + threadIdx.x; j < min(cols, (i + 1) * chunk); j += blockDim.x) v[j] = x[j] * inv;
  }
  const float* src = x != nullptr ? x : v;
  float acc = 0.f;
  if ((cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(wr) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    for (int j = threadIdx.x; j < (cols >> 2); j += blockDim.x) {
      const float4 a = reinterpret_cast<const float4*>(wr)[j];
      float4 b = reinterpret_cast<const float4*>(src)[j];
      if (x != nullptr) { b.x *= inv; b.y *= inv; b.z *= inv; b.w *= inv; }
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
      acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
  } else {
    for (int j = threadIdx.x; j < cols; j += blockDim.x) acc = fmaf(wr[j], x != nullptr ? src[j] * inv : src[j], acc);
  }
  acc = eg_block_sum(acc, red);
  if (threadIdx.x == 0) s[i] = acc;
}

// single block: if update_u: u = s / max(||s||, eps);  sigma = sum u[i]*s[i]
__global__ void __launch_bounds__(1024) sn_sigma_kernel(const float* __restrict__ s, float* __restrict__ u, int rows,
                                                        float eps, int update_u, float* __restrict__ sigma) {
  __shared__ float red[32];
  if (update_u) {
    float q = 0.f;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) q = fmaf(s[i], s[i], q);
    q = eg_block_sum(q, red);
    const float inv = 1.f / fmaxf(sqrtf(q), eps);
    for (int i = threadIdx.x; i < rows; i += blockDim.x) u[i] = s[i] * inv;
    __syncthreads();
  }
  float d = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) d = fmaf(u[i], s[i], d);
  d = eg_block_sum(d, red);
  if (threadIdx.x == 0) *sigma = d;
}

__global__ void __launch_bounds__(256) sn_scale_kernel(const float* __restrict__ W, const float* __restrict__ sigma,
                                                       float* __restrict__ out, int64_t n) {
  const float sg = *sigma;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    for (int64_t i = tid; i < (n >> 2); i += nth) {
      float4 a = reinterpret_cast<const float4*>(W)[i];
      a.x /= sg; a.y /= sg; a.z /= sg; a.w /= sg;
      reinterpret_cast<float4*>(out)[i] = a;
    }
  } else {
    for (int64_t i = tid; i < n; i += nth) out[i] = W[i] / sg;
  }
}

// part[block] = <a, b> over the block's grid-stride slice (no atomics: the total is summed in a fixed order)
__global__ void __launch_bounds__(256) sn_dot_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     int64_t n, float* __restrict__ part) {
  __shared__ float red[32];
  float s = 0.f;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)gridDim.x * blockDim.x) {
      const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
      s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      s = fmaf(a[i], b[i], s);
  }
  s = eg_block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// dW_orig[i,j] = dW[i,j]/sigma - (dot/sigma^2) * u[i] * v[j];  dot = sum of nparts partials (every block
// re-sums them in the same order); one block row-slice at a time so u[r] is a scalar per row
__global__ void __launch_bounds__(256) sn_bwd_kernel(const float* __restrict__ dW, const float* __restrict__ u,
                                                     const float* __restrict__ v, const float* __restrict__ sigma,
                                                     const float* __restrict__ part, int nparts, int rows, int cols,
                                                     float* __restrict__ out) {
  __shared__ float red[32];
  float d = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) d += part[i];
  d = eg_block_sum(d, red);
  const float sg = *sigma;
  const float coef = d / (sg * sg);
  const int64_t n = (int64_t)rows * cols;
  if ((cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(dW) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
    const int cv = cols >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / cv), c4 = (int)(i - (int64_t)r * cv);
      const float4 g = reinterpret_cast<const float4*>(dW)[i], vv = reinterpret_cast<const float4*>(v)[c4];
      const float cu = coef * u[r];
      float4 o;
      o.x = g.x / sg - cu * vv.x; o.y = g.y / sg - cu * vv.y; o.z = g.z / sg - cu * vv.z; o.w = g.w / sg - cu * vv.w;
      reinterpret_cast<float4*>(out)[i] = o;
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
      out[i] = dW[i] / sg - coef * u[r] * v[c];
    }
  }
}

// ---- very small matrices (<= 16 K weights: the first convs and the Linear heads of the dSprites / MNIST nets): the
// whole forward in ONE block instead of five launches.  Same arithmetic: W^T u (thread per column), normalise, W v
// (warp per row), normalise, sigma.  Measured (r02v): with the threshold at 128 K weights the dSprites step got SLOWER
// (5.68 -> 5.93 ms): one block walks a 64 x 1024 matrix three times with little memory parallelism, while the five
// launches of the multi-kernel path use the whole chip and sit on a side stream anyway.
constexpr int SN_SMALL_MAX = 16 * 1024;     // weights
constexpr int SN_SMALL_COLS = 4096, SN_SMALL_ROWS = 1024;
__global__ void __launch_bounds__(1024) sn_small_fwd_kernel(const float* __restrict__ W, float* __restrict__ u,
                                                            float* __restrict__ v, int rows, int cols, int do_pi, float eps,
                                                            float* __restrict__ sigma) {
  __shared__ float red[32];
  __shared__ float vs[SN_SMALL_COLS];
  __shared__ float ss[SN_SMALL_ROWS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (do_pi) {
    for (int i = tid; i < rows; i += 1024) ss[i] = u[i];
    __syncthreads();
    float q = 0.f;
    for (int j = tid; j < cols; j += 1024) {
      float acc = 0.f;
#pragma unroll 8
      for (int i = 0; i < rows; ++i) acc = fmaf(W[(int64_t)i * cols + j], ss[i], acc);
      vs[j] = acc;
      q = fmaf(acc, acc, q);
    }
    q = eg_block_sum(q, red);                       // (every thread gets the total)
    const float inv = 1.f / fmaxf(sqrtf(q), eps);
    for (int j = tid; j < cols; j += 1024) { const float x = vs[j] * inv; vs[j] = x; v[j] = x; }
  } else {
    for (int j = tid; j < cols; j += 1024) vs[j] = v[j];
  }
  __syncthreads();
  for (int i = warp; i < rows; i += 32) {
    const float* wr = W + (int64_t)i * cols;
    float acc = 0.f;
#pragma unroll 8
    for (int j = lane; j < cols; j += 32) acc = fmaf(wr[j], vs[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) ss[i] = acc;
  }
  __syncthreads();
  if (do_pi) {
    float q = 0.f;
    for (int i = tid; i < rows; i += 1024) q = fmaf(ss[i], ss[i], q);
    q = eg_block_sum(q, red);
    const float inv = 1.f / fmaxf(sqrtf(q), eps);
    float d = 0.f;
    for (int i = tid; i < rows; i += 1024) { const float x = ss[i] * inv; u[i] = x; d = fmaf(x, ss[i], d); }
    d = eg_block_sum(d, red);
    if (tid == 0) *sigma = d;
  } else {
    float d = 0.f;
    for (int i = tid; i < rows; i += 1024) d = fmaf(u[i], ss[i], d);
    d = eg_block_sum(d, red);
    if (tid == 0) *sigma = d;
  }
}

// backward of a small layer in one block: dot = <dW, W>, then dW_orig = dW / sigma - dot / sigma^2 * u v^T
__global__ void __launch_bounds__(1024) sn_small_bwd_kernel(const float* __restrict__ dW, const float* __restrict__ W,
                                                            const float* __restrict__ u, const float* __restrict__ v,
                                                            const float* __restrict__ sigma, int rows, int cols,
                                                            float* __restrict__ out) {
  __shared__ float red[32];
  const int n = rows * cols;
  float d = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) d = fmaf(dW[i], W[i], d);
  d = eg_block_sum(d, red);
  const float sg = *sigma;
  const float coef = d / (sg * sg);
  for (int i = threadIdx.x; i < n; i += 1024) {
    const int r = i / cols, c = i - r * cols;
    out[i] = dW[i] / sg - coef * u[r] * v[c];
  }
}

bool sn_small(int rows, int cols) {
  return (int64_t)rows * cols <= SN_SMALL_MAX && cols <= SN_SMALL_COLS && rows <= SN_SMALL_ROWS;
}

int grid_for(int64_t n) {
  int64_t b = (n + 1023) / 1024;
  const int64_t cap = 16 * (int64_t)eg_sm_count();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int eadgan_spectral_norm_fwd(const float* w_orig, int rows, int cols, float* u, float* v,
                                        int do_power_iter, float eps, float* sigma, float* w_sn,
                                        float* scratch, void* stream) {
  EG_REQUIRE(w_orig && u && v && sigma && scratch && rows > 0 && cols > 0, EADGAN_ERR_INVALID,
             "spectral_norm_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (sn_small(rows, cols)) {
    sn_small_fwd_kernel<<<1, 1024, 0, st>>>(w_orig, u, v, rows, cols, do_power_iter ? 1 : 0, eps, sigma);
    EG_LAUNCH_CHECK("sn_small_fwd_kernel");
    if (w_sn) {
      const int64_t n = (int64_t)rows * cols;
      sn_scale_kernel<<<grid_for(n), 256, 0, st>>>(w_orig, sigma, w_sn, n);
      EG_LAUNCH_CHECK("sn_scale_kernel");
    }
    return 0;
  }
  int col_tiles = (cols + 127) / 128;
  int splits = (4 * eg_sm_count() + col_tiles - 1) / col_tiles;
  int max_splits = (rows + 15) / 16;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int rpb = (rows + splits - 1) / splits;
  splits = (rows + rpb - 1) / rpb;
  float* s = scratch;                    // [rows]
  float* nrm = scratch + ((rows + 3) & ~3);   // [ceil(cols/256)] partials of ||W^T u||^2
  const int nparts = (cols + 255) / 256;
  float* t = nrm + ((nparts + 3) & ~3);  // [splits][cols] partials of W^T u (16-byte aligned with `scratch`)
  if (do_power_iter) {
    sn_wtu_kernel<<<dim3(col_tiles, splits), 128, 0, st>>>(w_orig, u, t, rows, cols, rpb);
    EG_LAUNCH_CHECK("sn_wtu_kernel");
    sn_vsum_kernel<<<nparts, 256, 0, st>>>(t, cols, splits, nrm);
    EG_LAUNCH_CHECK("sn_vsum_kernel");
  }
  sn_wv_kernel<<<rows, 256, 0, st>>>(w_orig, do_power_iter ? t : nullptr, nrm, nparts, eps, v, s, rows, cols);
  EG_LAUNCH_CHECK("sn_wv_kernel");
  sn_sigma_kernel<<<1, 1024, 0, st>>>(s, u, rows, eps, do_power_iter ? 1 : 0, sigma);
  EG_LAUNCH_CHECK("sn_sigma_kernel");
  if (w_sn) {
    const int64_t n = (int64_t)rows * cols;
    sn_scale_kernel<<<grid_for(n), 256, 0, st>>>(w_orig, sigma, w_sn, n);
    EG_LAUNCH_CHECK("sn_scale_kernel");
  }
  return 0;
}

extern "C" int eadgan_spectral_norm_bwd(const float* dw_sn, const float* w_orig, const float* u,
                                        const float* v, const float* sigma, int rows, int cols,
                                        float* dw_orig, float* scratch, void* stream) {
  EG_REQUIRE(dw_sn && w_orig && u && v && sigma && dw_orig && scratch && rows > 0 && cols > 0,
             EADGAN_ERR_INVALID, "spectral_norm_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)rows * cols;
  if (sn_small(rows, cols)) {
    sn_small_bwd_kernel<<<1, 1024, 0, st>>>(dw_sn, w_orig, u, v, sigma, rows, cols, dw_orig);
    EG_LAUNCH_CHECK("sn_small_bwd_kernel");
    return 0;
  }
  const int parts = grid_for(n);
  sn_dot_kernel<<<parts, 256, 0, st>>>(dw_sn, w_orig, n, scratch);
  EG_LAUNCH_CHECK("sn_dot_kernel");
  sn_bwd_kernel<<<grid_for(n), 256, 0, st>>>(dw_sn, u, v, sigma, scratch, parts, rows, cols, dw_orig);
  EG_LAUNCH_CHECK("sn_bwd_kernel");
  return 0;
}

/* floats of scratch the two entry points need (per-split / per-block partial sums live there) */
extern "C" size_t eadgan_spectral_norm_scratch_floats(int rows, int cols, int backward) {
  if (backward) return (size_t)grid_for((int64_t)rows * cols) + 8;
  return (size_t)((rows + 3) & ~3) + (size_t)(((cols + 255) / 256 + 3) & ~3) + (size_t)((rows + 15) / 16 + 1) * (size_t)cols + 8;
}

/* W_sn = W / sigma on its own: the forward entry point skips it when w_sn == NULL (the tcgen05 chain applies
 * 1/sigma in the conv epilogue); a consumer that does need the normalised weight materialises it with this. */
extern "C" int eadgan_spectral_norm_scale(const float* w_orig, const float* sigma, float* w_sn, long long n,
                                          void* stream) {
  EG_REQUIRE(w_orig && sigma && w_sn && n > 0, EADGAN_ERR_INVALID, "spectral_norm_scale: bad arguments");
  sn_scale_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(w_orig, sigma, w_sn, n);
  EG_LAUNCH_CHECK("sn_scale_kernel");
  return 0;
}

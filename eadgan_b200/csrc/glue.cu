// Affine-code glue of the training step as fused device kernels:
//   * the spatial transformer  F.grid_sample(x, F.affine_grid(theta, x.size()), padding_mode)  with
//     align_corners = False (transformation_2D.stn: celebA/EAD-GAN_celebA.py:149-153, dSprites/rp.py:200-211,
//     MNIST/EAD-GAN_rpqmnxy.py:180-188; 'border' everywhere except colored_dSprites/pxy_color.py:90 'zeros')
//     -- one kernel instead of a base-grid construction, a batched fp32 GEMM and the sampler;
//   * the closed-form recovery of relative affine parameters, "affine_regularzier"
//     (celebA/utils_rpqxy.py:82-116, dSprites/utils_rp.py:117-147) and the matrix rows fed to MNIST's
//     approximator (MNIST/utils_rpqmnxy.py:117-129).  The reference builds 3x3 matrices on the HOST
//     (8-9 synchronising copies per call); eadgan_b200/affine.py restates the algebra as ~60 tiny torch
//     ops (+ ~150 in backward).  Here ONE forward kernel evaluates the algebra in forward-mode dual numbers
//     (value + partials w.r.t. the 2 x K input codes) and stores the Jacobian; the backward kernel is J^T g.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// spatial transformer, forward only (no consumed gradient of the training steps flows through it)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stn_fwd_kernel(const float* __restrict__ img, const float* __restrict__ theta,
                                                      int n, int c, int h, int w, int border,
                                                      float* __restrict__ out) {
  const int64_t total = (int64_t)n * h * w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int b = (int)(i / ((int64_t)w * h));
    const float* t = theta + (int64_t)b * 6;
    // affine_grid, align_corners = False: pixel centres (2 i + 1) / size - 1
    const float xn = (2.f * x + 1.f) / w - 1.f, yn = (2.f * y + 1.f) / h - 1.f;
    const float gx = fmaf(t[0], xn, fmaf(t[1], yn, t[2]));
    const float gy = fmaf(t[3], xn, fmaf(t[4], yn, t[5]));
    // grid_sample un-normalisation, align_corners = False
    float ix = ((gx + 1.f) * w - 1.f) * 0.5f, iy = ((gy + 1.f) * h - 1.f) * 0.5f;
    if (border) {
      ix = fminf(fmaxf(ix, 0.f), (float)(w - 1));
      iy = fminf(fmaxf(iy, 0.f), (float)(h - 1));
    }
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float tx = ix - fx, ty = iy - fy;
    const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
    const bool vx0 = x0 >= 0 && x0 < w, vx1 = x1 >= 0 && x1 < w, vy0 = y0 >= 0 && y0 < h, vy1 = y1 >= 0 && y1 < h;
    for (int ch = 0; ch < c; ++ch) {
      const float* p = img + ((int64_t)b * c + ch) * h * w;
      float v = 0.f;
      if (vy0 && vx0) v += p[(int64_t)y0 * w + x0] * w00;
      if (vy0 && vx1) v += p[(int64_t)y0 * w + x1] * w01;
      if (vy1 && vx0) v += p[(int64_t)y1 * w + x0] * w10;
      if (vy1 && vx1) v += p[(int64_t)y1 * w + x1] * w11;
      out[((int64_t)b * c + ch) * h * w + (int64_t)y * w + x] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// forward-mode dual numbers with N partials
// ---------------------------------------------------------------------------------------------------
template <int N>
struct Dual {
  float v;
  float d[N];
};
template <int N> __device__ __forceinline__ Dual<N> dconst(float c) {
  Dual<N> r; r.v = c;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = 0.f;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> dvar(float c, int idx) {
  Dual<N> r = dconst<N>(c);
  r.d[idx] = 1.f;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; const float inv = 1.f / b.v; r.v = a.v * inv;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> affine1(const Dual<N>& a, float mul, float add) {  // a*mul + add
  Dual<N> r; r.v = a.v * mul + add;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * mul;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> dsin(const Dual<N>& a) {
  Dual<N> r; r.v = sinf(a.v); const float c = cosf(a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> dcos(const Dual<N>& a) {
  Dual<N> r; r.v = cosf(a.v); const float s = -sinf(a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> datan(const Dual<N>& a) {
  Dual<N> r; r.v = atanf(a.v); const float g = 1.f / (1.f + a.v * a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = g * a.d[i];
  return r;
}

constexpr float PI_F = 3.14159265358979323846f;

// [[a, b, tx], [c, d, ty], [0, 0, 1]] = R(theta) diag(p, q, 1) Skew(m, n) T(x, y)
template <int N>
struct Aff { Dual<N> a, b, c, d, tx, ty; };

template <int N>
__device__ __forceinline__ Aff<N> compose(const Dual<N>& theta, const Dual<N>& p, const Dual<N>& q, const Dual<N>* m,
                                          const Dual<N>* nn, const Dual<N>& x, const Dual<N>& y) {
  const Dual<N> cs = dcos(theta), sn = dsin(theta);
  Aff<N> A;
  if (m == nullptr) {
    A.a = cs * p; A.b = -(sn * q); A.c = sn * p; A.d = cs * q;
  } else {  // (R Z) [[1, m], [n, 1]]
    const Dual<N> cp = cs * p, sq = sn * q, sp = sn * p, cq = cs * q;
    A.a = cp - sq * (*nn); A.b = cp * (*m) - sq; A.c = sp + cq * (*nn); A.d = sp * (*m) + cq;
  }
  A.tx = A.a * x + A.b * y;
  A.ty = A.c * x + A.d * y;
  return A;
}
// rel = M2 inverse(M1), both affine: closed-form inverse (no pivoting, no host check)
template <int N>
__device__ __forceinline__ Aff<N> relative(const Aff<N>& A1, const Aff<N>& A2) {
  const Dual<N> det = A1.a * A1.d - A1.b * A1.c;
  const Dual<N> ia = A1.d / det, ib = -(A1.b / det), ic = -(A1.c / det), id = A1.a / det;
  const Dual<N> itx = -(ia * A1.tx + ib * A1.ty), ity = -(ic * A1.tx + id * A1.ty);
  Aff<N> R;
  R.a = A2.a * ia + A2.b * ic; R.b = A2.a * ib + A2.b * id;
  R.c = A2.c * ia + A2.d * ic; R.d = A2.c * ib + A2.d * id;
  R.tx = A2.a * itx + A2.b * ity + A2.tx; R.ty = A2.c * itx + A2.d * ity + A2.ty;
  return R;
}

enum { MODE_CELEBA = 0, MODE_DSPRITES = 1, MODE_MNIST = 2 };
template <int MODE> struct ModeCfg;
template <> struct ModeCfg<MODE_CELEBA> { static constexpr int K = 5, OUT = 5; };     // theta p q x y
template <> struct ModeCfg<MODE_DSPRITES> { static constexpr int K = 4, OUT = 4; };   // theta p x y
template <> struct ModeCfg<MODE_MNIST> { static constexpr int K = 7, OUT = 6; };      // theta p q m n x y -> 2x3 rows

template <int MODE>
__global__ void __launch_bounds__(128) relcode_fwd_kernel(const float* __restrict__ real, int64_t real_stride,
                                                          const float* __restrict__ trans, int64_t trans_stride, int n,
                                                          float* __restrict__ out, float* __restrict__ jac) {
  constexpr int K = ModeCfg<MODE>::K, OUT = ModeCfg<MODE>::OUT, N = 2 * K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Dual<N> c1[K], c2[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    c1[k] = dvar<N>(real[(int64_t)i * real_stride + k], k);
    c2[k] = dvar<N>(trans[(int64_t)i * trans_stride + k], K + k);
  }
  Dual<N> o[OUT];
  if (MODE == MODE_CELEBA) {
    const Aff<N> A1 = compose<N>(affine1(c1[0], PI_F / 9.f, 0.f), affine1(c1[1], 0.2f, 1.f), affine1(c1[2], 0.2f, 1.f),
                                 nullptr, nullptr, affine1(c1[3], 0.1f, 0.f), affine1(c1[4], 0.1f, 0.f));
    const Aff<N> A2 = compose<N>(affine1(c2[0], PI_F / 9.f, 0.f), affine1(c2[1], 0.2f, 1.f), affine1(c2[2], 0.2f, 1.f),
                                 nullptr, nullptr, affine1(c2[3], 0.1f, 0.f), affine1(c2[4], 0.1f, 0.f));
    const Aff<N> R = relative(A1, A2);
    // celebA/utils_rpqxy.py:95-113
    const Dual<N> num = affine1(R.a * R.c - R.b * R.d, 2.f, 0.f);
    const Dual<N> den = R.a * R.a + R.d * R.d - R.b * R.b - R.c * R.c;
    const Dual<N> th = affine1(datan(num / den), 0.5f, 0.f);
    const Dual<N> ct = dcos(th), st = dsin(th);
    const Dual<N> p = R.a * ct + R.c * st;
    const Dual<N> q = R.d * ct - R.b * st;
    const Dual<N> x = (R.tx * ct + R.ty * st) / p;
    const Dual<N> y = (R.ty * ct - R.tx * st) / q;
    o[0] = affine1(th, 9.f / PI_F, 0.f); o[1] = affine1(p, 5.f, -5.f); o[2] = affine1(q, 5.f, -5.f);
    o[3] = affine1(x, 10.f, 0.f); o[4] = affine1(y, 10.f, 0.f);
  } else if (MODE == MODE_DSPRITES) {
    const Dual<N> p1 = affine1(c1[1], 0.2f, 1.f), p2 = affine1(c2[1], 0.2f, 1.f);
    const Aff<N> A1 = compose<N>(affine1(c1[0], PI_F / 9.f, 0.f), p1, p1, nullptr, nullptr, affine1(c1[2], 0.1f, 0.f),
                                 affine1(c1[3], 0.1f, 0.f));
    const Aff<N> A2 = compose<N>(affine1(c2[0], PI_F / 9.f, 0.f), p2, p2, nullptr, nullptr, affine1(c2[2], 0.1f, 0.f),
                                 affine1(c2[3], 0.1f, 0.f));
    const Aff<N> R = relative(A1, A2);
    // dSprites/utils_rp.py:129-140
    const Dual<N> dm = R.c - R.b, sm = R.a + R.d;
    const Dual<N> th = datan(dm / sm);
    const Dual<N> ct = dcos(th), st = dsin(th);
    const Dual<N> p = affine1(ct * sm + st * dm, 0.5f, 0.f);
    const Dual<N> x = (R.tx * ct + R.ty * st) / p;
    const Dual<N> y = (R.ty * ct - R.tx * st) / p;
    o[0] = affine1(th, 9.f / PI_F, 0.f); o[1] = affine1(p, 5.f, -5.f); o[2] = affine1(x, 10.f, 0.f);
    o[3] = affine1(y, 10.f, 0.f);
  } else {
    const Dual<N> m1 = affine1(c1[3], 0.2f, 0.f), n1 = affine1(c1[4], 0.2f, 0.f);
    const Dual<N> m2 = affine1(c2[3], 0.2f, 0.f), n2 = affine1(c2[4], 0.2f, 0.f);
    const Aff<N> A1 = compose<N>(affine1(c1[0], PI_F / 9.f, 0.f), affine1(c1[1], 0.2f, 1.f), affine1(c1[2], 0.2f, 1.f), &m1,
                                 &n1, affine1(c1[5], 0.1f, 0.f), affine1(c1[6], 0.1f, 0.f));
    const Aff<N> A2 = compose<N>(affine1(c2[0], PI_F / 9.f, 0.f), affine1(c2[1], 0.2f, 1.f), affine1(c2[2], 0.2f, 1.f), &m2,
                                 &n2, affine1(c2[5], 0.1f, 0.f), affine1(c2[6], 0.1f, 0.f));
    const Aff<N> R = relative(A1, A2);
    o[0] = R.a; o[1] = R.b; o[2] = R.tx; o[3] = R.c; o[4] = R.d; o[5] = R.ty;
  }
#pragma unroll
  for (int k = 0; k < OUT; ++k) {
    out[(int64_t)i * OUT + k] = o[k].v;
#pragma unroll
    for (int j = 0; j < N; ++j) jac[((int64_t)i * OUT + k) * N + j] = o[k].d[j];
  }
}

// d_real[i][k] = sum_o g[i][o] J[i][o][k];  d_trans[i][k] = sum_o g[i][o] J[i][o][K + k]
__global__ void relcode_bwd_kernel(const float* __restrict__ g, const float* __restrict__ jac, int n, int K, int OUT,
                                   float* __restrict__ d_real, float* __restrict__ d_trans) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * 2 * K) return;
  const int i = idx / (2 * K), j = idx - i * 2 * K;
  float s = 0.f;
  for (int o = 0; o < OUT; ++o) s += g[(int64_t)i * OUT + o] * jac[((int64_t)i * OUT + o) * 2 * K + j];
  if (j < K) d_real[(int64_t)i * K + j] = s; else d_trans[(int64_t)i * K + (j - K)] = s;
}

}  // namespace

extern "C" int eadgan_stn_fwd(const float* img, const float* theta, int n, int c, int h, int w, int padding_border,
                              float* out, void* stream) {
  EG_REQUIRE(img && theta && out && n > 0 && c > 0 && h > 0 && w > 0, EADGAN_ERR_INVALID, "stn_fwd: bad arguments");
  const int64_t total = (int64_t)n * h * w;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 32 * eg_sm_count()) blocks = 32 * eg_sm_count();
  stn_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(img, theta, n, c, h, w, padding_border, out);
  EG_LAUNCH_CHECK("stn_fwd_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// F.affine_grid / F.grid_sample as separate differentiable operators (align_corners = False, bilinear,
// padding 'border' | 'zeros'): what the UNMODIFIED reference scripts call (transformation_2D.stn:
// celebA/EAD-GAN_celebA.py:149-153; dSprites/rp.py:200-211; colored_dSprites/pxy_color.py:82-96 'zeros').
// In rp.py their BACKWARD is on the executed path: the frozen Encoder_pxy is grad-tracked, so autograd runs
// grid_sample's gradient w.r.t. the sampled image and w.r.t. the grid, and affine_grid's w.r.t. theta
// (dSprites/rp.py:374-377,399-400).  Formulas follow ATen's grid_sampler_2d (bilinear) including the
// clip_coordinates_set_grad rule of 'border' (zero coordinate gradient where the coordinate was clamped).
// ---------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) affine_grid_fwd_kernel(const float* __restrict__ theta, int n, int h, int w,
                                                              float* __restrict__ grid) {
  const int64_t total = (int64_t)n * h * w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const float* t = theta + (i / ((int64_t)w * h)) * 6;
    const float xn = (2.f * x + 1.f) / w - 1.f, yn = (2.f * y + 1.f) / h - 1.f;
    reinterpret_cast<float2*>(grid)[i] = make_float2(fmaf(t[0], xn, fmaf(t[1], yn, t[2])), fmaf(t[3], xn, fmaf(t[4], yn, t[5])));
  }
}

// d_theta[b][i][j] = sum over pixels of d_grid[b][y][x][i] * base_j,  base = (xn, yn, 1).  One block per image,
// fixed-order tree reduction: deterministic.
__global__ void __launch_bounds__(256) affine_grid_bwd_kernel(const float* __restrict__ dgrid, int h, int w,
                                                              float* __restrict__ dtheta) {
  __shared__ float red[6][256];
  const int b = blockIdx.x;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float2* g = reinterpret_cast<const float2*>(dgrid) + (int64_t)b * h * w;
  for (int i = threadIdx.x; i < h * w; i += 256) {
    const int x = i % w, y = i / w;
    const float xn = (2.f * x + 1.f) / w - 1.f, yn = (2.f * y + 1.f) / h - 1.f;
    const float2 d = g[i];
    acc[0] += d.x * xn; acc[1] += d.x * yn; acc[2] += d.x;
    acc[3] += d.y * xn; acc[4] += d.y * yn; acc[5] += d.y;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) red[j][threadIdx.x] = acc[j];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int j = 0; j < 6; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x < 6) dtheta[b * 6 + threadIdx.x] = red[threadIdx.x][0];
}

struct Sample {            // bilinear footprint of one output pixel
  int x0, y0;              // north-west corner (may be out of range)
  float tx, ty;            // fractional offsets
  float mx, my;            // d(source coordinate) / d(grid coordinate): size / 2, or 0 where 'border' clamped
};
__device__ __forceinline__ Sample locate(float gx, float gy, int h, int w, int border) {
  Sample s;
  float ix = ((gx + 1.f) * w - 1.f) * 0.5f, iy = ((gy + 1.f) * h - 1.f) * 0.5f;
  s.mx = 0.5f * w; s.my = 0.5f * h;
  if (border) {            // clip_coordinates_set_grad
    if (ix <= 0.f) { ix = 0.f; s.mx = 0.f; } else if (ix >= (float)(w - 1)) { ix = (float)(w - 1); s.mx = 0.f; }
    if (iy <= 0.f) { iy = 0.f; s.my = 0.f; } else if (iy >= (float)(h - 1)) { iy = (float)(h - 1); s.my = 0.f; }
  }
  const float fx = floorf(ix), fy = floorf(iy);
  s.x0 = (int)fx; s.y0 = (int)fy; s.tx = ix - fx; s.ty = iy - fy;
  return s;
}

__global__ void __launch_bounds__(256) grid_sample_fwd_kernel(const float* __restrict__ img, const float* __restrict__ grid,
                                                              int n, int c, int h, int w, int oh, int ow, int border,
                                                              float* __restrict__ out) {
  const int64_t total = (int64_t)n * oh * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / ((int64_t)oh * ow));
    const float2 g = reinterpret_cast<const float2*>(grid)[i];
    const Sample s = locate(g.x, g.y, h, w, border);
    const int x1 = s.x0 + 1, y1 = s.y0 + 1;
    const float w00 = (1.f - s.tx) * (1.f - s.ty), w01 = s.tx * (1.f - s.ty), w10 = (1.f - s.tx) * s.ty, w11 = s.tx * s.ty;
    const bool vx0 = s.x0 >= 0 && s.x0 < w, vx1 = x1 >= 0 && x1 < w, vy0 = s.y0 >= 0 && s.y0 < h, vy1 = y1 >= 0 && y1 < h;
    const int64_t pix = i - (int64_t)b * oh * ow;
    for (int ch = 0; ch < c; ++ch) {
      const float* p = img + ((int64_t)b * c + ch) * h * w;
      float v = 0.f;
      if (vy0 && vx0) v += p[(int64_t)s.y0 * w + s.x0] * w00;
      if (vy0 && vx1) v += p[(int64_t)s.y0 * w + x1] * w01;
      if (vy1 && vx0) v += p[(int64_t)y1 * w + s.x0] * w10;
      if (vy1 && vx1) v += p[(int64_t)y1 * w + x1] * w11;
      out[((int64_t)b * c + ch) * oh * ow + pix] = v;
    }
  }
}

// d_img (scatter-add over the 4 corners: fp32 atomics, as ATen's own CUDA backward; pre-zeroed by the caller) and
// d_grid (one thread per output pixel, no atomics).  Either output may be NULL.
__global__ void __launch_bounds__(256) grid_sample_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ img,
                                                              const float* __restrict__ grid, int n, int c, int h, int w,
                                                              int oh, int ow, int border, float* __restrict__ dimg,
                                                              float* __restrict__ dgrid) {
  const int64_t total = (int64_t)n * oh * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / ((int64_t)oh * ow));
    const float2 g = reinterpret_cast<const float2*>(grid)[i];
    const Sample s = locate(g.x, g.y, h, w, border);
    const int x1 = s.x0 + 1, y1 = s.y0 + 1;
    const float w00 = (1.f - s.tx) * (1.f - s.ty), w01 = s.tx * (1.f - s.ty), w10 = (1.f - s.tx) * s.ty, w11 = s.tx * s.ty;
    const bool vx0 = s.x0 >= 0 && s.x0 < w, vx1 = x1 >= 0 && x1 < w, vy0 = s.y0 >= 0 && s.y0 < h, vy1 = y1 >= 0 && y1 < h;
    const int64_t pix = i - (int64_t)b * oh * ow;
    float gix = 0.f, giy = 0.f;
    for (int ch = 0; ch < c; ++ch) {
      const float go = gout[((int64_t)b * c + ch) * oh * ow + pix];
      const int64_t base = ((int64_t)b * c + ch) * h * w;
      if (dimg) {
        if (vy0 && vx0) atomicAdd(dimg + base + (int64_t)s.y0 * w + s.x0, go * w00);
        if (vy0 && vx1) atomicAdd(dimg + base + (int64_t)s.y0 * w + x1, go * w01);
        if (vy1 && vx0) atomicAdd(dimg + base + (int64_t)y1 * w + s.x0, go * w10);
        if (vy1 && vx1) atomicAdd(dimg + base + (int64_t)y1 * w + x1, go * w11);
      }
      if (dgrid) {
        const float* p = img + base;
        const float v00 = (vy0 && vx0) ? p[(int64_t)s.y0 * w + s.x0] : 0.f, v01 = (vy0 && vx1) ? p[(int64_t)s.y0 * w + x1] : 0.f;
        const float v10 = (vy1 && vx0) ? p[(int64_t)y1 * w + s.x0] : 0.f, v11 = (vy1 && vx1) ? p[(int64_t)y1 * w + x1] : 0.f;
        gix += go * ((v01 - v00) * (1.f - s.ty) + (v11 - v10) * s.ty);
        giy += go * ((v10 - v00) * (1.f - s.tx) + (v11 - v01) * s.tx);
      }
    }
    if (dgrid) reinterpret_cast<float2*>(dgrid)[i] = make_float2(gix * s.mx, giy * s.my);
  }
}

int blocks_for(int64_t total) {
  int blocks = (int)((total + 255) / 256);
  if (blocks > 32 * eg_sm_count()) blocks = 32 * eg_sm_count();
  return blocks < 1 ? 1 : blocks;
}

}  // namespace

extern "C" int eadgan_affine_grid_fwd(const float* theta, int n, int h, int w, float* grid, void* stream) {
  EG_REQUIRE(theta && grid && n > 0 && h > 0 && w > 0, EADGAN_ERR_INVALID, "affine_grid_fwd: bad arguments");
  affine_grid_fwd_kernel<<<blocks_for((int64_t)n * h * w), 256, 0, (cudaStream_t)stream>>>(theta, n, h, w, grid);
  EG_LAUNCH_CHECK("affine_grid_fwd_kernel");
  return 0;
}

extern "C" int eadgan_affine_grid_bwd(const float* dgrid, int n, int h, int w, float* dtheta, void* stream) {
  EG_REQUIRE(dgrid && dtheta && n > 0 && h > 0 && w > 0, EADGAN_ERR_INVALID, "affine_grid_bwd: bad arguments");
  affine_grid_bwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(dgrid, h, w, dtheta);
  EG_LAUNCH_CHECK("affine_grid_bwd_kernel");
  return 0;
}

extern "C" int eadgan_grid_sample_fwd(const float* img, const float* grid, int n, int c, int h, int w, int oh, int ow,
                                      int padding_border, float* out, void* stream) {
  EG_REQUIRE(img && grid && out && n > 0 && c > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, EADGAN_ERR_INVALID,
             "grid_sample_fwd: bad arguments");
  grid_sample_fwd_kernel<<<blocks_for((int64_t)n * oh * ow), 256, 0, (cudaStream_t)stream>>>(img, grid, n, c, h, w, oh, ow,
                                                                                          padding_border, out);
  EG_LAUNCH_CHECK("grid_sample_fwd_kernel");
  return 0;
}

extern "C" int eadgan_grid_sample_bwd(const float* gout, const float* img, const float* grid, int n, int c, int h, int w,
                                      int oh, int ow, int padding_border, float* dimg_zeroed, float* dgrid, void* stream) {
  EG_REQUIRE(gout && img && grid && (dimg_zeroed || dgrid) && n > 0 && c > 0 && h > 0 && w > 0 && oh > 0 && ow > 0,
             EADGAN_ERR_INVALID, "grid_sample_bwd: bad arguments");
  grid_sample_bwd_kernel<<<blocks_for((int64_t)n * oh * ow), 256, 0, (cudaStream_t)stream>>>(
      gout, img, grid, n, c, h, w, oh, ow, padding_border, dimg_zeroed, dgrid);
  EG_LAUNCH_CHECK("grid_sample_bwd_kernel");
  return 0;
}

extern "C" int eadgan_relcode_dims(int mode, int* k_in, int* k_out) {
  EG_REQUIRE(mode >= 0 && mode <= 2 && k_in && k_out, EADGAN_ERR_INVALID, "relcode_dims: bad mode %d", mode);
  *k_in = mode == 0 ? 5 : (mode == 1 ? 4 : 7);
  *k_out = mode == 0 ? 5 : (mode == 1 ? 4 : 6);
  return 0;
}

extern "C" int eadgan_relcode_fwd(int mode, const float* real, long long real_stride, const float* trans,
                                  long long trans_stride, int n, float* out, float* jac, void* stream) {
  EG_REQUIRE(real && trans && out && jac && n > 0, EADGAN_ERR_INVALID, "relcode_fwd: bad arguments");
  const int blocks = (n + 127) / 128;
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case 0: relcode_fwd_kernel<MODE_CELEBA><<<blocks, 128, 0, st>>>(real, real_stride, trans, trans_stride, n, out, jac); break;
    case 1: relcode_fwd_kernel<MODE_DSPRITES><<<blocks, 128, 0, st>>>(real, real_stride, trans, trans_stride, n, out, jac); break;
    case 2: relcode_fwd_kernel<MODE_MNIST><<<blocks, 128, 0, st>>>(real, real_stride, trans, trans_stride, n, out, jac); break;
    default: return eadgan_set_error(EADGAN_ERR_INVALID, "relcode_fwd: bad mode %d", mode);
  }
  EG_LAUNCH_CHECK("relcode_fwd_kernel");
  return 0;
}

extern "C" int eadgan_relcode_bwd(int mode, const float* g, const float* jac, int n, float* d_real, float* d_trans,
                                  void* stream) {
  int K, OUT;
  if (int e = eadgan_relcode_dims(mode, &K, &OUT)) return e;
  EG_REQUIRE(g && jac && d_real && d_trans && n > 0, EADGAN_ERR_INVALID, "relcode_bwd: bad arguments");
  const int total = n * 2 * K;
  relcode_bwd_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, jac, n, K, OUT, d_real, d_trans);
  EG_LAUNCH_CHECK("relcode_bwd_kernel");
  return 0;
}

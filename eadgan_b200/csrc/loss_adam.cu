// Loss forward/backward kernels (warp-shuffle reductions, one or a few blocks: these
// tensors are [B] .. [B,10], latency-bound) and the fused multi-tensor Adam step
// (HBM-bound: 28 B per parameter -- read p,g,m,v, write p,m,v).
// Reference call sites: torch.nn.BCELoss / MSELoss / CrossEntropyLoss at
// celebA/EAD-GAN_celebA.py:161-164,342,356-363,383-395; mutual_info_loss at
// dSprites/rp.py:225-232; torch.optim.Adam at celebA/EAD-GAN_celebA.py:211-217
// (op order of torch/optim/adam.py::_single_tensor_adam).
#include "common.cuh"

namespace {

constexpr int LB = 256;      // loss block size
constexpr int LCHUNK = 8192;  // elements per loss block

int loss_blocks(int64_t n) {
  int64_t b = (n + LCHUNK - 1) / LCHUNK;
  if (b > 256) b = 256;
  return (int)(b < 1 ? 1 : b);
}

__device__ __forceinline__ void loss_commit(float part, float* loss, float* red) {
  part = eg_block_sum(part, red);
  if (threadIdx.x == 0) {
    if (gridDim.x == 1) *loss = part; else atomicAdd(loss, part);
  }
}

__global__ void __launch_bounds__(LB) bce_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                     int64_t n, float* loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)LB + threadIdx.x; i < n; i += (int64_t)gridDim.x * LB) {
    const float pi = p[i], ti = t[i];
    const float l1 = fmaxf(logf(pi), -100.f), l0 = fmaxf(log1pf(-pi), -100.f);
    s += -(ti * l1 + (1.f - ti) * l0);
  }
  loss_commit(s / (float)n, loss, red);
}

__global__ void __launch_bounds__(LB) bce_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                     const float* __restrict__ gout, int64_t n, float* __restrict__ dp) {
  const float g = *gout / (float)n;
  for (int64_t i = blockIdx.x * (int64_t)LB + threadIdx.x; i < n; i += (int64_t)gridDim.x * LB) {
    const float pi = p[i];
    dp[i] = g * (pi - t[i]) / fmaxf((1.f - pi) * pi, 1e-12f);
  }
}

__global__ void __launch_bounds__(LB) mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     int64_t n, float* loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)LB + threadIdx.x; i < n; i += (int64_t)gridDim.x * LB) {
    const float d = a[i] - b[i];
    s = fmaf(d, d, s);
  }
  loss_commit(s / (float)n, loss, red);
}

__global__ void __launch_bounds__(LB) mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     const float* __restrict__ gout, int64_t n,
                                                     float* __restrict__ da, float* __restrict__ db) {
  const float g = 2.f * *gout / (float)n;
  for (int64_t i = blockIdx.x * (int64_t)LB + threadIdx.x; i < n; i += (int64_t)gridDim.x * LB) {
    const float v = g * (a[i] - b[i]);
    if (da) da[i] = v;
    if (db) db[i] = -v;
  }
}

// thread per row (cols <= 64 in every reference use: 10 or 3 classes)
__global__ void __launch_bounds__(LB) ce_fwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels,
                                                    int rows, int cols, float* loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int r = blockIdx.x * LB + threadIdx.x; r < rows; r += gridDim.x * LB) {
    const float* xr = x + (int64_t)r * cols;
    float mx = -INFINITY;
    for (int j = 0; j < cols; ++j) mx = fmaxf(mx, xr[j]);
    float z = 0.f;
    for (int j = 0; j < cols; ++j) z += expf(xr[j] - mx);
    s += (mx + logf(z)) - xr[labels[r]];
  }
  loss_commit(s / (float)rows, loss, red);
}

__global__ void __launch_bounds__(LB) ce_bwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels,
                                                    const float* __restrict__ gout, int rows, int cols,
                                                    float* __restrict__ dx) {
  const float g = *gout / (float)rows;
  for (int r = blockIdx.x * LB + threadIdx.x; r < rows; r += gridDim.x * LB) {
    const float* xr = x + (int64_t)r * cols;
    float mx = -INFINITY;
    for (int j = 0; j < cols; ++j) mx = fmaxf(mx, xr[j]);
    float z = 0.f;
    for (int j = 0; j < cols; ++j) z += expf(xr[j] - mx);
    const float inv = 1.f / z;
    const int lab = (int)labels[r];
    for (int j = 0; j < cols; ++j)
      dx[(int64_t)r * cols + j] = g * (expf(xr[j] - mx) * inv - (j == lab ? 1.f : 0.f));
  }
}

// mean_rows(-sum_j log(q+eps) c) + mean_rows(-sum_j log(c+eps) c)
__global__ void __launch_bounds__(LB) mi_fwd_kernel(const float* __restrict__ q, const float* __restrict__ c,
                                                    int rows, int cols, float* loss) {
  __shared__ float red[32];
  const int64_t n = (int64_t)rows * cols;
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)LB + threadIdx.x; i < n; i += (int64_t)gridDim.x * LB) {
    const float ci = c[i];
    s -= (logf(q[i] + 1e-8f) + logf(ci + 1e-8f)) * ci;
  }
  loss_commit(s / (float)rows, loss, red);
}

__global__ void __launch_bounds__(LB) mi_bwd_kernel(const float* __restrict__ q, const float* __restrict__ c,
                                                    const float* __restrict__ gout, int rows, int cols,
                                                    float* __restrict__ dq) {
  const int64_t n = (int64_t)rows * cols;
  const float g = *gout / (float)rows;
  for (int64_t i = blockIdx.x * (int64_t)LB + threadIdx.x; i < n; i += (int64_t)gridDim.x * LB)
    dq[i] = -g * c[i] / (q[i] + 1e-8f);
}

// ---------------- fused multi-tensor Adam ------------------------------------------------
constexpr int ADAM_CHUNK = 16384;  // elements per block
struct AdamLaunch {
  eadgan_adam_tensors t;
  int blk_start[EADGAN_ADAM_MAX_TENSORS + 1];
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float w1, float beta2,
                                         float w2, float bc2_sqrt, float eps, float neg_step) {
  // m.lerp_(g, 1-beta1)  (ATen lerp: small-weight / large-weight branches)
  const float diff = __fsub_rn(g, m);
  m = (fabsf(w1) < 0.5f) ? __fadd_rn(m, __fmul_rn(w1, diff)) : __fsub_rn(g, __fmul_rn(diff, __fsub_rn(1.f, w1)));
  // v.mul_(beta2).addcmul_(g, g, value=1-beta2)
  v = __fadd_rn(__fmul_rn(v, beta2), __fmul_rn(__fmul_rn(w2, g), g));
  // denom = sqrt(v)/sqrt(bc2) + eps ; p.addcdiv_(m, denom, value=-step_size)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(neg_step, m), denom));
}

// step_dev != NULL: the step count lives on the device (CUDA-graph replays advance it with
// adam_advance_kernel); the bias corrections are then formed here, in double, exactly like the host path
__global__ void __launch_bounds__(256) adam_kernel(const AdamLaunch L, float w1, float beta2, float w2, float eps,
                                                   float neg_step, float bc2_sqrt, float gscale,
                                                   const int64_t* __restrict__ step_dev, double lr, double beta1d,
                                                   double beta2d) {
  if (step_dev) {
    const double t = (double)*step_dev;
    neg_step = (float)(-(lr / (1.0 - pow(beta1d, t))));
    bc2_sqrt = (float)sqrt(1.0 - pow(beta2d, t));
  }
  int ti = 0;
  while (ti + 1 < L.t.count && (int)blockIdx.x >= L.blk_start[ti + 1]) ++ti;
  const int64_t off = (int64_t)(blockIdx.x - L.blk_start[ti]) * ADAM_CHUNK;
  const int64_t n = L.t.numel[ti];
  const int64_t end = off + ADAM_CHUNK < n ? off + ADAM_CHUNK : n;
  float* __restrict__ p = L.t.p[ti];
  const float* __restrict__ g = L.t.g[ti];
  float* __restrict__ m = L.t.m[ti];
  float* __restrict__ v = L.t.v[ti];
  const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (al) {
    const int64_t end4 = off + ((end - off) & ~(int64_t)3);
    for (int64_t i = off + 4 * (int64_t)threadIdx.x; i < end4; i += 4 * 256) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      if (gscale != 1.f) { gv.x *= gscale; gv.y *= gscale; gv.z *= gscale; gv.w *= gscale; }
      adam_one(pv.x, gv.x, mv.x, vv.x, w1, beta2, w2, bc2_sqrt, eps, neg_step);
      adam_one(pv.y, gv.y, mv.y, vv.y, w1, beta2, w2, bc2_sqrt, eps, neg_step);
      adam_one(pv.z, gv.z, mv.z, vv.z, w1, beta2, w2, bc2_sqrt, eps, neg_step);
      adam_one(pv.w, gv.w, mv.w, vv.w, w1, beta2, w2, bc2_sqrt, eps, neg_step);
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    }
    for (int64_t i = end4 + threadIdx.x; i < end; i += 256) {
      float gi = g[i];
      if (gscale != 1.f) gi *= gscale;
      adam_one(p[i], gi, m[i], v[i], w1, beta2, w2, bc2_sqrt, eps, neg_step);
    }
  } else {
    for (int64_t i = off + threadIdx.x; i < end; i += 256) {
      float gi = g[i];
      if (gscale != 1.f) gi *= gscale;
      adam_one(p[i], gi, m[i], v[i], w1, beta2, w2, bc2_sqrt, eps, neg_step);
    }
  }
}

__global__ void adam_advance_kernel(int64_t* step_dev) { *step_dev += 1; }

}  // namespace

#define LOSS_PROLOGUE(name, n)                                                          \
  EG_REQUIRE((n) > 0, EADGAN_ERR_INVALID, name ": empty input");                        \
  cudaStream_t st = (cudaStream_t)stream;                                               \
  const int blocks = loss_blocks(n);                                                    \
  if (blocks > 1) EG_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));

extern "C" int eadgan_bce_fwd(const float* p, const float* target, int64_t numel, float* loss, void* stream) {
  EG_REQUIRE(p && target && loss, EADGAN_ERR_INVALID, "bce_fwd: NULL argument");
  LOSS_PROLOGUE("bce_fwd", numel)
  bce_fwd_kernel<<<blocks, LB, 0, st>>>(p, target, numel, loss);
  EG_LAUNCH_CHECK("bce_fwd_kernel");
  return 0;
}
extern "C" int eadgan_bce_bwd(const float* p, const float* target, const float* gout, int64_t numel, float* dp,
                              void* stream) {
  EG_REQUIRE(p && target && gout && dp && numel > 0, EADGAN_ERR_INVALID, "bce_bwd: bad arguments");
  bce_bwd_kernel<<<loss_blocks(numel), LB, 0, (cudaStream_t)stream>>>(p, target, gout, numel, dp);
  EG_LAUNCH_CHECK("bce_bwd_kernel");
  return 0;
}
extern "C" int eadgan_mse_fwd(const float* a, const float* b, int64_t numel, float* loss, void* stream) {
  EG_REQUIRE(a && b && loss, EADGAN_ERR_INVALID, "mse_fwd: NULL argument");
  LOSS_PROLOGUE("mse_fwd", numel)
  mse_fwd_kernel<<<blocks, LB, 0, st>>>(a, b, numel, loss);
  EG_LAUNCH_CHECK("mse_fwd_kernel");
  return 0;
}
extern "C" int eadgan_mse_bwd(const float* a, const float* b, const float* gout, int64_t numel, float* da,
                              float* db, void* stream) {
  EG_REQUIRE(a && b && gout && (da || db) && numel > 0, EADGAN_ERR_INVALID, "mse_bwd: bad arguments");
  mse_bwd_kernel<<<loss_blocks(numel), LB, 0, (cudaStream_t)stream>>>(a, b, gout, numel, da, db);
  EG_LAUNCH_CHECK("mse_bwd_kernel");
  return 0;
}
extern "C" int eadgan_ce_fwd(const float* x, const int64_t* labels, int rows, int cols, float* loss, void* stream) {
  EG_REQUIRE(x && labels && loss && cols > 0, EADGAN_ERR_INVALID, "ce_fwd: bad arguments");
  LOSS_PROLOGUE("ce_fwd", (int64_t)rows)
  ce_fwd_kernel<<<blocks, LB, 0, st>>>(x, labels, rows, cols, loss);
  EG_LAUNCH_CHECK("ce_fwd_kernel");
  return 0;
}
extern "C" int eadgan_ce_bwd(const float* x, const int64_t* labels, const float* gout, int rows, int cols,
                             float* dx, void* stream) {
  EG_REQUIRE(x && labels && gout && dx && rows > 0 && cols > 0, EADGAN_ERR_INVALID, "ce_bwd: bad arguments");
  ce_bwd_kernel<<<loss_blocks(rows), LB, 0, (cudaStream_t)stream>>>(x, labels, gout, rows, cols, dx);
  EG_LAUNCH_CHECK("ce_bwd_kernel");
  return 0;
}
extern "C" int eadgan_mi_fwd(const float* q, const float* c, int rows, int cols, float* loss, void* stream) {
  EG_REQUIRE(q && c && loss && cols > 0, EADGAN_ERR_INVALID, "mi_fwd: bad arguments");
  LOSS_PROLOGUE("mi_fwd", (int64_t)rows * cols)
  mi_fwd_kernel<<<blocks, LB, 0, st>>>(q, c, rows, cols, loss);
  EG_LAUNCH_CHECK("mi_fwd_kernel");
  return 0;
}
extern "C" int eadgan_mi_bwd(const float* q, const float* c, const float* gout, int rows, int cols, float* dq,
                             void* stream) {
  EG_REQUIRE(q && c && gout && dq && rows > 0 && cols > 0, EADGAN_ERR_INVALID, "mi_bwd: bad arguments");
  mi_bwd_kernel<<<loss_blocks((int64_t)rows * cols), LB, 0, (cudaStream_t)stream>>>(q, c, gout, rows, cols, dq);
  EG_LAUNCH_CHECK("mi_bwd_kernel");
  return 0;
}

namespace {
int adam_launch(const eadgan_adam_tensors* t, double beta1, double beta2, double eps, double step_size,
                double bc2_sqrt, float grad_scale, const int64_t* step_dev, double lr, void* stream) {
  EG_REQUIRE(t && t->count > 0 && t->count <= EADGAN_ADAM_MAX_TENSORS, EADGAN_ERR_INVALID,
             "adam_step: tensor count must be in [1, %d]", EADGAN_ADAM_MAX_TENSORS);
  AdamLaunch L;
  L.t = *t;
  int blocks = 0;
  for (int i = 0; i < t->count; ++i) {
    EG_REQUIRE(t->p[i] && t->g[i] && t->m[i] && t->v[i] && t->numel[i] > 0, EADGAN_ERR_INVALID,
               "adam_step: tensor %d has a NULL pointer or no elements", i);
    L.blk_start[i] = blocks;
    blocks += (int)((t->numel[i] + ADAM_CHUNK - 1) / ADAM_CHUNK);
  }
  L.blk_start[t->count] = blocks;
  // scalars are rounded to fp32 exactly where torch rounds its Python doubles: 1-beta1 and
  // 1-beta2 are formed in double first (1 - 0.999 != 1.f - 0.999f)
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(L, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                        (float)eps, (float)(-step_size), (float)bc2_sqrt, grad_scale,
                                                        step_dev, lr, beta1, beta2);
  EG_LAUNCH_CHECK("adam_kernel");
  return 0;
}
}  // namespace

extern "C" int eadgan_adam_step(const eadgan_adam_tensors* t, double beta1, double beta2, double eps,
                                double step_size, double bc2_sqrt, float grad_scale, void* stream) {
  return adam_launch(t, beta1, beta2, eps, step_size, bc2_sqrt, grad_scale, nullptr, 0.0, stream);
}

extern "C" int eadgan_adam_step_dev(const eadgan_adam_tensors* t, double beta1, double beta2, double eps, double lr,
                                    const int64_t* step_dev, float grad_scale, void* stream) {
  EG_REQUIRE(step_dev != nullptr, EADGAN_ERR_INVALID, "adam_step_dev: NULL step counter");
  return adam_launch(t, beta1, beta2, eps, 0.0, 1.0, grad_scale, step_dev, lr, stream);
}

extern "C" int eadgan_adam_advance(int64_t* step_dev, void* stream) {
  EG_REQUIRE(step_dev != nullptr, EADGAN_ERR_INVALID, "adam_advance: NULL step counter");
  adam_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  EG_LAUNCH_CHECK("adam_advance_kernel");
  return 0;
}

// Generic fp32-accumulate SIMT implicit-GEMM convolution: any kernel size, stride and
// padding, arbitrary tensor strides, fp32 or bf16 storage.  This is the fp32 parity
// path (max rel err <= 1e-5 vs torch fp32) and the path for every geometry the
// tcgen05 kernels do not cover (3x3 MNIST convs, Cin=1/3 first layers, N=1/3 last
// layers, Linear heads).  Three directions share one register-tiled mainloop:
//   fprop : y[n,k,p,q]  = sum_{c,r,s} x[n,c,p*st-pad+r,q*st-pad+s] * w[k,c,r,s]
//   dgrad : dx[n,c,h,w] = sum_{k,r,s} dy[n,k,(h+pad-r)/st,(w+pad-s)/st] * w[k,c,r,s]
//   wgrad : dw[k,c,r,s] = sum_{n,p,q} dy[n,k,p,q] * x[n,c,p*st-pad+r,q*st-pad+s]
// Reference call sites: every nn.Conv2d / nn.ConvTranspose2d / nn.Linear of
// celebA/EAD-GAN_celebA.py:75-122, dSprites/rp.py:66-183, MNIST/EAD-GAN_rpqmnxy.py:77-163.
#include "common.cuh"
#include <cooperative_groups.h>

namespace {
namespace cg = cooperative_groups;

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct ConvArgs {
  eadgan_conv_desc d;
  eadgan_tensor4 a;  // gathered operand: x (fprop, wgrad) or dy (dgrad)
  eadgan_tensor4 o;  // output y / dx (fprop, dgrad) or the dense operand dy (wgrad)
  eadgan_tensor4 mask;  // optional: output *= mask_act'(mask[n,ch,y,x]) (fused activation backward)
  int mask_act;
  float mask_slope;
  const float* w;
  const float* bias;
  float* dw; float* partial;
  int act;
  float slope;
  int M, N, K;       // GEMM extents of this direction
  int m_per_split;   // wgrad only
  int ksplit;        // fprop / dgrad: CTAs of one thread-block cluster sharing an output tile along K (1 = none)
};

// ---- fprop / dgrad: C[M,N] = A_gather[M,K] * B[K,N] --------------------------------------
template <bool DGRAD>
__global__ void __launch_bounds__(NT) conv_gemm_kernel(const ConvArgs P) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float red[BM * BN];   // split-K only: this CTA's partial tile, read by the cluster's first CTA
  const eadgan_conv_desc& d = P.d;
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int rs = d.r * d.s;
  // extents of the map this GEMM's rows enumerate (output pixels)
  const int OH = DGRAD ? d.h : d.p, OW = DGRAD ? d.w : d.q;

  // A-load mapping: fixed row per thread
  const int am = t % BM;
  const int ak0 = t / BM;  // 0..3, k = ak0 + 4*i
  const int m = m0 + am;
  const bool m_ok = m < P.M;
  int ab = 0, aoy = 0, aox = 0;
  if (m_ok) {
    ab = m / (OH * OW);
    int rem = m - ab * OH * OW;
    aoy = rem / OW;
    aox = rem - aoy * OW;
  }
  const int64_t a_base = (int64_t)ab * P.a.sn;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int tx = t % 16, ty = t / 16;  // tx -> rows (m), ty -> cols (n)
  // nn.Linear (1x1 kernel on a 1x1 map): the gather is the identity -- A(m, kk) = a[m][kk], B as below with tap 0 --
  // and the per-element index arithmetic (two integer divisions per gathered element) is most of what a Linear head
  // executes: an 8-CTA launch with K = 1024 took 102 us for 67 MFLOP (ncu r02wd).  Same values, same order.
  const bool lin = rs == 1 && d.h == 1 && d.w == 1 && d.p == 1 && d.q == 1 && d.pad == 0;

  // Split K: a launch with few output tiles and a long reduction (a Linear head: 8 tiles, K = 1024) runs as clusters
  // of `ksplit` CTAs along z; CTA z reduces k tiles [z * per, (z + 1) * per), the partial tiles are summed through
  // distributed shared memory by the cluster's first CTA in rank order (deterministic) and only that CTA stores.
  const int k_tiles = (P.K + BK - 1) / BK;
  const int per = (k_tiles + P.ksplit - 1) / P.ksplit;
  const int k_begin = (P.ksplit > 1 ? (int)blockIdx.z * per : 0) * BK;
  const int k_end = P.ksplit > 1 ? min(P.K, ((int)blockIdx.z + 1) * per * BK) : P.K;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- gather A
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kl = ak0 + 4 * i;
      const int kk = k0 + kl;
      float v = 0.f;
      if (lin) {
        if (m_ok && kk < P.K) v = eg_ld(P.a.ptr, a_base + (int64_t)kk * P.a.sc, P.a.dtype);
      } else if (m_ok && kk < P.K) {
        const int ch = kk / rs;
        const int tap = kk - ch * rs;
        const int ky = tap / d.s, kx = tap - ky * d.s;
        if (!DGRAD) {
          const int iy = aoy * d.stride - d.pad + ky, ix = aox * d.stride - d.pad + kx;
          if (iy >= 0 && iy < d.h && ix >= 0 && ix < d.w)
            v = eg_ld(P.a.ptr, a_base + (int64_t)ch * P.a.sc + (int64_t)iy * P.a.sh + (int64_t)ix * P.a.sw,
                      P.a.dtype);
        } else {
          const int tyy = aoy + d.pad - ky, txx = aox + d.pad - kx;
          if (tyy >= 0 && txx >= 0) {
            const int oy = tyy / d.stride, ox = txx / d.stride;
            if (oy * d.stride == tyy && ox * d.stride == txx && oy < d.p && ox < d.q)
              v = eg_ld(P.a.ptr, a_base + (int64_t)ch * P.a.sc + (int64_t)oy * P.a.sh + (int64_t)ox * P.a.sw,
                        P.a.dtype);
          }
        }
      }
      As[kl][am] = v;
    }
    // ---- load B
    if (!DGRAD) {
      // B(kk, n) = w[n*K + kk]; consecutive threads -> consecutive kk
      const int kl = t % BK, nl0 = t / BK;  // nl = nl0 + 16*i
      const int kk = k0 + kl;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int nl = nl0 + 16 * i;
        const int nn = n0 + nl;
        float v = 0.f;
        if (kk < P.K && nn < P.N) v = P.w[(int64_t)nn * P.K + kk];
        Bs[kl][nl] = v;
      }
    } else {
      // B(kk, n) = w[(ko*c + n)*rs + tap], kk = ko*rs + tap
      const int kl = t % BK, nl0 = t / BK;
      const int kk = k0 + kl;
      const int ko = lin ? kk : kk / rs, tap = kk - ko * rs;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int nl = nl0 + 16 * i;
        const int nn = n0 + nl;
        float v = 0.f;
        if (kk < P.K && nn < P.N) v = P.w[((int64_t)ko * d.c + nn) * rs + tap];
        Bs[kl][nl] = v;
      }
    }
    __syncthreads();
    // two-level accumulation: each 16-deep k tile is summed on its own, then added to the running
    // total -- rounding error grows with sqrt(K/16) instead of sqrt(K) (K is up to 16384 here)
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kl = 0; kl < BK; ++kl) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kl][tx * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kl][ty * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }

  if (P.ksplit > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(tx * 4 + i) * BN + ty * 4 + j] = acc[i][j];
    cluster.sync();
    if (rank == 0) {
      for (int r = 1; r < P.ksplit; ++r) {
        const float* rr = cluster.map_shared_rank(red, r);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += rr[(tx * 4 + i) * BN + ty * 4 + j];
      }
    }
    cluster.sync();        // nobody leaves while its partial tile may still be read
    if (rank != 0) return;
  }

  // ---- epilogue: + bias, activation, strided store
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + tx * 4 + i;
    if (mm >= P.M) continue;
    const int b = mm / (OH * OW);
    const int rem = mm - b * OH * OW;
    const int oy = rem / OW, ox = rem - oy * OW;
    const int64_t base = (int64_t)b * P.o.sn + (int64_t)oy * P.o.sh + (int64_t)ox * P.o.sw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + ty * 4 + j;
      if (nn >= P.N) continue;
      float v = acc[i][j];
      if (P.bias) v += P.bias[nn];
      v = eg_act(v, P.act, P.slope);
      if (P.mask_act != EADGAN_ACT_NONE) {
        const float mv = eg_ld(P.mask.ptr, (int64_t)b * P.mask.sn + (int64_t)nn * P.mask.sc +
                                               (int64_t)oy * P.mask.sh + (int64_t)ox * P.mask.sw, P.mask.dtype);
        v *= eg_act_grad(mv, P.mask_act, P.mask_slope);
      }
      eg_st(P.o.ptr, base + (int64_t)nn * P.o.sc, P.o.dtype, v);
    }
  }
}

// ---- wgrad: dw[ko, kk] += sum_m dy[m, ko] * xg[m, kk] ------------------------------------
__global__ void __launch_bounds__(NT) conv_wgrad_kernel(const ConvArgs P) {
  __shared__ float Ps[BK][BM + 4];  // [m chunk][ko]
  __shared__ float Qs[BK][BN + 4];  // [m chunk][kk]
  const eadgan_conv_desc& d = P.d;
  const int t = threadIdx.x;
  const int ko0 = blockIdx.y * BM, kk0 = blockIdx.x * BN;
  const int rs = d.r * d.s;
  const int Mtot = d.n * d.p * d.q;
  const int m_begin = blockIdx.z * P.m_per_split;
  const int m_end = min(Mtot, m_begin + P.m_per_split);

  const int ml = t % BK;    // row of the chunk this thread loads
  const int cl0 = t / BK;   // column = cl0 + 16*i
  // precompute the 4 kk decodes (fixed per thread)
  int q_ch[4], q_ky[4], q_kx[4];
  bool q_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kk = kk0 + cl0 + 16 * i;
    q_ok[i] = kk < P.N;
    const int ch = kk / rs;
    const int tap = kk - ch * rs;
    q_ch[i] = ch;
    q_ky[i] = tap / d.s;
    q_kx[i] = tap - q_ky[i] * d.s;
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tx = t % 16, ty = t / 16;  // tx -> ko, ty -> kk

  for (int mc = m_begin; mc < m_end; mc += BK) {
    const int m = mc + ml;
    const bool ok = m < m_end;
    int b = 0, oy = 0, ox = 0;
    if (ok) {
      b = m / (d.p * d.q);
      const int rem = m - b * d.p * d.q;
      oy = rem / d.q;
      ox = rem - oy * d.q;
    }
    const int64_t dy_base = (int64_t)b * P.o.sn + (int64_t)oy * P.o.sh + (int64_t)ox * P.o.sw;
    const int64_t x_base = (int64_t)b * P.a.sn;
    const int iy0 = oy * d.stride - d.pad, ix0 = ox * d.stride - d.pad;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = cl0 + 16 * i;
      const int ko = ko0 + col;
      float pv = 0.f;
      if (ok && ko < P.M) pv = eg_ld(P.o.ptr, dy_base + (int64_t)ko * P.o.sc, P.o.dtype);
      Ps[ml][col] = pv;
      float qv = 0.f;
      if (ok && q_ok[i]) {
        const int iy = iy0 + q_ky[i], ix = ix0 + q_kx[i];
        if (iy >= 0 && iy < d.h && ix >= 0 && ix < d.w)
          qv = eg_ld(P.a.ptr, x_base + (int64_t)q_ch[i] * P.a.sc + (int64_t)iy * P.a.sh + (int64_t)ix * P.a.sw,
                     P.a.dtype);
      }
      Qs[ml][col] = qv;
    }
    __syncthreads();
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kl = 0; kl < BK; ++kl) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Ps[kl][tx * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Qs[kl][ty * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ko = ko0 + tx * 4 + i;
    if (ko >= P.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = kk0 + ty * 4 + j;
      if (kk >= P.N) continue;
      // split z writes its own slab of the workspace (or dw itself when there is one split): no atomics, so
      // the result is run-to-run deterministic; wgrad_sum_kernel adds the slabs in a fixed order
      float* dst = gridDim.z == 1 ? P.dw : P.partial + (int64_t)blockIdx.z * P.M * P.N;
      dst[(int64_t)ko * P.N + kk] = acc[i][j];
    }
  }
}

__global__ void wgrad_sum_kernel(const float* __restrict__ partial, int splits, int64_t numel, float* __restrict__ dw) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[(int64_t)sp * numel + i];
    dw[i] = s;
  }
}

// out[ch] = sum_{n,h,w} t[n,ch,h,w]; one block per channel
__global__ void channel_sum_kernel(eadgan_tensor4 t, int n, int c, int h, int w, float* out) {
  __shared__ float red[32];
  const int ch = blockIdx.x;
  const int64_t hw = (int64_t)h * w;
  const int64_t total = (int64_t)n * hw;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int64_t b = i / hw;
    const int64_t rem = i - b * hw;
    const int y = (int)(rem / w), x = (int)(rem - (int64_t)y * w);
    s += eg_ld(t.ptr, b * t.sn + (int64_t)ch * t.sc + (int64_t)y * t.sh + (int64_t)x * t.sw, t.dtype);
  }
  s = eg_block_sum(s, red);
  if (threadIdx.x == 0) out[ch] = s;
}

__global__ void copy4_kernel(eadgan_tensor4 src, eadgan_tensor4 dst, int n, int c, int h, int w,
                             int c_fast) {
  const int64_t total = (int64_t)n * c * h * w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int b, ch, y, x;
    int64_t r = i;
    if (c_fast) {  // enumerate channel fastest (NHWC-friendly)
      ch = (int)(r % c); r /= c;
      x = (int)(r % w); r /= w;
      y = (int)(r % h); r /= h;
      b = (int)r;
    } else {
      x = (int)(r % w); r /= w;
      y = (int)(r % h); r /= h;
      ch = (int)(r % c); r /= c;
      b = (int)r;
    }
    const float v = eg_ld(src.ptr, (int64_t)b * src.sn + (int64_t)ch * src.sc + (int64_t)y * src.sh + (int64_t)x * src.sw, src.dtype);
    eg_st(dst.ptr, (int64_t)b * dst.sn + (int64_t)ch * dst.sc + (int64_t)y * dst.sh + (int64_t)x * dst.sw, dst.dtype, v);
  }
}

// few output tiles and a long reduction: clusters of 2 / 4 / 8 CTAs along K (see conv_gemm_kernel)
int pick_ksplit(dim3 grid, int K) {
  const int ctas = (int)(grid.x * grid.y), k_tiles = (K + BK - 1) / BK;
  if (ctas > 37 || k_tiles < 16) return 1;
  int ks = 8;
  while (ks > 1 && (k_tiles / ks < 4 || ctas * ks > 296)) ks >>= 1;
  return ks;
}

template <bool DGRAD>
int launch_conv_gemm(ConvArgs& P, dim3 grid, cudaStream_t st) {
  P.ksplit = pick_ksplit(grid, P.K);
  if (P.ksplit == 1) {
    conv_gemm_kernel<DGRAD><<<grid, NT, 0, st>>>(P);
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid.x, grid.y, P.ksplit); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = P.ksplit;
  cfg.attrs = attr; cfg.numAttrs = 1;
  EG_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<DGRAD>, P));
  return 0;
}

int check_desc(const eadgan_conv_desc* d) {
  EG_REQUIRE(d != nullptr, EADGAN_ERR_INVALID, "conv desc is NULL");
  EG_REQUIRE(d->n > 0 && d->c > 0 && d->h > 0 && d->w > 0 && d->k > 0 && d->r > 0 && d->s > 0 &&
                 d->p > 0 && d->q > 0 && d->stride > 0 && d->pad >= 0,
             EADGAN_ERR_INVALID, "conv desc has non-positive extents");
  EG_REQUIRE((d->h + 2 * d->pad - d->r) / d->stride + 1 == d->p &&
                 (d->w + 2 * d->pad - d->s) / d->stride + 1 == d->q,
             EADGAN_ERR_INVALID, "conv desc: p,q inconsistent with h,w,r,s,stride,pad");
  return 0;
}

}  // namespace

extern "C" int eadgan_conv_fprop(const eadgan_conv_desc* d, const eadgan_tensor4* x, const float* w,
                                 const float* bias, int act, float slope, const eadgan_tensor4* y,
                                 const eadgan_tensor4* mask, int mask_act, float mask_slope, void* stream) {
  if (int e = check_desc(d)) return e;
  EG_REQUIRE(x && y && x->ptr && y->ptr && w, EADGAN_ERR_INVALID, "conv_fprop: NULL tensor");
  ConvArgs P{};
  P.d = *d; P.a = *x; P.o = *y; P.w = w; P.bias = bias; P.act = act; P.slope = slope;
  if (mask && mask_act != EADGAN_ACT_NONE) { P.mask = *mask; P.mask_act = mask_act; P.mask_slope = mask_slope; }
  P.M = d->n * d->p * d->q; P.N = d->k; P.K = d->c * d->r * d->s;
  dim3 grid((P.M + BM - 1) / BM, (P.N + BN - 1) / BN);
  if (int e = launch_conv_gemm<false>(P, grid, (cudaStream_t)stream)) return e;
  EG_LAUNCH_CHECK("conv_gemm_kernel<fprop>");
  return 0;
}

extern "C" int eadgan_conv_dgrad(const eadgan_conv_desc* d, const eadgan_tensor4* dy, const float* w,
                                 const float* bias, int act, float slope, const eadgan_tensor4* dx,
                                 const eadgan_tensor4* mask, int mask_act, float mask_slope, void* stream) {
  if (int e = check_desc(d)) return e;
  EG_REQUIRE(dy && dx && dy->ptr && dx->ptr && w, EADGAN_ERR_INVALID, "conv_dgrad: NULL tensor");
  ConvArgs P{};
  P.d = *d; P.a = *dy; P.o = *dx; P.w = w; P.bias = bias; P.act = act; P.slope = slope;
  if (mask && mask_act != EADGAN_ACT_NONE) { P.mask = *mask; P.mask_act = mask_act; P.mask_slope = mask_slope; }
  P.M = d->n * d->h * d->w; P.N = d->c; P.K = d->k * d->r * d->s;
  dim3 grid((P.M + BM - 1) / BM, (P.N + BN - 1) / BN);
  if (int e = launch_conv_gemm<true>(P, grid, (cudaStream_t)stream)) return e;
  EG_LAUNCH_CHECK("conv_gemm_kernel<dgrad>");
  return 0;
}

namespace {
// split of the reduction (pixel) dimension of the wgrad GEMM: enough CTAs to fill the machine, at least 4
// BK-chunks of work per split
void wgrad_split(const eadgan_conv_desc* d, int* splits, int* mps) {
  const int M = d->k, N = d->c * d->r * d->s, K = d->n * d->p * d->q;
  const int tiles = ((N + BN - 1) / BN) * ((M + BM - 1) / BM);
  const int target = 4 * eg_sm_count();
  int sp = (target + tiles - 1) / tiles;
  const int max_splits = (K + 4 * BK - 1) / (4 * BK);
  if (sp > max_splits) sp = max_splits;
  if (sp < 1) sp = 1;
  int m = (K + sp - 1) / sp;
  m = ((m + BK - 1) / BK) * BK;
  *splits = (K + m - 1) / m;
  *mps = m;
}
}  // namespace

extern "C" size_t eadgan_conv_wgrad_workspace(const eadgan_conv_desc* d) {
  if (!d || check_desc(d)) return 0;
  int splits, mps;
  wgrad_split(d, &splits, &mps);
  return splits > 1 ? (size_t)splits * d->k * d->c * d->r * d->s * sizeof(float) : 0;
}

extern "C" int eadgan_conv_wgrad(const eadgan_conv_desc* d, const eadgan_tensor4* x,
                                 const eadgan_tensor4* dy, float* dw, void* workspace, size_t ws_bytes,
                                 void* stream) {
  if (int e = check_desc(d)) return e;
  EG_REQUIRE(x && dy && x->ptr && dy->ptr && dw, EADGAN_ERR_INVALID, "conv_wgrad: NULL tensor");
  ConvArgs P{};
  P.d = *d; P.a = *x; P.o = *dy; P.dw = dw;
  P.M = d->k; P.N = d->c * d->r * d->s; P.K = d->n * d->p * d->q;
  int splits, mps;
  wgrad_split(d, &splits, &mps);
  const size_t need = splits > 1 ? (size_t)splits * P.M * P.N * sizeof(float) : 0;
  EG_REQUIRE(need == 0 || (workspace && ws_bytes >= need), EADGAN_ERR_WORKSPACE,
             "conv_wgrad: workspace %zu < %zu bytes", ws_bytes, need);
  P.partial = (float*)workspace;
  P.m_per_split = mps;
  dim3 grid((P.N + BN - 1) / BN, (P.M + BM - 1) / BM, splits);
  conv_wgrad_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(P);
  EG_LAUNCH_CHECK("conv_wgrad_kernel");
  if (splits > 1) {
    const int64_t numel = (int64_t)P.M * P.N;
    int blocks = (int)((numel + 255) / 256);
    if (blocks > 8 * eg_sm_count()) blocks = 8 * eg_sm_count();
    wgrad_sum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P.partial, splits, numel, dw);
    EG_LAUNCH_CHECK("wgrad_sum_kernel");
  }
  return 0;
}

extern "C" int eadgan_channel_sum(const eadgan_tensor4* t, int n, int c, int h, int w, float* out,
                                  void* stream) {
  EG_REQUIRE(t && t->ptr && out && n > 0 && c > 0 && h > 0 && w > 0, EADGAN_ERR_INVALID,
             "channel_sum: bad arguments");
  channel_sum_kernel<<<c, 256, 0, (cudaStream_t)stream>>>(*t, n, c, h, w, out);
  EG_LAUNCH_CHECK("channel_sum_kernel");
  return 0;
}

extern "C" int eadgan_copy4(const eadgan_tensor4* src, const eadgan_tensor4* dst, int n, int c,
                            int h, int w, void* stream) {
  EG_REQUIRE(src && dst && src->ptr && dst->ptr && n > 0 && c > 0 && h > 0 && w > 0,
             EADGAN_ERR_INVALID, "copy4: bad arguments");
  const int64_t total = (int64_t)n * c * h * w;
  int blocks = (int)((total + 255) / 256);
  const int cap = 32 * eg_sm_count();
  if (blocks > cap) blocks = cap;
  const int c_fast = (dst->sc == 1) ? 1 : 0;
  copy4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*src, *dst, n, c, h, w, c_fast);
  EG_LAUNCH_CHECK("copy4_kernel");
  return 0;
}

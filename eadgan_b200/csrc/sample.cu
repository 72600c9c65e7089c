// Device-side sampling of the per-iteration latent draws.
//
// The reference samples z ~ N(0,1), code ~ U(-1,1) and the class labels with NumPy on the HOST every iteration and
// copies them to the GPU (celebA/EAD-GAN_celebA.py:308-318, dSprites/rp.py:389-396,424-434).  Here they are drawn on
// the device with the counter-based generator Philox4x32-10 (Salmon et al., SC'11), so a captured training step needs
// no host RNG, no H2D copy of latents, and its stream is reproducible on the host word for word
// (oracle/philox_ref.py restates it in numpy; tests/test_sampling_gpu.py):
//
//   key     = (seed lo, seed hi)
//   counter = (g lo, g hi | stream << 24, step lo, step hi)      g = index of a group of 4 consecutive elements of the
//                                                                 GLOBAL [rows_global, cols] array, row-major
//   words x0..x3 of the group -> elements 4g .. 4g+3
//     uniform:  lo + (hi - lo) * (x >> 8) * 2^-24
//     randint:  (x * n) >> 32                      (int64 output)
//     normal:   Box-Muller on (x0, x1) and (x2, x3):  r = sqrt(-2 ln((x_a >> 8) + 1) * 2^-24)),  phi = 2 pi (x_b >> 8) 2^-24
//               -> r cos(phi), r sin(phi)
// Indexing by GLOBAL element makes the draws independent of how the batch is sharded: rank r of N asks for rows
// [r B/N, (r+1) B/N) and gets exactly the rows a single device would draw.
// `step` comes from a DEVICE int64 when step_dev != NULL (CUDA-graph replay: advanced by eadgan_adam_advance).
#include "common.cuh"

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

enum { KIND_UNIFORM = 0, KIND_NORMAL = 1, KIND_RANDINT = 2 };

struct SampleParams {
  uint64_t seed;
  const int64_t* step_dev;
  int64_t step_host;
  uint32_t stream;
  int64_t first, count;    // global element range [first, first + count)
  float lo, hi;
  int n;                   // randint: values in [0, n)
  void* out;               // float* (uniform, normal) or int64_t* (randint), indexed by (global element - first)
};

template <int KIND>
__global__ void __launch_bounds__(256) philox_kernel(const SampleParams P) {
  const int64_t g0 = P.first >> 2, g1 = (P.first + P.count + 3) >> 2;   // groups [g0, g1)
  const uint64_t step = (uint64_t)(P.step_dev ? *P.step_dev : P.step_host);
  for (int64_t g = g0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < g1; g += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)g, (uint32_t)((uint64_t)g >> 32) | (P.stream << 24), (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
    float v[4];
    int64_t iv[4];
    if (KIND == KIND_UNIFORM) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaf(P.hi - P.lo, (float)(c[j] >> 8) * 5.9604644775390625e-8f, P.lo);
    } else if (KIND == KIND_NORMAL) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float u1 = (float)((c[2 * j] >> 8) + 1u) * 5.9604644775390625e-8f;     // (0, 1]
        const float u2 = (float)(c[2 * j + 1] >> 8) * 5.9604644775390625e-8f;        // [0, 1)
        const float r = sqrtf(-2.f * logf(u1));
        float s, co;
        sincosf(6.283185307179586f * u2, &s, &co);
        v[2 * j] = r * co; v[2 * j + 1] = r * s;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) iv[j] = (int64_t)(((uint64_t)c[j] * (uint64_t)P.n) >> 32);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t e = 4 * g + j - P.first;
      if (e >= 0 && e < P.count) {
        if (KIND == KIND_RANDINT) reinterpret_cast<int64_t*>(P.out)[e] = iv[j];
        else reinterpret_cast<float*>(P.out)[e] = v[j];
      }
    }
  }
}

int launch(int kind, const SampleParams& P, cudaStream_t st) {
  const int64_t groups = ((P.first + P.count + 3) >> 2) - (P.first >> 2);
  int blocks = (int)((groups + 255) / 256);
  if (blocks > 8 * eg_sm_count()) blocks = 8 * eg_sm_count();
  if (blocks < 1) blocks = 1;
  if (kind == KIND_UNIFORM) philox_kernel<KIND_UNIFORM><<<blocks, 256, 0, st>>>(P);
  else if (kind == KIND_NORMAL) philox_kernel<KIND_NORMAL><<<blocks, 256, 0, st>>>(P);
  else philox_kernel<KIND_RANDINT><<<blocks, 256, 0, st>>>(P);
  EG_LAUNCH_CHECK("philox_kernel");
  return 0;
}

}  // namespace

extern "C" int eadgan_philox(int kind, unsigned long long seed, const int64_t* step_dev, long long step_host,
                             int stream_id, long long row0, long long rows, long long cols, float lo, float hi, int n,
                             void* out, void* stream) {
  EG_REQUIRE(out && rows > 0 && cols > 0 && row0 >= 0 && kind >= 0 && kind <= 2 && stream_id >= 0 && stream_id < 256,
             EADGAN_ERR_INVALID, "philox: bad arguments");
  EG_REQUIRE(kind != KIND_RANDINT || n > 0, EADGAN_ERR_INVALID, "philox: randint needs n > 0");
  SampleParams P{};
  P.seed = seed; P.step_dev = step_dev; P.step_host = step_host; P.stream = (uint32_t)stream_id;
  P.first = row0 * cols; P.count = rows * cols; P.lo = lo; P.hi = hi; P.n = n; P.out = out;
  EG_REQUIRE(((uint64_t)((P.first + P.count + 3) >> 2) >> 56) == 0, EADGAN_ERR_UNSUPPORTED, "philox: array too large");
  return launch(kind, P, (cudaStream_t)stream);
}

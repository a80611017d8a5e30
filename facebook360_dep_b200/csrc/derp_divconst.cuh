// Correctly rounded fp32 division by a loop-invariant divisor in three instructions.
// The filter kernels divide per tap by constants ( /3.0f, /(2 sigma^2), /65535.0f, /sigma^2:
// TemporalBilateralFilter.h:103-104,176-185 ) and the results must round like the CPU's IEEE divisions.
// The hardware-free IEEE division sequence costs ~10 instructions + a range branch; with the reciprocal
// known up front,    q = RN(a * rc);  r = fma(-q, c, a) (exact);  q' = fma(r, rc, q)
// is RN(a / c) (Markstein) whenever rc = RN(1/c) and nothing under/overflows.  The theorem's side
// condition is not taken on trust: divConstExact() checks a constant EXHAUSTIVELY (every mantissa of the
// dividend; the sequence is exponent-invariant in the normal range) on the device against __fdiv_rn the
// first time it is used, and the kernels fall back to the plain division for a constant that fails.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>

namespace derp {

struct DivConst {
  float c, rc;
  int fast;  // 1: the three-instruction sequence is exact for this divisor
};

__device__ __forceinline__ float divBy(float a, const DivConst k) {
  if (!k.fast) return __fdiv_rn(a, k.c);  // uniform branch
  const float q = __fmul_rn(a, k.rc);
  const float r = __fmaf_rn(-q, k.c, a);
  return __fmaf_rn(r, k.rc, q);
}

__global__ void divConstCheckKernel(float c, float rc, unsigned long long* bad) {
  const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;  // 2^23 mantissas
  if (m >= (1u << 23)) return;
  unsigned long long n = 0;
  // exponents spanning the magnitudes the filters produce (1e-10 .. 1e5) and both signs
  for (int e = 127 - 34; e <= 127 + 18; e += 13) {
    const float a = __uint_as_float(((unsigned)e << 23) | m);
    const float q = __fmul_rn(a, rc);
    const float r = __fmaf_rn(-q, c, a);
    const float f = __fmaf_rn(r, rc, q);
    n += (__float_as_uint(f) != __float_as_uint(__fdiv_rn(a, c)));
    n += (__float_as_uint(-f) != __float_as_uint(__fdiv_rn(-a, c)));
  }
  if (n) atomicAdd(bad, n);
}

// host: descriptor for divisor c; validated once per distinct value and device-independent (IEEE arithmetic)
inline cudaError_t makeDivConst(float c, cudaStream_t stream, DivConst* out) {
  static std::mutex mu;
  static std::map<unsigned, int> cache;
  unsigned key;
  memcpy(&key, &c, 4);
  out->c = c;
  out->rc = 1.0f / c;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      out->fast = it->second;
      return cudaSuccess;
    }
  }
  int fast = 0;
  if (c == c && c != 0 && fabsf(c) > 1e-30f && fabsf(c) < 1e30f) {
    unsigned long long* dBad = nullptr;
    cudaError_t e = cudaMalloc(&dBad, sizeof(*dBad));
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(dBad, 0, sizeof(*dBad), stream);
    divConstCheckKernel<<<(1u << 23) / 256, 256, 0, stream>>>(c, out->rc, dBad);
    unsigned long long bad = 1;
    e = cudaMemcpyAsync(&bad, dBad, sizeof(bad), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(dBad);
    if (e != cudaSuccess) return e;
    fast = bad == 0;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    cache[key] = fast;
  }
  out->fast = fast;
  return cudaSuccess;
}

}  // namespace derp

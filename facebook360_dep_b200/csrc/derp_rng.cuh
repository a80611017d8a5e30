// std::default_random_engine (= std::minstd_rand0) + std::uniform_real_distribution<float> as
// libstdc++ implements them (bits/random.h, bits/random.tcc: generate_canonical<float,24> draws ONE
// engine value r and returns float(r - 1) / 2^31, clamped below 1), with O(log n) skip-ahead so that
// randomProposal's per-row sequential stream (Derp.cpp:757-808) can be evaluated per pixel.
// tests/test_host_units.py checks the stream on the host against libstdc++ itself.
#pragma once

#if defined(__CUDACC__)
#define DERP_RNG_HD __host__ __device__ __forceinline__
#else
#define DERP_RNG_HD inline
#endif

namespace derp {

struct MinstdRand0 {
  unsigned x;
  static constexpr unsigned kM = 2147483647u;  // 2^31 - 1
  static constexpr unsigned kA = 16807u;

  DERP_RNG_HD static unsigned mulmod(unsigned a, unsigned b) {
    return (unsigned)(((unsigned long long)a * (unsigned long long)b) % (unsigned long long)kM);
  }
  // linear_congruential_engine::seed: x = s mod m, and 0 -> 1 because c == 0
  DERP_RNG_HD void seed(unsigned s) {
    x = s % kM;
    if (x == 0) x = 1;
  }
  // advance the state by n draws: x <- a^n x mod m
  DERP_RNG_HD void discard(unsigned long long n) {
    unsigned base = kA, acc = 1;
    while (n) {
      if (n & 1ull) acc = mulmod(acc, base);
      base = mulmod(base, base);
      n >>= 1;
    }
    x = mulmod(x, acc);
  }
  DERP_RNG_HD unsigned next() {
    x = mulmod(x, kA);
    return x;
  }
  // uniform_real_distribution<float>(a, b)(engine)
  DERP_RNG_HD float uniform(float a, float b) {
    const unsigned r = next();
    float u = (float)(r - 1u) * 4.656612873077392578125e-10f;  // / float(2147483646.0L) == 2^31
    if (u >= 1.0f) u = 0.99999994f;                            // nextafter(1.f, 0.f)
    return u * (b - a) + a;
  }
};

}  // namespace derp

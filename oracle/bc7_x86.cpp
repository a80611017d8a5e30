// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// The product's BC7 block encoder (facebook360_dep_b200/csrc/derp_bc7.cuh, the same functions the CUDA kernel runs)
// compiled as plain C++ with -DDERP_BC7_X86_ESTIMATES: `a / b` and 1 / sqrt become the RCPPS / RSQRTPS + Newton sequences
// the reference's ispc build emits (kernel.ispc compiled with --opt=fast-math, ISPC.cmake:4).  tests/test_bc7.py uses it to
// show that the encoder restates the reference's algorithm exactly: with the same estimate instructions the blocks are
// byte-identical to oracle/_ref's; the product (IEEE division) can differ from the reference only where the last bit of an
// estimate decides.  Never linked into or loaded by the product.
#define DERP_BC7_X86_ESTIMATES 1
#include "../facebook360_dep_b200/csrc/derp_bc7.cuh"

extern "C" int derp_x86_bc7_blocks(const uint8_t* rgba, int width, int height, uint8_t* blocks) {
  std::memset(blocks, 0, (size_t)width * height);
  derp::bc7::encodeSurfaceOnHost(rgba, width, height, blocks);
  return 0;
}

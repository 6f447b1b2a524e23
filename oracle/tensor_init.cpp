// tensor_init.cpp - restatement of tpp-run's deterministic input generators.
// TEST INFRASTRUCTURE (see xsmm_oracle.h): used so that the GPU path and the
// CPU oracle see exactly the data a Linux tpp-run would feed the kernels.
//
// Follows:
//   include/TPP/Transforms/Utils/TensorInitFloat.h:89-152  (generator classes)
//   lib/TPP/Transforms/Utils/TensorInitFloat.cpp:54-95     (fill order)
//   lib/TPP/Transforms/Utils/TensorInit.cpp:63-144         (one generator per
//       (type, dtype, seed), reused sequentially across tensors)
// The reference uses std::default_random_engine and
// std::normal_distribution<float>; both are implementation-defined, so this file
// must be built with g++/libstdc++ (what a Linux tpp-run is built with) to
// reproduce the stream. Pinned by test/Integration/xsmm-fusion.mlir:54-57
// (seed 123) in tests/test_oracle_golden.py.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <random>

extern "C" uint16_t xo_f32_to_bf16(float f);

namespace {

enum InitType { kConst = 0, kSimple = 1, kCont = 2, kRandom = 3, kNormal = 4 };

struct Generator {
  int type;
  int dtype; // 1 = f32, 2 = bf16
  std::default_random_engine engine;
  std::uniform_real_distribution<float> uniform{0.0f, 1.0f};
  std::normal_distribution<float> normal{0.0f, 0.2f};
  Generator(int t, int d, int seed) : type(t), dtype(d), engine(seed) {}
};

inline void store(int dtype, void *out, int64_t i, float v) {
  if (dtype == 1)
    static_cast<float *>(out)[i] = v;
  else
    static_cast<uint16_t *>(out)[i] = xo_f32_to_bf16(v);
}

} // namespace

extern "C" {

// type: 0 const(1.0), 1 simple(0.3,0.6,0.9), 2 cont(i/size), 3 random U(0,1),
//       4 normal N(0,0.2) clamped to [0,1]   (TensorInit.cpp:63-73)
void *ti_create(int type, int dtype, int seed) {
  return new Generator(type, dtype, seed);
}

void ti_destroy(void *h) { delete static_cast<Generator *>(h); }

// Fill the next tensor of `size` elements from this generator's stream
// (values are produced in f32 and rounded RNE to the element type,
// TensorInitFloat.cpp:36-52).
void ti_fill(void *h, int64_t size, void *out) {
  Generator &g = *static_cast<Generator *>(h);
  static const float simple[3] = {0.3f, 0.6f, 0.9f};
  for (int64_t i = 0; i < size; ++i) {
    float v;
    switch (g.type) {
    case kConst:
      v = 1.0f;
      break;
    case kSimple:
      v = simple[i % 3];
      break;
    case kCont:
      v = static_cast<float>(i) / static_cast<float>(size);
      break;
    case kRandom:
      v = g.uniform(g.engine);
      break;
    default: {
      float x = g.normal(g.engine);
      v = std::clamp(x, 0.0f, 1.0f);
    }
    }
    store(g.dtype, out, i, v);
  }
}

} // extern "C"

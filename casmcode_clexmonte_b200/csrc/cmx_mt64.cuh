// std::mt19937_64 and the libstdc++ distributions the reference draws from
// ([EXT] monte::RandomNumberGenerator; lotto::RandomGeneratorT, submodules/kmc-lotto/
// include/lotto/random.hpp:35-91), restated for the device.  Checked draw by draw against
// libstdc++ (tests: test_rng_stream_matches_libstdcxx).
#pragma once

#define MT_NN 312
#define MT_MM 156

struct Mt64 {
  unsigned long long mt[MT_NN];
  int idx;
};

__device__ inline void mt_seed(Mt64 &g, unsigned long long seed) {
  g.mt[0] = seed;
  for (int i = 1; i < MT_NN; ++i)
    g.mt[i] = 6364136223846793005ull * (g.mt[i - 1] ^ (g.mt[i - 1] >> 62)) + (unsigned long long)i;
  g.idx = MT_NN;
}

__device__ inline unsigned long long mt_next(Mt64 &g) {
  const unsigned long long UM = 0xFFFFFFFF80000000ull, LM = 0x7FFFFFFFull,
                           A = 0xB5026F5AA96619E9ull;
  if (g.idx >= MT_NN) {
    int i;
    for (i = 0; i < MT_NN - MT_MM; ++i) {
      unsigned long long x = (g.mt[i] & UM) | (g.mt[i + 1] & LM);
      g.mt[i] = g.mt[i + MT_MM] ^ (x >> 1) ^ ((x & 1ull) ? A : 0ull);
    }
    for (; i < MT_NN - 1; ++i) {
      unsigned long long x = (g.mt[i] & UM) | (g.mt[i + 1] & LM);
      g.mt[i] = g.mt[i + (MT_MM - MT_NN)] ^ (x >> 1) ^ ((x & 1ull) ? A : 0ull);
    }
    unsigned long long x = (g.mt[MT_NN - 1] & UM) | (g.mt[0] & LM);
    g.mt[MT_NN - 1] = g.mt[MT_MM - 1] ^ (x >> 1) ^ ((x & 1ull) ? A : 0ull);
    g.idx = 0;
  }
  unsigned long long x = g.mt[g.idx++];
  x ^= (x >> 29) & 0x5555555555555555ull;
  x ^= (x << 17) & 0x71D67FFFEDA60000ull;
  x ^= (x << 37) & 0xFFF7EEE000000000ull;
  x ^= (x >> 43);
  return x;
}

// libstdc++ std::generate_canonical<double,53>(mt19937_64): one draw,
// double(x) / 2^64, clamped below 1.
__device__ inline double mt_canonical(Mt64 &g) {
  double r = __ull2double_rn(mt_next(g)) * 5.421010862427522170037264004349708557128906250e-20;
  if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;
  return r;
}
// std::uniform_real_distribution<double>(0, max): canonical * (max - 0) + 0
__device__ inline double mt_real(Mt64 &g, double maxv) {
  return __dadd_rn(__dmul_rn(mt_canonical(g), __dsub_rn(maxv, 0.0)), 0.0);
}
// std::uniform_int_distribution<long>(0, maxv) for a 64-bit URNG: Lemire's
// nearly-divisionless method with 128-bit products (libstdc++ _S_nd).
__device__ inline long long mt_int(Mt64 &g, long long maxv) {
  unsigned long long urange = (unsigned long long)maxv;
  if (urange == 0xFFFFFFFFFFFFFFFFull) return (long long)mt_next(g);
  unsigned long long range = urange + 1ull;
  unsigned long long x = mt_next(g);
  unsigned long long low = x * range;
  unsigned long long high = __umul64hi(x, range);
  if (low < range) {
    unsigned long long threshold = (0ull - range) % range;
    while (low < threshold) {
      x = mt_next(g);
      low = x * range;
      high = __umul64hi(x, range);
    }
  }
  return (long long)high;
}

// lotto::RandomGeneratorT::sample_unit_interval (random.hpp:49-53,63-66):
// std::uniform_real_distribution<double>(nextafter(0, max), nextafter(1, max)), i.e. (0, 1]:
// canonical * (b - a) + a with a = 4.9e-324, b = 1 + 2^-52
__device__ inline double mt_unit_interval(Mt64 &g) {
  const double a = 4.9406564584124654e-324, b = 1.0000000000000002220446049250313080847263336181640625;
  return __dadd_rn(__dmul_rn(mt_canonical(g), __dsub_rn(b, a)), a);
}

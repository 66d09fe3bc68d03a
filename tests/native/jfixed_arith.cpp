// Host check of nix_b200/csrc/jfixed.cuh (the NIX_J_FIXED experiment): the header is compiled AS IS, the CUDA
// intrinsics it uses get one-line stand-ins (the adds of one "thread" at a time model any interleaving of atomics,
// because each atomicAdd is indivisible and the scheme never reads a word outside an atomic).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#define __device__
#define __forceinline__ inline
using std::fabs;
using std::fmax;
using std::ilogb;
using std::scalbn;
static inline long long __double2ll_rn(double x) { return std::llrint(x); } // round to nearest even, as cvt.rni.s64.f64
static inline long long __double_as_longlong(double x)
{
  long long v;
  std::memcpy(&v, &x, 8);
  return v;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v)
{
  unsigned o = *p;
  *p         = o + v;
  return o;
}
static inline float atomicAdd(float* p, float v)
{
  float o = *p;
  *p      = o + v;
  return o;
}
#include "jfixed.cuh"

static int fails = 0;
#define CHECK(c, ...)                                                                                                  \
  do {                                                                                                                 \
    if (!(c)) {                                                                                                        \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__);                                                                 \
      std::printf(__VA_ARGS__);                                                                                        \
      std::printf("\n");                                                                                               \
      fails++;                                                                                                         \
    }                                                                                                                  \
  } while (0)

static long long raw(double word) { return __double_as_longlong(word); }

int main()
{
  std::mt19937_64 gen(12345);
  // 1. the scale: a power of two that puts the largest contribution in [2^44, 2^45)
  for (double q : {1.0, -1.0, 1e-3, 3.7e5, 0.0123, -25.0}) {
    for (double r : {0.5, 2.0, 1.0, 7.3}) {
      const double s = jt_scale<double>(q, q * r, q * r * 0.9, q * r * 1.1), m = std::max({std::fabs(q), std::fabs(q * r), std::fabs(q * r * 1.1)});
      int          e = 0;
      CHECK(std::frexp(s, &e) == 0.5, "scale %g is not a power of two", s);
      CHECK(m * s >= std::ldexp(1.0, 44) && m * s < std::ldexp(1.0, 45), "scale %g for max %g", s, m);
    }
  }
  // 2. random contributions of both signs and many magnitudes: any order gives the same WORD, the word is the exact
  //    integer sum of the rounded contributions, and its value agrees with the fp64 sum to n/2 units
  for (int trial = 0; trial < 20; trial++) {
    const double q = std::ldexp(1.0 + (gen() % 1000) / 1000.0, (int)(gen() % 40) - 20) * ((trial & 1) ? -1 : 1);
    const double qd[3] = {q * 2.0, q * 0.7, q * 1.3};
    const double sc = jt_scale<double>(q, qd[0], qd[1], qd[2]), cmax = std::fabs(q) * 2.0;
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    std::vector<double> v(20000);
    for (auto& x : v) x = cmax * u(gen) * std::ldexp(1.0, -(int)(gen() % 30));
    double a = 0.0, b = 0.0;
    __int128 exact = 0;
    long double ref = 0.0L;
    for (double x : v) {
      jt_add<double>(&a, x, sc);
      exact += (__int128)std::llrint(x * sc);
      ref += (long double)x;
    }
    std::shuffle(v.begin(), v.end(), gen);
    for (double x : v) jt_add<double>(&b, x, sc);
    CHECK(raw(a) == raw(b), "order dependence: %lld vs %lld", raw(a), raw(b));
    CHECK((__int128)raw(a) == exact, "word is not the exact sum");
    const double got = jt_read<double>(&a, 1.0 / sc);
    CHECK(std::fabs(got - (double)ref) <= 0.5 * v.size() / sc, "value %.17g vs %.17Lg", got, ref);
    CHECK(std::fabs(got - (double)ref) <= 1e-12 * cmax * 10, "resolution: error %g against cmax %g", got - (double)ref, cmax);
  }
  // 3. carries and signs at the word boundary, in units of the fixed-point grid (scale 1)
  {
    double w = 0.0;
    jt_add<double>(&w, 4294967295.0, 1.0); // low word full
    jt_add<double>(&w, 1.0, 1.0);          // carry into the high word
    CHECK(raw(w) == 4294967296LL, "carry: %lld", raw(w));
    jt_add<double>(&w, -1.0, 1.0); // borrow back
    CHECK(raw(w) == 4294967295LL, "borrow: %lld", raw(w));
    jt_add<double>(&w, -4294967296.0, 1.0);
    CHECK(raw(w) == -1LL, "negative: %lld", raw(w));
    jt_add<double>(&w, -3.0e15, 1.0);
    jt_add<double>(&w, 1.0, 1.0);
    CHECK(raw(w) == -3000000000000000LL, "large negative: %lld", raw(w));
    CHECK(jt_read<double>(&w, 0.25) == -7.5e14, "read");
    double z = 0.0;
    jt_add<double>(&z, 0.2, 1.0); // rounds to zero: the word stays a clean zero
    CHECK(raw(z) == 0 && jt_read<double>(&z, 1.0) == 0.0, "zero");
  }
  // 4. many carries: 2^20 adds of 2^31 + 3 units
  {
    double w = 0.0;
    for (int i = 0; i < (1 << 20); i++) jt_add<double>(&w, 2147483651.0, 1.0);
    CHECK(raw(w) == 2147483651LL * (1LL << 20), "many carries: %lld", raw(w));
  }
  // 5. the fp32 instantiation is a plain add (unchanged behaviour)
  {
    float f = 1.5f;
    jt_add<float>(&f, 2.25f, 123.0f);
    CHECK(f == 3.75f && jt_read<float>(&f, 9.0f) == 3.75f, "float path");
  }
  std::printf(fails ? "FAILED %d\n" : "ok\n", fails);
  return fails ? 1 : 0;
}

#include <cstdint>
#include <thread>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
static void widen(const uint32_t* s, int64_t n, int64_t* d0, int64_t* d1) {
  for (int64_t i = 0; i < n; ++i) { d0[i] = s[2*i] & 0xFFFFFFu; d1[i] = s[2*i+1]; }
}
static void lutf(const uint8_t* s, int64_t n, double inv, float* d) {
  for (int64_t i = 0; i < n; ++i) d[i] = (float)((double)s[i] * inv);
}
template <class F> double run(int nt, int64_t n, F f) {
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) th.emplace_back([=]{ int64_t a = n * t / nt, b = n * (t+1) / nt; f(a, b); });
  for (auto& x : th) x.join();
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
int main() {
  const int64_t E = 20600000, X = 67108864;
  uint32_t* s = (uint32_t*)aligned_alloc(64, E * 8); int64_t* d0 = (int64_t*)aligned_alloc(64, E * 8); int64_t* d1 = (int64_t*)aligned_alloc(64, E * 8);
  uint8_t* c = (uint8_t*)aligned_alloc(64, X); float* x = (float*)aligned_alloc(64, X * 4);
  memset(s, 1, E*8); memset(d0, 0, E*8); memset(d1, 0, E*8); memset(c, 2, X); memset(x, 0, X*4);
  printf("hw threads %u\n", std::thread::hardware_concurrency());
  for (int nt : {1, 4, 8, 16, 32}) {
    double a = 1e9, b = 1e9;
    for (int r = 0; r < 4; ++r) {
      a = std::min(a, run(nt, E, [=](int64_t lo, int64_t hi){ widen(s + 2*lo, hi - lo, d0 + lo, d1 + lo); }));
      b = std::min(b, run(nt, X, [=](int64_t lo, int64_t hi){ lutf(c + lo, hi - lo, 0.25, x + lo); }));
    }
    printf("threads %2d  widen 20.6M edges %.2f ms   u8->f32 67M %.2f ms\n", nt, a, b);
  }
  return 0;
}

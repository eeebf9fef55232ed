// stdsort_test.cpp -- sloam_b200/csrc/dev_stdsort.h against this toolchain's std::sort.
//   g++ -std=c++17 -O2 tests/stdsort_test.cpp -o stdsort_test && ./stdsort_test
#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>
#include "../sloam_b200/csrc/dev_stdsort.h"

struct P { float k; int id; };

int main() {
  std::mt19937 rng(7);
  long cases = 0, bad = 0;
  for (int n = 0; n <= 300; ++n)
    for (int rep = 0; rep < (n <= 130 ? 400 : 40); ++rep) {
      const int levels = 1 + (int)(rng() % (rep % 5 == 0 ? 2 : (rep % 5 == 1 ? n + 1 : 8)));  // many ties
      std::vector<P> a(n);
      std::vector<float> key(n);
      for (int i = 0; i < n; ++i) { key[i] = (float)(rng() % levels); a[i] = {key[i], i}; }
      if (rep % 7 == 3) std::sort(key.begin(), key.end());              // presorted input
      if (rep % 7 == 4) std::sort(key.begin(), key.end(), std::greater<float>());
      for (int i = 0; i < n; ++i) a[i] = {key[i], i};
      std::sort(a.begin(), a.end(), [](const P &x, const P &y) { return x.k < y.k; });
      std::vector<int16_t> p(n);
      for (int i = 0; i < n; ++i) p[i] = (int16_t)i;
      sb::StdSort<int16_t> s(p.data(), key.data());
      s.sort(n);
      if (n > 0) {  // prefix variant: the first r positions of std::sort's result
        const int r = 1 + (int)(rng() % n);
        std::vector<int16_t> q(n);
        for (int i = 0; i < n; ++i) q[i] = (int16_t)i;
        sb::StdSort<int16_t> sq(q.data(), key.data());
        const int bound = sq.sort_prefix(n, r);
        if (bound < r) { ++bad; }
        for (int i = 0; i < r; ++i) if (q[i] != a[i].id) { ++bad; if (bad < 5) std::printf("prefix mismatch n=%d r=%d at %d\n", n, r, i); break; }
      }
      {  // packed variant must give the same permutation
        std::vector<unsigned long long> e(n);
        for (int i = 0; i < n; ++i) { unsigned u; __builtin_memcpy(&u, &key[i], 4); e[i] = ((unsigned long long)u << 32) | (unsigned)i; }
        sb::StdSortPacked sp{e.data(), sb::PackedLess{}};
        sp.sort(n);
        for (int i = 0; i < n; ++i) if ((int)(e[i] & 0xFFFFFFFFu) != a[i].id) { ++bad; break; }
      }
      ++cases;
      for (int i = 0; i < n; ++i)
        if (p[i] != a[i].id) { ++bad; if (bad < 5) std::printf("mismatch n=%d rep=%d at %d\n", n, rep, i); break; }
    }
  // the reference's three consecutive sorts (by x, then y, then z) on one array
  for (int rep = 0; rep < 2000; ++rep) {
    const int n = 17 + (int)(rng() % 112);
    struct Q { float x, y, z; int id; };
    std::vector<Q> a(n);
    std::vector<float> x(n), y(n), z(n);
    for (int i = 0; i < n; ++i) { x[i] = (float)(rng() % 9); y[i] = (float)(rng() % 5); z[i] = (float)(rng() % 3); a[i] = {x[i], y[i], z[i], i}; }
    std::sort(a.begin(), a.end(), [](const Q &p1, const Q &p2) { return p1.x < p2.x; });
    std::sort(a.begin(), a.end(), [](const Q &p1, const Q &p2) { return p1.y < p2.y; });
    std::sort(a.begin(), a.end(), [](const Q &p1, const Q &p2) { return p1.z < p2.z; });
    std::vector<int16_t> p(n);
    for (int i = 0; i < n; ++i) p[i] = (int16_t)i;
    { sb::StdSort<int16_t> s1(p.data(), x.data()); s1.sort(n); }
    { sb::StdSort<int16_t> s2(p.data(), y.data()); s2.sort(n); }
    { sb::StdSort<int16_t> s3(p.data(), z.data()); s3.sort(n); }
    ++cases;
    for (int i = 0; i < n; ++i)
      if (p[i] != a[i].id) { ++bad; break; }
  }
  std::printf("%ld cases, %ld mismatches\n%s\n", cases, bad, bad ? "FAILED" : "OK");
  return bad ? 1 : 0;
}

// Robust camera mean of computeCost (Derp.cpp:200-215):
//     keep = max(1, n-2); std::nth_element(SSDs, SSDs+keep, SSDs+n); cost = sum(SSDs[0..keep).second)
// std::nth_element leaves the kept elements in an implementation-defined order and the fp32 sum
// depends on that order in its last bit.  To stay bit-exact with the reference as built with GCC's
// libstdc++ (the reference's toolchain, Dockerfile: Ubuntu + g++), this header reproduces the
// permutation libstdc++'s introselect produces (median-of-3 pivot to first, unguarded partition,
// 2*lg(n) depth limit with heap-select fallback, final insertion sort on <=3 elements), operating
// on (first, second) pairs with std::pair's lexicographic operator<.
// tests/test_host_units.py checks it on the host against std::nth_element itself.
#pragma once

#include <limits>

#if defined(__CUDACC__)
#define DERP_SEL_HD __host__ __device__ __forceinline__
#else
#define DERP_SEL_HD inline
#endif

namespace derp {

// Element access is abstracted so that the same code runs on plain arrays (host unit test, device fallback in
// local memory) and on the strided per-thread slots in shared memory the cost kernels use: local-memory
// arrays get evicted from L1 by the texel gathers and every compare then waits on L2 (26 % of the sweep's
// stall samples before this change, profiles/README.md).
struct PairVal {
  float a, b;  // first (biased SSD), second (unbiased SSD)
};

struct ArrayPairs {  // two separate float arrays
  float* a;
  float* b;
  DERP_SEL_HD PairVal get(int i) const { return PairVal{a[i], b[i]}; }
  DERP_SEL_HD void set(int i, PairVal v) const {
    a[i] = v.a;
    b[i] = v.b;
  }
};

DERP_SEL_HD bool pairLess(float a0, float b0, float a1, float b1) {
  return a0 < a1 || (!(a1 < a0) && b0 < b1);
}
DERP_SEL_HD bool pairLess(PairVal x, PairVal y) { return pairLess(x.a, x.b, y.a, y.b); }

template <class V>
DERP_SEL_HD bool lessAt(const V& v, int i, int j) {
  return pairLess(v.get(i), v.get(j));
}

template <class V>
DERP_SEL_HD void swapAt(const V& v, int i, int j) {
  const PairVal x = v.get(i), y = v.get(j);
  v.set(i, y);
  v.set(j, x);
}

// std::__insertion_sort on [first, last)
template <class V>
DERP_SEL_HD void insertionSort(const V& v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    const PairVal val = v.get(i);
    if (pairLess(val, v.get(first))) {
      for (int k = i; k > first; --k) v.set(k, v.get(k - 1));
      v.set(first, val);
    } else {
      int last2 = i, next = i - 1;
      while (pairLess(val, v.get(next))) {
        v.set(last2, v.get(next));
        last2 = next;
        --next;
      }
      v.set(last2, val);
    }
  }
}

// std::__adjust_heap (max-heap under pairLess) on [first, first+len)
template <class V>
DERP_SEL_HD void adjustHeap(const V& v, int first, int holeIndex, int len, PairVal val) {
  const int topIndex = holeIndex;
  int secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (lessAt(v, first + secondChild, first + (secondChild - 1))) secondChild--;
    v.set(first + holeIndex, v.get(first + secondChild));
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    v.set(first + holeIndex, v.get(first + (secondChild - 1)));
    holeIndex = secondChild - 1;
  }
  // __push_heap
  int parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && pairLess(v.get(first + parent), val)) {
    v.set(first + holeIndex, v.get(first + parent));
    holeIndex = parent;
    parent = (holeIndex - 1) / 2;
  }
  v.set(first + holeIndex, val);
}

// std::__heap_select(first, middle, last)
template <class V>
DERP_SEL_HD void heapSelect(const V& v, int first, int middle, int last) {
  const int len = middle - first;
  if (len >= 2) {  // __make_heap
    int parent = (len - 2) / 2;
    while (true) {
      adjustHeap(v, first, parent, len, v.get(first + parent));
      if (parent == 0) break;
      parent--;
    }
  }
  for (int i = middle; i < last; ++i)
    if (lessAt(v, i, first)) {  // __pop_heap(first, middle, i)
      const PairVal val = v.get(i);
      v.set(i, v.get(first));
      adjustHeap(v, first, 0, len, val);
    }
}

// std::nth_element(v, v+nth, v+n) with libstdc++'s algorithm.
template <class V>
DERP_SEL_HD void nthElement(const V& v, int nth, int n) {
  if (n == 0 || nth == n) return;
  int first = 0, last = n;
  int depth = 0;  // 2 * floor(log2(n))
  for (int t = n; t > 1; t >>= 1) depth += 2;
  while (last - first > 3) {
    if (depth == 0) {
      heapSelect(v, first, nth + 1, last);
      swapAt(v, first, nth);
      return;
    }
    --depth;
    // __unguarded_partition_pivot
    const int mid = first + (last - first) / 2;
    {  // __move_median_to_first(first, first+1, mid, last-1)
      const int a = first + 1, b = mid, c = last - 1;
      const PairVal va = v.get(a), vb = v.get(b), vc = v.get(c);
      int m;
      if (pairLess(va, vb)) {
        if (pairLess(vb, vc)) m = b;
        else if (pairLess(va, vc)) m = c;
        else m = a;
      } else if (pairLess(va, vc)) {
        m = a;
      } else if (pairLess(vb, vc)) {
        m = c;
      } else {
        m = b;
      }
      swapAt(v, first, m);
    }
    int lo = first + 1, hi = last;
    const PairVal pivot = v.get(first);
    while (true) {
      while (pairLess(v.get(lo), pivot)) ++lo;
      --hi;
      while (pairLess(pivot, v.get(hi))) --hi;
      if (!(lo < hi)) break;
      swapAt(v, lo, hi);
      ++lo;
    }
    const int cut = lo;
    if (cut <= nth) first = cut;
    else last = cut;
  }
  insertionSort(v, first, last);
}

// cost numerator of computeCost: sum of the `keep` kept unbiased SSDs in nth_element's order.
template <class V>
DERP_SEL_HD float robustSum(const V& v, int n, int keep) {
  nthElement(v, keep, n);
  float cost = 0;
  for (int i = 0; i < keep; ++i) cost += v.get(i).b;
  return cost;
}

DERP_SEL_HD float robustSum(float* a, float* b, int n, int keep) { return robustSum(ArrayPairs{a, b}, n, keep); }

// ---- table-driven selection for 4 <= n <= 8 -----------------------------------------------------------
// nth_element's sequence of swaps depends only on the outcomes of its comparisons.  When the first keys are pairwise
// distinct, pairLess never looks at the second key and every outcome is determined by the relative order of the keys
// BY POSITION, i.e. by the permutation that sorts them.  So for small n the final arrangement can be looked up
// instead of executed: index = the permutation's inversion vector R_j = #{i < j : a_i > a_j} in the mixed radix
// sum_j R_j * j! (radix j+1 at digit j, independent of n, so slots >= n padded with +inf contribute 0), entry = the
// positions of the first `keep` elements after nth_element, in order, 3 bits each.  The table (46 224 entries for
// n = 4..8) is produced on the host by running nthElement above on every permutation.  The evaluation is
// branch-free: 28 compares, 7 multiply-adds, one table load, `keep` adds — against a few hundred data-dependent,
// divergent instructions for the general algorithm (13 % of the sweep's instructions, profiles/README.md).
// Ties or NaNs among the first keys (where the second key or the comparison order matters) return false and the
// caller runs the general algorithm.
constexpr int kSelTabMinN = 4, kSelTabMaxN = 8;
constexpr int kSelTabSize = 24 + 120 + 720 + 5040 + 40320;
DERP_SEL_HD int selTabOffset(int n) { return n == 4 ? 0 : (n == 5 ? 24 : (n == 6 ? 144 : (n == 7 ? 864 : 5904))); }

// K = number of slots examined (n <= K <= 8): the index weights j! do not depend on n and padded slots contribute
// nothing, so a warp whose lanes all have n <= 6 can run the 15-compare K = 6 instance instead of the 28-compare one.
template <int K, class V>
DERP_SEL_HD bool robustSumTable(const V& v, int n, int keep, const unsigned* __restrict__ tab, float* out) {
#if defined(__CUDA_ARCH__)
  const float inf = __int_as_float(0x7f800000);
#else
  const float inf = std::numeric_limits<float>::infinity();
#endif
  static_assert(K >= kSelTabMinN && K <= kSelTabMaxN, "slot count");
  float a[K];
#pragma unroll
  for (int i = 0; i < K; ++i) a[i] = i < n ? v.get(i).a : inf;
  bool bad = false;
  int idx = 0, fact = 1;
#pragma unroll
  for (int j = 1; j < K; ++j) {
    fact *= j;  // j!
    int r = 0;
#pragma unroll
    for (int i = 0; i < j; ++i) {
      r += (a[j] < a[i]) ? 1 : 0;
      bad |= (a[i] == a[j]) && (j < n);
    }
    bad |= !(a[j] == a[j]);
    idx += r * fact;
  }
  bad |= !(a[0] == a[0]);
  if (bad) return false;
#if defined(__CUDA_ARCH__)
  const unsigned entry = __ldg(tab + selTabOffset(n) + idx);
#else
  const unsigned entry = tab[selTabOffset(n) + idx];
#endif
  float cost = 0;
  for (int k = 0; k < keep; ++k) cost += v.get((int)((entry >> (3 * k)) & 7u)).b;
  *out = cost;
  return true;
}

template <class V>
DERP_SEL_HD bool robustSumTable(const V& v, int n, int keep, const unsigned* __restrict__ tab, float* out) {
  return robustSumTable<kSelTabMaxN>(v, n, keep, tab, out);
}

// Host: the table robustSumTable reads.  Every permutation of n distinct keys (by inversion-vector index) goes
// through nthElement; the second key carries the original position so the arrangement can be read back.
inline void buildSelectTable(unsigned* tab /* kSelTabSize entries */) {
  for (int n = kSelTabMinN; n <= kSelTabMaxN; ++n) {
    const int keep = n - 2;
    int nfact = 1;
    for (int j = 2; j <= n; ++j) nfact *= j;
    for (int idx = 0; idx < nfact; ++idx) {
      // decode R_j, then place element j so that exactly R_j of the earlier ones are greater
      double key[kSelTabMaxN];
      for (int j = 0; j < n; ++j) {
        const int radix = j + 1;
        int fj = 1;
        for (int t = 2; t <= j; ++t) fj *= t;
        const int r = (idx / fj) % radix;
        // earlier keys sorted ascending; the new key goes above (j - r) of them
        double sorted[kSelTabMaxN];
        for (int i = 0; i < j; ++i) sorted[i] = key[i];
        for (int x = 1; x < j; ++x) {
          const double t = sorted[x];
          int y = x - 1;
          while (y >= 0 && sorted[y] > t) {
            sorted[y + 1] = sorted[y];
            --y;
          }
          sorted[y + 1] = t;
        }
        const int below = j - r;
        const double lo = below == 0 ? (j == 0 ? 0.0 : sorted[0] - 1.0) : sorted[below - 1];
        const double hi = below == j ? (j == 0 ? 1.0 : sorted[j - 1] + 1.0) : sorted[below];
        key[j] = below == 0 ? lo : (below == j ? hi : 0.5 * (lo + hi));
      }
      // ranks as exactly representable floats
      float a[kSelTabMaxN], b[kSelTabMaxN];
      for (int i = 0; i < n; ++i) {
        int rank = 0;
        for (int k = 0; k < n; ++k) rank += key[k] < key[i];
        a[i] = (float)rank;
        b[i] = (float)i;
      }
      nthElement(ArrayPairs{a, b}, keep, n);
      unsigned entry = 0;
      for (int k = 0; k < keep; ++k) entry |= (unsigned)(int)b[k] << (3 * k);
      tab[selTabOffset(n) + idx] = entry;
    }
  }
}

}  // namespace derp

// Robust camera mean of computeCost (Derp.cpp:200-215):
//     keep = max(1, n-2); std::nth_element(SSDs, SSDs+keep, SSDs+n); cost = sum(SSDs[0..keep).second)
// std::nth_element leaves the kept elements in an implementation-defined order and the fp32 sum
// depends on that order in its last bit.  To stay bit-exact with the reference as built with GCC's
// libstdc++ (the reference's toolchain, Dockerfile: Ubuntu + g++), this header reproduces the
// permutation libstdc++'s introselect produces (median-of-3 pivot to first, unguarded partition,
// 2*lg(n) depth limit with heap-select fallback, final insertion sort on <=3 elements), operating
// on (first, second) pairs with std::pair's lexicographic operator<.
// tests/test_select.py checks it on the host against std::nth_element itself.
#pragma once

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

}  // namespace derp

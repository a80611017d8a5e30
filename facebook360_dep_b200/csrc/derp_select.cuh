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

struct PairRef {
  float* a;  // first  (biased SSD)
  float* b;  // second (unbiased SSD)
};

DERP_SEL_HD bool pairLess(float a0, float b0, float a1, float b1) {
  return a0 < a1 || (!(a1 < a0) && b0 < b1);
}

DERP_SEL_HD bool lessAt(const PairRef& v, int i, int j) { return pairLess(v.a[i], v.b[i], v.a[j], v.b[j]); }

DERP_SEL_HD void swapAt(const PairRef& v, int i, int j) {
  const float ta = v.a[i], tb = v.b[i];
  v.a[i] = v.a[j];
  v.b[i] = v.b[j];
  v.a[j] = ta;
  v.b[j] = tb;
}

// std::__insertion_sort on [first, last)
DERP_SEL_HD void insertionSort(const PairRef& v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    const float va = v.a[i], vb = v.b[i];
    if (pairLess(va, vb, v.a[first], v.b[first])) {
      for (int k = i; k > first; --k) {
        v.a[k] = v.a[k - 1];
        v.b[k] = v.b[k - 1];
      }
      v.a[first] = va;
      v.b[first] = vb;
    } else {
      int last2 = i, next = i - 1;
      while (pairLess(va, vb, v.a[next], v.b[next])) {
        v.a[last2] = v.a[next];
        v.b[last2] = v.b[next];
        last2 = next;
        --next;
      }
      v.a[last2] = va;
      v.b[last2] = vb;
    }
  }
}

// std::__adjust_heap (max-heap under pairLess) on [first, first+len)
DERP_SEL_HD void adjustHeap(const PairRef& v, int first, int holeIndex, int len, float va, float vb) {
  const int topIndex = holeIndex;
  int secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (lessAt(v, first + secondChild, first + (secondChild - 1))) secondChild--;
    v.a[first + holeIndex] = v.a[first + secondChild];
    v.b[first + holeIndex] = v.b[first + secondChild];
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    v.a[first + holeIndex] = v.a[first + (secondChild - 1)];
    v.b[first + holeIndex] = v.b[first + (secondChild - 1)];
    holeIndex = secondChild - 1;
  }
  // __push_heap
  int parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && pairLess(v.a[first + parent], v.b[first + parent], va, vb)) {
    v.a[first + holeIndex] = v.a[first + parent];
    v.b[first + holeIndex] = v.b[first + parent];
    holeIndex = parent;
    parent = (holeIndex - 1) / 2;
  }
  v.a[first + holeIndex] = va;
  v.b[first + holeIndex] = vb;
}

// std::__heap_select(first, middle, last)
DERP_SEL_HD void heapSelect(const PairRef& v, int first, int middle, int last) {
  const int len = middle - first;
  if (len >= 2) {  // __make_heap
    int parent = (len - 2) / 2;
    while (true) {
      adjustHeap(v, first, parent, len, v.a[first + parent], v.b[first + parent]);
      if (parent == 0) break;
      parent--;
    }
  }
  for (int i = middle; i < last; ++i)
    if (lessAt(v, i, first)) {  // __pop_heap(first, middle, i)
      const float va = v.a[i], vb = v.b[i];
      v.a[i] = v.a[first];
      v.b[i] = v.b[first];
      adjustHeap(v, first, 0, len, va, vb);
    }
}

// std::nth_element(v, v+nth, v+n) with libstdc++'s algorithm.
DERP_SEL_HD void nthElement(const PairRef& v, int nth, int n) {
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
      if (lessAt(v, a, b)) {
        if (lessAt(v, b, c)) swapAt(v, first, b);
        else if (lessAt(v, a, c)) swapAt(v, first, c);
        else swapAt(v, first, a);
      } else if (lessAt(v, a, c)) {
        swapAt(v, first, a);
      } else if (lessAt(v, b, c)) {
        swapAt(v, first, c);
      } else {
        swapAt(v, first, b);
      }
    }
    int lo = first + 1, hi = last;
    const float pa = v.a[first], pb = v.b[first];
    while (true) {
      while (pairLess(v.a[lo], v.b[lo], pa, pb)) ++lo;
      --hi;
      while (pairLess(pa, pb, v.a[hi], v.b[hi])) --hi;
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
DERP_SEL_HD float robustSum(float* a, float* b, int n, int keep) {
  PairRef v{a, b};
  nthElement(v, keep, n);
  float cost = 0;
  for (int i = 0; i < keep; ++i) cost += b[i];
  return cost;
}

}  // namespace derp

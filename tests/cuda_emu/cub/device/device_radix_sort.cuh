// tests/cuda_emu: host stand-in for cub::DeviceRadixSort::SortPairs (stable, ascending, keys compared on bits [begin_bit, end_bit))
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>
#include <cuda_runtime.h>
namespace cub {
struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, int n, int begin_bit = 0, int end_bit = sizeof(K) * 8,
                               cudaStream_t = nullptr) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    const unsigned long long mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1ull);
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      return (((unsigned long long)(unsigned)kin[a] & mask) >> begin_bit) < (((unsigned long long)(unsigned)kin[b] & mask) >> begin_bit); });
    for (int i = 0; i < n; i++) { kout[i] = kin[order[i]]; vout[i] = vin[order[i]]; }
    return cudaSuccess;
  }
};
}  // namespace cub

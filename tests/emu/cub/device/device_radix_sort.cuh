// tests/emu/cub/device/device_radix_sort.cuh -- host stand-in for cub::DeviceRadixSort::SortPairs as csrc/sparse.cu
// calls it: a STABLE sort of (key, value) pairs by the key bits [begin_bit, end_bit).  Test infrastructure only.
#pragma once
#include <algorithm>
#include <cstddef>
#include <numeric>
#include <vector>
namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static int SortPairs(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, int n, int begin_bit, int end_bit, void* = nullptr) {
        if (!tmp) { bytes = 16; return 0; }
        const unsigned long long mask = end_bit - begin_bit >= 64 ? ~0ull : ((1ull << (end_bit - begin_bit)) - 1ull);
        auto key = [&](int i) { return ((unsigned long long)kin[i] >> begin_bit) & mask; };
        std::vector<int> order((size_t)n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
        std::vector<K> ks((size_t)n); std::vector<V> vs((size_t)n);
        for (int i = 0; i < n; ++i) { ks[(size_t)i] = kin[order[(size_t)i]]; vs[(size_t)i] = vin[order[(size_t)i]]; }
        for (int i = 0; i < n; ++i) { kout[i] = ks[(size_t)i]; vout[i] = vs[(size_t)i]; }
        return 0;
    }
};
}  // namespace cub

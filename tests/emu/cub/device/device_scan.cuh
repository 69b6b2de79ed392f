// tests/emu/cub/device/device_scan.cuh -- host stand-in for the two cub::DeviceScan calls of csrc/sparse.cu
// (same two-phase calling convention: a null temporary buffer asks for its size).  Test infrastructure only.
#pragma once
#include <cstddef>
namespace cub {
struct DeviceScan {
    template <class In, class Out>
    static int InclusiveSum(void* tmp, size_t& bytes, In in, Out out, int n, void* = nullptr) {
        if (!tmp) { bytes = 16; return 0; }
        long long run = 0;
        for (int i = 0; i < n; ++i) { run += in[i]; out[i] = (int)run; }
        return 0;
    }
    template <class In, class Out>
    static int ExclusiveSum(void* tmp, size_t& bytes, In in, Out out, int n, void* = nullptr) {
        if (!tmp) { bytes = 16; return 0; }
        long long run = 0;
        for (int i = 0; i < n; ++i) { const auto v = in[i]; out[i] = (int)run; run += v; }
        return 0;
    }
};
}  // namespace cub

// Shared plumbing of the C-ABI translation units: error string, launch counter, status helpers.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/smcnuts_b200.h"
#include "common.cuh"
#include "models.cuh"

namespace smcb {

std::string& last_error_ref();
extern std::atomic<long long> g_launches;

inline int fail(const char* where, const char* what) {
    last_error_ref() = std::string(where) + ": " + what;
    return -1;
}
inline int check_launch(const char* where) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(where, cudaGetErrorString(e));
    return 0;
}
#define SMCB_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) return ::smcb::fail(#call, cudaGetErrorString(e_)); \
    } while (0)
#define SMCB_REQUIRE(cond, msg)                          \
    do {                                                 \
        if (!(cond)) return ::smcb::fail(__func__, msg); \
    } while (0)

struct Model {
    ModelDesc desc;
    double* d_data;
};

int device_sm_count();

// grid for a grid-stride elementwise / reduction kernel: a multiple of the SM count
inline int stride_grid(long long n, int threads, int per_sm) {
    long long want = (n + threads - 1) / threads;
    long long cap = (long long)device_sm_count() * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace smcb

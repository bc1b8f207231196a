// Host-side helpers shared by the C-ABI translation units: error reporting, TMA descriptor
// construction through the driver entry point (no link-time libcuda dependency), launch counting.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

namespace pk {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
int sm_count();

// Encode a tiled bf16 tensor map (rank 2..5), 128B swizzle, zero OOB fill.
//   dims[0] is the contiguous dimension; strides_bytes[i] is the stride of dims[i+1].
bool make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box);

#define PK_CHECK(cond, ...)      \
  do {                           \
    if (!(cond)) {               \
      pk::set_error(__VA_ARGS__); \
      return 1;                  \
    }                            \
  } while (0)

#define PK_LAUNCH_CHECK(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      pk::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));         \
      return 2;                                                                      \
    }                                                                                \
    pk::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
  } while (0)

}  // namespace pk

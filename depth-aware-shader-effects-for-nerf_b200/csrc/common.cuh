// Shared host/device helpers for libnerfw_sm100.so.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "nerfw.h"

namespace nerfw {

// thread-local error text returned by nerfw_last_error()
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define NERFW_REQUIRE(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      ::nerfw::set_error(__VA_ARGS__);        \
      return NERFW_EINVAL;                    \
    }                                         \
  } while (0)

#define NERFW_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::nerfw::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return NERFW_ECUDA;                                                             \
    }                                                                                 \
  } while (0)

// after a <<<>>> launch
#define NERFW_LAUNCHED()                      \
  do {                                        \
    ::nerfw::count_launch();                  \
    NERFW_CUDA(cudaGetLastError());           \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count();  // cached multiprocessor count of the current device

// True the first time it is called for the current device with this (per call site) mask: function attributes such as the
// dynamic shared-memory limit are per device, so launchers set them once per device rather than once per process.
bool first_use_on_device(unsigned long long& mask);

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nerfw

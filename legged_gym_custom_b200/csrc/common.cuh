// Shared host-side helpers of libb200gym.so (error text, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/b200gym.h"

void b200_set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)   \
  do {                              \
    if (!(cond)) {                  \
      b200_set_error(__VA_ARGS__);  \
      return -1;                    \
    }                               \
  } while (0)

#define B200_CHECK_LAUNCH(name)                                             \
  do {                                                                      \
    cudaError_t err_ = cudaGetLastError();                                  \
    if (err_ != cudaSuccess) {                                              \
      b200_set_error("%s: %s", name, cudaGetErrorString(err_));             \
      return (int)err_;                                                     \
    }                                                                       \
  } while (0)

struct B200Env {
  B200EnvParams p;
  int device;
  unsigned long long* phase_trace;   // device [ceil(num_envs / 8)][8] or NULL (b200_env_set_phase_trace)
  int prefetch_history;       // b200_env_set_prefetch: L2 prefetch of the history rows at kernel entry
  int force_generic_layout;   // tests: run the layout-generic kernel variant even for the go2 layout
};

// Host-side helpers shared by all translation units of libctclip_b200.so:
// error reporting behind the C ABI, launch checks, TMA tensor-map encoding through the
// driver entry point (resolved lazily so the library loads on a machine without libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define CTCLIP_OK 0
#define CTCLIP_ERR_ARG 1
#define CTCLIP_ERR_CUDA 2
#define CTCLIP_ERR_DRIVER 3
#define CTCLIP_ERR_UNSUPPORTED 4

namespace ctb {

void set_error(const char* fmt, ...);  // defined in capi.cu
int num_sms();                         // cached cudaDevAttrMultiProcessorCount of current device

#define CTB_CHECK_ARG(cond, ...)    \
  do {                              \
    if (!(cond)) {                  \
      ctb::set_error(__VA_ARGS__);  \
      return CTCLIP_ERR_ARG;        \
    }                               \
  } while (0)

#define CTB_CUDA(expr)                                                                 \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ctb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CTCLIP_ERR_CUDA;                                                          \
    }                                                                                  \
  } while (0)

#define CTB_LAUNCH_CHECK() CTB_CUDA(cudaGetLastError())

// Encode a 2-D tiled tensor map. dims/strides follow cuTensorMapEncodeTiled conventions:
// inner = contiguous dimension (elements), pitch_bytes = byte stride between consecutive outer rows.
int encode_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base,
                   uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner,
                   uint32_t box_outer, CUtensorMapSwizzle swz);

// rank <= 5 tiled tensor map, dims[0] = contiguous dimension, strides_bytes[i] = byte stride of dims[i+1] (rank-1 entries)
int encode_tmap_nd(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace ctb

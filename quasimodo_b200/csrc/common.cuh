// common.cuh -- shared scaffolding for the quasimodo_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/quasimodo_b200.h"

#define QM_WARP 32

struct qm_scratch {
    void  *ptr = nullptr;
    size_t cap = 0;
};

struct qm_ctx {
    int device = -1;
    int sm_count = 0;
    std::string err;
    // scratch arenas grown on demand (never shrunk); index = purpose
    qm_scratch scratch[8];
    cudaStream_t own_stream = nullptr;
};

// grow-only scratch; contents are NOT preserved across a growth
int qm_scratch_reserve(qm_ctx *ctx, int which, size_t bytes, void **out);
int qm_fail(qm_ctx *ctx, int code, const char *fmt, ...);

#define QM_CUDA(ctx, call)                                                                         \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return qm_fail((ctx), QM_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,           \
                           cudaGetErrorString(e__));                                               \
    } while (0)

static __device__ __forceinline__ int qm_lane() { return threadIdx.x & 31; }

// Internal (non-ABI) helpers shared by the translation units behind include/rg_b200.h.
#pragma once
#include <cuda_runtime.h>

int rg_fail(const char* fmt, ...);   // records the thread-local error string, returns 1
void rg_count_launch(int n);         // bookkeeping behind rg_launch_count()

#define RG_CU(expr)                                                                            \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return rg_fail("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

// Stream-ordered scratch (cudaMallocAsync) is served from the device's default pool; by default the pool
// hands memory back to the driver at every synchronisation, which makes the next allocation cost ~0.3 ms.
void rg_keep_mempool();

// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tapqir_b200.h"

namespace tq {

void set_error(const char* fmt, ...);
int cuda_status(cudaError_t err, const char* what);

// Number of SMs of the current device (cached per process; 148 on B200).
int sm_count();

#define TQ_CHECK_ARG(cond, msg)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::tq::set_error("%s: %s", __func__, msg); \
            return TQ_ERR_ARG;                  \
        }                                       \
    } while (0)

#define TQ_LAUNCH_CHECK(what)                                    \
    do {                                                         \
        int _st = ::tq::cuda_status(cudaGetLastError(), what);   \
        if (_st != TQ_OK) return _st;                            \
    } while (0)

// Location of minibatch unit u in the device store.
struct UnitIndex {
    int ni, fi, c;     // minibatch coordinates
    int64_t patch;     // linear (AOI, frame, channel) index into the store
    int aoi;           // store AOI index
};

__device__ __forceinline__ UnitIndex locate_unit(int64_t u, int fb, int C, int F,
                                                 const int32_t* __restrict__ ndx,
                                                 const int32_t* __restrict__ fdx) {
    UnitIndex r;
    r.c = (int)(u % C);
    const int64_t nf = u / C;
    r.fi = (int)(nf % fb);
    r.ni = (int)(nf / fb);
    r.aoi = ndx ? ndx[r.ni] : r.ni;
    const int f = fdx ? fdx[r.fi] : r.fi;
    r.patch = ((int64_t)r.aoi * F + f) * C + r.c;
    return r;
}

// same for a unit index known to fit 32 bits (unsigned 32-bit divisions; none at all when C == 1)
__device__ __forceinline__ UnitIndex locate_unit32(uint32_t u, int fb, int C, int F,
                                                   const int32_t* __restrict__ ndx,
                                                   const int32_t* __restrict__ fdx) {
    UnitIndex r;
    uint32_t nf = u;
    r.c = 0;
    if (C != 1) { nf = u / (uint32_t)C; r.c = (int)(u - nf * (uint32_t)C); }
    r.ni = (int)(nf / (uint32_t)fb);
    r.fi = (int)(nf - (uint32_t)r.ni * (uint32_t)fb);
    r.aoi = ndx ? ndx[r.ni] : r.ni;
    const int f = fdx ? fdx[r.fi] : r.fi;
    r.patch = ((int64_t)r.aoi * F + f) * C + r.c;
    return r;
}

}  // namespace tq

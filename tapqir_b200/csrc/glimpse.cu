// Ingestion of raw movie frames (row N4 of SURVEY.md section 8f): AOI cropping and the camera-offset
// histogram, straight from the bytes of the .glimpse files.
//
// Replaces the per-frame / per-AOI Python loop of imscroll/glimpse_reader.py:354-381 and the frame decoding
// of :168-186 (np.fromfile(">i2") + 2**15).  Integer / index work, HBM-bound: every frame byte is read once
// (coalesced 16-bit loads, byte swap in registers), every patch pixel written once; the histogram goes through
// shared-memory-free global atomics on a 65536-entry table (the offset region is a few thousand pixels/frame).
#include "common.cuh"

namespace tq {

// big-endian int16 + 2^15  ->  value in [0, 65535]          glimpse_reader.py:181-186
__device__ __forceinline__ uint32_t decode_pixel(uint16_t raw) {
    const uint16_t sw = (uint16_t)((raw << 8) | (raw >> 8));
    return (uint32_t)((int32_t)(int16_t)sw + 32768);
}

// round half to even, like Python's round() on a float                 glimpse_reader.py:365-366
__device__ __forceinline__ int round_half_even(double v) { return (int)rint(v); }

// one block per (AOI, frame of the chunk): P*P pixels, threads stride over them
__global__ void crop_aois_kernel(const uint16_t* __restrict__ frames, int H, int W, int Fc, int f0,
                                 const double* __restrict__ aoi_xy, const double* __restrict__ drift, int N, int F,
                                 int P, uint16_t* __restrict__ patches, double* __restrict__ target_xy,
                                 int* __restrict__ status) {
    const int n = blockIdx.x, fl = blockIdx.y;
    const int f = f0 + fl;
    if (n >= N || fl >= Fc || f >= F) return;
    // raw_target_xy = aoiinfo[x, y] + cumdrift[dx, dy]                 :338-341
    const double rx = aoi_xy[2 * n + 0] + drift[2 * f + 0];
    const double ry = aoi_xy[2 * n + 1] + drift[2 * f + 1];
    const double half = 0.5 * (double)(P - 1);
    const int shiftx = round_half_even(rx - half), shifty = round_half_even(ry - half);
    if (threadIdx.x == 0) {
        target_xy[((int64_t)n * F + f) * 2 + 0] = rx - (double)shiftx;
        target_xy[((int64_t)n * F + f) * 2 + 1] = ry - (double)shifty;
        // numpy would wrap negative starts and truncate at the far edge; the reference's later asserts assume neither
        if (shiftx < 0 || shifty < 0 || shiftx + P > W || shifty + P > H) atomicOr(status, 1);
    }
    if (shiftx < 0 || shifty < 0 || shiftx + P > W || shifty + P > H) return;
    const uint16_t* img = frames + (int64_t)fl * H * W;
    uint16_t* out = patches + ((int64_t)n * F + f) * P * P;
    for (int p = threadIdx.x; p < P * P; p += blockDim.x) {
        const int r = p / P, c = p - r * P;
        out[p] = (uint16_t)decode_pixel(img[(int64_t)(shifty + r) * W + shiftx + c]);
    }
}

// counts[value] += occurrences in the offset_P x offset_P region of every frame of the chunk        :356-362
__global__ void offset_hist_kernel(const uint16_t* __restrict__ frames, int H, int W, int Fc, int ox, int oy, int oP,
                                   unsigned long long* __restrict__ counts) {
    const int fl = blockIdx.y;
    if (fl >= Fc) return;
    const uint16_t* img = frames + (int64_t)fl * H * W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < oP * oP; p += gridDim.x * blockDim.x) {
        const int r = p / oP, c = p - r * oP;
        atomicAdd(&counts[decode_pixel(img[(int64_t)(oy + r) * W + ox + c])], 1ull);
    }
}

}  // namespace tq

using namespace tq;

extern "C" int tq_crop_aois(const void* frames_raw, int H, int W, int Fc, int f0, const double* aoi_xy, const double* drift,
                            int N, int F, int P, void* patches, double* target_xy, int* status, void* stream) {
    TQ_CHECK_ARG(H > 0 && W > 0 && Fc >= 0 && f0 >= 0 && N >= 0 && F >= 0 && P > 0 && P <= H && P <= W, "bad sizes");
    TQ_CHECK_ARG(f0 + Fc <= F, "frame chunk exceeds the movie");
    if (Fc == 0 || N == 0) return TQ_OK;
    TQ_CHECK_ARG(frames_raw && aoi_xy && drift && patches && target_xy && status, "NULL pointer");
    TQ_CHECK_ARG(Fc <= 65535, "at most 65535 frames per call");
    crop_aois_kernel<<<dim3(N, Fc), 64, 0, (cudaStream_t)stream>>>((const uint16_t*)frames_raw, H, W, Fc, f0, aoi_xy, drift, N, F, P,
                                                                    (uint16_t*)patches, target_xy, status);
    TQ_LAUNCH_CHECK("crop_aois_kernel launch");
    return TQ_OK;
}

extern "C" int tq_offset_hist(const void* frames_raw, int H, int W, int Fc, int offset_x, int offset_y, int offset_P,
                              void* counts, void* stream) {
    TQ_CHECK_ARG(H > 0 && W > 0 && Fc >= 0 && offset_P > 0, "bad sizes");
    TQ_CHECK_ARG(offset_x >= 0 && offset_y >= 0 && offset_x + offset_P <= W && offset_y + offset_P <= H, "offset region outside the frame");
    if (Fc == 0) return TQ_OK;
    TQ_CHECK_ARG(frames_raw && counts, "NULL pointer");
    TQ_CHECK_ARG(Fc <= 65535, "at most 65535 frames per call");
    const int blocks = (offset_P * offset_P + 255) / 256;
    offset_hist_kernel<<<dim3(blocks < 64 ? blocks : 64, Fc), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)frames_raw, H, W, Fc, offset_x,
                                                                                             offset_y, offset_P, (unsigned long long*)counts);
    TQ_LAUNCH_CHECK("offset_hist_kernel launch");
    return TQ_OK;
}

/*
 * tapqir_b200 -- C ABI of the B200-native cosmos SVI hot path.
 *
 * The reference (gelles-brandeis/tapqir) is pure Python and has no FFI of its own; the entry
 * points below are what a binding for this path would call, one per reference function /
 * method they replace (cited as file:line in the reference tree).  INTEGRATION.md shows the
 * ctypes stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless marked "host";
 *   - nothing is allocated or freed inside; work is enqueued on `stream` (a cudaStream_t,
 *     0 = legacy default stream) and the call returns without synchronising;
 *   - `dtype` selects the arithmetic type of all floating-point buffers of the call:
 *     TQ_F32 (production) or TQ_F64 (the reference CLI's dtype, main.py:428);
 *   - return value: 0 on success; TQ_ERR_* otherwise, with tq_last_error() giving the text.
 *     The Python layer turns TQ_ERR_OOM into CudaOutOfMemoryError (exceptions.py:33-39) and
 *     anything else into RuntimeError;
 *   - re-entrant per stream, no global mutable state except the last-error string (thread local).
 *
 * Patch indexing: a "unit" is one (AOI, frame, channel) PxP patch.  Minibatch unit
 *   u = (ni * fb + fi) * C + c   reads dataset patch (ndx[ni], fdx[fi], c);  ndx / fdx == NULL
 *   mean the identity (0..nb-1 / 0..fb-1).  Per-unit arrays are laid out (nb, fb, C) and
 *   per-spot arrays (K, nb, fb, C), K = 2.
 */
#ifndef TAPQIR_B200_H
#define TAPQIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQ_K 2  /* spots per patch (cosmos K) */
#define TQ_M 4  /* enumerated spot-presence configurations, 2^K */

enum { TQ_F32 = 0, TQ_F64 = 1 };
enum { TQ_PIX_U16 = 0, TQ_PIX_F32 = 1, TQ_PIX_F64 = 2 };
enum {
    TQ_OK = 0,
    TQ_ERR_ARG = 1,      /* bad argument (shape, dtype, NULL) */
    TQ_ERR_CUDA = 2,     /* CUDA runtime error other than OOM */
    TQ_ERR_OOM = 3,      /* cudaErrorMemoryAllocation */
    TQ_ERR_UNSUPPORTED = 4
};

/* Device-resident dataset view + minibatch selection (host struct, passed by pointer).
 * Replaces CosmosDataset.fetch (utils/dataset.py:140-151) and OffsetData (dataset.py:18-37). */
typedef struct {
    int32_t nb, fb, C;          /* minibatch AOIs, minibatch frames, channels */
    int32_t F;                  /* frames in the stored dataset (row stride of the AOI axis) */
    int32_t P;                  /* patch edge, 2 <= P <= 32 */
    int32_t O;                  /* offset bins */
    int32_t pixtype;            /* TQ_PIX_* of `pixels` */
    const int32_t* ndx;         /* (nb,) AOI indices into the store, or NULL */
    const int32_t* fdx;         /* (fb,) frame indices, or NULL */
    const void* pixels;         /* (Nt, F, C, P, P) */
    const void* xy;             /* (Nt, F, C, 2) target x, y; dtype */
    const uint8_t* is_ontarget; /* (Nt,) */
    const uint8_t* mask;        /* (Nt,) */
    const void* offset_samples; /* (O,) dtype */
    const void* offset_logits;  /* (O,) dtype: log of eps-clamped weights (dataset.py:27-29) */
} tq_patch_view;

int tq_version(void);
const char* tq_last_error(void);

/* gaussian_spots (distributions/util.py:15-64).
 * height,width,x,y: (K, U); target_xy: (U, 2); m: (K, U) or NULL; out: (U, K, P, P). */
int tq_gaussian_spots(int dtype, int64_t U, int P, const void* height, const void* width,
                      const void* x, const void* y, const void* target_xy, const void* m,
                      void* out, void* stream);

/* KSMOGN.log_prob (distributions/ksmogn.py:187-238) for NM spot-presence configurations.
 * height,width,x,y: (K, U); background: (U,); gain: device scalar; mcfg: (NM, K) device table
 * (NM in {1, 4}); logp out: (NM, U). */
int tq_ksmogn_fwd(int dtype, const tq_patch_view* view, const void* height, const void* width,
                  const void* x, const void* y, const void* background, const void* gain,
                  const void* mcfg, int NM, void* logp, void* stream);

/* Reverse mode of the above (what autograd derives in the reference): given W = dLoss/dlogp
 * (NM, U) writes g_height.. g_y (K, U), g_background (U,) and g_rate (U,), the per-unit
 * derivative w.r.t. 1/gain (sum it and multiply by -1/gain^2 for d/dgain).  logp may be NULL. */
int tq_ksmogn_fwd_bwd(int dtype, const tq_patch_view* view, const void* height, const void* width,
                      const void* x, const void* y, const void* background, const void* gain,
                      const void* mcfg, int NM, const void* W, void* logp, void* g_height,
                      void* g_width, void* g_x, void* g_y, void* g_background, void* g_rate,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAPQIR_B200_H */

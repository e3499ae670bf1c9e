/* hdiff_b200 LAB library (libhdiff_b200_lab.so, `python -m hdiff_b200.build --lab`): hardware probes and timing
 * experiments used while designing the kernels (scripts/probe_*.py, scripts/conv_clock.py; results in profiles/).
 * NOT part of the product ABI: libhdiff_b200.so exports none of these, and is compiled without -DHDIFF_LAB, so the
 * HDIFF_CONV_DBG timing modes do not exist in it.  The lab library contains the whole product library plus these. */
#ifndef HDIFF_B200_LAB_H
#define HDIFF_B200_LAB_H
#include "hdiff_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* timing experiments (HDIFF_CONV_DBG=4): (clock64, globaltimer ns) at the start / end of CTA 0 of the last hd_conv_tc launch */
int hd_conv_dbg_read(long long* out4);
/* ---- hardware probe (scripts/probe_shift.py): tcgen05 A operand starting at an arbitrary 128-byte row of a swizzled box ---- */
int hd_probe_shift(const void* x, const void* w, float* out, int shift, int mode, hd_stream_t stream);
/* ---- hardware probe (scripts/probe_pair.py): C[256][N] = A[256][K] B[N][K]^T by a CTA pair; mode 1 = each CTA on its own
 *      (tcgen05.mma.cta_group::1), mode 2 = one M = 256 MMA stream for the pair (cta_group::2, each CTA holds half of B);
 *      out accumulates `reps` identical products, cycles[2] = clock64 span of the MMA sequence per CTA ---- */
int hd_probe_pair(const void* a, const void* b, float* out, int N, int K, int mode, int reps, long long* cycles, int shift,
                  int fill, int cper, int nclusters, int ring, hd_stream_t stream);
/* how far the issuing thread can run ahead of the tensor pipe: `groups` x (12 unrolled MMAs + `gap` idle cycles) */
int hd_probe_queue(const void* a, const void* b, int N, int groups, int gap, int issuers, int second_warp, long long* cycles,
                   hd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif

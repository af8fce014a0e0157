/* osb_synth.c -- exports the synthetic stack generator (include/opensot_b200_synth.h) from
 * libopensot_b200.so for bench.py and the tests. Pure host C, no CUDA. */
#include "opensot_b200_synth.h"

int osb_synth_shape(const char* name, double eps_factor, osb_stack_desc* d)
{
    return synth_shape(name, eps_factor, d);
}

int osb_synth_fill(const osb_stack_desc* d, uint64_t seed, uint64_t first, int32_t count, double tight,
                   double* blobs)
{
    osb_layout L;
    if (osb_blob_layout(d, &L) != OSB_SUCCESS) return OSB_E_INVALID;
    for (int32_t i = 0; i < count; ++i)
        synth_fill(d, &L, seed, first + (uint64_t)i, tight, blobs + (size_t)i * L.stride);
    return OSB_SUCCESS;
}

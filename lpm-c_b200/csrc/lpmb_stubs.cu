// lpmb_stubs.cu -- entry points declared in include/lpmb200.h whose kernels are not built yet.
// Each fails loudly (LPMB_ERR_UNSUPPORTED); none falls back to the CPU.
#include "lpmb_internal.cuh"

#define LPMB_STUB(name)                                               \
    lpmb_set_error(name ": not implemented in this build");           \
    return LPMB_ERR_UNSUPPORTED

extern "C" int lpmb_dist_unique_id(void *) { LPMB_STUB("lpmb_dist_unique_id"); }
extern "C" int lpmb_dist_init(lpmb_ctx *, const void *, int, int) { LPMB_STUB("lpmb_dist_init"); }
extern "C" int lpmb_dist_set_slab(lpmb_ctx *, long long, long long, long long, int, int) { LPMB_STUB("lpmb_dist_set_slab"); }

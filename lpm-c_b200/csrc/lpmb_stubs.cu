// lpmb_stubs.cu -- entry points declared in include/lpmb200.h whose kernels are not built yet.
// Each fails loudly (LPMB_ERR_UNSUPPORTED); none falls back to the CPU.
#include "lpmb_internal.cuh"

#define LPMB_STUB(name)                                               \
    lpmb_set_error(name ": not implemented in this build");           \
    return LPMB_ERR_UNSUPPORTED

extern "C" int lpmb_set_neighbors(lpmb_ctx *, const int *, const int *) { LPMB_STUB("lpmb_set_neighbors"); }
extern "C" int lpmb_build_topology(lpmb_ctx *, double, double) { LPMB_STUB("lpmb_build_topology"); }
extern "C" int lpmb_fd_stiffness(lpmb_ctx *, int) { LPMB_STUB("lpmb_fd_stiffness"); }
extern "C" int lpmb_calc_kntv(lpmb_ctx *, const double *, int) { LPMB_STUB("lpmb_calc_kntv"); }
extern "C" int lpmb_compute_dl(lpmb_ctx *) { LPMB_STUB("lpmb_compute_dl"); }
extern "C" int lpmb_bond_force(lpmb_ctx *, int, int) { LPMB_STUB("lpmb_bond_force"); }
extern "C" int lpmb_switch_state(lpmb_ctx *, int) { LPMB_STUB("lpmb_switch_state"); }
extern "C" int lpmb_update_rr(lpmb_ctx *, double *, double *) { LPMB_STUB("lpmb_update_rr"); }
extern "C" int lpmb_update_damage(lpmb_ctx *, int, int *, int *, int) { LPMB_STUB("lpmb_update_damage"); }
extern "C" int lpmb_update_crack(lpmb_ctx *) { LPMB_STUB("lpmb_update_crack"); }
extern "C" int lpmb_newton_iteration(lpmb_ctx *, int, int, double, double, int, int *, double *) { LPMB_STUB("lpmb_newton_iteration"); }
extern "C" int lpmb_dist_unique_id(void *) { LPMB_STUB("lpmb_dist_unique_id"); }
extern "C" int lpmb_dist_init(lpmb_ctx *, const void *, int, int) { LPMB_STUB("lpmb_dist_init"); }
extern "C" int lpmb_dist_set_slab(lpmb_ctx *, long long, long long, long long, int, int) { LPMB_STUB("lpmb_dist_set_slab"); }

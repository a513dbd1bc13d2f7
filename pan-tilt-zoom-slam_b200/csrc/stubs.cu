// stubs.cu - entry points declared in include/ptzba.h whose implementation file is not linked yet.
#include "common.h"
#define NOT_YET(ctx) ptzba_fail((ctx), PTZBA_ERR_STATE, "%s: not implemented in this build", __func__)
void ptzba_comm_release(ptzba_ctx*) {}
extern "C" int ptzba_ekf_update(ptzba_ctx* ctx, const ptzba_ekf_params*, int, double*, double*, double*, double*, int, const double*, const int32_t*, int32_t*) { return NOT_YET(ctx); }
extern "C" int ptzba_ekf_batch_create(ptzba_ctx* ctx, const ptzba_ekf_params*, int, int, int, const double*, const double*, ptzba_ekf_batch**) { return NOT_YET(ctx); }
extern "C" void ptzba_ekf_batch_destroy(ptzba_ekf_batch*) {}
extern "C" int ptzba_ekf_batch_step(ptzba_ekf_batch*, int, const double*, const int32_t*, const int32_t*, int32_t*) { return PTZBA_ERR_STATE; }
extern "C" int ptzba_ekf_batch_get(ptzba_ekf_batch*, double*, double*, double*) { return PTZBA_ERR_STATE; }
extern "C" int ptzba_ekf_batch_get_cov(ptzba_ekf_batch*, int, double*) { return PTZBA_ERR_STATE; }
extern "C" int ptzba_comm_unique_id(ptzba_ctx* ctx, void*) { return NOT_YET(ctx); }
extern "C" int ptzba_comm_init(ptzba_ctx* ctx, const void*, int, int) { return NOT_YET(ctx); }
extern "C" int ptzba_comm_allreduce_f64(ptzba_ctx* ctx, double*, int64_t) { return NOT_YET(ctx); }

// stubs.cu - entry points declared in include/ptzba.h whose implementation file is not linked yet.
#include "common.h"
#define NOT_YET(ctx) ptzba_fail((ctx), PTZBA_ERR_STATE, "%s: not implemented in this build", __func__)
void ptzba_comm_release(ptzba_ctx*) {}
extern "C" int ptzba_comm_unique_id(ptzba_ctx* ctx, void*) { return NOT_YET(ctx); }
extern "C" int ptzba_comm_init(ptzba_ctx* ctx, const void*, int, int) { return NOT_YET(ctx); }
extern "C" int ptzba_comm_allreduce_f64(ptzba_ctx* ctx, double*, int64_t) { return NOT_YET(ctx); }

// phovo_batch.cu -- batch-of-pairs extension (placeholder until the persistent kernel lands).
#include "phovo_ctx.h"

struct phovo_batch_state { int dummy; };
void phovo_batch_release(phovo_ctx* ctx) { delete ctx->batch; ctx->batch = nullptr; }

extern "C" int phovo_batch_align(phovo_ctx* ctx, int, int, int, const uint8_t*, const void*, int, double,
                                 const uint8_t*, const double*, double*, int32_t*) {
  return ctx ? ctx->fail(PHOVO_E_UNSUPPORTED, "batch path not built yet") : PHOVO_E_INVALID;
}
extern "C" int phovo_batch_align_device(phovo_ctx* ctx, int, int, int, const uint8_t*, const void*, int, double,
                                        const uint8_t*, const double*, double*, int32_t*) {
  return ctx ? ctx->fail(PHOVO_E_UNSUPPORTED, "batch path not built yet") : PHOVO_E_INVALID;
}
extern "C" int phovo_batch_set_record_stats(phovo_ctx* ctx, int) { return ctx ? PHOVO_OK : PHOVO_E_INVALID; }
extern "C" int phovo_batch_get_iter_stats(const phovo_ctx*, int, int, phovo_iter_stats*) { return PHOVO_E_UNSUPPORTED; }
extern "C" int phovo_batch_num_iter_stats(const phovo_ctx*, int) { return 0; }

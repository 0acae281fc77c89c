#include "binning.cuh"
namespace gudni_bin {
int binScene(gudni_ctx* ctx, const gudni_shape_entry*, int) { return ctxFail(ctx, GUDNI_ERR_STATE, "binning not built yet"); }
}

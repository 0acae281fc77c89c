// strands.cuh — launcher of level 3 (outlines -> geometry heap + shape entries); see strands.cu.
#pragma once
struct gudni_ctx;
namespace gudni_launch {
int strandTableInit(gudni_ctx* ctx);
int buildStrands(gudni_ctx* ctx, const void* devShapes, int nShapes, const void* devOutlines, const void* devPairs,
                 const void* devTransforms);
}  // namespace gudni_launch

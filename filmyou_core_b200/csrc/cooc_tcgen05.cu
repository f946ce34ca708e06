// cooc_tcgen05.cu -- config 3: item-item co-occurrence counts.  Placeholder entry
// points until the int8 tcgen05 GEMM lands; they fail loudly instead of falling back.
#include "../../include/filmyou_rm2.h"

extern "C" int fy_cooc_counts(fy_rm2_ctx*, int32_t, int32_t, int32_t*, double*) { return FY_E_UNSUPPORTED; }
extern "C" int fy_cooc_topk(fy_rm2_ctx*, int32_t, int32_t*, int32_t*, int32_t*) { return FY_E_UNSUPPORTED; }

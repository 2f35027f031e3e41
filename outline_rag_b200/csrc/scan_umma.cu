// placeholder until the tcgen05 scan lands: never selected.
#include "scan_umma.h"
namespace orx {
struct UmmaPlan { int device; };
UmmaPlan *umma_plan_create(int device) { return new UmmaPlan{device}; }
void umma_plan_destroy(UmmaPlan *p) { delete p; }
void umma_plan_invalidate(UmmaPlan *) {}
bool umma_should_use(const UmmaPlan *, int, uint32_t) { return false; }
const char *umma_last_error() { return "tcgen05 scan not built"; }
int umma_search(UmmaPlan *, int, const void *, const float *, const double *, const orx_id *, uint32_t,
                const float *, const float *, const __nv_bfloat16 *, const QueryPrep *, int, int, orx_id *,
                double *, int *, int *, cudaStream_t, uint64_t *) { return ORX_ERR_INVALID; }
}

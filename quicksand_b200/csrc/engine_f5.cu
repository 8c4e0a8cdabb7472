// FUNNEL family (config c5: Neal's funnel)
#include "engine_inst.cuh"
namespace qsb {
extern const EngineVariant g_variants_f5[] = {
    VariantImpl<FAM_FUNNEL, 1, 16>::make(), VariantImpl<FAM_FUNNEL, 32, 4>::make(), VariantImpl<FAM_FUNNEL, 32, 32>::make()};
extern const int g_nvariants_f5 = 3;
}  // namespace qsb

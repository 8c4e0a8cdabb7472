// EIGHT_SCHOOLS family (config c4: hierarchical normal, J groups)
#include "engine_inst.cuh"
namespace qsb {
extern const EngineVariant g_variants_f4[] = {
    VariantImpl<FAM_EIGHT_SCHOOLS, 1, 16>::make(), VariantImpl<FAM_EIGHT_SCHOOLS, 32, 4>::make(), VariantImpl<FAM_EIGHT_SCHOOLS, 32, 32>::make()};
extern const int g_nvariants_f4 = 3;
}  // namespace qsb

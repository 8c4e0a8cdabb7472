// INDEP family (independent ERPs with constant parameters; test/testsuite.t:621-691, 719-730, 776-823)
#include "engine_inst.cuh"
namespace qsb {
extern const EngineVariant g_variants_f1[] = {
    VariantImpl<FAM_INDEP, 1, 1>::make(), VariantImpl<FAM_INDEP, 1, 10>::make(), VariantImpl<FAM_INDEP, 32, 4>::make()};
extern const int g_nvariants_f1 = 3;
}  // namespace qsb

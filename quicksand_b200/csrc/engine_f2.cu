// GAUSS_PREC family (config c2: 10-D correlated Gaussian; thread per chain, state in registers)
#include "engine_inst.cuh"
namespace qsb {
extern const EngineVariant g_variants_f2[] = {
    VariantImpl<FAM_GAUSS_PREC, 1, 4>::make(), VariantImpl<FAM_GAUSS_PREC, 1, 10>::make(), VariantImpl<FAM_GAUSS_PREC, 1, 16>::make()};
extern const int g_nvariants_f2 = 3;
}  // namespace qsb

// GAUSS_PREC family (config c2: 10-D correlated Gaussian; thread per chain, state in registers)
#include "engine_inst.cuh"
namespace qsb {
extern const EngineVariant g_variants_f2[] = {
    VariantImpl<FAM_GAUSS_PREC, 1, 4>::make(), VariantImpl<FAM_GAUSS_PREC, 1, 10>::make(), VariantImpl<FAM_GAUSS_PREC, 1, 16>::make(),
    // four lanes per chain (8 chains per warp): a quarter of the state per thread, four times the warps -- what fills the
    // machine at config c2's 65 536 chains (thread per chain: 0.86 waves of 13 warps per SM)
    VariantImpl<FAM_GAUSS_PREC, 4, 3>::make(), VariantImpl<FAM_GAUSS_PREC, 4, 4>::make()};
extern const int g_nvariants_f2 = 5;
}  // namespace qsb

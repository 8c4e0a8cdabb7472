// INDEP family, thread per chain with up to 32 components (vector-valued dirichlet choices need the thread-per-chain
// mapping); its own translation unit so that it compiles in parallel with the others
#include "engine_inst.cuh"
namespace qsb {
extern const EngineVariant g_variants_f1w[] = {VariantImpl<FAM_INDEP, 1, 32>::make()};
extern const int g_nvariants_f1w = 1;
}  // namespace qsb

// Host interface of the Bernoulli-logit GLM path (config c3: Bayesian logistic regression under HMC).
// Implemented in glm_hmc.cu; used by capi.cu only.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace qsb {

struct GlmModel {
  int N = 0, d = 0;        // observations, regression dimension
  int Npad = 0;            // rows padded to a whole number of row blocks
  int S = 0;               // row stride of the staged X in doubles (>= d, S % 8 == 4: conflict-free fragment loads; 114 when perm)
  int Dp = 0;              // d padded to a multiple of 8 (DMMA n-tiles)
  double prior_sigma = 0;
  bool perm = false;       // c3's shape (13 dimension tiles, d <= 100): columns of Xp permuted, dimension 8J + n in column 14 n + J (glm_hmc.cu: PERMX)
  double* Xp = nullptr;    // HBM [Npad][S], zero padded
  double* Xraw = nullptr;  // HBM [N][d] staging of qsb_model_update_data when Xp is column-permuted
  double* yv = nullptr;    // HBM [Npad]: 0.0 / 1.0, 2.0 marks a padding row
  uint8_t* ybin_dev = nullptr;  // HBM [N] raw labels, staging of qsb_model_update_data
};

struct GlmSampler;

int glm_model_create(GlmModel* m, const double* X, const uint8_t* y, int N, int d, double prior_sigma, std::string& why);
void glm_model_destroy(GlmModel* m);
int glm_model_update(GlmModel* m, const double* X, const uint8_t* y, void* stream, std::string& why);

// kind: 1 HMC (stepSize, numSteps), 2 Drift (scale), 3 HARM (scale) -- QSB_KERNEL_*
GlmSampler* glm_sampler_create(GlmModel* m, int kind, double stepSize, uint32_t numSteps, double scale, bool adapt, uint64_t seed,
                               uint32_t chain_begin, uint32_t numchains, uint32_t sample_chains, uint32_t flags, void* stream, std::string& why);
void glm_sampler_destroy(GlmSampler* s);
int glm_sampler_init(GlmSampler* s, std::string& why);
int glm_sampler_set_state(GlmSampler* s, const double* x, std::string& why);
int glm_sampler_get_state(GlmSampler* s, double* x, std::string& why);
int glm_sampler_begin(GlmSampler* s, uint64_t max_samples, uint64_t burnin, uint64_t lag, uint64_t numiters, std::string& why);
int glm_sampler_advance(GlmSampler* s, uint64_t iters, std::string& why);
int glm_sampler_set_temperatures(GlmSampler* s, const double* temps, uint64_t n, std::string& why);
void glm_sampler_moments_view(GlmSampler* s, const double** mom, uint64_t* half, uint32_t* batch);
void glm_sampler_sample_view(GlmSampler* s, const double** samples, const double** samp_lp, int* G, uint32_t* SC, int* retdim);
int glm_sampler_stats(GlmSampler* s, uint64_t* made, uint64_t* acc, uint64_t* grad_evals, uint64_t* leapfrogs, uint64_t* launches,
                      double* seconds, std::string& why);
void glm_sampler_health(GlmSampler* s, uint64_t* nonfinite, uint64_t* divergent);
int glm_sampler_fetch_record(GlmSampler* s, double* state, int32_t* accept, double* dh, double* step, double* leap, std::string& why);
void glm_sampler_set_timing(GlmSampler* s, bool on);
int glm_sampler_kernel_time(GlmSampler* s, double* seconds, uint64_t* launches, std::string& why);
int glm_layout(int ntiles, int nblocks, uint32_t C, int nsm, std::vector<int>& start, std::vector<int>& first, std::vector<int>& np);
int glm_chains_per_tile();
int glm_rows_per_block();
int glm_debug_leapfrog(GlmModel* m, int n, const double* x, const double* p, const double* g, const double* eps, int last,
                       double* x_out, double* p_out, double* g_out, double* lp_out, std::string& why);
int glm_logp_grad(GlmModel* m, const double* x, int n, double* lp_dual, double* grad, double* lp_float, std::string& why);

}  // namespace qsb

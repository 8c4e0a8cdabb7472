/* qsb200.h -- plain-C boundary of the B200 many-chain MCMC engine (libqsb200.so).
 *
 * This header is what quicksand's Terra host code loads with terralib.includec (the
 * reference's own FFI mechanism: qs/lib/ad.t:3, qs/lib/random.t:2-7, qs/lib/util.t:45-61,
 * qs/mcmc.t:12-17).  It is plain C: fixed-width integers, plain pointers and sizes,
 * opaque handles, no C++ or torch types.  Every entry point returns 0 on success and a
 * non-zero code on failure (message through qsb_last_error()); nothing aborts or throws
 * across the boundary -- the Terra shim (qs/b200.t) turns a non-zero return into the
 * reference's print + abort behaviour (qs/lib/std.t:35-43).
 *
 * The boundary sits at the *method* level of the reference,
 *     method(program) -> terra(samples: &S.Vector(SampleType(program)))   (qs/infer.t:57-69,
 *                                                                          qs/mcmc.t:45,95-103)
 * and not at the per-iteration kernel:next level (qs/mcmc.t:63): one FFI call per chain
 * per iteration would forfeit every fusion the device path depends on.
 *
 * Threading: one host thread per sampler (the reference is single-threaded with
 * process-global state: qs/trace.t:452-454, qs/lib/ad.t:9,19).  Memory: the library never
 * allocates memory the caller frees and never frees caller memory (qs/infer.t:127
 * "Caller is responsible for the memory of the returned vector").
 */
#ifndef QSB200_H
#define QSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSB_VERSION 100

/* ---- model families: fixed-structure continuous programs.
 * A quicksand program of one of these shapes (all choices {struc=false}) is described by a
 * POD descriptor; anything else (structural choices, discrete ERPs) is rejected on the
 * Terra side at qs.MCMC(...)(program) time. */
enum {
  QSB_FAM_INDEP = 1,         /* K independent ERPs with constant params   (test/testsuite.t:621-691, 719-730) */
  QSB_FAM_GAUSS_PREC = 2,    /* x_i ~ gaussian(0,1); factor(-0.5 x'Ax)     (config c2) */
  QSB_FAM_LOGISTIC = 3,      /* theta_j ~ gaussian(0,s); flip.observe(y_i, 1/(1+exp(-x_i.theta)))  (config c3) */
  QSB_FAM_EIGHT_SCHOOLS = 4, /* mu, logtau, theta_j ~ gaussian(mu, exp(logtau)); gaussian.observe(y_j, theta_j, sigma_j) (c4) */
  QSB_FAM_FUNNEL = 5         /* v ~ gaussian(0,3); x_i ~ gaussian(0, exp(v/2))  (config c5) */
};

/* ERP kinds with their bounding transforms as bound in qs/erpdefs.t:
 * gaussian -> Bounds.None (46-57), gamma -> Lower(0) (61-69), beta -> LowerUpper(0,1) (92-100),
 * uniform -> LowerUpper(lo,hi) (28-36). */
enum { QSB_ERP_GAUSSIAN = 0, QSB_ERP_GAMMA = 1, QSB_ERP_BETA = 2, QSB_ERP_UNIFORM = 3,
       /* dirichlet -> Bounds.UnitSimplex (erpdefs.t:179-212, erp.t:200-254): a K-vector choice is K consecutive
        * components of this kind with erp_p0 = alpha_k and erp_p1 = k (index inside the choice, 0..K-1); component K-1
        * of the unconstrained vector is the reference's unused stick coordinate.  gammamv / betamv (erpdefs.t:72-88,
        * 103-125) are host-side reparametrisations of gamma / beta. */
       QSB_ERP_DIRICHLET = 4,
       /* Bounds.Upper (erp.t:158-173), which the reference offers to user-defined ERPs (makeRandomChoice) and binds to none of
        * its own: value = -gamma(shape = erp_p0, scale = erp_p1) < 0, upper bound 0. */
       QSB_ERP_NEGGAMMA = 5 };

/* transition kernels: qs/hmc.t:57-66, qs/drift.t:63-73, qs/harm.t:16-22; TRACEMH = TraceMHKernel{doStruct=false}
 * (qs/mcmc.t:134-193): one random choice per iteration, gaussian ERPs drift with their own sigma (erpdefs.t:46-57), the
 * others resample from their prior (erp.t:18-38); the structural half of that kernel is out of scope. */
enum { QSB_KERNEL_HMC = 1, QSB_KERNEL_DRIFT = 2, QSB_KERNEL_HARM = 3, QSB_KERNEL_TRACEMH = 4 };

typedef struct qsb_model qsb_model;       /* opaque; owns device copies of the model data */
typedef struct qsb_sampler qsb_sampler;   /* opaque; owns the SoA chain state in HBM */

/* All pointers are HOST pointers, read during qsb_model_create and not retained. */
typedef struct {
  int32_t family;          /* QSB_FAM_* */
  int32_t dim;             /* number of real components n (unconstrained dimension) */
  int32_t nobs;            /* LOGISTIC: N rows;  EIGHT_SCHOOLS: J (dim = J+2);  otherwise 0 */
  int32_t ret_mode;        /* INDEP: 0 = return sum(values)/ret_div (scalar), 1 = return the value vector */
  double ret_div;
  double prior_sigma;      /* LOGISTIC: prior standard deviation */
  const int32_t* erp_kind; /* INDEP [dim] */
  const double* erp_p0;    /* INDEP [dim]  first ERP parameter  (mu / shape / a / lo) */
  const double* erp_p1;    /* INDEP [dim]  second ERP parameter (sigma / scale / b / hi) */
  const double* A;         /* GAUSS_PREC [dim*dim] row-major */
  const double* X;         /* LOGISTIC [nobs*dim] row-major */
  const uint8_t* ybin;     /* LOGISTIC [nobs] 0/1 */
  const double* y;         /* EIGHT_SCHOOLS [nobs] */
  const double* sigma;     /* EIGHT_SCHOOLS [nobs] */
  const int32_t* erp_lexid;/* INDEP [dim] or NULL: call-site id of each component's choice (erp.t:318-321); choices made by one
                            * ERP call inside a loop share an id.  NULL = every choice is its own call site.  Only
                            * DriftKernel{lexicalScaleSharing=true} reads it.  The other families have fixed sites:
                            * GAUSS_PREC / LOGISTIC one loop; EIGHT_SCHOOLS mu, logtau, theta-loop; FUNNEL v, x-loop. */
  const double* erp_fmu;   /* INDEP [dim] or NULL: a Gaussian factor on the value of scalar choice i right after it,          */
  const double* erp_fsigma;/*   qs.factor(gaussian.logprob(v_i, fmu_i, fsigma_i)) == qs.gaussian.observe(v_i, fmu_i, fsigma_i)
                            *   (trace.t:1099-1107; erp.t:687-696; testsuite.t:394-442).  fsigma_i <= 0: no factor.            */
  /* Latent ERP parameters (INDEP, all three NULL = every parameter constant).  Parameter q (0/1) of scalar choice i is
   *   f( erp_p{q}[i] + erp_pcoef[2i+q] * value_j ),  j = erp_pref[2i+q] < i an earlier scalar choice (-1: constant),
   *   f = identity (erp_ptf[2i+q] = 0) or exp (1)
   * -- the parameter expressions of a program such as  mu = gaussian(0,1); x = gaussian(mu, exp(logtau)).  The reference
   * re-evaluates them on every run and hands them to RandomChoiceT:update (trace.t:948-994, erp.t:576-610); in a dual trace the
   * tape differentiates the log-density with respect to them, including d/dk of the 5-coefficient log-gamma (distrib.t:84-106).
   * gaussian, gamma and beta choices only (a uniform's parameters are its bounds); thread per chain up to 32 components, warp per
   * chain (general evaluator on scratch memory) up to the INDEP family's 128. */
  const int32_t* erp_pref;  /* [2*dim] */
  const double* erp_pcoef;  /* [2*dim] */
  const int32_t* erp_ptf;   /* [2*dim] or NULL (all identity) */
  /* Hard constraints (INDEP, both NULL = none): the statement  qs.condition(value_i > erp_cond_lo[i] and value_i <
   * erp_cond_hi[i])  right after scalar choice i (trace.t:794-796, 1103-1107; -inf / +inf = no bound on that side).  The trace
   * is initialised by rejection (trace.t:612-623: the program is re-run from scratch until every condition holds) and every
   * kernel's accept test is `conditionsSatisfied and log(random()) < ...` (hmc.t:164, drift.t:158, harm.t:93, mcmc.t:181). */
  const double* erp_cond_lo; /* [dim] */
  const double* erp_cond_hi; /* [dim] */
} qsb_model_desc;

/* One transition kernel; nspecs > 1 means MixtureKernel(kernels, weights) (qs/mcmc.t:199-291). */
typedef struct {
  int32_t kind;        /* QSB_KERNEL_* */
  int32_t doAdapt;     /* doStepSizeAdapt (hmc.t:61-62) / doScaleAdapt (drift.t:66-67, harm.t:19-20) */
  double stepSize;     /* HMC: stepSize (hmc.t:59) */
  uint64_t numSteps;   /* HMC: numSteps (hmc.t:60) */
  double scale;        /* Drift / HARM: scale (drift.t:65, harm.t:18) */
  double weight;       /* mixture weight; ignored when nspecs == 1 */
  int32_t lexicalScaleSharing; /* Drift: choices from one call site share one adapted bandwidth (drift.t:58-60, 71-73, 101-115) */
  int32_t reserved;
} qsb_kernel_spec;

enum {
  QSB_FLAG_MOMENTS = 1u,     /* keep running statistics of the kept samples INSTEAD of the samples: per chain and return-value
                              * dimension Welford moments of the two halves of the run (split-Rhat) and of sqrt(max_samples)-long
                              * batch means (ESS), 56 bytes in all, updated in the transition kernels.  qsb_sampler_diagnostics then
                              * works without stored draws (fetch / queries have nothing to read); see its layout note. */
  QSB_FLAG_RECORD = 2u,      /* keep a per-iteration record (state, accept, dH, step, kernel index) -- parity tests */
  QSB_FLAG_RECORD_LEAP = 4u, /* additionally record the position after every leapfrog step (single HMC kernel) */
  QSB_FLAG_TIME_KERNEL = 8u, /* bracket every launch of the dominant kernel with CUDA events (roofline measurement) */
  QSB_FLAG_BALANCED = 16u,   /* LOGISTIC: cut the gradient kernel's work into one equal-cost run of row blocks per SM.  This is the
                              * DEFAULT since round 2 (the flag is accepted and changes nothing): one full wave for any chain count,
                              * 2-3 partial sums per chain tile.  A chain's draws then depend, in the last bits, on how many chains
                              * the sampler holds and where the chain sits (still within 1e-10 per step of any other grouping). */
  QSB_FLAG_INVARIANT = 32u   /* LOGISTIC: the (chain tile x 12 row splits) grid instead: the grouping of a chain's row blocks into
                              * partial sums depends on the data shape alone, so a chain's draws are bit-identical for ANY
                              * partition of the chains over samplers / GPUs (every other family is invariant by construction).
                              * Costs wave quantisation: ~11 % at 1 024 chains per GPU, <1 % at 8 192. */
};

typedef struct {
  uint64_t seed;           /* Philox key; the sub-stream of a draw is (seed, global chain id, iteration, slot) */
  uint32_t chain_begin;    /* global id of this sampler's first chain (multi-GPU shards: rank * numchains) */
  uint32_t numchains;      /* chains owned by this sampler; numchains = 1 reproduces the reference's single chain */
  uint32_t sample_chains;  /* leading chains whose samples are stored (0 = all) */
  uint32_t flags;          /* QSB_FLAG_* */
  int32_t device;          /* CUDA device ordinal (-1 = current) */
  int32_t reserved;
  void* stream;            /* cudaStream_t to launch on (NULL = default stream) */
} qsb_run_config;

typedef struct {
  uint64_t props_made, props_accepted;    /* KernelPropStats summed over chains and sub-kernels (mcmc.t:116-125) */
  uint64_t sub_made[8], sub_accepted[8];  /* per sub-kernel of a mixture (mcmc.t:256-273) */
  uint64_t grad_evals;                    /* log-density gradient evaluations (= leapfrog steps + re-gathers) */
  uint64_t leapfrog_steps;                /* HMCKernel:step calls inside :next, all chains */
  uint64_t iterations;                    /* MCMC iterations done so far (per chain) */
  uint64_t kernel_launches;               /* CUDA kernels launched by this sampler so far */
  double seconds;                         /* device time of the last qsb_sampler_run / advance (CUDA events) */
  /* chain health, summed over chains since the last init / set_state: transitions whose accept threshold was not finite (a NaN
   * or infinite log-density: the proposal left the support or overflowed; always rejected), and HMC transitions whose energy
   * error exceeded 1000 in magnitude (divergent trajectories: the step size is too large for the local curvature) */
  uint64_t nonfinite, divergent;
} qsb_stats;

int qsb_version(void);
const char* qsb_last_error(void);
/* SHA-256 (hex) over the CUDA sources and this header the loaded binary was compiled from (quicksand_b200/build.py
 * passes it in at compile time): ties a measured library to a source tree. */
const char* qsb_source_hash(void);

/* ---- models */
int qsb_model_create(const qsb_model_desc* desc, int32_t device, qsb_model** out);
int qsb_model_destroy(qsb_model* m);
/* Refresh the observation data of an existing model in place from HOST buffers (same shapes as at creation;
 * LOGISTIC: X, ybin; EIGHT_SCHOOLS: y, sigma).  Asynchronous on `stream` when the buffers are pinned. */
int qsb_model_update_data(qsb_model* m, const qsb_model_desc* desc, void* stream);
int qsb_model_dim(const qsb_model* m);      /* n: real components          (trace.t:830-905) */
int qsb_model_retdim(const qsb_model* m);   /* length of the return value  (infer.t:47-49)   */

/* ---- samplers: replaces TraceType.salloc():init(true) + KernelType.salloc():init() (mcmc.t:55-56) */
int qsb_sampler_create(qsb_model* m, const qsb_kernel_spec* specs, int32_t nspecs,
                       const qsb_run_config* cfg, qsb_sampler** out);
int qsb_sampler_destroy(qsb_sampler* s);

/* Rejection-free forward-sampling initialisation of every chain (trace.t:592-624, erp.t:567-572). */
int qsb_sampler_init(qsb_sampler* s);
/* Explicit initial values, ERP option {init=...} (erp.t:57-61, 677-679): x is HOST [numchains][dim],
 * unconstrained coordinates.  Also used by tests to re-synchronise with the oracle. */
int qsb_sampler_set_state(qsb_sampler* s, const double* x);
int qsb_sampler_get_state(qsb_sampler* s, double* x);   /* HOST [numchains][dim], unconstrained */

/* doMCMC (mcmc.t:50-90): iters = burnin + nsamps*lag; iteration i is kept iff i >= burnin and
 * i % lag == 0.  Restarts the iteration counter at 0 and discards previously kept samples. */
int qsb_sampler_run(qsb_sampler* s, uint64_t nsamps, uint64_t burnin, uint64_t lag);
/* Continue the same run by `iters` more iterations (iteration counter keeps counting); every
 * iteration with (i >= burnin and i % lag == 0) under the parameters of qsb_sampler_begin is kept. */
int qsb_sampler_begin(qsb_sampler* s, uint64_t max_samples, uint64_t burnin, uint64_t lag, uint64_t numiters);
int qsb_sampler_advance(qsb_sampler* s, uint64_t iters);
uint64_t qsb_sampler_num_samples(const qsb_sampler* s);

/* AnnealingKernel(kernel, annealSched) (mcmc.t:297-319): temps is a HOST array, temps[i] = annealSched(i, numiters)
 * evaluated by the caller for iterations i = 0..n-1 (the schedule is a Terra function on the host side; later
 * iterations reuse temps[n-1]).  Iteration i sets the trace temperature to temps[i] before the kernel runs
 * (trace.t:652-655): every log-density evaluated during it is divided by the temperature (trace.t:737-741), cached
 * values keep the temperature they were computed under, kept samples store logprob * temperature (mcmc.t:69).
 * n = 0 (or temps = NULL) removes the schedule (temperature 1).  Call between qsb_sampler_init and run/advance. */
int qsb_sampler_set_temperatures(qsb_sampler* s, const double* temps, uint64_t n);

/* Copy kept samples into caller-owned memory laid out like S.Vector(Sample(T))._data:
 * element = { double value[retdim]; double logprob; double loglikelihood; }  (infer.t:13-18),
 * element (c - chain_begin_local) * (sample_end - sample_begin) + (k - sample_begin) at dst + that * stride.
 * stride_bytes >= (retdim+2)*8.  Chains are sampler-local indices. */
int qsb_sampler_fetch_samples(qsb_sampler* s, uint32_t chain_begin, uint32_t chain_end,
                              uint64_t sample_begin, uint64_t sample_end, void* dst, size_t stride_bytes);
/* The same without blocking the host: the transposing gather is queued on the sampler's stream, the device-to-host copy
 * on a second stream of the library (two staging buffers alternate), so the copy of one step's draws overlaps the
 * transitions queued after it.  dst should be page-locked (the caller pins it; Terra: cudaHostRegister on Vector._data)
 * and must stay untouched until qsb_sampler_fetch_sync returns. */
int qsb_sampler_fetch_samples_async(qsb_sampler* s, uint32_t chain_begin, uint32_t chain_end,
                                    uint64_t sample_begin, uint64_t sample_end, void* dst, size_t stride_bytes);
int qsb_sampler_fetch_sync(qsb_sampler* s);
/* Only the values, as a dense HOST array [chain][sample][retdim] (no logprob fields). */
int qsb_sampler_fetch_values(qsb_sampler* s, uint32_t chain_begin, uint32_t chain_end,
                             uint64_t sample_begin, uint64_t sample_end, double* dst);

int qsb_sampler_stats(qsb_sampler* s, qsb_stats* out);
/* QSB_FLAG_TIME_KERNEL: summed device time and launch count of the dominant kernel (the log-density/gradient
 * contraction for LOGISTIC, the fused transition kernel otherwise) since the last call; resets the counters. */
int qsb_sampler_kernel_time(qsb_sampler* s, double* seconds, uint64_t* launches);
/* Switch QSB_FLAG_TIME_KERNEL on / off for the following advances.  (The event pairs sit between the launches of a
 * transition and suppress the programmatic-dependent-launch overlap of the LOGISTIC path: measure throughput with timing
 * off and the kernel's own duration in a separate pass with timing on.) */
int qsb_sampler_set_kernel_timing(qsb_sampler* s, int32_t on);

/* Cross-chain sufficient statistics of the kept samples, for split-Rhat / pooled moments / ESS.
 * dst receives, per return-value dimension j (retdim of them), 12 doubles:
 *   [0] number of half-chains   [1] draws per half-chain
 *   [2] sum over half-chains of the half-chain mean   [3] sum of (half-chain mean)^2
 *   [4] sum of the half-chain sample variances (n-1 denominators)
 *   [5] sum of all draws   [6] sum of all draws^2
 *   [7] number of chains M   [8] draws per chain N
 *   [9] sum over chains of the chain mean   [10] sum of (chain mean)^2   [11] sum of the chain sample variances (N-1)
 * followed (if maxlag > 0) by retdim*(maxlag+1) doubles: sum over chains of the lag-t autocovariance
 * numerators sum_i (x_i - m_c)(x_{i+t} - m_c), t = 0..maxlag, m_c = the chain's mean.
 * These are plain sums, so shards on different GPUs combine with one all-reduce(sum).
 * If dst_is_device != 0, dst is a device pointer on the sampler's device (e.g. a torch tensor
 * handed to NCCL); otherwise a host pointer.
 * With QSB_FLAG_MOMENTS the same 12 sums per dimension are formed from the running statistics (the halves are the
 * kept-sample ranges [0, max_samples/2) and [max_samples/2, n): run the sampler to max_samples for equal halves), maxlag is
 * ignored and 3 doubles per dimension follow instead of the autocovariances: [0] sum over chains of the sample variance of
 * the chain's completed batch means, [1] the batch size b, [2] completed batches per chain -- ESS = M N s^2 / (b * mean
 * batch-mean variance).  Total retdim * 15 doubles; [1], [2] are constants like header entries [1], [8]. */
int qsb_sampler_diagnostics(qsb_sampler* s, int32_t maxlag, double* dst, int32_t dst_is_device);

/* ---- queries (qs/infer.t:128-264) evaluated on the device, per chain, over the kept samples of a sampler (no host
 * round trip of the samples) or over a caller-owned HOST vector of Sample elements laid out as qsb_sampler_fetch_samples
 * writes it ([nchains][nsamps] elements of stride_bytes).  Chains are sampler-local indices [chain_begin, chain_end).
 *
 * Expectation(doVariance) (infer.t:146-187), reference arithmetic: the mean starts at value[0], weights
 * w_i = exp(loglikelihood_i) are summed from i = 1 (mean = sum(v)/(N-1) for MCMC samples), the variance is
 * sum_i |value_i - mean|^2 / N.  mean: HOST [nchains][retdim]; var: HOST [nchains] or NULL (doVariance = false). */
int qsb_query_expectation(qsb_sampler* s, uint32_t chain_begin, uint32_t chain_end, double* mean, double* var);
int qsb_query_expectation_host(const void* samples, size_t stride_bytes, uint32_t nchains, uint64_t nsamps, int32_t retdim,
                               double* mean, double* var);
/* MAP (infer.t:191-206): the first sample of largest logprob.  value: HOST [nchains][retdim]; logprob: HOST [nchains] or NULL. */
int qsb_query_map(qsb_sampler* s, uint32_t chain_begin, uint32_t chain_end, double* value, double* logprob);
int qsb_query_map_host(const void* samples, size_t stride_bytes, uint32_t nchains, uint64_t nsamps, int32_t retdim,
                       double* value, double* logprob);
/* Autocorrelation(mean, variance) (infer.t:213-264): ac[t] = sum_{i<N-t} (v_i - m).(v_{i+t} - m) / ((N-t) * variance),
 * t = 0..N-1.  mean: HOST [retdim] shared by all chains, or NULL = each chain's own Expectation(true) (then `variance`
 * is ignored).  ac: HOST [nchains][nsamps]. */
int qsb_query_autocorrelation(qsb_sampler* s, uint32_t chain_begin, uint32_t chain_end, const double* mean, double variance, double* ac);
int qsb_query_autocorrelation_host(const void* samples, size_t stride_bytes, uint32_t nchains, uint64_t nsamps, int32_t retdim,
                                   const double* mean, double variance, double* ac);

/* Per-iteration record (QSB_FLAG_RECORD): HOST arrays, any may be NULL.
 *   state [numchains][iters][dim]   accept, kidx [numchains][iters] (int32)   dh, step [numchains][iters]
 *   leap  [numchains][iters][numSteps][dim]  (QSB_FLAG_RECORD_LEAP) */
int qsb_sampler_fetch_record(qsb_sampler* s, double* state, int32_t* accept, double* dh, double* step,
                             int32_t* kidx, double* leap);

/* ---- multi-GPU (SURVEY.md 8b Threading, 8e): chains are block-partitioned over the devices of ONE box; a chain's random
 * stream is keyed by its GLOBAL id, so the partition does not change its draws.  There is no data-path collective; the only
 * exchange is the all-reduce(sum) of the diagnostics' sufficient statistics, done by the library over its own NCCL
 * communicator (NVLink / NVSwitch).  The reference is single-threaded (qs/trace.t:452-454), so two shapes are offered:
 *
 * (1) qsb_group: ONE host thread drives several devices -- what a Terra process calls.  qs.MCMC(kernel, {numchains=,
 *     devices={0,1,..}}) maps onto it (qs/mcmc.t:36-43 parameter table + the new `devices`).  Every call fans out to one
 *     sampler per device (asynchronous launches on per-device streams) and joins; results are indexed by global chain.
 * (2) qsb_comm: one process per device (torchrun / MPI style); each process owns a qsb_sampler with chain_begin = its
 *     offset and joins a communicator through a 128-byte id created by rank 0 and distributed by the caller. */
typedef struct qsb_group qsb_group;
typedef struct qsb_comm qsb_comm;
#define QSB_COMM_ID_BYTES 128

/* cfg->numchains = TOTAL chains, cfg->chain_begin = global id of the first; cfg->device / cfg->stream are ignored (each
 * shard runs on its own stream of its device).  devices[i] may repeat (several shards on one device: used by tests on a
 * single-GPU box; the reduction then runs as a device kernel instead of NCCL).  desc as for qsb_model_create: the model
 * data are replicated on every device. */
int qsb_group_create(const qsb_model_desc* desc, const qsb_kernel_spec* specs, int32_t nspecs, const qsb_run_config* cfg,
                     const int32_t* devices, int32_t ndevices, qsb_group** out);
int qsb_group_destroy(qsb_group* g);
int32_t qsb_group_size(const qsb_group* g);
qsb_sampler* qsb_group_sampler(qsb_group* g, int32_t shard);      /* borrowed; shard i owns a contiguous block of chains */
int qsb_group_init(qsb_group* g);
int qsb_group_run(qsb_group* g, uint64_t nsamps, uint64_t burnin, uint64_t lag);
int qsb_group_begin(qsb_group* g, uint64_t max_samples, uint64_t burnin, uint64_t lag, uint64_t numiters);
int qsb_group_advance(qsb_group* g, uint64_t iters);
int qsb_group_sync(qsb_group* g);                                  /* wait for every device */
/* chains are indices into the group's [0, numchains) */
int qsb_group_fetch_samples(qsb_group* g, uint32_t chain_begin, uint32_t chain_end, uint64_t sample_begin, uint64_t sample_end,
                            void* dst, size_t stride_bytes);
int qsb_group_stats(qsb_group* g, qsb_stats* out);                 /* counters summed over shards, seconds = max */
/* qsb_sampler_diagnostics of every shard, all-reduced (NCCL ncclSum over the group's communicator): dst is HOST.
 * (Here and in qsb_sampler_diagnostics_allreduce pass maxlag = -1 for samplers created with QSB_FLAG_MOMENTS: the
 * 15-doubles-per-dimension layout.) */
int qsb_group_diagnostics(qsb_group* g, int32_t maxlag, double* dst);

int qsb_comm_unique_id(uint8_t id[QSB_COMM_ID_BYTES]);            /* rank 0; hand the bytes to the other ranks */
int qsb_comm_create(const uint8_t id[QSB_COMM_ID_BYTES], int32_t rank, int32_t nranks, int32_t device, qsb_comm** out);
int qsb_comm_destroy(qsb_comm* c);
/* in-place sum over ranks of n doubles in DEVICE memory, on `stream` of the communicator's device */
int qsb_comm_allreduce_sum(qsb_comm* c, double* buf, size_t n, void* stream);
/* qsb_sampler_diagnostics of this rank's shard followed by the all-reduce; dst is HOST, identical on every rank */
int qsb_sampler_diagnostics_allreduce(qsb_sampler* s, qsb_comm* c, int32_t maxlag, double* dst);

/* ---- measurement support */
/* FP64 peak of the device the caller runs on, measured in-process in about `seconds` of device time: which = 0 dependent
 * chains of DFMA in registers, 1 mma.sync.m8n8k4.f64 (DMMA).  Returns TFLOP/s in *tflops (roofline denominators of
 * bench.py: MEASURED_PEAKS.json carries no fp64 figure). */
int qsb_measure_fp64_peak(int32_t which, double seconds, int32_t device, double* tflops);

/* ---- test hooks (parity with the CPU oracle) */
/* Host-only (no device needed): the balanced decomposition the LOGISTIC gradient kernel would use for `numchains` chains and
 * `nobs` observations on a device of `nsm` SMs.  start [ncta+1] (row-block units, chain-tile-major), first / np [ntiles];
 * *ntiles, *nblocks, *ncta receive the sizes; arrays may be NULL to query the sizes (capacity = entries available). */
int qsb_debug_glm_layout(uint32_t numchains, int32_t nobs, int32_t nsm, int32_t capacity, int32_t* start, int32_t* first, int32_t* np,
                         int32_t* ntiles, int32_t* nblocks, int32_t* ncta);
/* Teacher-forced leapfrog: exactly one HMCKernel:step (qs/hmc.t:307-323) of n independent injected states, taken with the
 * device functions the transition kernels use (LOGISTIC: the tensor-core gradient kernel).  HOST arrays: x, p [n][dim]
 * unconstrained positions / momenta, eps [n] step sizes, g [n][dim] the cached gradient at x (NULL: recomputed at x).
 * last != 0 runs the variant a trajectory uses on its last step and returns the log-density at x' in lp_out [n];
 * last == 0 the interior-step variant (lp_out = 0).  x_out, p_out, g_out [n][dim]: position, momentum and gradient after
 * the step.  Replaying the oracle's recorded states through this makes every leapfrog step of an adaptive run comparable. */
int qsb_debug_leapfrog(qsb_model* m, int32_t n, const double* x, const double* p, const double* g, const double* eps, int32_t last,
                       double* x_out, double* p_out, double* g_out, double* lp_out);
/* log-density and gradient of the model at n points x (HOST [n][dim], unconstrained), evaluated by the
 * device functions the kernels use: lp_dual (with log-Jacobians, what HMC sees: erp.t:623-640), its
 * gradient grad [n][dim], and lp_float (what Drift/HARM see: no Jacobian, erp.t:361). */
int qsb_logp_grad(qsb_model* m, const double* x, int32_t n, double* lp_dual, double* grad, double* lp_float);
/* Device density functions, same selector as the oracle: 1 uniform(x,lo,hi) 2 gaussian(x,mu,sigma)
 * 3 gamma(x,k,theta) 4 beta(x,a,b) 9 log_gamma(x) 0 bernoulli(val,p); n tuples of 3 doubles. */
int qsb_logprob(int32_t which, const double* args, int32_t n, double* out);
/* The injected uniform stream (replaces qs/lib/random.t:5,11) and the Leva Gaussian sampler on it
 * (qs/distrib.t:61-74), evaluated on the device. */
int qsb_uniforms(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot, int32_t n, double* out);
int qsb_gaussians(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot0, int32_t n, double mu, double sigma, double* out);

#ifdef __cplusplus
}
#endif
#endif /* QSB200_H */
